"""Render / loss engine: drop-in for T_NeRF_Full_2/Eval_Tools_2.py (get_PV :13-16, create_solor_rays_uniform :42-108,
All_in_One_Eval :111-459) and misc.py:234-261 (sample_pt_coarse, zero_invalid_pts).

Same constructor, methods, result-dict keys and `{name: [value, weight]}` loss dict as the reference.  Inputs may
arrive as CPU float32 tensors exactly like the reference's DataLoader rows: this shim owns the H2D copy.  Sampling,
the network and compositing run in the sm_100a library; only the O(N) loss arithmetic stays in torch.
Extra keyword arguments (`jitter=`, `solar=`, `solar_jitter=`) inject the random draws of the reference for parity
tests; when omitted the same draws are made from the same global RNGs in the same order as the reference.
"""
import os

import numpy as np
import torch as t

from . import ops
from .adaptive_loss import AdaptiveLossFunction as _OwnAdaptiveLoss
from .geometry import world_angle_2_local_vec

# get_loss without the prior: head activations + compositing + the solar-pass sums run in fused kernels on the RAW heads
# (ops.heads_composite / ops.solar_loss) instead of ~65 element-wise torch launches and their autograd nodes.
# SNB_FAST_LOSS=0 keeps the general path (eval / eval_Rho_Only dictionaries) for A/B tests.
FAST_LOSS = os.environ.get("SNB_FAST_LOSS", "1") != "0"
# The loss terms that only read the image pass (adaptive colour loss, sky / albedo regularisers: ~200 launches of a few
# microseconds, forward and again in backward) are enqueued on a side stream while the solar pass - 3 ms of GEMMs that do not
# depend on them - runs on the main one; inside the captured step graph this is a fork / join.  SNB_LOSS_OVERLAP=0 keeps one stream.
LOSS_OVERLAP = os.environ.get("SNB_LOSS_OVERLAP", "1") != "0"
# One kernel each way for ALL the O(N) loss terms of the default configuration (Barron colour loss, solar terms, no prior,
# one rank per BatchNorm batch): csrc/loss.cu.  SNB_FUSED_TAIL=0 keeps the torch arithmetic (parity tests compare the two).
FUSED_TAIL = os.environ.get("SNB_FUSED_TAIL", "1") != "0"


def _dev(x, device):
    return x.to(device=device, dtype=t.float32, non_blocking=True)


def sample_ts(n_course, eval_mode, include_end_pt=False, jitter=None):
    """misc.py:236-241, on the host exactly like the reference (torch CPU linspace / rand) -> float32 [n]."""
    if include_end_pt is False or eval_mode is False:
        ts = t.linspace(0, 1, n_course + 1)[0:-1].clone()
    else:
        ts = t.linspace(0, 1, n_course)
    if eval_mode is False:
        r = t.rand(n_course) if jitter is None else t.as_tensor(jitter, dtype=t.float32).cpu()
        ts += 1 / n_course * r
    return ts


def sample_pt_coarse(pt_tops, pt_bots, n_course, eval_mode, include_end_pt=False, jitter=None, device=None, ts=None):
    """misc.py:234-247 -> pts [N,S,3], deltas [N,S,1] (on the device).  `ts` (device float32 [S]) overrides the host-built
    sample fractions: the CUDA-graph training step keeps them in a static device buffer refreshed before each replay."""
    device = device or (pt_tops.device if pt_tops.is_cuda else t.device("cuda"))
    if ts is None:
        ts = sample_ts(n_course, eval_mode, include_end_pt, jitter).to(device)
    pts, deltas = ops.sample_rays(_dev(pt_tops, device), _dev(pt_bots, device), ts)
    return pts, deltas.unsqueeze(-1)


class zero_invalid_pts():
    """misc.py:249-261."""

    def __init__(self, X=(-1, 1), Y=(-1, 1), Z=(-1, 1)):
        self.X, self.Y, self.Z = X, Y, Z

    def __call__(self, Xs):
        return ~((Xs <= 1).all(-1) & (Xs >= -1).all(-1))


def get_PV(Rhos, Deltas):
    """Eval_Tools_2.py:13-16 on [N,S,1] tensors (exclusive-prefix transmittance)."""
    if not Rhos.is_cuda:
        raise ops._lib.SeasonNerfCudaError("season_nerf_b200.get_PV needs CUDA tensors (no CPU fallback)")
    N, S = Rhos.shape[0], Rhos.shape[1]
    rho = Rhos.reshape(N, S).float().contiguous()
    dl = Deltas.reshape(N, S).float().contiguous()
    z3 = t.zeros(N, S, 3, device=rho.device)
    PV = ops.composite_fwd(rho, dl, z3, t.zeros_like(rho), t.zeros(N, 3, device=rho.device))[0]
    return PV.reshape(N, S, 1).to(Rhos.dtype)


def owns_global_min(local_min):
    """1.0 where this rank's per-channel batch minimum is the minimum over all data-parallel ranks, else 0.0 (no gradient)"""
    import torch.distributed as dist
    g_min = local_min.detach().clone()
    dist.all_reduce(g_min, op=dist.ReduceOp.MIN)
    return (local_min.detach() == g_min).float()


class create_solor_rays_uniform():
    """Eval_Tools_2.py:42-108 with the per-ray Python loop vectorised (same RNG draws in the same order)."""

    def __init__(self, W2L_H, WCW, base_vecs=None):
        self.W2L = W2L_H
        self.WC = WCW
        self._base_vecs = base_vecs
        self._use_base_vecs = False      # the reference overrides this to False (:48)

    def __call__(self, n, include_times=False):
        az_el = np.random.random(n * 2).reshape([n, 2]) * np.array([[360, 89]]) + np.array([[-180, 1]])
        vec = world_angle_2_local_vec(az_el[:, 1], az_el[:, 0], self.WC, self.W2L)
        delta = 2 * (vec / vec[:, 2::])
        starts = t.ones([n, 3])
        starts[:, 0] = t.tensor(2.0) * t.rand(n) + t.tensor(-1.0)
        starts[:, 1] = t.tensor(2.0) * t.rand(n) + t.tensor(-1.0)
        ends = (starts - t.tensor(delta)).float()
        vec = t.tensor(vec).float()
        if include_times is False:
            return starts, ends, vec
        f = t.rand([n, 2]) * 2 * np.pi
        times = t.stack([t.cos(f[:, 0]), t.sin(f[:, 0]), t.cos(f[:, 1]), t.sin(f[:, 1])], 1)
        return starts, ends, vec, times, az_el

    def on_device(self, n, device, include_times=False, generator=None, out=None):
        """Same distribution as __call__, drawn and built ON THE DEVICE: azimuth / elevation (float64), the start
        positions and the year / day fractions come from torch's CUDA generator (NOT the reference's numpy / CPU-torch
        streams: a run is reproducible under torch.cuda.manual_seed, not draw-identical to the reference), the geometry
        of :75-87 runs in one kernel (ops.solar_rays).  Nothing is computed on the host and nothing is copied.
        -> starts, ends, vec (, times) float32 device tensors; `out` = preallocated tensors to fill."""
        device = t.device(device)
        sc = self.__dict__.get("_dev_scale")
        if sc is None or sc[0] != device:
            # built from scalar fills: no host->device copy, so the first call may happen inside a CUDA-graph capture
            mul, add = t.empty(1, 2, dtype=t.float64, device=device), t.empty(1, 2, dtype=t.float64, device=device)
            mul[:, 0].fill_(360.), mul[:, 1].fill_(89.), add[:, 0].fill_(-180.), add[:, 1].fill_(1.)
            sc = self._dev_scale = (device, mul, add)
        az_el = t.rand(n, 2, dtype=t.float64, device=device, generator=generator).mul_(sc[1]).add_(sc[2])
        u_xy = t.rand(n, 2, device=device, generator=generator)
        u_t = t.rand(n, 2, device=device, generator=generator) if include_times else None
        return ops.solar_rays(self.WC, self.W2L, az_el, u_xy, u_t, out=out)

    def create_given_vec(self, n, solar_angle_vec, include_times=False):
        delta = 2 * (solar_angle_vec / solar_angle_vec[2::])
        starts = t.ones([n, 3])
        starts[:, 0] = t.tensor(2.0) * t.rand(n) + t.tensor(-1.0)
        starts[:, 1] = t.tensor(2.0) * t.rand(n) + t.tensor(-1.0)
        ends = (starts - t.tensor(np.expand_dims(delta, 0))).float()
        vec = t.stack([t.tensor(solar_angle_vec).float()] * n, 0)
        if include_times is False:
            return starts, ends, vec
        f = t.rand([n, 2]) * 2 * np.pi
        times = t.stack([t.cos(f[:, 0]), t.sin(f[:, 0]), t.cos(f[:, 1]), t.sin(f[:, 1])], 1)
        return starts, ends, vec, times


class All_in_One_Eval():
    def __init__(self, args, device, n_steps, use_prior, ada_loss, H, WC, base_solar_vecs=None):
        self.device = t.device(device)
        if self.device.type != "cuda":
            raise ops._lib.SeasonNerfCudaError("season_nerf_b200.All_in_One_Eval needs a CUDA device (no CPU fallback)")
        self.args = args
        self.n_steps = n_steps
        self.MSE_loss = t.nn.MSELoss()
        self.smooth_L1 = t.nn.SmoothL1Loss()
        self.use_prior = use_prior
        self.use_reg = args.Use_Reg
        self.use_classic_solar = args.Solar_Type_2
        self.use_MSE_loss = args.Use_MSE_loss
        self.ada_loss = ada_loss
        self.solar_creation_tool = create_solor_rays_uniform(H, WC, base_solar_vecs)
        # False (default): solar rays are drawn on the host from the reference's RNG streams, in the reference's order.
        # True: drawn and built on the device (create_solor_rays_uniform.on_device) - no host work, no H2D copy.
        self.solar_on_device = False
        # > 1: the batch of a step is spread over this many data-parallel ranks (train.TrainStep(sync_bn=True)); the one loss
        # term that is not a mean over rays - the per-channel minimum of Albedo_Color (:374-377) - is then taken over all ranks
        self.sync_world = 1
        self.Sigmoid = t.nn.Sigmoid()
        self.BCE_loss = t.nn.BCELoss()

    # ---------------------------------------------------------------------------------------------------
    def _shade(self, Rho, deltas, Col, Vis, Sky_ray):
        """PV/PE/PS + colour for one density field (Eval_Tools_2.py:187-215)."""
        N, S = Rho.shape[0], Rho.shape[1]
        PV, PE, PS, albedo, rendered, _ = ops.composite(Rho.reshape(N, S), deltas.reshape(N, S), Col,
                                                        Vis.reshape(N, S), Sky_ray, self.use_classic_solar)
        return PV.unsqueeze(-1), PE.unsqueeze(-1), PS.unsqueeze(-1), albedo, rendered

    def eval(self, data_dict, Network, current_step, train_mode, jitter=None, ts=None):
        """Eval_Tools_2.py:165-252."""
        S = self.args.n_samples
        dev = self.device
        Xs, deltas = sample_pt_coarse(data_dict["Top"], data_dict["Bot"], S, not train_mode, jitter=jitter, device=dev, ts=ts)
        N = Xs.shape[0]
        sun, tim = _dev(data_dict["Sun_Angle"], dev), _dev(data_dict["Time_Encoded"], dev)
        pos, vis, adj, sky, cl = Network.forward_rays(Xs.reshape(-1, 3), sun, tim, S)
        Rho, Col, Vis, Sky, Cls, Adj = Network._mix(pos, vis, adj, sky, cl, S, True, True)
        Sky_ray = Network.Sigmoid(sky)
        Col, Rho, Vis = Col.reshape(N, S, -1), Rho.reshape(N, S, 1), Vis.reshape(N, S, 1)
        Sky, Cls, Adj = Sky.reshape(N, S, -1), Cls.reshape(N, S, -1), Adj.reshape(N, S, -1)
        PV, PE, PS, Albedo, Rendered = self._shade(Rho, deltas, Col, Vis, Sky_ray)
        R = {"Rendered_Col": Rendered, "PE": PE, "PV": PV, "PS": PS, "Solar_Vis": Vis, "Sky_Col": Sky, "Classes": Cls,
             "Adjust": Adj, "Rho": Rho, "Col": Col, "Col_Adj": -1, "deltas": deltas, "sample_pts": Xs,
             "Albedo_Color": Albedo}
        if self.use_prior:
            trust = self._trust(current_step)
            with t.no_grad():
                Rho_S = Network.Supervised_Sample(Xs.reshape(-1, 3), deltas.reshape(-1, 1)).reshape(N, S, 1)
            PV_S, PE_S, PS_S, Albedo_S, Rend_S = self._shade(Rho_S, deltas, Col, Vis, Sky_ray)
            Rho_M = Rho * trust + Rho_S * (1 - trust)
            PV_M, PE_M, PS_M, Albedo_M, Rend_M = self._shade(Rho_M, deltas, Col, Vis, Sky_ray)
            if not self.use_classic_solar:
                # Eval_Tools_2.py:214,231,243: Solar_Vis3 is computed ONCE, from the network's own (unmerged) PS, and shades
                # all three colours; its gradient reaches rho through that PS only
                sv3 = self.Sigmoid((t.sum(Vis.detach() * PS, 1) - .2) * 30)
                shade = sv3 + (1 - sv3) * Sky_ray
                Rend_S, Rend_M = Albedo_S * shade, Albedo_M * shade
            R.update({"PV_Supervised": PV_S, "PE_Supervised": PE_S, "PS_Supervised": PS_S,
                      "Rendered_Col_Supervised": Rend_S, "PV_Merged": PV_M, "PE_Merged": PE_M, "PS_Merged": PS_M,
                      "Rendered_Col_Merged": Rend_M, "Rho_Merged": Rho_M, "Albedo_Color": Albedo_M})
        return R

    def full_eval(self, data_dict, Network, current_step):
        """Eval_Tools_2.py:127-163 (eval-mode sampling, raw colour activated here)."""
        S = self.args.n_samples
        dev = self.device
        Xs, deltas = sample_pt_coarse(data_dict["Top"], data_dict["Bot"], S, True, device=dev)
        N = Xs.shape[0]
        sun, tim = _dev(data_dict["Sun_Angle"], dev), _dev(data_dict["Time_Encoded"], dev)
        pos, vis, adj, sky, cl = Network.forward_rays(Xs.reshape(-1, 3), sun, tim, S)
        Rho, Base, Vis, Sky, Cls, Adj = Network._mix(pos, vis, adj, sky, cl, S, True, True)
        Sky_ray = Network.Sigmoid(sky)
        Base, Rho, Vis = Base.reshape(N, S, -1), Rho.reshape(N, S, 1), Vis.reshape(N, S, 1)
        # the reference applies Sigmoid to the already-activated colour here (:155,158)
        PV, PE, PS, _, Rendered = self._shade(Rho, deltas, self.Sigmoid(Base), Vis, Sky_ray)
        return {"Rendered_Col": Rendered, "PE": PE, "PV": PV, "PS": PS, "Solar_Vis": Vis,
                "Sky_Col": Sky.reshape(N, S, -1), "Classes": Cls.reshape(N, S, -1), "Adjust": Adj.reshape(N, S, -1),
                "Rho": Rho, "Col": Base, "deltas": deltas, "sample_pts": Xs}

    def eval_Rho_Only(self, data_dict, Network, train_mode, current_step=0, jitter=None, ts=None):
        """Eval_Tools_2.py:297-337."""
        S = self.args.n_samples
        dev = self.device
        Xs, deltas = sample_pt_coarse(data_dict["Top"], data_dict["Bot"], S, not train_mode, include_end_pt=True,
                                      jitter=jitter, device=dev, ts=ts)
        N = Xs.shape[0]
        sun = _dev(data_dict["Sun_Angle"], dev)
        rho_raw, vis_raw, sky_raw = Network.forward_rays(Xs.reshape(-1, 3), sun, None, S, mode="solar")
        Rho = Network.Softplus(rho_raw).reshape(N, S, 1)
        Vis = Network.Sigmoid(vis_raw).reshape(N, S, 1)
        Sky = sky_raw.repeat_interleave(S, 0).reshape(N, S, -1)     # RAW sky (T_NeRF_net_v2.py:157)
        if self.use_prior:
            trust = self._trust(current_step)
            with t.no_grad():
                # Eval_Tools_2.py:319-334: the prior-DSM density replaces rho only at points inside the cube.  Shape-static
                # form of the reference's boolean indexing (CUDA-graph capturable): look every point up with clamped
                # coordinates, keep the look-up where the point is inside
                Xs2, d2 = Xs.reshape(-1, 3), deltas.reshape(-1, 1)
                good = t.all((Xs2 <= 1.) * (Xs2 >= -1.), 1, keepdim=True)
                sup = Network.Supervised_Sample(t.clamp(Xs2, -1., 1.), d2)
                Rho_S = t.where(good, sup, Rho.reshape(-1, 1).detach()).reshape(N, S, 1)
            Rho = Rho * trust + Rho_S * (1 - trust)
        PE = 1 - t.exp(-Rho * deltas)
        PV = get_PV(Rho.detach(), deltas) if not Rho.requires_grad else self._pv_autograd(Rho, deltas)
        return {"PE": PE, "PV_Exact": PV, "Solar_Vis": Vis, "Sky_Col": Sky}

    def _eval_fast(self, data_dict, Network, train_mode, jitter=None, ts=None):
        """eval() reduced to what get_loss reads without the prior (Rendered_Col, Albedo_Color, the per-ray sky colour), from
        the raw heads in ONE kernel (Eval_Tools_2.py:165-215)."""
        S = self.args.n_samples
        dev = self.device
        Xs, deltas = sample_pt_coarse(data_dict["Top"], data_dict["Bot"], S, not train_mode, jitter=jitter, device=dev, ts=ts)
        N = Xs.shape[0]
        sun, tim = _dev(data_dict["Sun_Angle"], dev), _dev(data_dict["Time_Encoded"], dev)
        pos, vis, adj, sky, cl = Network.forward_rays(Xs.reshape(-1, 3), sun, tim, S)
        if sky.shape[0] != N:
            sky = sky.expand(N, sky.shape[1])
        if cl.shape[0] != N:
            cl = cl.expand(N, cl.shape[1])
        albedo, rendered, sky_act, _ = ops.heads_composite(pos, vis, adj, sky, cl, deltas.reshape(N, S), self.use_classic_solar)
        return {"Rendered_Col": rendered, "Albedo_Color": albedo, "Sky_Col": sky_act}

    def _solar_fast(self, data_dict, Network, train_mode, jitter=None, ts=None):
        """eval_Rho_Only + the two solar sums of get_loss (Eval_Tools_2.py:297-337, :353-368) -> err [N], absorb [N]"""
        S = self.args.n_samples
        dev = self.device
        Xs, deltas = sample_pt_coarse(data_dict["Top"], data_dict["Bot"], S, not train_mode, include_end_pt=True,
                                      jitter=jitter, device=dev, ts=ts)
        N = Xs.shape[0]
        sun = _dev(data_dict["Sun_Angle"], dev)
        rho_raw, vis_raw, _ = Network.forward_rays(Xs.reshape(-1, 3), sun, None, S, mode="solar")
        return ops.solar_loss(rho_raw.detach(), vis_raw, deltas.reshape(N, S))

    def _trust(self, current_step):
        """trust = step / n_steps (Eval_Tools_2.py:218); inside a CUDA-graph capture the device scalar
        `trust_tensor` (set by the graphed training step, refreshed before each replay) stands in for the Python number"""
        tt = getattr(self, "trust_tensor", None)
        if tt is not None and t.cuda.is_current_stream_capturing():
            return tt
        return current_step / self.n_steps

    def _pv_autograd(self, Rho, deltas):
        N, S = Rho.shape[0], Rho.shape[1]
        z = t.zeros(N, S, device=Rho.device)
        PV = ops.composite(Rho.reshape(N, S), deltas.reshape(N, S), t.zeros(N, S, 3, device=Rho.device), z,
                           t.zeros(N, 3, device=Rho.device), False)[0]
        return PV.unsqueeze(-1)

    def _get_exact_solar(self, world_pts, sun_angle, Network):
        """Eval_Tools_2.py:255-269 for all sample points of ONE ray."""
        dev = self.device
        world_pts = _dev(world_pts, dev)
        sun_angle = _dev(sun_angle, dev)
        n = world_pts.shape[0]
        tops = ops.solar_tops(world_pts, sun_angle.tolist(), f64=False)
        sub = {"Top": tops, "Bot": world_pts, "Sun_Angle": sun_angle.reshape(1, 3).expand(n, 3),
               "Time_Encoded": t.ones(n, 4)}
        r = self.eval_Rho_Only(sub, Network, False)
        return r["PV_Exact"][:, -1], r["Solar_Vis"][:, -1]

    def eval_exact_solar(self, data_dict, Network, current_step, train_mode):
        """Eval_Tools_2.py:273-295; the reference's per-ray Python loop is batched over all rays of the call."""
        R = self.eval(data_dict, Network, current_step, train_mode)
        R["Est_Solar_Vis"] = R["Solar_Vis"].clone()
        dev = self.device
        pts = R["sample_pts"]
        N, S = pts.shape[0], pts.shape[1]
        sun = _dev(data_dict["Sun_Angle"], dev)
        with t.no_grad():
            sun_pts = sun.repeat_interleave(S, 0)
            p = pts.reshape(-1, 3)
            K = (1 - p[:, 2]) / sun_pts[:, 2]                                   # :257 (float32)
            tops = p + K.unsqueeze(1) * sun_pts                                 # :258
            sub = {"Top": tops, "Bot": p, "Sun_Angle": sun_pts, "Time_Encoded": t.ones(p.shape[0], 4)}
            r = self.eval_Rho_Only(sub, Network, False)
            R["Solar_Vis"] = r["PV_Exact"][:, -1].reshape(N, S, 1)
        R["Col_Adj"] = (R["Solar_Vis"] + (1 - R["Solar_Vis"]) * R["Sky_Col"]) * R["Col"]
        if self.use_classic_solar:
            R["Rendered_Col"] = t.sum(R["PS"] * R["Col"] * (R["Solar_Vis"] + (1 - R["Solar_Vis"]) * R["Sky_Col"]), 1)
        else:
            sv3 = self.Sigmoid((t.sum(R["Solar_Vis"] * R["PS"], 1) - .2) * 30)
            R["Rendered_Col"] = t.sum(R["PS"] * R["Col"], 1) * (sv3 + (1 - sv3) * t.mean(R["Sky_Col"], 1))
        return R

    # ---------------------------------------------------------------------------------------------------
    def _fused_tail_usable(self, train_mode):
        """the fused loss kernels cover the default configuration: Barron colour loss of THIS package (the quadrature buffers
        and the latent parameterisation are read directly), solar terms on, no prior DSM, no cross-rank albedo minimum"""
        a = self.args
        return bool(FUSED_TAIL and a.Use_Solar and not self.use_prior and self.use_MSE_loss is not True and self.sync_world == 1
                    and isinstance(self.ada_loss, _OwnAdaptiveLoss) and self.ada_loss.latent_alpha.shape[-1] == 3)

    def _image_terms(self, out, gt, train_mode):
        """the terms of get_loss that read the image pass only (Eval_Tools_2.py:370-443), keyed for get_loss's assembly:
        regularisers Sky_Color_Var / Albedo_Color (:370-390), colour terms (:401-443), `scale` = mean colour-loss scale ** 2"""
        args = self.args
        T = {}
        if args.Use_Solar and args.Solar_Type_2 is False:
            alb = out["Albedo_Color"]
            sk_alb, _ = t.min(alb, 0)
            m = (sk_alb < .2).float()           # branch-free: no host sync (.item() stalls)
            if self.sync_world > 1 and train_mode:
                # global minimum: only the rank that owns it contributes (the mean over ranks of the gradient
                # all-reduce then equals the single-batch term  sum_c (1 - min_c/.2)^2 / N_total)
                m = m * owns_global_min(sk_alb)
            alb_loss = t.sum(m * (1. - sk_alb / .2) ** 2) / alb.shape[0]
            # Sky_Col is [N,S,3] in the reference (one colour per ray, repeated over its samples); the fast path keeps
            # the [N,3] rows: sum / numel is the same mean
            sk = (out["Sky_Col"] - .5) / .5
            sk_loss = t.sum(t.relu(sk) ** 2) / float(np.prod(sk.shape))
            if self.use_prior:
                sk_loss = sk_loss.detach()
            T["Sky_Color_Var"], T["Albedo_Color"] = sk_loss, alb_loss
        merged = "Rendered_Col_Merged" if (self.use_prior and train_mode) else "Rendered_Col"
        if self.use_MSE_loss is True:
            T["Color"] = self.MSE_loss(out[merged], gt)
            if self.use_prior:
                T["Alpha_Adjust"] = self.MSE_loss(out["PE"], out["PE_Supervised"].detach())
        else:
            diff = out["Rendered_Col"] - gt
            if self.use_prior:
                a0, a1 = self.ada_loss[0], self.ada_loss[1]
                adiff = (out["PE"] - out["PE_Supervised"].detach()).reshape([-1, 1])
                T["Alpha_Adjust_ada"] = t.mean(a1.lossfun(adiff))
                T["Color_ada"] = t.mean(a0.lossfun(diff))
                T["Color_alpha"] = t.mean(a0.alpha().detach())
                T["Color_width"] = t.mean(a0.scale().detach())
                T["Alpha_Adjust_mse"] = self.MSE_loss(out["PE"], out["PE_Supervised"].detach())
                T["scale"] = t.mean(a0.scale().detach()) ** 2
                T["Alpha_alpha"] = t.mean(a1.alpha().detach())
                T["Alpha_width"] = t.mean(a1.scale().detach())
            else:
                a0 = self.ada_loss
                T["Color_ada"] = t.mean(a0.lossfun(diff))
                T["Color_alpha"] = t.mean(a0.alpha().detach())
                T["Color_width"] = t.mean(a0.scale().detach())
                T["scale"] = t.mean(a0.scale().detach()) ** 2
            with t.no_grad():
                T["Color_mse"] = self.MSE_loss(out[merged], gt).detach()
        return T

    def get_loss(self, data_dict, Network, current_step, train_mode, jitter=None, solar=None, solar_jitter=None,
                 ts=None, solar_ts=None):
        """Eval_Tools_2.py:340-459."""
        n_rays = data_dict["Top"].shape[0]
        device = self.device
        args = self.args
        Loss = {}
        weight = {"Color": 1.0, "Solar_Correction": args.sc_lambda, "Alpha_Adjust": 1.}
        from . import network as _nw
        overlap = (_nw.OVERLAP and args.Use_Solar and train_mode and not self.use_prior and n_rays >= 64
                   and hasattr(Network, "stage_weights") and getattr(Network, "precision", None) == "bf16")
        if overlap:
            # The image pass and the solar pass are independent until the losses: they run on two streams, so the
            # tensor-bound GEMMs of one overlap the HBM-bound activation passes of the other.  The solar pass is enqueued
            # second (BatchNorm running statistics keep the reference's update order through per-module events) but only
            # waits for the point where the weights were staged.
            Network.stage_weights()
            Network._bn_order = {}
            main = t.cuda.current_stream()
            fork = t.cuda.Event()
            fork.record(main)
        fast = (FAST_LOSS and not self.use_prior and not overlap and hasattr(Network, "forward_rays")
                and ops.heads_composite_usable(args.n_samples, getattr(Network, "n_classes", 99)))
        sol = sol_fast = img_terms = None
        fused_tail = False
        self._fused_total = None
        try:
            if fast:
                out = self._eval_fast(data_dict, Network, train_mode, jitter=jitter, ts=ts)
                fused_tail = self._fused_tail_usable(train_mode)
                if LOSS_OVERLAP and train_mode and args.Use_Solar and n_rays >= 64 and not fused_tail:
                    gt = _dev(data_dict["GT_Color"], device)
                    main_s = t.cuda.current_stream()
                    side_s = _nw.side_stream(("loss", main_s.cuda_stream))
                    side_s.wait_stream(main_s)
                    for v in list(out.values()) + [gt]:
                        v.record_stream(side_s)
                    with t.cuda.stream(side_s):
                        img_terms = self._image_terms(out, gt, train_mode)
                    join = (main_s, side_s)
            else:
                out = self.eval(data_dict, Network, current_step, train_mode, jitter=jitter, ts=ts)
            if args.Use_Solar:
                if solar is None and self.solar_on_device:
                    starts, ends, svec, stime = self.solar_creation_tool.on_device(n_rays, device, include_times=True)
                elif solar is None:
                    starts, ends, svec, stime, _ = self.solar_creation_tool(n_rays, include_times=True)
                else:
                    starts, ends, svec, stime = solar
                sdict = {"Top": starts, "Bot": ends, "Sun_Angle": svec, "Time_Encoded": stime}
                if fast:
                    sol_fast = self._solar_fast(sdict, Network, train_mode, jitter=solar_jitter, ts=solar_ts)
                elif overlap:
                    side = _nw.side_stream(("solar", main.cuda_stream))
                    side.wait_event(fork)
                    with t.cuda.stream(side):
                        sol = self.eval_Rho_Only(sdict, Network, train_mode, current_step, jitter=solar_jitter, ts=solar_ts)
                    main.wait_stream(side)
                    for v in sol.values():
                        if isinstance(v, t.Tensor):
                            v.record_stream(main)
                else:
                    sol = self.eval_Rho_Only(sdict, Network, train_mode, current_step, jitter=solar_jitter, ts=solar_ts)
        finally:
            if overlap:
                Network._bn_order = None
        if img_terms is not None:        # join: the main stream consumes the side stream's scalars from here on
            join[0].wait_stream(join[1])
            for v in img_terms.values():
                v.record_stream(join[0])
        gt = _dev(data_dict["GT_Color"], device)
        if fused_tail and sol_fast is not None:
            # every O(N) term, its weight and the weighted total in ONE kernel (and one for all their gradients)
            a0 = self.ada_loss
            T = ops.loss_tail(out["Rendered_Col"], gt, out["Albedo_Color"], out["Sky_Col"], sol_fast[0], sol_fast[1], a0.alpha(),
                              a0.scale(), a0._th, a0._w, args.sc_lambda, bool(args.Solar_Type_2))
            Loss["Solar_Correction"] = [T["Solar_Correction"], T["solar_weight"]]
            Loss["Solar_Correction_2"] = [T["Solar_Correction_2"] if args.Solar_Type_2 else T["Solar_Correction_2"].detach(),
                                          T["solar_weight"]]
            if args.Solar_Type_2 is False:
                Loss["Sky_Color_Var"] = [T["Sky_Color_Var"], weight["Solar_Correction"]]
                Loss["Albedo_Color"] = [T["Albedo_Color"], weight["Solar_Correction"]]
            Loss["Color_ada"] = [T["Color_ada"], weight["Color"]]
            Loss["Color_alpha"] = [T["Color_alpha"], 1.]
            Loss["Color_width"] = [T["Color_width"], 1.]
            Loss["Color"] = [T["Color"], weight["Color"]]
            self._fused_total = T["total"]       # = sum of term * weight over the dictionary (train.TrainStep back-propagates it)
            return Loss
        img_terms = self._image_terms(out, gt, train_mode) if img_terms is None else img_terms
        if args.Use_Solar:
            if sol_fast is not None:
                err, absorb = t.mean(sol_fast[0]), t.mean(sol_fast[1])
            else:
                err = t.mean(t.sum((sol["Solar_Vis"] - sol["PV_Exact"].detach()) ** 2, 1))
                absorb = t.mean(1 - t.sum(sol["PE"].detach() * sol["PV_Exact"].detach() * sol["Solar_Vis"], 1))
            Loss["Solar_Correction"] = [err, weight["Solar_Correction"]]
            Loss["Solar_Correction_2"] = [absorb if args.Solar_Type_2 else absorb.detach(), weight["Solar_Correction"]]
            if args.Solar_Type_2 is False:
                Loss["Sky_Color_Var"] = [img_terms["Sky_Color_Var"], weight["Solar_Correction"]]
                Loss["Albedo_Color"] = [img_terms["Albedo_Color"], weight["Solar_Correction"]]
        for k in ("Color", "Alpha_Adjust", "Alpha_Adjust_ada", "Color_ada", "Color_alpha", "Color_width", "Alpha_Adjust_mse",
                  "Alpha_alpha", "Alpha_width", "Color_mse"):
            if k in img_terms:
                name = {"Alpha_Adjust_mse": "Alpha_Adjust", "Color_mse": "Color"}.get(k, k)
                Loss[name] = [img_terms[k], {"Color": weight["Color"], "Color_ada": weight["Color"], "Color_mse": weight["Color"],
                                             "Alpha_Adjust": weight["Alpha_Adjust"], "Alpha_Adjust_ada": weight["Alpha_Adjust"],
                                             "Alpha_Adjust_mse": weight["Alpha_Adjust"]}.get(k, 1.)]
        if "scale" in img_terms and args.Use_Solar:
            Loss["Solar_Correction"][1] = Loss["Solar_Correction"][1] / img_terms["scale"]
            Loss["Solar_Correction_2"][1] = Loss["Solar_Correction_2"][1] / img_terms["scale"]
        return Loss
