"""Drop-in `T_NeRF` network (reference: T_NeRF_Full_2/T_NeRF_net_v2.py:20-204, T_NeRF_Full_2/G_NeRF.py:6-157,
misc.py:105-194) whose dense work runs in the sm_100a library.

Same constructor, same module tree and therefore the same 94 state_dict keys / shapes, same public methods and
return tuples as the reference, `.train()/.eval()` switch BatchNorm semantics, parameters are ordinary
`nn.Parameter`s usable by `torch.optim`, and every differentiable output supports autograd.  The modules are
parameter containers: the arithmetic is scheduled by `_Pass` below as hand-written CUDA kernels
(GEMM -> column statistics -> sin(BN(.)) per layer, with a hand-scheduled backward) or, in eval mode on whole
rays, by the fused tcgen05 kernel (season_nerf_b200/fused.py).

precision: "bf16" (production: bf16 operands, fp32 accumulate on tcgen05) or "fp32" (validation build).
"""
from math import sqrt

import numpy as np
import torch as t
from torch import nn

from . import ops

OMEGA_0 = 30.0
# activation work fused into the tcgen05 GEMM epilogues (forward sin of layers without BatchNorm; cos / BatchNorm-backward
# reductions in the input-gradient GEMM).  SNB_FUSE_EPILOGUES=0 keeps the stand-alone element-wise kernels (A/B tests).
import os as _os
FUSE_EPILOGUES = _os.environ.get("SNB_FUSE_EPILOGUES", "1") != "0"
# Optional: independent work on parallel CUDA streams (weight-gradient GEMMs next to the input-gradient chain; the solar
# pass next to the image pass, see engine.get_loss).  Measured on B200 (scripts/overlap_probe.py, bench.py): a trunk GEMM
# and an activation pass that share the SMs finish in 290 us instead of 303 us back to back - both stream 403 MB matrices
# and the step is HBM-bound - and the whole step gains 0.6 %.  Off by default (SNB_OVERLAP=1 enables).
OVERLAP = _os.environ.get("SNB_OVERLAP", "0") == "1"
# Consumer-side activation: a train-mode BatchNorm layer whose output feeds exactly one full-width BatchNorm layer (and whose
# activated output nobody else needs - the no-grad trunk of the solar pass) skips its sin pass; the NEXT layer's GEMM applies
# sin(a*z + c) to the operand tile in shared memory (ops.gemm_stats_xf).  SNB_XFORM=0 keeps the stand-alone pass (A/B tests).
XFORM = _os.environ.get("SNB_XFORM", "1") != "0"
XFORM_GRAD = _os.environ.get("SNB_XFORM_GRAD", "1") != "0"        # also in passes with a backward (activated operand written back)
_side_streams = {}
_consts = {}
BULK_STAGE = _os.environ.get("SNB_BULK_STAGE", "1") != "0"      # one launch for the bf16 copies of all weights of a pass
_capture_epoch = [0]        # bumped by train.TrainStep before every CUDA-graph capture (see _Pass._wc)


def const_vec(n, value, device):
    """cached read-only float32 [n] vector of `value` (identity affine of layers without BatchNorm, zero offsets): created
    once per (n, value, device) instead of a fill launch per layer and pass"""
    key = (int(n), float(value), str(device))
    v = _consts.get(key)
    if v is None:
        v = t.full((int(n),), float(value), device=device, dtype=t.float32)
        if not t.cuda.is_current_stream_capturing():      # a vector born inside a graph capture is only filled by replays
            _consts[key] = v
    return v


def _sync_world(net):
    """world size of a BatchNorm-synchronised data-parallel run (train.TrainStep(sync_bn=True)), else 1"""
    w = getattr(net, "_sync_bn", None)
    return int(w) if w else 1


def _allreduce_pair(a, b):
    """column statistics of all ranks: one small all-reduce (sum) of the stacked pair -> (a_all, b_all)"""
    import torch.distributed as dist
    pair = t.stack([a, b])
    dist.all_reduce(pair, op=dist.ReduceOp.SUM)
    return pair[0], pair[1]


def side_stream(key):
    """one lazily created side stream per (purpose, parent stream)"""
    st = _side_streams.get(key)
    if st is None:
        st = _side_streams[key] = t.cuda.Stream()
    return st


def _r8(x):
    return (x + 7) // 8 * 8


class PE_Encode(nn.Module):
    """misc.py:105-139 (parameter-free).  Kept as a module so `net.G_NeRF_net.PE_encoder(X)` keeps working."""

    def __init__(self, n, use_Extend_encoding, scale=np.pi / 2):
        super().__init__()
        self.n = n
        self.use_Extended = use_Extend_encoding

    def forward(self, X):
        if not self.use_Extended:
            raise NotImplementedError("season_nerf_b200 implements the extended encoding the reference network uses")
        X = X.float().contiguous()
        out = t.empty(X.shape[0], X.shape[1] * (2 * self.n + 1), device=X.device, dtype=t.float32)
        return ops.pe_encode(X, self.n, out)


class SineLayer(nn.Module):
    """misc.py:148-194: sin(norm(omega_0 * linear(x))); BatchNorm1d(momentum=0.01) iff use_norm and not is_first."""

    def __init__(self, in_features, out_features, bias=True, is_first=False, omega_0=30, use_norm=False):
        super().__init__()
        self.omega_0 = omega_0
        self.is_first = is_first
        self.in_features = in_features
        self.linear = nn.Linear(in_features, out_features, bias=bias)
        self.init_weights()
        if is_first is False and use_norm:
            self.norm = nn.BatchNorm1d(out_features, momentum=0.01)
        else:
            self.norm = nn.Identity()

    def init_weights(self):
        with t.no_grad():
            if self.is_first:
                self.linear.weight.uniform_(-1 / self.in_features, 1 / self.in_features)
            else:
                b = sqrt(6 / self.in_features) / self.omega_0
                self.linear.weight.uniform_(-b, b)

    @property
    def has_bn(self):
        return isinstance(self.norm, nn.BatchNorm1d)

    def forward(self, input):
        net = _SingleLayerNet(self)
        return _run(net, "single", input, None, None, 1, getattr(self, "precision", "fp32"))[0]


class _SingleLayerNet:
    """adapter so that a lone SineLayer can run through the executor"""

    def __init__(self, layer):
        self.layer = layer
        self.training = layer.training

    def parameters(self):
        return self.layer.parameters()


class G_NeRF_Net_Classic(nn.Module):
    """G_NeRF.py:6-71 module tree (non-SIREN2 branch)."""

    def __init__(self, layer_width=512, expand_size_pose=10, expand_size_solar_angle=4, num_out_channels=3,
                 extended_encoding=False):
        super().__init__()
        if not extended_encoding or expand_size_pose == 0 or expand_size_solar_angle == 0:
            raise NotImplementedError("only the extended positional encoding used by T_NeRF is implemented")
        self.ignore_hue = False
        self.use_norm = True
        self.use_SIREN2 = False
        self._expand_size_pose = expand_size_pose
        self._expand_size_solar_angle = expand_size_solar_angle
        lw, lw2, lw4 = layer_width, max(layer_width // 2, 1), max(layer_width // 4, 1)
        input_size = 3 * (expand_size_pose * 2 + 1)
        input_size_solar = 3 * (expand_size_solar_angle * 2 + 1)
        self.PE_encoder = PE_Encode(expand_size_pose, True)
        self.PE_encoder_solar = PE_Encode(expand_size_solar_angle, True)
        self.fc1 = SineLayer(input_size, lw, is_first=True)
        self.fc2 = SineLayer(lw, lw, use_norm=True)
        self.fc3 = SineLayer(lw, lw, use_norm=True)
        self.fc4 = SineLayer(lw, lw, use_norm=True)
        self.fc5 = SineLayer(lw + input_size, lw, is_first=False, use_norm=True)
        self.fc6 = SineLayer(lw, lw, use_norm=True)
        self.fc7 = SineLayer(lw, lw, use_norm=True)
        self.fc8 = SineLayer(lw, lw, use_norm=True)
        self.fc9 = SineLayer(lw, lw2, use_norm=True)
        self.fc10Col = nn.Linear(lw2, num_out_channels)
        self.fc10Sigma = nn.Linear(lw2, 1)
        self._inv_delta = 1
        self.fc_solar_1 = SineLayer(input_size_solar + lw2, lw2, is_first=True)
        self.fc_solar_2 = SineLayer(lw2, lw2)
        self.fc_solar_3 = SineLayer(lw2, lw2)
        self.fc_solar_4 = nn.Linear(lw2, 1)
        self.fc_sky_color_1 = SineLayer(input_size_solar, lw4, is_first=True)
        self.fc_sky_color_2 = nn.Linear(lw4, 3)
        self.num_out_channels = num_out_channels
        self.sig = nn.Sigmoid()
        self.SoftPlus = nn.Softplus()
        self.precision = "bf16"

    # -- public methods of the reference that callers use directly (Eval_funcs.py:286-287,314) ---------------
    def forward_Sigma_Only(self, X):
        """G_NeRF.py:74-77."""
        return self.SoftPlus(_run(self._net(), "sigma", X, None, None, 1)[0]) * self._inv_delta

    def forward_color_only(self, X):
        """G_NeRF.py:154-157."""
        return self.sig(_run(self._net(), "color", X, None, None, 1)[0])

    def forward_Position(self, X):
        """G_NeRF.py:93-98 -> X_Encode, rho_raw, col_raw."""
        pos, xenc = _run(self._net(), "position", X, None, None, 1)
        return xenc, pos[:, 0:1], pos[:, 1:]

    def _net(self):
        return _GOnly(self)


class _GOnly:
    """adapter: the position-only schedules (sigma / color / position) need nothing outside the G-net"""

    def __init__(self, g):
        self.G_NeRF_net = g
        self.layer_width = g.fc1.linear.out_features
        self.training = g.training
        self.precision = g.precision

    def parameters(self):
        return self.G_NeRF_net.parameters()


class T_NeRF(nn.Module):
    """T_NeRF_net_v2.py:20-204."""

    def __init__(self, layer_width, n_classes=4, HM=np.array([[0], [0]]), precision="bf16"):
        super().__init__()
        self.allow_other_temporal_adjust = False
        self.hm = t.tensor(HM, requires_grad=False)
        self._hm_const = t.tensor(self.hm.shape).reshape([1, 2]) - 1
        self.G_NeRF_net = G_NeRF_Net_Classic(layer_width=layer_width, extended_encoding=True)
        self.Time_Enocder = PE_Encode(2, use_Extend_encoding=True)
        self.time_layer_1 = SineLayer(4 * 2 + 2, layer_width, is_first=True)
        self.time_layer_2 = SineLayer(layer_width, layer_width)
        self.get_class_layer = nn.Linear(layer_width, n_classes)
        self.adjust_layer_1 = SineLayer(layer_width // 2, layer_width)
        self.adjust_layer_2 = SineLayer(layer_width, layer_width)
        self.adjust_layer_3 = SineLayer(layer_width, layer_width)
        self.adjust_col = nn.Linear(layer_width, n_classes * 3)
        self.adjust_rho = nn.Linear(layer_width, n_classes)              # unused heads: present for load_state_dict
        self.adjust_solar_vis = nn.Linear(layer_width, n_classes)
        self.adjust_sky_col = nn.Linear(layer_width, n_classes * 3)
        self.n_classes = n_classes
        self.layer_width = layer_width
        self.SoftMax = nn.Softmax(1)
        self.Softplus = nn.Softplus()
        self.Sigmoid = nn.Sigmoid()
        self.batch_params_freeze = False
        self._ignore_solar = False
        self.precision = precision
        self._fused_cache = None

    @property
    def precision(self):
        return self.G_NeRF_net.precision

    @precision.setter
    def precision(self, p):
        if p not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.G_NeRF_net.precision = p

    def ignore_solar(self):
        self._ignore_solar = True

    def _process_time(self, Time):
        return Time[:, 0:2]

    # ---- per-ray interface used by the render / loss engine (time, sky and class branches once per ray) ----
    def forward_rays(self, pts, sun, time, S, mode="full"):
        """pts [N*S,3]; sun [N,3] (or [1,3]: one direction for all rays); time [N,>=2] (or [1,..]).
        Returns the RAW heads: full -> (pos [M,4], vis [M,1], adj [M,12], sky [N,3], class logits [N,C]);
        solar -> (rho [M,1], vis [M,1], sky [N,3]); sigma -> (rho [M,1],).
        Eval mode without autograd runs the fused tcgen05 kernel (bf16, layer_width 512); training / fp32
        validation run the layer-wise schedule."""
        from . import fused
        N = pts.shape[0] // S
        if fused.usable(self, pts) and mode in ("full", "solar", "sigma"):
            return fused.forward_rays(self, pts, sun, time, S, mode)
        if sun is not None and sun.shape[0] == 1 and N > 1:
            sun = sun.expand(N, sun.shape[1])
        if time is not None and time.shape[0] == 1 and N > 1:
            time = time.expand(N, time.shape[1])
        return _run(self, mode, pts, sun, time, S)

    def stage_weights(self, modes=("full", "solar")):
        """bf16 copies of every weight the given passes use, staged on the current stream (before passes fork onto
        parallel streams, so that all of them find the cache filled)."""
        for mode in modes:
            run = _Pass(self, mode, self.precision)
            for sp in run.specs:
                run._wc(sp)

    def _fused_ready(self):
        from . import fused
        return fused.usable(self, None)

    def _mix(self, pos, vis, adj, sky, cls_logits, S, mix, act_col):
        M = pos.shape[0]
        cls = self.SoftMax(cls_logits)
        clsp = cls if S == 1 else cls.repeat_interleave(S, 0)
        skyp = sky if S == 1 else sky.repeat_interleave(S, 0)
        Adj = adj.reshape(M, self.n_classes, -1)
        Rho = self.Softplus(pos[:, 0:1])
        Col = pos[:, 1:]
        if mix:
            Adjust_col = t.sum(Adj * clsp.unsqueeze(2), 1)
            Col = self.Sigmoid(Col + Adjust_col) if act_col else Col
        else:
            Adjust_col = Adj
        return Rho, Col, self.Sigmoid(vis), self.Sigmoid(skyp), clsp, Adjust_col

    # ---- per-point public API (reference signatures) -------------------------------------------------------
    def forward(self, X, Solar_Angle, Time):
        """T_NeRF_net_v2.py:75-105."""
        pos, vis, adj, sky, cl = _run(self, "full", X, Solar_Angle, Time, 1)
        return self._mix(pos, vis, adj, sky, cl, 1, True, True)

    def forward_seperate(self, X, Solar_Angle, Time):
        """T_NeRF_net_v2.py:131-151: raw col, unmixed Adj [M,C,3]."""
        pos, vis, adj, sky, cl = _run(self, "full", X, Solar_Angle, Time, 1)
        return self._mix(pos, vis, adj, sky, cl, 1, False, False)

    def forward_full_eval(self, X, Solar_Angle, Time):
        """T_NeRF_net_v2.py:184-204 (same tuple as forward_seperate)."""
        return self.forward_seperate(X, Solar_Angle, Time)

    def forward_Solar(self, X, Solar_Angle, Time):
        """T_NeRF_net_v2.py:154-157: trunk under no_grad, solar + sky heads with grad, RAW sky."""
        rho, vis, sky = _run(self, "solar", X, Solar_Angle, None, 1)
        return self.Softplus(rho), self.Sigmoid(vis), sky

    def get_class_only(self, Time):
        """T_NeRF_net_v2.py:160-163."""
        return self.SoftMax(_run(self, "class", None, None, Time, 1)[0])

    def forward_Classic_Sigma_Only(self, X):
        """T_NeRF_net_v2.py:169-170."""
        return self.G_NeRF_net.forward_Sigma_Only(X)

    def approx_Solar(self, X, X_solar, Time):
        """T_NeRF_net_v2.py:107-128."""
        n = X.shape[0]
        pos, xenc = _run(self, "position", t.cat([X, X_solar], 0), None, None, 1)
        cl = _run(self, "class", None, None, Time, 1)[0]
        adj = _run(self, "adjust", xenc[0:n], None, None, 1)[0]
        cls = self.SoftMax(cl)
        Adjust_col = t.sum(adj.reshape(n, self.n_classes, -1) * cls.unsqueeze(2), 1)
        Rho = self.Softplus(pos[:, 0:1])
        Col = self.Sigmoid(pos[0:n, 1:] + Adjust_col)
        return Rho[0:n], Rho[n::], Col, cls, Adjust_col

    def Supervised_Sample(self, world_pts, delta):
        """T_NeRF_net_v2.py:175-181 (the reference runs this on the CPU with a D2H + H2D round trip per step): one kernel on
        the device; the height map (a plain float64 CPU attribute, :28-29) is uploaded once per version."""
        dev = world_pts.device
        if dev.type != "cuda":
            raise ops._lib.SeasonNerfCudaError("season_nerf_b200.T_NeRF.Supervised_Sample runs on CUDA only (no CPU fallback)")
        cached = self.__dict__.get("_hm_dev")
        if cached is None or cached[0] != dev or cached[1] is not self.hm or cached[2] != self.hm._version:
            cached = self.__dict__["_hm_dev"] = (dev, self.hm, self.hm._version, self.hm.to(device=dev, dtype=t.float64).contiguous())
        return ops.supervised_sample(world_pts.float(), delta.float().reshape(-1), cached[3])


# =========================================================================================================
# executor
# =========================================================================================================
class _Spec:
    __slots__ = ("name", "kind", "inp", "in_col0", "kp", "out", "out_col0", "lin", "layer", "need_dx", "grad")

    def __init__(self, name, kind, inp, in_col0, kp, out, out_col0, lin, layer=None, need_dx=True, grad=True):
        self.name, self.kind, self.inp, self.in_col0, self.kp = name, kind, inp, in_col0, kp
        self.out, self.out_col0, self.lin, self.layer, self.need_dx, self.grad = out, out_col0, lin, layer, need_dx, grad

    @property
    def n_out(self):
        return sum(l.out_features for l in self.lin)


def _plan(net, mode):
    """Layer schedule.  Buffers (row-major activation matrices):
       cat5  [M, lw+64]  = [h4 | enc(63) | 0]   (fc1 reads the enc slice, fc5 the whole row)      G_NeRF.py:81-86
       cats1 [M, lw2+32] = [X_Encode | enc_sun(27) | 0]                                           G_NeRF.py:101-102
    """
    if mode == "single":
        L = net.layer
        return [_Spec("layer", "sine", "x", 0, _r8(L.in_features), "y", 0, [L.linear], L, need_dx=True)], None
    g = net.G_NeRF_net
    lw = net.layer_width
    S = lambda *a, **k: _Spec(*a, **k)
    trunk_grad = mode not in ("solar",)
    tg = dict(grad=trunk_grad)
    trunk = [
        S("fc1", "sine", "cat5", lw, 64, "h1", 0, [g.fc1.linear], g.fc1, need_dx=False, **tg),
        S("fc2", "sine", "h1", 0, lw, "h2", 0, [g.fc2.linear], g.fc2, **tg),
        S("fc3", "sine", "h2", 0, lw, "h3", 0, [g.fc3.linear], g.fc3, **tg),
        S("fc4", "sine", "h3", 0, lw, "cat5", 0, [g.fc4.linear], g.fc4, **tg),
        S("fc5", "sine", "cat5", 0, lw + 64, "h5", 0, [g.fc5.linear], g.fc5, **tg),
        S("fc6", "sine", "h5", 0, lw, "h6", 0, [g.fc6.linear], g.fc6, **tg),
        S("fc7", "sine", "h6", 0, lw, "h7", 0, [g.fc7.linear], g.fc7, **tg),
        S("fc8", "sine", "h7", 0, lw, "h8", 0, [g.fc8.linear], g.fc8, **tg),
        S("fc9", "sine", "h8", 0, lw, "cats1", 0, [g.fc9.linear], g.fc9, **tg),
    ]
    lw2 = g.fc9.linear.out_features
    pos = S("pos", "linear", "cats1", 0, lw2, "pos", 0, [g.fc10Sigma, g.fc10Col], **tg)
    sigma = S("sigma", "linear", "cats1", 0, lw2, "pos", 0, [g.fc10Sigma], **tg)
    color = S("color", "linear", "cats1", 0, lw2, "pos", 0, [g.fc10Col], **tg)
    solar = [
        S("fc_solar_1", "sine", "cats1", 0, lw2 + 32, "s1", 0, [g.fc_solar_1.linear], g.fc_solar_1, need_dx=(mode == "full")),
        S("fc_solar_2", "sine", "s1", 0, lw2, "s2", 0, [g.fc_solar_2.linear], g.fc_solar_2),
        S("fc_solar_3", "sine", "s2", 0, lw2, "s3", 0, [g.fc_solar_3.linear], g.fc_solar_3),
        S("fc_solar_4", "linear", "s3", 0, lw2, "vis", 0, [g.fc_solar_4]),
    ]
    sky = [
        S("fc_sky_color_1", "sine", "senc", 0, 32, "k1", 0, [g.fc_sky_color_1.linear], g.fc_sky_color_1, need_dx=False),
        S("fc_sky_color_2", "linear", "k1", 0, g.fc_sky_color_1.linear.out_features, "sky", 0, [g.fc_sky_color_2]),
    ]
    time, adjust = [], []
    if mode in ("full", "class", "adjust"):
        time = [
            S("time_layer_1", "sine", "tenc", 0, 16, "t1", 0, [net.time_layer_1.linear], net.time_layer_1, need_dx=False),
            S("time_layer_2", "sine", "t1", 0, lw, "t2", 0, [net.time_layer_2.linear], net.time_layer_2),
            S("get_class_layer", "linear", "t2", 0, lw, "cls", 0, [net.get_class_layer]),
        ]
        adjust = [
            S("adjust_layer_1", "sine", "cats1", 0, lw2, "a1", 0, [net.adjust_layer_1.linear], net.adjust_layer_1),
            S("adjust_layer_2", "sine", "a1", 0, lw, "a2", 0, [net.adjust_layer_2.linear], net.adjust_layer_2),
            S("adjust_layer_3", "sine", "a2", 0, lw, "a3", 0, [net.adjust_layer_3.linear], net.adjust_layer_3),
            S("adjust_col", "linear", "a3", 0, lw, "adj", 0, [net.adjust_col]),
        ]
    if mode == "full":
        # solar branch of the image pass is forward-only in effect (vis is detached by the caller's colour
        # formula when not classic) but autograd decides: gradients flow if vis_raw receives one.
        return trunk + [pos] + solar + sky + time + adjust, ("pos", "vis", "adj", "sky", "cls")
    if mode == "solar":
        return trunk + [sigma] + solar + sky, ("pos", "vis", "sky")
    if mode == "sigma":
        return trunk + [sigma], ("pos",)
    if mode == "color":
        return trunk + [color], ("pos",)
    if mode == "position":
        return trunk + [pos], ("pos", "cats1")
    if mode == "class":
        return time, ("cls",)
    if mode == "sky":
        return sky, ("sky",)
    if mode == "adjust":
        adjust[0] = S("adjust_layer_1", "sine", "x", 0, lw2, "a1", 0, [net.adjust_layer_1.linear], net.adjust_layer_1)
        return adjust, ("adj",)
    raise ValueError(mode)


_RAY_BUFS = ("senc", "k1", "sky", "tenc", "t1", "t2", "cls")


class _Pass:
    """One forward (and optionally backward) sweep over a layer schedule with explicit buffers."""

    def __init__(self, net, mode, precision):
        self.net, self.mode = net, mode
        self.dt = t.bfloat16 if precision == "bf16" else t.float32
        self.specs, self.outs = _plan(net, mode)
        self.saved = {}
        self.bufs = {}

    # -- weights in compute dtype, K padded to the buffer width (zero columns) --
    def _wc(self, spec):
        """(weight in the compute dtype, K zero-padded to the operand width; float32 bias).  Staged once per parameter
        version: the image pass and the solar pass of one training step share it (the cache lives on the first nn.Linear
        of the spec; a CUDA-graph replay bumps the version counters, see train.TrainStep)."""
        key = (tuple((l.weight._version, l.bias._version, l.weight.data_ptr()) for l in spec.lin), self.dt, spec.kp)
        hit = spec.lin[0].__dict__.get("_snb_wc")
        capturing = t.cuda.is_current_stream_capturing()
        # inside a CUDA-graph capture the staging MUST be recorded (a warm cache would freeze the weights of every replay
        # while the in-graph optimiser updates the fp32 parameters); the first pass of a capture restages, later passes of
        # the same capture (solar pass after the image pass) reuse what that capture staged
        if hit is not None and hit[0] == key and (not capturing or hit[3] == _capture_epoch[0]):
            return hit[1], hit[2]
        if BULK_STAGE and self.dt == t.bfloat16 and self.mode in ("full", "solar") and not self.__dict__.get("_bulk_done"):
            # first stale layer of a training pass: stage EVERY layer of the pass in one launch (29 per-layer converts, their
            # zero fills and concatenations were 0.15 ms of a step)
            self._bulk_done = True
            self._bulk_stage()
            hit = spec.lin[0].__dict__.get("_snb_wc")
            if hit is not None and hit[0] == key:
                return hit[1], hit[2]
        W = t.cat([l.weight for l in spec.lin], 0) if len(spec.lin) > 1 else spec.lin[0].weight
        if spec.kp == W.shape[1]:
            Wc = t.empty(W.shape[0], spec.kp, device=W.device, dtype=self.dt)
        else:
            Wc = t.zeros(W.shape[0], spec.kp, device=W.device, dtype=self.dt)
        ops.convert(W.detach(), Wc[:, :W.shape[1]])
        b = t.cat([l.bias for l in spec.lin], 0) if len(spec.lin) > 1 else spec.lin[0].bias
        b = b.detach().float().contiguous()
        spec.lin[0].__dict__["_snb_wc"] = (key, Wc, b, _capture_epoch[0] if capturing else None)
        return Wc, b

    def _bulk_stage(self):
        """bf16 copies of the weights of every stale layer of this pass: ONE zero-filled flat buffer (K padding), ONE kernel.
        Fresh storage per staging, like the per-layer path: a backward pass still holds the copies it saved."""
        capturing = t.cuda.is_current_stream_capturing()
        todo, total = [], 0
        for sp in self.specs:
            key = (tuple((l.weight._version, l.bias._version, l.weight.data_ptr()) for l in sp.lin), self.dt, sp.kp)
            hit = sp.lin[0].__dict__.get("_snb_wc")
            if hit is not None and hit[0] == key and (not capturing or hit[3] == _capture_epoch[0]):
                continue
            if any(l.weight.dtype != t.float32 or not l.weight.is_contiguous() for l in sp.lin):
                continue                                  # the per-layer path handles it
            todo.append((sp, key, total))
            total += sp.n_out * sp.kp
        if not todo:
            return
        flat = t.zeros(total, device=todo[0][0].lin[0].weight.device, dtype=self.dt)
        pairs = []
        for sp, key, off in todo:
            Wc = flat[off:off + sp.n_out * sp.kp].view(sp.n_out, sp.kp)
            r0 = 0
            for l in sp.lin:
                pairs.append((l.weight.detach(), Wc[r0:r0 + l.out_features, :l.in_features]))
                r0 += l.out_features
            b = t.cat([l.bias for l in sp.lin], 0) if len(sp.lin) > 1 else sp.lin[0].bias
            b = b.detach().float().contiguous()
            sp.lin[0].__dict__["_snb_wc"] = (key, Wc, b, _capture_epoch[0] if capturing else None)
        ops.stage_weights(pairs)

    def _buf(self, name, rows, width, dtype=None, zero=False):
        if name not in self.bufs:
            mk = t.zeros if zero else t.empty
            self.bufs[name] = mk(rows, width, device=self.dev, dtype=dtype or self.dt)
        return self.bufs[name]

    def _stats_slot(self, n):
        """[2, n] float32 zeros carved from one zero-filled pool per sweep (forward / backward): the fused GEMM epilogues
        accumulate their column sums into it"""
        pool = self.__dict__.get("_spool")
        if pool is None or self._spool_off + 2 * n > pool.numel():
            total = sum(2 * s_.n_out for s_ in self.specs if s_.kind == "sine")
            pool = self._spool = t.zeros(max(total, 2 * n), device=self.dev, dtype=t.float32)
            self._spool_off = 0
        out = pool[self._spool_off:self._spool_off + 2 * n].view(2, n)
        self._spool_off += 2 * n
        return out

    def _width(self, name):
        lw = getattr(self.net, "layer_width", None)
        if name == "cat5":
            return lw + 64
        if name == "cats1":
            return lw // 2 + 32
        raise KeyError(name)

    def forward(self, X, sun, time, S, keep):
        net = self.net
        self.keep = keep
        self.x_requires_grad = bool(X is not None and X.requires_grad and self.mode in ("single", "adjust"))
        dev = self.dev = (X if X is not None else (time if time is not None else sun)).device
        self.S = S
        M = X.shape[0] if X is not None else 0
        N = (M // S) if X is not None else (time.shape[0] if time is not None else sun.shape[0])
        self.M, self.N = M, N
        training = net.training
        names = {s.name for s in self.specs}
        if self.mode == "single":
            L = net.layer
            xin = self._buf("x", M, _r8(L.in_features), zero=True)
            ops.convert(X.float(), xin[:, :L.in_features])
        elif self.mode == "adjust":
            xin = self._buf("x", M, X.shape[1])
            ops.convert(X.float().contiguous(), xin)
        elif "fc1" in names:
            g = net.G_NeRF_net
            cat5 = self._buf("cat5", M, self._width("cat5"))
            ops.pe_encode(X.float().contiguous(), g._expand_size_pose, cat5, col0=net.layer_width, pad_to=64)
            self._buf("cats1", M, self._width("cats1"))
        if "fc_solar_1" in names:
            sun = sun.float().contiguous()
            lw2 = net.layer_width // 2
            if S == 1:
                ops.pe_encode(sun, net.G_NeRF_net._expand_size_solar_angle, self.bufs["cats1"], col0=lw2, pad_to=32)
            else:   # the direction is per ray: encode N rows, broadcast over the S samples of each ray
                senc_ray = t.empty(N, 32, device=dev, dtype=self.dt)
                ops.pe_encode(sun, net.G_NeRF_net._expand_size_solar_angle, senc_ray, pad_to=32)
                self.bufs["cats1"].view(N, S, -1)[:, :, lw2:lw2 + 32] = senc_ray.unsqueeze(1)
        if "fc_sky_color_1" in names:
            sun = sun.float().contiguous()
            if X is None:
                N = self.N = sun.shape[0]
            senc = self._buf("senc", N, 32)
            ops.pe_encode(sun, net.G_NeRF_net._expand_size_solar_angle, senc, pad_to=32)
        if "time_layer_1" in names:
            tenc = self._buf("tenc", N, 16)
            ops.pe_encode(time[:, 0:2].float().contiguous(), 2, tenc, pad_to=16)
        pending = {}        # buffer name -> (Z, a, c): the producer's activation is applied by its consumer (XFORM)
        for si, sp in enumerate(self.specs):
            rows = N if sp.out in _RAY_BUFS else M
            pend = pending.pop(sp.inp, None)
            Xv = None if pend is not None else self.bufs[sp.inp][:, sp.in_col0:sp.in_col0 + sp.kp]
            Wc, b = self._wc(sp)
            n_out = sp.n_out
            if sp.kind == "linear":
                out = self._buf(sp.out, rows, 16 if n_out <= 16 else _r8(n_out), dtype=t.float32)
                ops.gemm(Xv, Wc, out[:, :n_out], bias=b, alpha=1.0)
                if keep and sp.grad:
                    self.saved[sp.name] = (Wc,)
                continue
            Z = t.empty(rows, n_out, device=dev, dtype=self.dt)
            st = None
            if pend is not None:                      # A = sin(a_prev * Z_prev + c_prev), formed in shared memory
                st = ops.gemm_stats_xf(pend[0], pend[1], pend[2], Wc, Z, bias=b, alpha=OMEGA_0, stats=self._stats_slot(n_out),
                                       Y=pend[3])    # Y: the activated operand, written back only if a backward pass needs it
                if st is None:                        # shape not taken by the CTA-pair kernels: materialise the activation
                    Xv = self._buf(sp.inp, rows, pend[0].shape[1])
                    ops.sine_fwd(pend[0], pend[1], pend[2], Xv)
            if st is None and training and sp.layer.has_bn:          # batch statistics fused into the GEMM epilogue (CTA-pair kernel)
                st = ops.gemm_stats(Xv, Wc, Z, bias=b, alpha=OMEGA_0, stats=self._stats_slot(n_out))
            if self._defer_activation(si, sp, rows, training, keep, st):
                a, c, mean, invstd = self._affine(sp.layer, Z, rows, training, st)
                nx = [s_ for s_ in self.specs[si + 1:] if s_.inp == sp.out][0]
                Ybuf = self._buf(sp.out, rows, n_out) if (keep and nx.grad) else None     # operand of nx's weight gradient
                pending[sp.out] = (Z, a, c, Ybuf)
                if keep and sp.grad:
                    self.saved[sp.name] = (Wc, Z, a, c, mean, invstd)
                continue
            if sp.out in ("cat5", "cats1"):
                Y = self.bufs[sp.out][:, sp.out_col0:sp.out_col0 + n_out]
            else:
                Y = self._buf(sp.out, rows, n_out)
            if not sp.layer.has_bn and FUSE_EPILOGUES and ops.gemm_sine_fwd(Xv, Wc, Z, Y, bias=b, alpha=OMEGA_0):
                a, c, mean, invstd = self._affine(sp.layer, Z, rows, training, None)     # identity: sin already applied
            else:
                if st is None:
                    ops.gemm(Xv, Wc, Z, bias=b, alpha=OMEGA_0)
                a, c, mean, invstd = self._affine(sp.layer, Z, rows, training, st)
                ops.sine_fwd(Z, a, c, Y)
            if keep and sp.grad:
                self.saved[sp.name] = (Wc, Z, a, c, mean, invstd)
        res = []
        for o in self.outs:
            if o == "cats1":
                res.append(self.bufs["cats1"][:, :net.layer_width // 2].float())
            elif o in ("pos", "vis", "adj", "sky", "cls"):
                sp = [s for s in self.specs if s.out == o][0]
                res.append(self.bufs[o][:, :sp.n_out])
            else:
                res.append(self.bufs[o].float())
        if self.mode == "single":
            res = [self.bufs["y"].float()]
        if not keep:
            self.bufs = {k: v for k, v in self.bufs.items() if k in ("pos", "vis", "adj", "sky", "cls", "y")}
        return res

    def _defer_activation(self, si, sp, rows, training, keep, st):
        """True if layer `sp`'s sin(BatchNorm(.)) can be left to its consumer's GEMM (ops.gemm_stats_xf): train-mode BatchNorm
        with fused statistics, bf16, and exactly one consumer - a full-width train-mode BatchNorm layer reading the whole
        buffer (no concatenation, not a network output).  If a backward pass follows, the consumer's kernel writes the
        activated operand back (it is the B operand of the consumer's weight gradient): the resident-A kernel's shapes only."""
        if not (XFORM and training and st is not None and sp.layer.has_bn and self.dt == t.bfloat16 and rows >= 256):
            return False
        if sp.out in ("cat5", "cats1") or sp.out in (self.outs or ()) or _sync_world(self.net) > 1:
            return False
        users = [s_ for s_ in self.specs[si + 1:] if s_.inp == sp.out]
        if len(users) != 1:
            return False
        nx = users[0]
        if not (nx.kind == "sine" and nx.layer.has_bn and nx.in_col0 == 0 and nx.kp == sp.n_out and nx.n_out >= 128 and nx.kp <= 1024):
            return False
        if keep and nx.grad:       # Y must be written back: gemm_tc3.cu (N = 256 / 512, K <= 512 in steps of 64)
            return nx.n_out in (256, 512) and nx.kp <= 512 and nx.kp % 64 == 0 and XFORM_GRAD
        return True

    def _affine(self, layer, Z, rows, training, stats=None):
        """fold BatchNorm1d(momentum=.01, eps=1e-5) into y = a*z + c  (misc.py:169-170)."""
        dev = Z.device
        n = Z.shape[1]
        if not layer.has_bn:
            return const_vec(n, 1.0, dev), const_vec(n, 0.0, dev), None, None
        bn = layer.norm
        if training:
            s, ss = stats if stats is not None else ops.col_stats(Z)
            world = _sync_world(self.net)
            if world > 1:
                # SyncBN: the batch of a data-parallel step is the union of the ranks' rays (equal shards): sums of z and
                # z^2 over all ranks, so that every rank normalises with - and stores - the same statistics
                s, ss = _allreduce_pair(s, ss)
                rows = rows * world
            with t.no_grad():
                # two passes of one step may run on parallel streams: the running statistics of a module are updated in
                # program order (image pass, then solar pass - misc.py:169-170 is not commutative in the momentum update)
                order = getattr(self.net, "_bn_order", None)
                if order is not None:
                    cur = t.cuda.current_stream()
                    ev = order.get(id(bn))
                    if ev is not None:
                        cur.wait_event(ev)
                res = ops.bn_finalize(s, ss, rows, bn)       # one launch: batch statistics, running update, folded affine
                if order is not None:
                    ev = t.cuda.Event()
                    ev.record(cur)
                    order[id(bn)] = ev
                return res
        else:
            mean = bn.running_mean.detach().float()
            invstd = 1.0 / t.sqrt(bn.running_var.detach().float() + bn.eps)
        a = bn.weight.detach().float() * invstd
        c = bn.bias.detach().float() - mean * a
        return a.contiguous(), c.contiguous(), mean.contiguous(), invstd.contiguous()

    # ------------------------------------------------------------------------------------------------------
    def backward(self, gouts):
        """gouts: gradients of self.outs (None allowed).  Returns {parameter: grad}."""
        net, dt, dev = self.net, self.dt, self.dev
        training = net.training
        grads = {}
        gbuf = {}   # activation-gradient buffers keyed by buffer name

        def gb(name, rows, width):
            # every column that is ever read (the dY slice of the producing layer) is written by the first input-gradient
            # GEMM (accumulate=0) or by a convert below: no zero fill of these M x width matrices
            if name not in gbuf:
                gbuf[name] = t.empty(rows, width, device=dev, dtype=dt)
            return gbuf[name]

        gout = dict(zip(self.outs, gouts))
        if self.mode == "single" and gouts[0] is not None:
            ops.convert(gouts[0].float().contiguous(), gb("y", self.M, gouts[0].shape[1]))
        if gout.get("cats1") is not None:   # forward_Position's X_Encode output carries a gradient
            G = gb("cats1", self.M, self.bufs["cats1"].shape[1])
            ops.convert(gout["cats1"].float().contiguous(), G[:, :net.layer_width // 2])
        dw_total = sum(s_.n_out * s_.kp for s_ in self.specs if s_.grad and s_.name in self.saved)
        dw_flat = t.zeros(dw_total, device=dev, dtype=t.float32)      # ONE fill for all split-K weight-gradient targets
        dw_off = 0
        cur = t.cuda.current_stream()
        wstream = side_stream(("wgrad", cur.cuda_stream)) if (OVERLAP and self.M >= 4096) else None
        if wstream is not None:
            wstream.wait_stream(cur)
        self._spool = None      # fresh statistics pool for the backward sweep
        fused_bwd = {}      # sine layer name -> (sum g, sum g*xhat) left by the consumer's fused input-gradient GEMM
        producers = {(s_.out, s_.out_col0): s_ for s_ in self.specs if s_.kind == "sine"}
        n_readers = {}
        for s_ in self.specs:
            if s_.grad and s_.name in self.saved and ((s_.need_dx and s_.inp != "x") or (s_.inp == "x" and self.x_requires_grad)):
                n_readers[s_.inp] = n_readers.get(s_.inp, 0) + 1
        for sp in reversed(self.specs):
            if not sp.grad or sp.name not in self.saved:
                continue
            rows = self.N if sp.out in _RAY_BUFS else self.M
            n_out = sp.n_out
            fused_db = None
            Xv = self.bufs[sp.inp][:, sp.in_col0:sp.in_col0 + sp.kp]
            if sp.kind == "linear":
                g = gout.get(sp.out)
                if g is None:
                    continue
                (Wc,) = self.saved[sp.name]
                dZ = t.zeros(rows, 16 if n_out <= 16 else _r8(n_out), device=dev, dtype=dt)
                g = g.float().contiguous()
                ops.convert(g, dZ[:, :n_out])
                dZv = dZ[:, :n_out]
                alpha = 1.0
                bn_train = False
                fused_db = g.sum(0)                # bias gradient of a head: column sums of the incoming float32 gradient
            else:
                if sp.out not in gbuf:
                    continue
                Wc, Z, a, c, mean, invstd = self.saved[sp.name]
                dY = gbuf[sp.out][:, sp.out_col0:sp.out_col0 + n_out]
                bn_train = sp.layer.has_bn and training
                fused = fused_bwd.pop(sp.name, None)
                if fused is not None:
                    # the input-gradient GEMM that produced dY already multiplied by cos(.) and reduced the columns
                    sg, sgx = fused
                    dZ = dY
                    if sp.layer.has_bn:
                        self._acc(grads, sp.layer.norm.weight, sgx)      # local sums: the gradient all-reduce adds the ranks
                        self._acc(grads, sp.layer.norm.bias, sg)
                        world = _sync_world(net) if bn_train else 1
                        if world > 1:
                            sg, sgx = _allreduce_pair(sg, sgx)           # SyncBN backward: batch means over all ranks
                        ops.bn_bwd_apply(dY, Z, a, mean, invstd, sg, sgx, dZ, scale=(1.0 / (rows * world)) if bn_train else 0.0)
                    else:
                        fused_db = sg
                elif bn_train:
                    dZ = t.empty(rows, n_out, device=dev, dtype=dt)
                    sg, sgx = ops.sine_bwd_reduce(dY, Z, a, c, mean, invstd)
                    bn = sp.layer.norm
                    self._acc(grads, bn.weight, sgx.float())
                    self._acc(grads, bn.bias, sg.float())
                    world = _sync_world(net)
                    if world > 1:
                        sg, sgx = _allreduce_pair(sg, sgx)
                    ops.sine_bwd_apply(dY, Z, a, c, dZ, mean, invstd, (sg / (rows * world)).float().contiguous(),
                                       (sgx / (rows * world)).float().contiguous())
                else:
                    dZ = t.empty(rows, n_out, device=dev, dtype=dt)
                    if sp.layer.has_bn:  # eval-mode BN: plain affine, parameters still get gradients
                        sg, sgx = ops.sine_bwd_reduce(dY, Z, a, c, mean, invstd)
                        self._acc(grads, sp.layer.norm.weight, sgx.float())
                        self._acc(grads, sp.layer.norm.bias, sg.float())
                    ops.sine_bwd_apply(dY, Z, a, c, dZ)
                dZv = dZ
                alpha = OMEGA_0
            # bias gradient: alpha * column sums of dZ (analytically zero in front of a train-mode BatchNorm)
            if bn_train:
                db = t.zeros(n_out, device=dev)
            elif fused_db is not None:
                db = alpha * fused_db
            else:
                db = alpha * ops.col_stats(dZv)[0].float()
            # weight gradient dW[n_out, kp] = alpha * dZ^T . X   (split-K, fp32 atomics into zeros)
            dW = dw_flat[dw_off:dw_off + n_out * sp.kp].view(n_out, sp.kp)
            dw_off += n_out * sp.kp
            if wstream is not None:
                # the weight gradient only feeds the optimiser: it runs beside the input-gradient chain of the next layers
                ev = t.cuda.Event()
                ev.record(cur)
                wstream.wait_event(ev)
                with t.cuda.stream(wstream):
                    ops.gemm(dZv, Xv, dW, alpha=alpha, accumulate=2, a_t=True, b_t=True)
            else:
                ops.gemm(dZv, Xv, dW, alpha=alpha, accumulate=2, a_t=True, b_t=True)
            r0 = 0
            for lin in sp.lin:
                r1 = r0 + lin.out_features
                self._acc(grads, lin.weight, dW[r0:r1, :lin.in_features])
                self._acc(grads, lin.bias, db[r0:r1])
                r0 = r1
            # input gradient dX[rows, kin] (+)= alpha * dZ . W
            if (sp.need_dx and sp.inp != "x") or (sp.inp == "x" and self.x_requires_grad):
                kin = sp.kp if sp.inp not in ("cat5", "cats1") else (net.layer_width if sp.inp == "cat5" else net.layer_width // 2)
                first = sp.inp not in gbuf
                width = self.bufs[sp.inp].shape[1]
                G = gb(sp.inp, rows, width)
                pr = producers.get((sp.inp, sp.in_col0))
                st = None
                if (FUSE_EPILOGUES and first and pr is not None and pr.grad and pr.name in self.saved and pr.n_out == kin
                        and n_readers.get(sp.inp, 0) == 1 and dt == t.bfloat16):
                    # sole consumer of a sine layer's output: its cos / BatchNorm-backward column sums ride in this epilogue
                    _, Zp, ap, cp, meanp, invstdp = self.saved[pr.name]
                    if meanp is None:
                        meanp, invstdp = const_vec(ap.shape[0], 0.0, dev), const_vec(ap.shape[0], 1.0, dev)
                    st = ops.gemm_sine_bwd(dZv, Wc[:, :kin], G[:, sp.in_col0:sp.in_col0 + kin], Zp, ap, cp, meanp, invstdp, alpha=alpha,
                                           stats=self._stats_slot(kin))
                    if st is not None:
                        fused_bwd[pr.name] = st
                if st is None:
                    ops.gemm(dZv, Wc[:, :kin], G[:, sp.in_col0:sp.in_col0 + kin], alpha=alpha, accumulate=0 if first else 1, b_t=True)
        if wstream is not None:
            cur.wait_stream(wstream)
        self.gbuf = gbuf
        return grads

    @staticmethod
    def _acc(grads, p, g):
        if not p.requires_grad:
            return
        g = g.reshape(p.shape).to(p.dtype)
        grads[p] = g if p not in grads else grads[p] + g


class _NetFn(t.autograd.Function):
    @staticmethod
    def forward(ctx, run, X, sun, time, S, n_params, *params):
        outs = run.forward(X, sun, time, S, keep=True)
        # outputs that receive no gradient (e.g. the solar-visibility head of the image pass: vis is detached in the
        # colour formula, Eval_Tools_2.py:214) must arrive as None, not as materialised zeros - otherwise their whole
        # branch would be back-propagated with zero gradients
        ctx.set_materialize_grads(False)
        ctx.run = run
        ctx.params = params
        ctx.x_cols = X.shape[1] if run.x_requires_grad else 0
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        run = ctx.run
        grads = run.backward(gouts)
        gx = run.gbuf["x"][:, :ctx.x_cols].float() if (ctx.x_cols and "x" in run.gbuf) else None
        run.saved.clear()
        run.bufs.clear()
        return (None, gx, None, None, None, None) + tuple(grads.get(p) for p in ctx.params)


def _run(net, mode, X, sun, time, S, precision=None):
    """Dispatch one network pass.  Fails loudly for non-CUDA tensors: there is no CPU path."""
    probe = X if X is not None else (time if time is not None else sun)
    if not probe.is_cuda:
        raise ops._lib.SeasonNerfCudaError("season_nerf_b200.T_NeRF runs on CUDA only (got a %s tensor); "
                                           "move the module and its inputs to the GPU" % probe.device)
    precision = precision or getattr(net, "precision", "bf16")
    run = _Pass(net, mode, precision)
    params = [p for p in net.parameters()]
    need_grad = t.is_grad_enabled() and (any(p.requires_grad for p in params) or (X is not None and X.requires_grad))
    if not need_grad:
        with t.no_grad():
            return run.forward(X, sun, time, S, keep=False)
    return _NetFn.apply(run, X, sun, time, S, len(params), *params)
