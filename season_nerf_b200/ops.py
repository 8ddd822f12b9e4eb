"""Tensor-level wrappers over the C ABI (include/season_nerf_b200.h).  torch is used only for device memory,
streams and autograd bookkeeping; every arithmetic op on the hot path is a kernel of the sm_100a library.
All wrappers require CUDA tensors: there is no CPU fallback."""
import ctypes as C

import torch

from . import _lib
from ._lib import BF16, F32, F64, check

_DT = {torch.float32: F32, torch.bfloat16: BF16, torch.float64: F64}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _cuda(t, dtype=None, name="tensor"):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.SeasonNerfCudaError("season_nerf_b200: %s must be a CUDA tensor (no CPU fallback)" % name)
    if dtype is not None and t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t


def _mat(t, name="matrix"):
    """2-D, unit inner stride, arbitrary row pitch -> (tensor, ld)."""
    _cuda(t, name=name)
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise ValueError("%s must be 2-D with contiguous rows" % name)
    ld = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))
    return t, int(ld)


# ---- sampling ------------------------------------------------------------------------------------------
def sample_rays(top, bot, ts, zero_oob=False, want_pts=True):
    """misc.py:234-261.  top, bot [N,3] f32; ts [S] f32 (host-built like the reference) -> pts [N,S,3], deltas [N,S]."""
    top = _cuda(top, torch.float32, "top").contiguous()
    bot = _cuda(bot, torch.float32, "bot").contiguous()
    ts = _cuda(ts, torch.float32, "ts").contiguous()
    N, S = top.shape[0], ts.shape[0]
    pts = torch.empty(N, S, 3, device=top.device, dtype=torch.float32) if want_pts else None
    deltas = torch.empty(N, S, device=top.device, dtype=torch.float32)
    if N == 0:
        return pts, deltas
    check(_lib.load().snb_sample_rays(_ptr(top), _ptr(bot), _ptr(ts), N, S, int(zero_oob), _ptr(pts), _ptr(deltas), _stream()))
    return pts, deltas


def camera_rays(P, device, rows=None, cols=None, grid=None, z_top=1.0, z_bot=-1.0, bounds=None, want_xy64=False):
    """P_img_Pinhole.invert_P (pre_NeRF/P_Img.py:133-147) at z_top / z_bot for a pixel list (rows, cols: int arrays) or
    the raster `grid=(H, W, ds)`; -> tops, bots [n,3] f32, good [n] bool (None without bounds), xy64 [n,4] f64 or None."""
    import numpy as np
    Pm = np.ascontiguousarray(np.asarray(P, dtype=np.float64).reshape(12))
    Pp = (C.c_double * 12)(*Pm.tolist())
    if rows is not None:
        r = torch.as_tensor(np.asarray(rows), dtype=torch.int32).to(device).contiguous()
        c = torch.as_tensor(np.asarray(cols), dtype=torch.int32).to(device).contiguous()
        n, W, ds = r.shape[0], 0, 0
    else:
        H, W, ds = grid
        r = c = None
        n = H * W
    tops = torch.empty(n, 3, device=device, dtype=torch.float32)
    bots = torch.empty(n, 3, device=device, dtype=torch.float32)
    good = torch.empty(n, device=device, dtype=torch.uint8) if bounds is not None else None
    xy64 = torch.empty(n, 4, device=device, dtype=torch.float64) if want_xy64 else None
    bp = (C.c_double * 4)(*[float(v) for v in np.asarray(bounds, dtype=np.float64).reshape(4)]) if bounds is not None else None
    check(_lib.load().snb_camera_rays(Pp, _ptr(r), _ptr(c), n, int(W), int(ds), float(z_top), float(z_bot), bp, _ptr(tops),
                                      _ptr(bots), _ptr(xy64), _ptr(good), _stream()))
    return tops, bots, (good.bool() if good is not None else None), xy64


def solar_rays(world_center, W2L_H, az_el, u_xy, u_time=None, out=None):
    """create_solor_rays_uniform.__call__ (Eval_Tools_2.py:72-108) from drawn random numbers, on the device:
    az_el [n,2] f64 degrees, u_xy [n,2] f32 uniforms, u_time [n,2] f32 uniforms (optional)
    -> starts, ends, vec [n,3] f32 (, times [n,4] f32).  `out`: preallocated (starts, ends, vec, times) to write into."""
    import numpy as np
    az_el = _cuda(az_el, torch.float64, "az_el").contiguous()
    u_xy = _cuda(u_xy, torch.float32, "u_xy").contiguous()
    n, dev = az_el.shape[0], az_el.device
    if u_time is not None:
        u_time = _cuda(u_time, torch.float32, "u_time").contiguous()
    if out is None:
        mk = lambda w: torch.empty(n, w, device=dev, dtype=torch.float32)
        out = (mk(3), mk(3), mk(3), mk(4) if u_time is not None else None)
    starts, ends, vec, times = out[0], out[1], out[2], (out[3] if u_time is not None else None)
    for x in (starts, ends, vec, times):
        if x is not None and not (x.is_cuda and x.is_contiguous() and x.dtype == torch.float32 and x.shape[0] == n):
            raise ValueError("solar_rays: out tensors must be contiguous float32 CUDA tensors with n rows")
    if n == 0:                                   # empty tensors have no data pointer to hand over
        return (starts, ends, vec, times) if u_time is not None else (starts, ends, vec)
    wc = (C.c_double * 3)(*[float(v) for v in np.asarray(world_center, dtype=np.float64).reshape(3)])
    Hm = (C.c_double * 16)(*np.asarray(W2L_H, dtype=np.float64).reshape(16).tolist())
    check(_lib.load().snb_solar_rays(wc, Hm, _ptr(az_el), _ptr(u_xy), _ptr(u_time), n, _ptr(starts), _ptr(ends), _ptr(vec),
                                     _ptr(times), _stream()))
    return (starts, ends, vec, times) if u_time is not None else (starts, ends, vec)


_K_HIT = None


def supervised_sample(pts, delta, hm):
    """T_NeRF.Supervised_Sample (T_NeRF_net_v2.py:175-181): pts [M,3] f32, delta [M,1] f32, hm [H,W] f64 (device) -> [M,1] f32"""
    global _K_HIT
    if _K_HIT is None:
        _K_HIT = float(-torch.log(1 - torch.tensor(0.99, dtype=torch.float32)))      # the reference's CPU float32 log
    pts = _cuda(pts, torch.float32, "pts").contiguous()
    delta = _cuda(delta, torch.float32, "delta").contiguous()
    hm = _cuda(hm, torch.float64, "hm").contiguous()
    M = pts.shape[0]
    out = torch.empty(M, 1, device=pts.device, dtype=torch.float32)
    if M:
        check(_lib.load().snb_supervised_sample(_ptr(pts), _ptr(delta), _ptr(hm), int(hm.shape[0]), int(hm.shape[1]), _K_HIT, M,
                                                _ptr(out), _stream()))
    return out


def solar_tops(pts, sun_vec, f64=True):
    """mg_Img_Eval.py:57-60 (f64) / Eval_Tools_2.py:255-258 (f32).  pts [M,3] -> tops [M,3]."""
    pts = _cuda(pts, torch.float32, "pts").contiguous().reshape(-1, 3)
    sun = (C.c_double * 3)(*[float(v) for v in sun_vec])
    tops = torch.empty_like(pts)
    check(_lib.load().snb_solar_tops(_ptr(pts), pts.shape[0], sun, int(f64), _ptr(tops), _stream()))
    return tops


def march_transmittance(rho, deltas):
    rho = _cuda(rho, torch.float32).contiguous()
    deltas = _cuda(deltas, torch.float32).contiguous()
    S = rho.shape[-1]
    Mrows = rho.numel() // S
    out = torch.empty(Mrows, device=rho.device, dtype=torch.float32)
    if Mrows == 0:
        return out
    check(_lib.load().snb_march_transmittance(_ptr(rho), _ptr(deltas), Mrows, S, _ptr(out), _stream()))
    return out


# ---- compositing ---------------------------------------------------------------------------------------
def composite_fwd(rho, deltas, col, vis, sky, classic=False, want_pv=True):
    """rho, deltas, vis [N,S]; col [N,S,3]; sky [N,3] or [N,S,3] -> dict (Eval_Tools_2.py:187-215)."""
    N, S = rho.shape
    per_sample = sky.dim() == 3
    dev = rho.device
    mk = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
    PV, PE, PS = (mk(N, S), mk(N, S), mk(N, S)) if want_pv else (None, None, None)
    albedo, rendered, vsum = mk(N, 3), mk(N, 3), mk(N)
    if N == 0:                                  # empty tensors have no data pointer to hand over
        return PV, PE, PS, albedo, rendered, vsum
    check(_lib.load().snb_composite_fwd(_ptr(rho), _ptr(deltas), _ptr(col), _ptr(vis), _ptr(sky), int(per_sample), N, S,
                                        int(classic), _ptr(PV), _ptr(PE), _ptr(PS), _ptr(albedo), _ptr(rendered),
                                        _ptr(vsum), _stream()))
    return PV, PE, PS, albedo, rendered, vsum


class _Composite(torch.autograd.Function):
    """Alpha compositing with the transmittance scan; vis is detached unless classic (Eval_Tools_2.py:214)."""

    @staticmethod
    def forward(ctx, rho, deltas, col, vis, sky, classic):
        rho, deltas, col, vis, sky = [_cuda(x, torch.float32).contiguous() for x in (rho, deltas, col, vis, sky)]
        PV, PE, PS, albedo, rendered, vsum = composite_fwd(rho, deltas, col, vis, sky, classic)
        ctx.save_for_backward(rho, deltas, col, vis, sky)
        ctx.set_materialize_grads(False)       # unused outputs (PV / PE / PS when only the colour is in the loss) stay None
        ctx.classic = classic
        ctx.mark_non_differentiable(vsum)
        return PV, PE, PS, albedo, rendered, vsum

    @staticmethod
    def backward(ctx, dPV, dPE, dPS, d_albedo, d_rendered, _dvs):
        rho, deltas, col, vis, sky = ctx.saved_tensors
        N, S = rho.shape
        per_sample = sky.dim() == 3
        c = lambda g: None if g is None else g.contiguous()
        dPV, dPE, dPS, d_albedo, d_rendered = c(dPV), c(dPE), c(dPS), c(d_albedo), c(d_rendered)
        d_rho, d_col = torch.empty_like(rho), torch.empty_like(col)
        d_sky = torch.zeros_like(sky)
        d_vis = torch.zeros_like(vis) if ctx.classic else None
        check(_lib.load().snb_composite_bwd(_ptr(rho), _ptr(deltas), _ptr(col), _ptr(vis), _ptr(sky), int(per_sample), N, S,
                                            int(ctx.classic), _ptr(d_rendered), _ptr(d_albedo), _ptr(dPE), _ptr(dPV),
                                            _ptr(dPS), _ptr(d_rho), _ptr(d_col), _ptr(d_sky), _ptr(d_vis), _stream()))
        return d_rho, None, d_col, d_vis, d_sky, None


def composite(rho, deltas, col, vis, sky, classic=False):
    """-> PV, PE, PS [N,S], albedo [N,3], rendered [N,3], vis_sum [N]; differentiable in rho, col, sky (and vis if classic)."""
    return _Composite.apply(rho, deltas, col, vis, sky, bool(classic))


def _rows(x, cols, name):
    """2-D float32 CUDA matrix whose rows hold `cols` contiguous floats -> (tensor, row pitch in floats); a strided view
    (the layer-wise path keeps its narrow heads in 16-float-wide rows) is passed through without a copy"""
    _cuda(x, torch.float32, name)
    if x.dim() == 1:
        x = x.unsqueeze(1)
    if x.dim() != 2 or x.shape[1] != cols or (cols > 1 and x.stride(1) != 1) or (x.data_ptr() % 16 and cols >= 4):
        x = x.reshape(x.shape[0], -1).contiguous()
        if x.shape[1] != cols:
            raise ValueError("%s must have %d columns, got %s" % (name, cols, tuple(x.shape)))
    ld = x.stride(0) if x.shape[0] > 1 else max(cols, x.stride(0))
    if ld < cols or (cols >= 4 and ld % 4):
        x = x.contiguous()
        ld = cols
    return x, int(ld)


HEADS_MAX_S, HEADS_MAX_C = 128, 4


def heads_composite_usable(S, C):
    return S <= HEADS_MAX_S and 1 <= C <= HEADS_MAX_C


def heads_composite_fwd(pos4, vis_raw, adj, sky_raw, cls_logits, deltas, classic=False, want_pv=False):
    """raw heads -> (albedo [N,3], rendered [N,3], sky_act [N,3], vis_sum [N], PV, PE, PS [N,S] or None)
    (T_NeRF_net_v2.py:87-98 activations + Eval_Tools_2.py:187-215 compositing in one kernel)."""
    N, S = deltas.shape[0], deltas.shape[1]
    C = cls_logits.shape[1]
    pos4, ldp = _rows(pos4, 4, "pos4")
    vis_raw, ldv = _rows(vis_raw, 1, "vis_raw")
    adj, lda = _rows(adj, 3 * C, "adj")
    sky_raw = _cuda(sky_raw, torch.float32, "sky_raw").contiguous()
    cls_logits = _cuda(cls_logits, torch.float32, "cls_logits").contiguous()
    deltas = _cuda(deltas, torch.float32, "deltas").contiguous()
    dev = deltas.device
    mk = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
    albedo, rendered, sky_act, vsum = mk(N, 3), mk(N, 3), mk(N, 3), mk(N)
    PV, PE, PS = (mk(N, S), mk(N, S), mk(N, S)) if want_pv else (None, None, None)
    if N:
        check(_lib.load().snb_heads_composite_fwd(_ptr(pos4), ldp, _ptr(vis_raw), ldv, _ptr(adj), lda, _ptr(sky_raw),
                                                  _ptr(cls_logits), _ptr(deltas), N, S, C, int(classic), _ptr(albedo),
                                                  _ptr(rendered), _ptr(sky_act), _ptr(vsum), _ptr(PV), _ptr(PE), _ptr(PS), _stream()))
    return albedo, rendered, sky_act, vsum, PV, PE, PS


class _HeadsComposite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos4, vis_raw, adj, sky_raw, cls_logits, deltas, classic):
        albedo, rendered, sky_act, vsum, _, _, _ = heads_composite_fwd(pos4, vis_raw, adj, sky_raw, cls_logits, deltas, classic)
        ctx.save_for_backward(pos4, vis_raw, adj, sky_raw, cls_logits, deltas)
        ctx.set_materialize_grads(False)
        ctx.classic = classic
        ctx.mark_non_differentiable(vsum)
        return albedo, rendered, sky_act, vsum

    @staticmethod
    def backward(ctx, d_albedo, d_rendered, d_sky_act, _dvs):
        pos4, vis_raw, adj, sky_raw, cls_logits, deltas = ctx.saved_tensors
        N, S = deltas.shape[0], deltas.shape[1]
        C = cls_logits.shape[1]
        M = N * S
        p4, ldp = _rows(pos4, 4, "pos4")
        vr, ldv = _rows(vis_raw, 1, "vis_raw")
        ad, lda = _rows(adj, 3 * C, "adj")
        c = lambda g: None if g is None else g.float().contiguous()
        d_albedo, d_rendered, d_sky_act = c(d_albedo), c(d_rendered), c(d_sky_act)
        dev = deltas.device
        mk = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        d_pos, d_adj, d_sky, d_cls = mk(M, 4), mk(M, 3 * C), mk(N, 3), mk(N, C)
        d_vis = mk(M, 1) if ctx.classic else None
        # contiguous copies are bound to names: a temporary inside the argument list would be freed - and its block handed to
        # the next temporary - before the kernel runs
        sky_c, cls_c, dl_c = sky_raw.contiguous(), cls_logits.contiguous(), deltas.contiguous()
        check(_lib.load().snb_heads_composite_bwd(_ptr(p4), ldp, _ptr(vr), ldv, _ptr(ad), lda, _ptr(sky_c),
                                                  _ptr(cls_c), _ptr(dl_c), N, S, C, int(ctx.classic),
                                                  _ptr(d_albedo), _ptr(d_rendered), _ptr(d_sky_act), _ptr(d_pos), _ptr(d_vis),
                                                  _ptr(d_adj), _ptr(d_sky), _ptr(d_cls), _stream()))
        return (d_pos.view_as(pos4) if pos4.shape == d_pos.shape else d_pos.reshape(pos4.shape),
                None if d_vis is None else d_vis.reshape(vis_raw.shape), d_adj.reshape(adj.shape), d_sky, d_cls, None, None)


def heads_composite(pos4, vis_raw, adj, sky_raw, cls_logits, deltas, classic=False):
    """differentiable in pos4, adj, sky_raw, cls_logits (and vis_raw if classic) -> albedo, rendered, sky_act [N,3], vis_sum [N]"""
    return _HeadsComposite.apply(pos4, vis_raw, adj, sky_raw, cls_logits, deltas, bool(classic))


class _SolarLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rho_raw, vis_raw, deltas):
        N, S = deltas.shape[0], deltas.shape[1]
        rr, ldr = _rows(rho_raw, 1, "rho_raw")
        vr, ldv = _rows(vis_raw, 1, "vis_raw")
        deltas = _cuda(deltas, torch.float32, "deltas").contiguous()
        err = torch.empty(N, device=deltas.device, dtype=torch.float32)
        absorb = torch.empty(N, device=deltas.device, dtype=torch.float32)
        if N:
            check(_lib.load().snb_solar_loss_fwd(_ptr(rr), ldr, _ptr(vr), ldv, _ptr(deltas), N, S, _ptr(err), _ptr(absorb), _stream()))
        ctx.save_for_backward(rho_raw, vis_raw, deltas)
        ctx.set_materialize_grads(False)
        return err, absorb

    @staticmethod
    def backward(ctx, g_err, g_abs):
        rho_raw, vis_raw, deltas = ctx.saved_tensors
        N, S = deltas.shape[0], deltas.shape[1]
        rr, ldr = _rows(rho_raw, 1, "rho_raw")
        vr, ldv = _rows(vis_raw, 1, "vis_raw")
        c = lambda g: None if g is None else g.float().contiguous()
        g_err, g_abs = c(g_err), c(g_abs)
        d_vis = torch.empty(N * S, 1, device=deltas.device, dtype=torch.float32)
        check(_lib.load().snb_solar_loss_bwd(_ptr(rr), ldr, _ptr(vr), ldv, _ptr(deltas), N, S, _ptr(g_err), _ptr(g_abs), _ptr(d_vis), _stream()))
        return None, d_vis.reshape(vis_raw.shape), None


def solar_loss(rho_raw, vis_raw, deltas):
    """solar pass of get_loss from the raw heads -> (err [N], absorb [N]); the only gradient is w.r.t. vis_raw (PV and PE
    are detached in both terms, Eval_Tools_2.py:353-368)"""
    return _SolarLoss.apply(rho_raw, vis_raw, deltas)


class _LossTail(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rendered, gt, albedo, sky, err, absorb, alpha, scale, theta, qw, sc_lambda, type2):
        N = rendered.shape[0]
        f = lambda x: _cuda(x, torch.float32, "loss input").contiguous()
        rendered, gt, albedo, sky, err, absorb = f(rendered), f(gt), f(albedo), f(sky), f(err).reshape(-1), f(absorb).reshape(-1)
        a3, s3 = f(alpha).reshape(-1), f(scale).reshape(-1)
        if tuple(rendered.shape) != (N, 3) or tuple(gt.shape) != (N, 3) or tuple(albedo.shape) != (N, 3) or tuple(sky.shape) != (N, 3) \
                or err.shape[0] != N or absorb.shape[0] != N or a3.shape[0] != 3 or s3.shape[0] != 3:
            raise ValueError("loss_tail: rendered / gt / albedo / sky must be [N,3], err / absorb [N], alpha / scale 3 values")
        theta, qw = _cuda(theta, torch.float64, "theta").contiguous(), _cuda(qw, torch.float64, "qw").contiguous()
        if theta.shape[0] != 768 or qw.shape[0] != 768:
            raise ValueError("loss_tail: the quadrature rule has 768 nodes")
        buf = torch.empty(48, device=rendered.device, dtype=torch.float32)
        vals, aux = buf[:16], buf[16:]
        check(_lib.load().snb_loss_tail_fwd(_ptr(rendered), _ptr(gt), _ptr(albedo), _ptr(sky), _ptr(err), _ptr(absorb), _ptr(a3), _ptr(s3),
                                            _ptr(theta), _ptr(qw), N, float(sc_lambda), int(bool(type2)), _ptr(vals), _ptr(aux), _stream()))
        ctx.save_for_backward(rendered, gt, sky, a3, s3, buf)
        ctx.meta = (N, float(sc_lambda), int(bool(type2)), tuple(alpha.shape), tuple(scale.shape))
        ctx.set_materialize_grads(False)
        v = vals.unbind(0)
        # differentiable: Color_ada, Solar_Correction, Solar_Correction_2, Sky_Color_Var, Albedo_Color, total
        nd = (v[1], v[2], v[3], v[8], v[9])
        ctx.mark_non_differentiable(*nd)
        return (v[0], v[4], v[5], v[6], v[7], v[10]) + nd

    @staticmethod
    def backward(ctx, g_color, g_err, g_abs, g_sky, g_alb, g_total, *unused):
        rendered, gt, sky, a3, s3, buf = ctx.saved_tensors
        N, sc_lambda, type2, a_shape, s_shape = ctx.meta
        out = torch.empty(11 * N + 6, device=rendered.device, dtype=torch.float32)
        d_r, d_a, d_s = out[:3 * N].view(N, 3), out[3 * N:6 * N].view(N, 3), out[6 * N:9 * N].view(N, 3)
        d_e, d_b, d_al, d_sc = out[9 * N:10 * N], out[10 * N:11 * N], out[11 * N:11 * N + 3], out[11 * N + 3:]
        g = [None if x is None else x.float().contiguous() for x in (g_color, g_err, g_abs, g_sky, g_alb, g_total)]
        check(_lib.load().snb_loss_tail_bwd(_ptr(rendered), _ptr(gt), _ptr(sky), _ptr(a3), _ptr(s3), _ptr(buf[16:]), _ptr(buf[:16]),
                                            *[_ptr(x) for x in g], N, sc_lambda, type2, _ptr(d_r), _ptr(d_a), _ptr(d_s), _ptr(d_e),
                                            _ptr(d_b), _ptr(d_al), _ptr(d_sc), _stream()))
        return d_r, None, d_a, d_s, d_e, d_b, d_al.view(a_shape), d_sc.view(s_shape), None, None, None, None


def loss_tail(rendered, gt, albedo, sky, err, absorb, alpha, scale, theta, qw, sc_lambda, solar_type2=False):
    """The O(N) loss terms of a training step (Eval_Tools_2.py:353-443, Barron colour loss, no prior) in one kernel each way.
    -> dict of 0-dim tensors: Color_ada, Solar_Correction, Solar_Correction_2, Sky_Color_Var, Albedo_Color, total
    (differentiable w.r.t. rendered, albedo, sky, err, absorb, alpha, scale), Color_alpha, Color_width, Color (mse), scale_sq,
    solar_weight = sc_lambda / scale_sq (values only)."""
    o = _LossTail.apply(rendered, gt, albedo, sky, err, absorb, alpha, scale, theta, qw, float(sc_lambda), bool(solar_type2))
    keys = ("Color_ada", "Solar_Correction", "Solar_Correction_2", "Sky_Color_Var", "Albedo_Color", "total", "Color_alpha",
            "Color_width", "Color", "scale_sq", "solar_weight")
    return dict(zip(keys, o))


def cli_composite(rho, deltas, base, vis, adj, cls, exact_vis=None):
    """mg_Img_Eval.py:123-190 sums in float64.  Inputs f32 or f64 device tensors; cls [C] f64."""
    N, S = rho.shape[0], rho.shape[1]
    Cn = adj.shape[2]
    dt = _DT[rho.dtype]
    dev = rho.device
    mk = lambda *s: torch.empty(*s, device=dev, dtype=torch.float64)
    base_img, season, extreme, raw = mk(N, 3), mk(N, 3), mk(Cn, N, 3), mk(N)
    raw_e = mk(N) if exact_vis is not None else None
    ins = [x.contiguous() for x in (rho, deltas, base, vis, adj)]
    ev = None if exact_vis is None else exact_vis.contiguous()
    if any(x.dtype != rho.dtype for x in ins) or (ev is not None and ev.dtype != rho.dtype):
        raise TypeError("cli_composite: all components (exact_vis included) must share one dtype, got %s"
                        % [str(x.dtype) for x in ins + ([ev] if ev is not None else [])])
    cls = _cuda(cls, torch.float64).contiguous()
    if N == 0:
        return base_img, season, extreme, raw, raw_e
    check(_lib.load().snb_cli_composite(*[_ptr(x) for x in ins], _ptr(cls), _ptr(ev), dt, N, S, Cn, _ptr(base_img),
                                        _ptr(season), _ptr(extreme), _ptr(raw), _ptr(raw_e), _stream()))
    return base_img, season, extreme, raw, raw_e


def render_composite_raw(pos4, vis_raw, adj, deltas, cls, exact_vis=None):
    """CLI output-image sums from the raw heads of a block of rays (float64): -> season [N,3], raw_shadow [N], raw_shadow_exact
    [N] or None.  pos4 [N*S,4], vis_raw [N*S], adj [N*S,C*3] float32 contiguous; deltas [N,S]; cls [C] float64."""
    N, S = deltas.shape[0], deltas.shape[1]
    Cn = cls.shape[0]
    pos4 = _cuda(pos4, torch.float32, "pos4").contiguous()
    vis_raw = _cuda(vis_raw, torch.float32, "vis_raw").contiguous()
    adj = _cuda(adj, torch.float32, "adj").contiguous()
    deltas = _cuda(deltas, torch.float32, "deltas").contiguous()
    cls = _cuda(cls, torch.float64, "cls").contiguous()
    ev = None if exact_vis is None else _cuda(exact_vis, torch.float32, "exact_vis").contiguous()
    mk = lambda *s: torch.empty(*s, device=deltas.device, dtype=torch.float64)
    season, raw = mk(N, 3), mk(N)
    raw_e = mk(N) if ev is not None else None
    if N:
        check(_lib.load().snb_render_composite_raw(_ptr(pos4), _ptr(vis_raw), _ptr(adj), _ptr(deltas), _ptr(cls), _ptr(ev), N, S, Cn,
                                                   _ptr(season), _ptr(raw), _ptr(raw_e), _stream()))
    return season, raw, raw_e


def cli_classic_shadow(rho, deltas, base, vis, adj, sky, cls):
    """mg_Img_Eval.py:166-181: sum_s PS * sigmoid(base + cls . adj) * (vis + (1 - vis) * sky) -> [N,3] f64.
    rho, deltas, vis [N,S]; base, sky [N,S,3]; adj [N,S,C,3] (all f32 or all f64 device tensors); cls [C] f64."""
    N, S = rho.shape[0], rho.shape[1]
    Cn = adj.shape[2]
    out = torch.empty(N, 3, device=rho.device, dtype=torch.float64)
    ins = [x.contiguous() for x in (rho, deltas, base, vis, adj, sky)]
    if any(x.dtype != rho.dtype for x in ins):
        raise TypeError("cli_classic_shadow: all components must share one dtype")
    cls = _cuda(cls, torch.float64).contiguous()
    check(_lib.load().snb_cli_classic_shadow(*[_ptr(x) for x in ins], _ptr(cls), _DT[rho.dtype], N, S, Cn, _ptr(out), _stream()))
    return out


# limits of snb_year_sweep (csrc/composite.cu): checked here so that callers fail before any rendering work
YEAR_SWEEP_MAX_S, YEAR_SWEEP_MAX_C, YEAR_SWEEP_MAX_CLS_BYTES = 128, 4, 40 * 1024


def year_sweep(rho, deltas, base, adj, cls, shade=None, out=None, ps_weight=None):
    """mg_Img_Eval.py:192-228 recombination for T class vectors at once -> [T,N,3] f64 (times shade [N,3] f64 if given).
    ps_weight [N,S] (same dtype as rho) multiplies PS per sample (classic-shadow alignment, mg_Img_Eval.py:448-449)."""
    N, S = rho.shape[0], rho.shape[1]
    Cn, T = adj.shape[2], cls.shape[0]
    if out is None:
        out = torch.empty(T, N, 3, device=rho.device, dtype=torch.float64)
    if S > YEAR_SWEEP_MAX_S or Cn > YEAR_SWEEP_MAX_C or T * Cn * 8 > YEAR_SWEEP_MAX_CLS_BYTES:
        raise ValueError("year_sweep: the register-resident recombination kernel takes S <= %d samples per ray, C <= %d classes and "
                         "T*C*8 <= %d bytes of class vectors per launch (got S=%d, C=%d, T=%d); split T or render with fewer samples"
                         % (YEAR_SWEEP_MAX_S, YEAR_SWEEP_MAX_C, YEAR_SWEEP_MAX_CLS_BYTES, S, Cn, T))
    ins = [x.contiguous() for x in (rho, deltas, base, adj)]
    if any(x.dtype != rho.dtype for x in ins):
        raise TypeError("year_sweep: all components must share one dtype")
    cls = _cuda(cls, torch.float64).contiguous()
    if shade is not None:
        shade = _cuda(shade, torch.float64, "shade").contiguous()
    if ps_weight is not None:
        ps_weight = _cuda(ps_weight, rho.dtype, "ps_weight").contiguous()
    if N == 0 or T == 0:
        return out
    check(_lib.load().snb_year_sweep(*[_ptr(x) for x in ins], _ptr(cls), _ptr(shade), _ptr(ps_weight), _DT[rho.dtype], N, S, Cn, T,
                                     _ptr(out), _stream()))
    return out


def year_sweep_raw(pos4, deltas, adj, cls, shade=None, out=None):
    """year_sweep fed by the raw pos4 [N*S,4] (sigma, base colour logits) and adj [N*S,C*3] of the network -> [T,N,3] f64"""
    N, S = deltas.shape[0], deltas.shape[1]
    T, Cn = cls.shape[0], cls.shape[1]
    if S > YEAR_SWEEP_MAX_S or Cn > YEAR_SWEEP_MAX_C or T * Cn * 8 > YEAR_SWEEP_MAX_CLS_BYTES:
        raise ValueError("year_sweep_raw: S <= %d, C <= %d, T*C*8 <= %d bytes per launch (got S=%d, C=%d, T=%d)"
                         % (YEAR_SWEEP_MAX_S, YEAR_SWEEP_MAX_C, YEAR_SWEEP_MAX_CLS_BYTES, S, Cn, T))
    pos4 = _cuda(pos4, torch.float32, "pos4").contiguous()
    adj = _cuda(adj, torch.float32, "adj").contiguous()
    deltas = _cuda(deltas, torch.float32, "deltas").contiguous()
    cls = _cuda(cls, torch.float64, "cls").contiguous()
    if shade is not None:
        shade = _cuda(shade, torch.float64, "shade").contiguous()
    if out is None:
        out = torch.empty(T, N, 3, device=deltas.device, dtype=torch.float64)
    if N and T:
        check(_lib.load().snb_year_sweep_raw(_ptr(pos4), _ptr(deltas), _ptr(adj), _ptr(cls), _ptr(shade), N, S, Cn, T, _ptr(out), _stream()))
    return out


# ---- dense-layer building blocks -------------------------------------------------------------------------
def pe_encode(x, n_freq, out, col0=0, pad_to=None):
    """misc.py:105-139 into columns [col0, col0+D*(2n+1)) of `out` (zero padded to pad_to columns)."""
    x, ldx = _mat(_cuda(x, torch.float32, "x"))
    out, ldo = _mat(out, "out")
    D = x.shape[1]
    width = D * (2 * n_freq + 1)
    pad_to = width if pad_to is None else pad_to
    check(_lib.load().snb_pe_encode(_ptr(x), ldx, x.shape[0], D, n_freq, _ptr(out), _DT[out.dtype], ldo, col0, pad_to, _stream()))
    return out


def gemm(A, B, out, bias=None, alpha=1.0, accumulate=0, a_t=False, b_t=False):
    """out[M,N] = alpha*(A.B^T + bias) (+out).  a_t: A stored [K,M]; b_t: B stored [K,N].
    accumulate: 0 store, 1 add, 2 split-K atomic add into a pre-zeroed f32 `out`."""
    A, lda = _mat(A, "A")
    B, ldb = _mat(B, "B")
    out, ldc = _mat(out, "out")
    if A.dtype != B.dtype:
        raise TypeError("gemm operands must share a dtype")
    M, N = out.shape
    K = A.shape[0] if a_t else A.shape[1]
    Kb = B.shape[0] if b_t else B.shape[1]
    if K != Kb or (A.shape[1] if a_t else A.shape[0]) != M or (B.shape[1] if b_t else B.shape[0]) != N:
        raise ValueError("gemm shape mismatch: A%s B%s out%s a_t=%s b_t=%s" % (tuple(A.shape), tuple(B.shape), tuple(out.shape), a_t, b_t))
    if bias is not None:
        _cuda(bias, torch.float32, "bias")
    check(_lib.load().snb_gemm(_ptr(A), lda, int(a_t), _ptr(B), ldb, int(b_t), _ptr(out), ldc, _ptr(bias), float(alpha),
                               int(accumulate), M, N, K, _DT[A.dtype], _DT[out.dtype], _stream()))
    return out


def gemm_stats(A, B, out, bias=None, alpha=1.0, stats=None):
    """out[M,N] (bf16) = alpha*(A.B^T + bias) with the column sum / sum of squares of the stored values fused into the
    GEMM epilogue.  -> (sum[N], sumsq[N]) float32, or None when the CTA-pair kernel does not take the shape (the
    caller then runs gemm + col_stats)."""
    A, lda = _mat(A, "A")
    B, ldb = _mat(B, "B")
    out, ldc = _mat(out, "out")
    M, N = out.shape
    K = A.shape[1]
    if A.dtype != torch.bfloat16 or B.dtype != torch.bfloat16 or out.dtype != torch.bfloat16:
        return None
    if B.shape[1] != K or A.shape[0] != M or B.shape[0] != N:
        raise ValueError("gemm_stats shape mismatch: A%s B%s out%s" % (tuple(A.shape), tuple(B.shape), tuple(out.shape)))
    if bias is not None:
        _cuda(bias, torch.float32, "bias")
    s = stats if stats is not None else torch.zeros(2, N, device=out.device, dtype=torch.float32)
    rc = _lib.load().snb_gemm_stats(_ptr(A), lda, _ptr(B), ldb, _ptr(out), ldc, _ptr(bias), float(alpha), M, N, K, _ptr(s), _stream())
    if rc == -2:
        return None
    check(rc)
    return s[0], s[1]


def gemm_stats_xf(Zprev, xa, xc, B, out, bias=None, alpha=1.0, stats=None, Y=None):
    """gemm_stats with A = sin(xa * Zprev + xc) formed in shared memory (consumer-side activation of the previous SIREN
    layer: no activation pass, the activated matrix is not read from HBM); Y [M,K] bf16 (optional) receives the activated
    operand for a later weight gradient.  -> (sum[N], sumsq[N]) or None when the shape is not taken."""
    Zprev, lda = _mat(Zprev, "Zprev")
    B, ldb = _mat(B, "B")
    out, ldc = _mat(out, "out")
    M, N = out.shape
    K = Zprev.shape[1]
    if Zprev.dtype != torch.bfloat16 or B.dtype != torch.bfloat16 or out.dtype != torch.bfloat16:
        return None
    if B.shape[1] != K or Zprev.shape[0] != M or B.shape[0] != N or xa.shape[0] != K or xc.shape[0] != K:
        raise ValueError("gemm_stats_xf shape mismatch: Zprev%s B%s out%s xa%s" % (tuple(Zprev.shape), tuple(B.shape), tuple(out.shape), tuple(xa.shape)))
    xa = _cuda(xa, torch.float32, "xa").contiguous()
    xc = _cuda(xc, torch.float32, "xc").contiguous()
    if bias is not None:
        _cuda(bias, torch.float32, "bias")
    s = stats if stats is not None else torch.zeros(2, N, device=out.device, dtype=torch.float32)
    ldy = 0
    if Y is not None:
        Y, ldy = _mat(Y, "Y")
        if Y.dtype != torch.bfloat16 or tuple(Y.shape) != (M, K):
            raise ValueError("gemm_stats_xf: Y must be bf16 [M,K]")
    rc = _lib.load().snb_gemm_stats_xf(_ptr(Zprev), lda, _ptr(xa), _ptr(xc), _ptr(B), ldb, _ptr(out), ldc, _ptr(bias), float(alpha),
                                       M, N, K, _ptr(s), _ptr(Y), ldy, _stream())
    if rc == -2:
        return None
    check(rc)
    return s[0], s[1]


def gemm_sine_fwd(A, B, Z, Y, bias=None, alpha=1.0):
    """Z = alpha*(A.B^T + bias), Y = sin(Z) in ONE kernel (SIREN layer without BatchNorm, bf16).  -> False when the
    CTA-pair kernel does not take the shape (caller: gemm + sine_fwd)."""
    A, lda = _mat(A, "A")
    B, ldb = _mat(B, "B")
    Z, ldz = _mat(Z, "Z")
    Y, ldy = _mat(Y, "Y")
    if not (A.dtype == B.dtype == Z.dtype == Y.dtype == torch.bfloat16):
        return False
    M, N = Z.shape
    K = A.shape[1]
    if B.shape[1] != K or A.shape[0] != M or B.shape[0] != N or tuple(Y.shape) != (M, N):
        raise ValueError("gemm_sine_fwd shape mismatch: A%s B%s Z%s Y%s" % (tuple(A.shape), tuple(B.shape), tuple(Z.shape), tuple(Y.shape)))
    if bias is not None:
        _cuda(bias, torch.float32, "bias")
    rc = _lib.load().snb_gemm_sine_fwd(_ptr(A), lda, _ptr(B), ldb, _ptr(Z), ldz, _ptr(Y), ldy, _ptr(bias), float(alpha), M, N, K,
                                       _stream())
    if rc == -2:
        return False
    check(rc)
    return True


def gemm_sine_bwd(dZn, W, G, Z, a, c, mean, invstd, alpha=1.0, stats=None):
    """G = alpha*(dZn . W) * cos(a*Z + c) with the column sums (sum G, sum G*xhat) fused into the epilogue of the
    input-gradient GEMM.  dZn [M,K]; W [K,N] (the next layer's weight, rows = its outputs); G, Z [M,N] bf16.
    -> (sum_g [N], sum_g_xhat [N]) float32, or None when the CTA-pair kernel does not take the shape."""
    dZn, lda = _mat(dZn, "dZn")
    W, ldw = _mat(W, "W")
    G, ldg = _mat(G, "G")
    Z, ldz = _mat(Z, "Z")
    if not (dZn.dtype == W.dtype == G.dtype == Z.dtype == torch.bfloat16):
        return None
    M, N = G.shape
    K = dZn.shape[1]
    if W.shape[0] != K or W.shape[1] != N or dZn.shape[0] != M or tuple(Z.shape) != (M, N):
        raise ValueError("gemm_sine_bwd shape mismatch: dZn%s W%s G%s Z%s" % (tuple(dZn.shape), tuple(W.shape), tuple(G.shape), tuple(Z.shape)))
    for v in (a, c, mean, invstd):
        _cuda(v, torch.float32, "column vector")
    s = stats if stats is not None else torch.zeros(2, N, device=G.device, dtype=torch.float32)
    rc = _lib.load().snb_gemm_sine_bwd(_ptr(dZn), lda, _ptr(W), ldw, _ptr(G), ldg, _ptr(Z), ldz, _ptr(a), _ptr(c), _ptr(mean),
                                       _ptr(invstd), float(alpha), M, N, K, _ptr(s), _stream())
    if rc == -2:
        return None
    check(rc)
    return s[0], s[1]


def bn_bwd_apply(G, Z, a, mean, invstd, k1, k2, dZ, scale=1.0):
    """dZ = a*(G - scale*k1 - xhat*scale*k2) (may alias G); scale = 1/rows when k1, k2 are the raw column sums."""
    G, ldg = _mat(G, "G")
    Z, ldz = _mat(Z, "Z")
    dZ, ldo = _mat(dZ, "dZ")
    check(_lib.load().snb_bn_bwd_apply(_ptr(G), ldg, _ptr(Z), ldz, _ptr(a), _ptr(mean), _ptr(invstd), _ptr(k1), _ptr(k2),
                                       float(scale), _ptr(dZ), ldo, Z.shape[0], Z.shape[1], _DT[Z.dtype], _stream()))
    return dZ


def bn_finalize(s, ss, rows, bn):
    """train-mode BatchNorm1d bookkeeping in one launch -> (a, c, mean, invstd) float32 [N]; updates the module's running
    statistics and batch counter in place (misc.py:169-170)."""
    N = s.shape[0]
    out = torch.empty(4, N, device=s.device, dtype=torch.float32)
    nb = bn.num_batches_tracked
    check(_lib.load().snb_bn_finalize(_ptr(s), _ptr(ss), _DT[s.dtype], int(rows), N, _ptr(bn.weight.detach()), _ptr(bn.bias.detach()),
                                      _ptr(bn.running_mean), _ptr(bn.running_var), _ptr(nb) if nb is not None and nb.is_cuda else None,
                                      float(bn.momentum), float(bn.eps), _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(out[3]),
                                      _stream()))
    return out[0], out[1], out[2], out[3]


def col_stats(Z):
    """-> (sum[N], sumsq[N]) float64 over the rows of Z."""
    Z, ldz = _mat(Z, "Z")
    N = Z.shape[1]
    s = torch.empty(2, N, device=Z.device, dtype=torch.float64)
    check(_lib.load().snb_col_stats(_ptr(Z), _DT[Z.dtype], ldz, Z.shape[0], N, _ptr(s[0]), _ptr(s[1]), _stream()))
    return s[0], s[1]


def sine_fwd(Z, a, c, Y):
    Z, ldz = _mat(Z, "Z")
    Y, ldy = _mat(Y, "Y")
    check(_lib.load().snb_sine_fwd(_ptr(Z), ldz, _ptr(a), _ptr(c), _ptr(Y), ldy, Z.shape[0], Z.shape[1], _DT[Z.dtype], _stream()))
    return Y


def sine_bwd_reduce(dY, Z, a, c, mean, invstd):
    dY, ldd = _mat(dY, "dY")
    Z, ldz = _mat(Z, "Z")
    N = Z.shape[1]
    s = torch.empty(2, N, device=Z.device, dtype=torch.float64)
    check(_lib.load().snb_sine_bwd_reduce(_ptr(dY), ldd, _ptr(Z), ldz, _ptr(a), _ptr(c), _ptr(mean), _ptr(invstd), Z.shape[0],
                                          N, _DT[Z.dtype], _ptr(s[0]), _ptr(s[1]), _stream()))
    return s[0], s[1]


def sine_bwd_apply(dY, Z, a, c, dZ, mean=None, invstd=None, k1=None, k2=None):
    dY, ldd = _mat(dY, "dY")
    Z, ldz = _mat(Z, "Z")
    dZ, ldo = _mat(dZ, "dZ")
    check(_lib.load().snb_sine_bwd_apply(_ptr(dY), ldd, _ptr(Z), ldz, _ptr(a), _ptr(c), _ptr(mean), _ptr(invstd), _ptr(k1),
                                         _ptr(k2), _ptr(dZ), ldo, Z.shape[0], Z.shape[1], _DT[Z.dtype], _stream()))
    return dZ


def stage_weights(pairs):
    """[(src fp32 [r, c] (row pitch = stride(0)), dst bf16 [r, c] view)] -> all copies in ONE launch (csrc/elementwise.cu)"""
    n = len(pairs)
    if n == 0:
        return
    if n > 48:
        stage_weights(pairs[:48])
        return stage_weights(pairs[48:])
    for s_, d_ in pairs:
        if not (s_.is_cuda and d_.is_cuda and s_.dtype == torch.float32 and d_.dtype == torch.bfloat16 and s_.dim() == 2
                and tuple(s_.shape) == tuple(d_.shape) and s_.stride(1) == 1 and d_.stride(1) == 1):
            raise ValueError("stage_weights: pairs of fp32 source / bf16 destination matrices of equal shape, unit column stride")
    P, I = C.c_void_p * n, C.c_int * n
    check(_lib.load().snb_stage_weights(P(*[s_.data_ptr() for s_, _ in pairs]), P(*[d_.data_ptr() for _, d_ in pairs]),
                                        I(*[s_.shape[0] for s_, _ in pairs]), I(*[s_.shape[1] for s_, _ in pairs]),
                                        I(*[s_.stride(0) for s_, _ in pairs]), I(*[d_.stride(0) for _, d_ in pairs]), n, _stream()))


def convert(src, dst):
    src, lds = _mat(src, "src")
    dst, ldd = _mat(dst, "dst")
    check(_lib.load().snb_convert(_ptr(src), _DT[src.dtype], lds, _ptr(dst), _DT[dst.dtype], ldd, src.shape[0], src.shape[1], _stream()))
    return dst
