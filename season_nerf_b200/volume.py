"""Dense sigma / colour volume consumers ("next" row 4 of the scope table): drop-in for `gen_results` and the expected
height map at the head of `eval_HM` (T_NeRF_Eval_Utils/Eval_funcs.py:268-313).  The H x W x S grid of sample points is
built on the device, evaluated by the fused tcgen05 network (density and raw colour come out of the same pass) and reduced
along the vertical axis in float64 like the reference's numpy - the arrays leave the device once."""
import numpy as np
import torch as t

from . import ops


def _axis(n, flip=False):
    """XYZ * s * 2 - 1 of Eval_funcs.py:272-285 for one axis, float64, same evaluation order; z is negated afterwards."""
    a = np.arange(n) * (1 / n) * 2 - 1
    return -a if flip else a


def _grid_points(H, W, S, device):
    ax = [t.tensor(_axis(H), dtype=t.float64, device=device), t.tensor(_axis(W), dtype=t.float64, device=device),
          t.tensor(_axis(S, flip=True), dtype=t.float64, device=device)]
    g = t.stack(t.meshgrid(*ax, indexing="ij"), -1)          # [H,W,S,3] float64
    return g.reshape(-1, 3).float()                           # the reference's t.tensor(xyz).float()


def _volume(network, H, W, S, device, want_col):
    device = t.device(device)
    if device.type != "cuda":
        raise ops._lib.SeasonNerfCudaError("season_nerf_b200 evaluates volumes on CUDA only (no CPU fallback)")
    was_training = network.training
    network.eval()                                            # Eval_funcs.py:281
    pts = _grid_points(H, W, S, device)
    M = pts.shape[0]
    rho = t.empty(M, device=device, dtype=t.float32)
    col = t.empty(M, 3, device=device, dtype=t.float32) if want_col else None
    sun = t.tensor([[0.0, 0.0, 1.0]], device=device)
    tim = t.tensor([[1.0, 0.0, 1.0, 0.0]], device=device)
    step = (1 << 22) // S * S
    with t.no_grad():
        for i in range(0, M, step):
            e = min(i + step, M)
            if want_col:       # one pass gives sigma and the raw colour logits (forward_Classic_Sigma_Only + forward_color_only)
                pos = network.forward_rays(pts[i:e], sun, tim, S, mode="full")[0]
                rho[i:e] = network.Softplus(pos[:, 0])
                col[i:e] = network.Sigmoid(pos[:, 1:4])
            else:
                rho[i:e] = network.Softplus(network.forward_rays(pts[i:e], None, None, S, mode="sigma")[0][:, 0])
        y = rho.double().reshape(H, W, S) * (2 / S)           # all_Rhos * delta, float64 like the reference's numpy
        P_E = 1 - t.exp(-y)
        P_Vis = t.exp(-(t.cumsum(y, 2) - y))                  # exclusive cumulative sum (concat zero, drop last)
        P_Surf = P_E * P_Vis
    network.train(was_training)
    return rho.reshape(H, W, S, 1), P_E, P_Vis, P_Surf, (col.reshape(H, W, S, 3) if want_col else None)


def gen_results(network, img_shape, n_samples, device, max_batch_size=None):
    """Eval_funcs.py:268-296 -> all_Rhos [H,W,S,1], P_E, P_Vis, P_Surf [H,W,S], all_Cols [H,W,S,3] (float64 numpy).
    `max_batch_size` is accepted for compatibility; chunking is by device memory."""
    rho, pe, pv, ps, col = _volume(network, img_shape[0], img_shape[1], n_samples, device, True)
    f = lambda x: x.double().cpu().numpy()
    return f(rho), f(pe), f(pv), f(ps), f(col)


def height_map(network, shape, n_samples, device, h_range=None):
    """Expected surface height of eval_HM (Eval_funcs.py:298-313): sum(P_Surf * linspace(1,-1,S)) / sum(P_Surf), in the
    normalised cube, or in metres when h_range = (h0, h1) is given (:341-342)."""
    _, _, _, ps, _ = _volume(network, shape[0], shape[1], n_samples, device, False)
    z = t.tensor(np.linspace(1, -1, n_samples), dtype=t.float64, device=ps.device).reshape(1, 1, -1)
    hm = (t.sum(ps * z, 2) / t.sum(ps, 2)).cpu().numpy()
    if h_range is not None:
        hm = (hm + 1) / 2 * (h_range[1] - h_range[0]) + h_range[0]
    return hm


def confidence_range(P_Surf, h_range):
    """The 67 % range loop of eval_HM (Eval_funcs.py:315-331), for all columns at once on the device: starting at the mode of
    the surface PDF the window grows by one sample on each side (clamped) until it holds >= 0.67 of the mass or covers the
    column.  -> conf_range [H,W,3] float64 = (z0, z1, (z1 - z0) / S * (h1 - h0)).  The window mass of step k is the difference of
    two float64 prefix sums; the reference sums the slice - identical up to the last bit of a comparison against 0.67."""
    pdf = P_Surf / t.sum(P_Surf, 2, keepdim=True)
    H, W, S = pdf.shape
    start = t.argmax(pdf, 2)                                            # first maximum, like np.argmax
    csum = t.cat([t.zeros(H, W, 1, dtype=pdf.dtype, device=pdf.device), t.cumsum(pdf, 2)], 2)       # [H,W,S+1]
    k = t.arange(S + 1, device=pdf.device).reshape(1, 1, -1)
    z0 = (start.unsqueeze(-1) - k).clamp_min(0)                         # window of growth step k
    z1 = (start.unsqueeze(-1) + 1 + k).clamp_max(S)
    value = t.gather(csum, 2, z1) - t.gather(csum, 2, z0)
    # step 0 is the mode alone; the loop stops at the first step whose window reaches 0.67 or spans the column.  A NaN PDF
    # (all-zero column) never satisfies `value < .67`: the reference leaves the window at the mode
    done = ~(value < .67) | ((z0 == 0) & (z1 == S))
    first = t.argmax(done.to(t.int8), 2, keepdim=True)
    a, b = t.gather(z0, 2, first).squeeze(-1).double(), t.gather(z1, 2, first).squeeze(-1).double()
    return t.stack([a, b, (b - a) / S * (h_range[1] - h_range[0])], -1)


def eval_HM(network, GT, h_range, n_samples, device, max_batch_size=None):
    """Eval_funcs.py:298-395: expected height map of the density volume, its 67 % confidence range, the mean-shifted height
    map in metres and the error scores against the ground-truth DSM *before alignment*.
    -> (Imgs {"GT", "Est_HM_no_Shift", "Conf_Range"}, scores_before {"MAE", "RMSE", "Acc_1_m", "Median"}, conf_stats
    (nanmean, nanmedian of the range in metres - the two numbers the reference prints)).  The shift / rotation alignment
    search that follows in the reference (:397 ff., scipy image warps on the host) is evaluation tooling and is not part of
    this package; the first two return values are the reference's first two."""
    GT = np.asarray(GT, dtype=np.float64)
    _, _, _, ps, _ = _volume(network, GT.shape[0], GT.shape[1], n_samples, device, False)
    z = t.tensor(np.linspace(1, -1, n_samples), dtype=t.float64, device=ps.device).reshape(1, 1, -1)
    est = (t.sum(ps * z, 2) / t.sum(ps, 2)).cpu().numpy()
    conf = confidence_range(ps, h_range).cpu().numpy()
    h0, h1 = h_range[0], h_range[1]
    est = (est + 1) / 2 * (h1 - h0) + h0
    GTm = (GT + 1) / 2 * (h1 - h0) + h0
    est = est + np.nanmean((GTm - est).ravel())
    diff = est - GTm
    diff = np.ravel(diff[diff == diff])
    scores = {"MAE": np.mean(np.abs(diff)), "RMSE": np.sqrt(np.mean(diff ** 2)), "Acc_1_m": np.sum(np.abs(diff) <= 1) / diff.shape[0],
              "Median": np.median(np.abs(diff))}
    return {"GT": GTm, "Est_HM_no_Shift": est, "Conf_Range": conf}, scores, (np.nanmean(conf[:, :, 2]), np.nanmedian(conf[:, :, 2]))
