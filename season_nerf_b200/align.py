"""Seasonal alignment of cached render components to a target image ("next" row 3 of the scope table): drop-in for
`Grad_Descent_Seasonal_Align_v3` / `_grad_descent_v3` / `_grad_descent_v3_classic_shadows`
(T_NeRF_Eval_Utils/mg_Img_Eval.py:349-475).

The reference loops over 367 candidate times; each iteration recomposites the seasonal colour of every ray and solves a
closed-form least squares for the sky colour.  Here the 367 recompositions are ONE launch of the fused year-sweep kernel
(components read once, csrc/composite.cu) and the least squares / scores are batched reductions over [T, N, 3] on the
device; only the winning class vector, sky colour and time leave it."""
import numpy as np
import torch as t

from . import ops
from .render import _device_components


def _grad_descent_v3(Results_Dict, target_img, t0, network, device):
    """mg_Img_Eval.py:354-414 -> (class vector [C] cpu, sky colour [1,1,3] float32 cpu, best time of year)."""
    device = t.device(device)
    ts = t.tensor([t0] + list(np.linspace(0, 1, 366))).float()
    ts_scaled = t.stack([t.cos(ts * 2 * np.pi), t.sin(ts * 2 * np.pi), t.cos(ts * 2 * np.pi), t.sin(ts * 2 * np.pi)], 1)
    with t.no_grad():
        tv = network.get_class_only(ts_scaled.to(device))                                  # [T,C]
        ip = np.asarray(Results_Dict["Image_Points_in_GT_Img"])
        GT = t.tensor(np.asarray(target_img)[ip[:, 0], ip[:, 1]]).float().to(device).double()  # [N,3]
        rho, dl, base, vis, adj = _device_components(Results_Dict, ["Rho", "Deltas", "Base_Col", "Est_Solar_Vis", "Adjust_col"])
        N, S = rho.shape[0], rho.shape[1]
        cls = tv.double().contiguous()
        # A[t] = sum_s PS * sigmoid(Base + sum_c Adjust_c * tv[t,c]) for all T candidates at once          (:392)
        A = ops.year_sweep(rho.reshape(N, S), dl.reshape(N, S), base, adj, cls)            # [T,N,3] float64
        # Solar_Vis = sigmoid((sum_s PS * Est_Solar_Vis - .2) * 30)                                          (:381)
        _, _, _, raw, _ = ops.cli_composite(rho.reshape(N, S), dl.reshape(N, S), base, vis.reshape(N, S), adj, cls[0].contiguous())
        SV = t.sigmoid((raw - .2) * 30).unsqueeze(1)                                       # [N,1]
        good = (SV < .99)[:, 0]
        Ag, SVg, GTg = A[:, good], SV[good].unsqueeze(0), GT[good].unsqueeze(0)
        Y = GTg - Ag * SVg                                                                 # (:393-394)
        X = (1 - SVg) * Ag
        sky = t.clamp(t.sum(X * Y, 1) / t.sum(X * X, 1), 0, 1)                             # [T,3]  (:395-397)
        R = A * (SV.unsqueeze(0) + (1 - SV.unsqueeze(0)) * sky.unsqueeze(1))               # (:398-399)
        scores = t.mean((R - GT.unsqueeze(0)) ** 2, (1, 2))                                # MSELoss over rays and channels
        best = int(t.argmin(scores))
    return tv[best].float().cpu(), sky[best].float().reshape(1, 1, 3).cpu(), ts[best].item()


def _grad_descent_v3_classic_shadows(Results_Dict, target_img, t0, network, device):
    """mg_Img_Eval.py:416-475: the solar visibility shades every sample inside the colour sum, so each candidate time
    needs two recompositions, sum_s PS*col_t*vis and sum_s PS*col_t - two launches of the year-sweep kernel (the first
    with the per-sample weight vis) - and the sky least squares runs over all rays."""
    device = t.device(device)
    ts = t.tensor([t0] + list(np.linspace(0, 1, 366))).float()
    ts_scaled = t.stack([t.cos(ts * 2 * np.pi), t.sin(ts * 2 * np.pi), t.cos(ts * 2 * np.pi), t.sin(ts * 2 * np.pi)], 1)
    with t.no_grad():
        tv = network.get_class_only(ts_scaled.to(device))                                  # [T,C]
        ip = np.asarray(Results_Dict["Image_Points_in_GT_Img"])
        GT = t.tensor(np.asarray(target_img)[ip[:, 0], ip[:, 1]]).float().to(device).double()  # [N,3]
        rho, dl, base, vis, adj = _device_components(Results_Dict, ["Rho", "Deltas", "Base_Col", "Est_Solar_Vis", "Adjust_col"])
        N, S = rho.shape[0], rho.shape[1]
        cls = tv.double().contiguous()
        A_vis = ops.year_sweep(rho.reshape(N, S), dl.reshape(N, S), base, adj, cls, ps_weight=vis.reshape(N, S))   # [T,N,3]
        A_all = ops.year_sweep(rho.reshape(N, S), dl.reshape(N, S), base, adj, cls)
        Y = GT.unsqueeze(0) - A_vis                                                        # (:448)
        X = A_all - A_vis                                                                  # sum_s PS*col*(1 - vis)  (:449)
        xx = t.sum(X * X, 1)                                                               # [T,3]
        if not bool((xx > 0).all()):
            # the reference drops such channels (:450-452) and then fails on the shape of its own broadcast (:454-455)
            raise ValueError("classic-shadow alignment: a colour channel has no unshaded contribution (sum X*X == 0)")
        sky = t.clamp(t.sum(X * Y, 1) / xx, 0, 1)                                          # [T,3]
        R = A_vis + X * sky.unsqueeze(1)                                                   # sum_s PS*col*(vis + (1-vis)*sky)
        scores = t.mean((R - GT.unsqueeze(0)) ** 2, (1, 2))
        best = int(t.argmin(scores))
    return tv[best].float().cpu(), sky[best].float().reshape(1, 1, 3).cpu(), ts[best].item()


def Grad_Descent_Seasonal_Align_v3(Results_Dict, target_img, t0, network, device, batch_size=15000, steps=100,
                                   use_classic_shadows=False):
    """mg_Img_Eval.py:349-353 (batch_size / steps are unused by the reference's v3 as well)."""
    if use_classic_shadows:
        return _grad_descent_v3_classic_shadows(Results_Dict, target_img, t0, network, device)
    return _grad_descent_v3(Results_Dict, target_img, t0, network, device)
