"""Solar-visibility evaluation on columns of points ("next" row 4 of the scope table): drop-in for `eval_shadow_data`,
`Test_Shadow_Points` and `shadow_anaylysis` (T_NeRF_Eval_Utils/mg_Shadow_Eval.py:72-163).

For every sun angle and every ground point the reference marches one ray through the cube along the sun direction
(`sample_pt_coarse`, eval mode, out-of-cube steps zeroed), evaluates `forward_Solar` and turns the density into the exact
transmittance `get_PV`, one (angle, chunk of ground points) at a time with a host round trip each.  Here ALL angles x ground
points form one batch of rays: positions from the bit-exact sampling kernel, the fused network in its solar program (sun
direction per ray), the transmittance scan kernel - results leave the device once."""
import numpy as np
import torch as t

from . import ops
from .engine import sample_ts
from .geometry import world_angle_2_local_vec


def eval_shadow_data(shadow_net, shadow_angles, ground_points, Z_points, world_center_LLA, W2L_H, max_batch_size, device):
    """mg_Shadow_Eval.py:72-104 -> Results_Vis_Exact [A,G,Z,1], Results_Vis_Est [A,G,Z,1], Results_Sky_Col [A,3] (float64 numpy;
    the sky colour is the RAW head, as forward_Solar returns it, T_NeRF_net_v2.py:157).  `max_batch_size` is accepted for
    compatibility; chunking is by device memory."""
    device = t.device(device)
    if device.type != "cuda":
        raise ops._lib.SeasonNerfCudaError("season_nerf_b200.eval_shadow_data runs on CUDA only (no CPU fallback)")
    shadow_angles, ground_points = np.asarray(shadow_angles, dtype=np.float64), np.asarray(ground_points, dtype=np.float64)
    A, G, Z = shadow_angles.shape[0], ground_points.shape[0], int(Z_points)
    vec0 = np.array([world_angle_2_local_vec(shadow_angles[i, 0], shadow_angles[i, 1], world_center_LLA, W2L_H) for i in range(A)])
    vec = vec0 / vec0[:, -1::]                                                                    # :81
    gp3 = np.expand_dims(np.concatenate([ground_points, np.zeros([G, 1])], 1), 0)
    tops = t.tensor(gp3 + np.expand_dims(vec, 1)).float().reshape(A * G, 3).to(device)            # :82-83 (float64 sum, then .float())
    bots = t.tensor(gp3 - np.expand_dims(vec, 1)).float().reshape(A * G, 3).to(device)
    sun = t.tensor(vec0).float().to(device).repeat_interleave(G, 0)                               # :89: the UN-normalised vector
    ts = sample_ts(Z, eval_mode=True).to(device)
    was_training = shadow_net.training
    shadow_net.eval()
    N = A * G
    PV = t.empty(N, Z, device=device, dtype=t.float32)
    Vis = t.empty(N, Z, device=device, dtype=t.float32)
    Sky = t.empty(N, 3, device=device, dtype=t.float32)
    step = max(1, (1 << 22) // Z)
    with t.no_grad():
        for i in range(0, N, step):
            e = min(i + step, N)
            pts, deltas = ops.sample_rays(tops[i:e], bots[i:e], ts, zero_oob=True)               # :87-88
            rho_raw, vis_raw, sky_raw = shadow_net.forward_rays(pts.reshape(-1, 3), sun[i:e], None, Z, mode="solar")
            rho = shadow_net.Softplus(rho_raw).reshape(e - i, Z)
            z3 = t.zeros(e - i, Z, 3, device=device)
            PV[i:e] = ops.composite_fwd(rho.contiguous(), deltas, z3, t.zeros_like(rho), t.zeros(e - i, 3, device=device))[0]   # get_PV, :96
            Vis[i:e] = shadow_net.Sigmoid(vis_raw).reshape(e - i, Z)
            Sky[i:e] = sky_raw if sky_raw.shape[0] == e - i else sky_raw.expand(e - i, 3)
    shadow_net.train(was_training)
    f = lambda x: x.double().cpu().numpy()
    return f(PV).reshape(A, G, Z, 1), f(Vis).reshape(A, G, Z, 1), f(Sky).reshape(A, G, 3)[:, 0, :]


def shadow_anaylysis(Ground_Points, Solar_el_az, Results_Dict):
    """mg_Shadow_Eval.py:134-163 (numpy scores of a result dict; the reference's spelling)."""
    E, V = Results_Dict["Exact_Vis"], Results_Dict["Est_Vis"]
    direct_loss = np.mean((E - V) ** 2)
    avg_error = np.mean(np.abs(E - V))
    thresh_GT, thresh_Est = E > .5, V > .5
    TP = np.sum(thresh_GT * thresh_Est)
    TN = np.sum((~thresh_GT) * (~thresh_Est))
    FP = np.sum((~thresh_GT) * thresh_Est)
    FN = np.sum(thresh_GT * (~thresh_Est))
    with np.errstate(divide="ignore", invalid="ignore"):
        Acc = (TP + TN) / (TP + TN + FP + FN)
        Prec_Sun = np.float64(TP) / (TP + FP)
        Recall_Sun = np.float64(TP) / (TP + FN)
        Prec_Shadow = np.float64(TN) / (TN + FN)
        Recall_Shadow = np.float64(TN) / (TN + FP)
    Surf_Dist = np.sum(thresh_GT, 2) - np.sum(thresh_Est, 2)
    avg_offset = np.mean(np.abs(Surf_Dist))
    return {"Acc": Acc, "Prec_Sun": Prec_Sun, "Recall_Sun": Recall_Sun, "Prec_Shadow": Prec_Shadow, "Recall_Shadow": Recall_Shadow,
            "Loss": direct_loss, "Avg_Error": avg_error, "Avg_Offset": avg_offset}


def Test_Shadow_Points(shadow_net, training_points, testing_points, close_walking_points, all_walking_points, ground_points,
                       world_center_LLA, W2L_H, device, Z_points=96, max_batch_size=15000, full_return=True):
    """mg_Shadow_Eval.py:107-131."""
    names = (("Training", training_points), ("Testing", testing_points), ("Near_Walk", close_walking_points), ("Full_Walk", all_walking_points))
    S = {"Ground_Points": ground_points, "Sun_El_Az": {k: v for k, v in names}}
    for (k, pts), rk in zip(names, ("Training_Results", "Testing_Results", "Near_Results", "Full_Results")):
        ve, vs, sk = eval_shadow_data(shadow_net, pts, ground_points, Z_points, world_center_LLA, W2L_H, max_batch_size, device)
        S[rk] = {"Exact_Vis": ve, "Est_Vis": vs, "Sky_Col": sk}
    if full_return is False:
        S = {short: shadow_anaylysis(S["Ground_Points"], S["Sun_El_Az"][k], S[rk]) for short, k, rk in
             (("Training", "Training", "Training_Results"), ("Testing", "Testing", "Testing_Results"),
              ("Near", "Near_Walk", "Near_Results"), ("Full", "Full_Walk", "Full_Results"))}
    return S
