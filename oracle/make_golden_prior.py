"""Golden fixture for the DSM-guided (use_prior) section from the UNMODIFIED reference: `All_in_One_Eval.eval` with
`use_prior=True` (Eval_Tools_2.py:217-246: the supervised / merged densities are shaded with the Solar_Vis3 of the
network's own, unmerged PS) and `get_loss` + backward with `--Use_MSE_loss` and the prior, where `Rendered_Col_Merged`
is the training target (Eval_Tools_2.py:399-405).
Run in the build container: python -m oracle.make_golden_prior"""
import os
import sys

import numpy as np
import torch as t

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import season_oracle as so            # noqa: E402
from oracle.make_golden import S, rays, ref_net, save  # noqa: E402
from oracle.ref_import import import_reference    # noqa: E402

PRIOR_KEYS = ["Rendered_Col", "Albedo_Color", "PS", "PV_Supervised", "PE_Supervised", "PS_Supervised", "Rendered_Col_Supervised",
              "PV_Merged", "PE_Merged", "PS_Merged", "Rendered_Col_Merged", "Rho_Merged"]


def main():
    ref = import_reference()
    t.set_num_threads(8)
    P0 = so.init_params(seed=0, perturb_bn=True)
    hm = (t.rand(16, 16, generator=t.Generator().manual_seed(61)) * 1.2 - 0.6).numpy()
    out = {"hm": hm}
    # ---- eval(), eval mode and train mode (batch-statistic BatchNorm), trust = 30 / 100 ----
    d = rays(6, 61)
    out.update({"in_" + k: v for k, v in d.items()})
    tool = ref.All_in_One_Eval(so.default_args(), t.device("cpu"), 100, True, None, so.oma_w2l_h(), so.OMA_W2C)
    net = ref_net(ref, P0, hm=hm)
    with t.no_grad():
        R = tool.eval(d, net, 30, False)
    out.update({"ev_" + k: R[k] for k in PRIOR_KEYS})
    netT = ref_net(ref, P0, hm=hm, train=True)
    t.manual_seed(62)
    jit = t.rand(S)
    t.manual_seed(62)
    with t.no_grad():
        R = tool.eval(d, netT, 30, True)
    out["jitter"] = jit
    out.update({"tr_" + k: R[k] for k in PRIOR_KEYS})
    save("engine_eval_prior", **out)

    # ---- get_loss + backward: MSE colour loss on Rendered_Col_Merged, prior on ----
    n, seed = 8, 63
    a = so.default_args(Use_MSE_loss=True)
    nt = ref_net(ref, P0, hm=hm, train=True)
    tl = ref.All_in_One_Eval(a, t.device("cpu"), 100, True, None, so.oma_w2l_h(), so.OMA_W2C)
    dd = rays(n, seed)
    t.manual_seed(seed)
    np.random.seed(seed)
    jit = t.rand(S)
    az_el = np.random.random(n * 2).reshape([n, 2]) * np.array([[360, 89]]) + np.array([[-180, 1]])
    sx, sy = t.rand(n), t.rand(n)
    fr = t.rand(n, 2) * 2 * np.pi
    sjit = t.rand(S)
    t.manual_seed(seed)
    np.random.seed(seed)
    L = tl.get_loss(dd, nt, 30, True)
    tot = 0
    for k in L:
        tot = tot + L[k][0] * L[k][1]
    tot.backward()
    vec = np.array([ref.world_angle_2_local_vec(az_el[i][1], az_el[i][0], so.OMA_W2C, so.oma_w2l_h()) for i in range(n)])
    starts = t.ones(n, 3)
    starts[:, 0] = 2 * sx - 1
    starts[:, 1] = 2 * sy - 1
    ends = (starts - 2 * (vec / vec[:, 2::])).float()
    stimes = t.stack([t.cos(fr[:, 0]), t.sin(fr[:, 0]), t.cos(fr[:, 1]), t.sin(fr[:, 1])], 1)
    out = {"in_" + k: v for k, v in dd.items()}
    out.update(hm=hm, jitter=jit, solar_jitter=sjit, s_top=starts, s_bot=ends, s_sun=t.tensor(vec).float(), s_time=stimes,
               total=tot.detach())
    for k in L:
        out["loss_" + k] = t.as_tensor(L[k][0]).detach()
        out["w_" + k] = np.float32(float(L[k][1]))
    small = ["G_NeRF_net.fc10Sigma.weight", "G_NeRF_net.fc10Col.weight", "G_NeRF_net.fc2.norm.weight", "G_NeRF_net.fc9.linear.bias",
             "adjust_col.weight"]
    names, norms = [], []
    for k, prm in nt.named_parameters():
        names.append(k)
        norms.append(0.0 if prm.grad is None else float(prm.grad.norm()))
        if k in small:
            out["grad_" + k] = prm.grad
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array(norms, dtype=np.float64)
    sdd = nt.state_dict()
    out["fc2_rv"] = sdd["G_NeRF_net.fc2.norm.running_var"]
    save("loss_prior_mse", **out)


if __name__ == "__main__":
    main()
