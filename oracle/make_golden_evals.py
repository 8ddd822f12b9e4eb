"""Golden fixtures for the sigma-only volume consumers ("next" row 4), from the UNMODIFIED reference:
  * `eval_shadow_data` (T_NeRF_Eval_Utils/mg_Shadow_Eval.py:72-104): exact (marched) vs estimated solar visibility of columns
    of points above a grid of ground points, for a list of sun angles;
  * `eval_HM` (T_NeRF_Eval_Utils/Eval_funcs.py:298-395): expected height map, the 67 % confidence range loop (:315-330 - its
    mean / median are only PRINTED by the reference, so they are captured from stdout), the shifted height map and the
    scores before alignment (the shift / rotation search after :395 is evaluation tooling and not pinned).
Run in the build container: python -m oracle.make_golden_evals"""
import contextlib
import io
import os
import sys

import numpy as np
import torch as t

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import season_oracle as so            # noqa: E402
from oracle.make_golden import ref_net, save      # noqa: E402
from oracle.ref_import import import_reference    # noqa: E402


def main():
    ref = import_reference()
    t.set_num_threads(8)
    from T_NeRF_Eval_Utils.mg_Shadow_Eval import eval_shadow_data, shadow_anaylysis
    from T_NeRF_Eval_Utils.Eval_funcs import eval_HM
    P0 = so.init_params(seed=0, perturb_bn=True)
    net = ref_net(ref, P0)
    angles = np.array([[60., 140.], [25., 300.], [85., 10.]])
    gp = np.stack(np.meshgrid(np.linspace(-1, 1, 3), np.linspace(-1, 1, 2), indexing="ij"), -1).reshape([-1, 2])
    ve, vs, sk = eval_shadow_data(net, angles, gp, 96, so.OMA_W2C, so.oma_w2l_h(), 15000, t.device("cpu"))
    sa = shadow_anaylysis(gp, angles, {"Exact_Vis": ve, "Est_Vis": vs})
    out = dict(angles=angles, ground_points=gp, vis_exact=ve, vis_est=vs, sky_col=sk,
               sa_keys=np.array(sorted(sa.keys())), sa_vals=np.array([float(sa[k]) for k in sorted(sa.keys())], dtype=np.float64))
    g = t.Generator().manual_seed(97)
    GT = (t.rand(6, 5, generator=g) * 1.2 - 0.6).numpy().astype(np.float64)
    GT[2, 3] = np.nan
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        Imgs, before, after = eval_HM(net, GT.copy(), (300., 370.), 96, t.device("cpu"), 5000)
    first = buf.getvalue().splitlines()[0].split()
    out.update(hm_GT=GT, hm_GT_m=Imgs["GT"], hm_est_no_shift=Imgs["Est_HM_no_Shift"], conf_mean=np.float64(first[0]),
               conf_median=np.float64(first[1]), before_keys=np.array(sorted(before.keys())),
               before_vals=np.array([float(before[k]) for k in sorted(before.keys())], dtype=np.float64))
    np.savez_compressed(os.path.join(os.path.dirname(HERE), "tests", "golden", "evals.npz"), **out)
    print("wrote evals", first, before)


if __name__ == "__main__":
    main()
