"""Golden fixture for the remaining public entry points of the hot path, from the UNMODIFIED reference:
`All_in_One_Eval.full_eval` (Eval_Tools_2.py:127-163, also with classic solar), `T_NeRF.approx_Solar`
(T_NeRF_net_v2.py:107-128), `T_NeRF.forward_full_eval` (:184-204), `G_NeRF_Net_Classic.forward_Position` (G_NeRF.py:93-98)
and `create_solor_rays_uniform.create_given_vec` (Eval_Tools_2.py:50-70).
Run in the build container: python -m oracle.make_golden_api"""
import os
import sys

import numpy as np
import torch as t

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import season_oracle as so            # noqa: E402
from oracle.make_golden import rays, ref_net, save  # noqa: E402
from oracle.ref_import import import_reference    # noqa: E402


def main():
    ref = import_reference()
    P0 = so.init_params(seed=0, perturb_bn=True)
    net = ref_net(ref, P0)
    out = {}
    g = t.Generator().manual_seed(41)
    M = 160
    X = t.rand(M, 3, generator=g) * 2 - 1
    Xs = t.rand(M, 3, generator=g) * 2 - 1
    sun = t.nn.functional.normalize(t.rand(M, 3, generator=g) + t.tensor([-.5, -.5, .2]), dim=1)
    f = t.rand(M, generator=g)
    Time = t.stack([t.cos(2 * np.pi * f), t.sin(2 * np.pi * f), t.ones(M), t.zeros(M)], 1)
    with t.no_grad():
        ap = net.approx_Solar(X, Xs, Time)
        fe = net.forward_full_eval(X, sun, Time)
        fp = net.G_NeRF_net.forward_Position(X)
    out.update(X=X, Xs=Xs, sun=sun, Time=Time)
    out.update({"ap_%d" % i: v for i, v in enumerate(ap)})
    out.update({"fe_%d" % i: v for i, v in enumerate(fe)})
    out.update({"fp_%d" % i: v for i, v in enumerate(fp)})
    d = rays(6, 51)
    out.update({"in_" + k: v for k, v in d.items()})
    keys = ["Rendered_Col", "PE", "PV", "PS", "Solar_Vis", "Sky_Col", "Classes", "Adjust", "Rho", "Col", "deltas", "sample_pts"]
    for tag, classic in (("full", False), ("fullc", True)):
        tool = ref.All_in_One_Eval(so.default_args(Solar_Type_2=classic), t.device("cpu"), 100, False, None, so.oma_w2l_h(), so.OMA_W2C)
        with t.no_grad():
            R = tool.full_eval(d, net, 0)
        out.update({"%s_%s" % (tag, k): R[k] for k in keys})
    # create_given_vec: torch global generator draws (start positions, then the time fractions)
    gen = ref.create_solor_rays_uniform(so.oma_w2l_h(), so.OMA_W2C)
    vec = ref.world_angle_2_local_vec(40.0, 120.0, so.OMA_W2C, so.oma_w2l_h())
    t.manual_seed(9)
    st, en, sv, tm = gen.create_given_vec(12, vec, include_times=True)
    out.update(gv_vec=vec, gv_starts=st, gv_ends=en, gv_sun=sv, gv_times=tm)
    save("api_extra", **out)


if __name__ == "__main__":
    main()
