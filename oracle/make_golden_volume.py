"""Golden fixture for the dense-volume consumers ("next" row 4): the UNMODIFIED reference `gen_results`
(T_NeRF_Eval_Utils/Eval_funcs.py:268-296) on the seeded network.  Run in the build container: python -m oracle.make_golden_volume"""
import os
import sys

import numpy as np
import torch as t

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import season_oracle as so            # noqa: E402
from oracle.ref_import import import_reference    # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    ref = import_reference()
    from T_NeRF_Eval_Utils import Eval_funcs
    P0 = so.init_params(seed=0, perturb_bn=True)
    net = ref.T_NeRF(512, 4)
    net.load_state_dict({k: v.clone() for k, v in P0.items()}, strict=True)
    shape, S = (6, 5), 24
    rho, pe, pv, ps, col = Eval_funcs.gen_results(net, shape, S, t.device("cpu"), 1000)
    hm = np.sum(ps * np.linspace(1, -1, S).reshape([1, 1, -1]), 2) / np.sum(ps, 2)          # Eval_funcs.py:313
    np.savez_compressed(os.path.join(OUT, "gen_results.npz"), shape=np.array(shape), S=np.array(S), rho=rho.astype(np.float32),
                        P_E=pe.astype(np.float32), P_Vis=pv.astype(np.float32), P_Surf=ps.astype(np.float32),
                        col=col.astype(np.float32), height=hm.astype(np.float32))
    print("wrote gen_results", rho.shape, float(rho.max()))


if __name__ == "__main__":
    main()
