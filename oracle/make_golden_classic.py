"""Golden fixture for the classic-shadow conventions: `get_imgs_from_Img_Dict(..., use_classic_shadows=True)`
(T_NeRF_Eval_Utils/mg_Img_Eval.py:166-181) and `_grad_descent_v3_classic_shadows` (:416-475), both run UNMODIFIED on the
component arrays already stored by the cli_render / cli_render_exact / season_align fixtures (which the reference
produced itself).  Run in the build container: python -m oracle.make_golden_classic"""
import os
import sys

import numpy as np
import torch as t

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import season_oracle as so            # noqa: E402
from oracle.ref_import import import_reference    # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _components(g, prefix):
    return {k[len(prefix):]: (g[k].astype(np.float64) if g[k].dtype == np.float32 else g[k]) for k in g.files if k.startswith(prefix)}


def main():
    ref = import_reference()
    me = ref.mg_Img_Eval
    out = {}
    g = np.load(os.path.join(OUT, "cli_render.npz"))
    R = me.get_imgs_from_Img_Dict(_components(g, "d_"), (5, 6, 96), True)
    out["sa_classic"] = R["Shadow_Adjust"]
    g = np.load(os.path.join(OUT, "cli_render_exact.npz"))
    R = me.get_imgs_from_Img_Dict(_components(g, "d_"), (2, 3, 96), True)
    out["sa_classic_x"], out["sae_classic_x"] = R["Shadow_Adjust"], R["Shadow_Adjust_Exact"]
    # alignment with classic shadows on the season_align components, same network as make_golden_align.py
    g = np.load(os.path.join(OUT, "season_align.npz"))
    P0 = {k: v.clone() for k, v in so.init_params(seed=0, perturb_bn=True).items()}
    P0["get_class_layer.weight"] *= float(g["class_scale"])
    net = ref.T_NeRF(512, 4)
    net.load_state_dict(P0, strict=True)
    net.eval()
    D = _components(g, "D_")
    adj, sky, best_t = me._grad_descent_v3_classic_shadows(D, g["target"], float(g["t0"]), net, t.device("cpu"))
    out.update(adj_vec=adj.numpy(), sky=sky.numpy(), best_t=np.array(best_t))
    np.savez_compressed(os.path.join(OUT, "cli_classic.npz"), **out)
    print("wrote cli_classic: best_t %.6f sky %s" % (best_t, sky.numpy().ravel()))


if __name__ == "__main__":
    main()
