"""Golden fixture for the seasonal-alignment search ("next" row 3): the UNMODIFIED reference `_grad_descent_v3`
(T_NeRF_Eval_Utils/mg_Img_Eval.py:354-414) on cached components of a small synthetic view whose target image is the
reference's own render at a known time of year.  Run in the build container: python -m oracle.make_golden_align"""
import os
import sys

import numpy as np
import torch as t

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import season_oracle as so            # noqa: E402
from oracle.ref_import import import_reference    # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
S = 32


def main():
    ref = import_reference()
    me = ref.mg_Img_Eval
    P0 = so.init_params(seed=0, perturb_bn=True)
    # sharpen the time dependence of the random network so that the search has a clear minimum
    P0 = {k: v.clone() for k, v in P0.items()}
    P0["get_class_layer.weight"] *= 40.0
    net = ref.T_NeRF(512, 4)
    net.load_state_dict(P0, strict=True)
    net.eval()
    size = (7, 6, S)
    dev = t.device("cpu")
    D = me.component_render_by_dir(net, [80, 0], [45, 135], 0.3, size, so.OMA_W2C, so.oma_w2l_h(), dev, include_exact_solar=False)
    D["Image_Points_in_GT_Img"] = D["Image_Points"]
    # target = the render at t* = 200/365 with a known sky colour
    t_star = np.linspace(0, 1, 366)[200]
    te = t.tensor([[np.cos(t_star * 2 * np.pi), np.sin(t_star * 2 * np.pi), np.cos(t_star * 2 * np.pi), np.sin(t_star * 2 * np.pi)]]).float()
    with t.no_grad():
        cv = net.get_class_only(te).numpy().astype(np.float64)
    imgs = me.get_imgs_from_Img_Dict_t_step(D, size, cv)
    target = np.nan_to_num(imgs[0])
    adj, sky, best_t = me._grad_descent_v3(D, target, 0.3, net, dev)
    keep = ["Rho", "Deltas", "Base_Col", "Adjust_col", "Est_Solar_Vis", "Sky_Col", "Output_class", "Image_Points", "Image_Points_in_GT_Img"]
    np.savez_compressed(os.path.join(OUT, "season_align.npz"), size=np.array(size), t0=np.array(0.3), t_star=np.array(t_star),
                        target=target, class_scale=np.array(40.0), adj_vec=adj.numpy(), sky=sky.numpy(), best_t=np.array(best_t),
                        **{"D_" + k: (np.asarray(D[k]).astype(np.float32) if np.asarray(D[k]).dtype == np.float64 else np.asarray(D[k])) for k in keep})
    print("wrote season_align: best_t %.6f (t* %.6f) sky %s adj %s" % (best_t, t_star, sky.numpy().ravel(), adj.numpy()))


if __name__ == "__main__":
    main()
