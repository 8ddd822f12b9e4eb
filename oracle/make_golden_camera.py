"""Golden fixture for the camera-ray generator ("next" row 1 of the scope table): runs the UNMODIFIED reference
`P_img_Pinhole.invert_P` (pre_NeRF/P_Img.py:133-147) and the bounds filters of `component_render_by_P`
(T_NeRF_Eval_Utils/mg_Img_Eval.py:74-94) / `setup_quick_loader` (mg_Pt_holder.py:169-187) on a synthetic OMA_281-like
affine camera.  float64 arrays are stored as float64 (the CUDA kernel must reproduce them bit for bit).
Run in the build container only:  python -m oracle.make_golden_camera"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import season_oracle as so            # noqa: E402
from oracle.ref_import import import_reference    # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    import_reference()                                      # stubs gdal / rpcm / astropy, puts the reference on sys.path
    from pre_NeRF.P_Img import P_img_Pinhole
    P = so.synthetic_camera_P(seed=7)
    cam = object.__new__(P_img_Pinhole)                     # the constructor needs a satellite image + RPC (absent offline)
    cam.P = P.copy()
    cam.norm_P()                                            # P_Img.py:128-131
    H_img, W_img = 2048, 2048
    # component_render_by_P pixel list (mg_Img_Eval.py:76-77), 37 x 29 output
    out_size = (37, 29)
    XY = np.stack(np.meshgrid(np.linspace(0, H_img - 1, out_size[0]), np.linspace(0, W_img - 1, out_size[1]), indexing="ij"), -1)
    XY = np.round(XY).astype(int).reshape([-1, 2])
    xt, yt, _ = cam.invert_P(XY[:, 0], XY[:, 1], 1.)
    xb, yb, _ = cam.invert_P(XY[:, 0], XY[:, 1], -1.)
    tops = np.stack([xt, yt, np.ones_like(xt)], -1)
    bots = np.stack([xb, yb, -np.ones_like(xb)], -1)
    good = (tops[:, 0] >= -1) * (tops[:, 1] <= 1) * (bots[:, 0] >= -1) * (bots[:, 1] <= 1) * \
           (tops[:, 1] >= -1) * (tops[:, 0] <= 1) * (bots[:, 1] >= -1) * (bots[:, 0] <= 1)
    # setup_quick_loader raster (mg_Pt_holder.py:169-187) with downscale 16 and model bounds
    DS = 32
    shp = np.array([H_img, W_img]) // DS
    bounds_model = np.array([[-1., 1.], [-1., 1.], [-1., 1.]])
    XYg = np.stack([np.repeat(np.arange(0, shp[0]), shp[1]), np.tile(np.arange(0, shp[1]), shp[0])], -1)
    Zt = np.ones([XYg.shape[0]]) * bounds_model[2, 1]
    Zb = np.ones([XYg.shape[0]]) * bounds_model[2, 0]
    topsg = np.stack(cam.invert_P(XYg[:, 0] * DS, XYg[:, 1] * DS, Zt))
    botsg = np.stack(cam.invert_P(XYg[:, 0] * DS, XYg[:, 1] * DS, Zb))
    goodg = (topsg[0] <= bounds_model[0, 1]) * (bounds_model[0, 0] <= topsg[0]) * \
            (topsg[1] <= bounds_model[1, 1]) * (bounds_model[1, 0] <= topsg[1]) * \
            (botsg[0] <= bounds_model[0, 1]) * (bounds_model[0, 0] <= botsg[0]) * \
            (botsg[1] <= bounds_model[1, 1]) * (bounds_model[1, 0] <= botsg[1])
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "camera_rays.npz"), P=cam.P, img_shape=np.array([H_img, W_img]),
                        out_size=np.array(out_size), XY=XY.astype(np.int64), tops64=tops, bots64=bots, good=good,
                        tops32=tops.astype(np.float32), bots32=bots.astype(np.float32),
                        DS=np.array(DS), grid_xy64=np.stack([topsg[0], topsg[1], botsg[0], botsg[1]], -1), grid_good=goodg)
    print("wrote camera_rays: list %d rays (%d inside), grid %d rays (%d inside)" % (XY.shape[0], good.sum(), XYg.shape[0], goodg.sum()))


if __name__ == "__main__":
    main()
