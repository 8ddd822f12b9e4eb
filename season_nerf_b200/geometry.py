"""Host geometry helpers (float64 numpy, vectorised).  reference: all_NeRF/mg_unit_converter.py:5-9,29-34,59-68,
T_NeRF_Full_2/Quick_Run.py:9-12."""
import numpy as np

_R_EARTH_KM = 6378.137


def LLA_get_vec(LLA_Center, theta_deg, rho_deg):
    """mg_unit_converter.py:59-68; theta/rho may be arrays -> [..., 3]."""
    theta_deg, rho_deg = np.asarray(theta_deg, dtype=np.float64), np.asarray(rho_deg, dtype=np.float64)
    Y = np.cos(np.deg2rad(theta_deg))
    X = np.sin(np.deg2rad(theta_deg))
    Z = np.tan(np.deg2rad(rho_deg)) * np.sqrt(X ** 2 + Y ** 2)
    norm = np.sqrt(X ** 2 + Y ** 2 + Z ** 2) / 1000
    X, Y, Z = X / norm, Y / norm, Z / norm
    dLat = Y / (1000. * _R_EARTH_KM)                                             # lat_lon_shift, :29-34
    dLon = X / (1000. * _R_EARTH_KM * np.cos(np.deg2rad(LLA_Center[0])))
    return np.stack([LLA_Center[0] + np.rad2deg(dLat), LLA_Center[1] + np.rad2deg(dLon), LLA_Center[2] + Z], -1)


def world_angle_2_local_vec(world_el, world_az, world_center, World2Local_H):
    """mg_unit_converter.py:5-9.  Scalars -> [3]; arrays of n angles -> [n,3]."""
    lla = LLA_get_vec(world_center, world_az, world_el)
    H = np.asarray(World2Local_H, dtype=np.float64)
    hom = np.concatenate([lla, np.ones(lla.shape[:-1] + (1,))], -1)
    temp = (hom @ H.T)[..., 0:3]
    return temp / np.sqrt(np.sum(temp ** 2, -1, keepdims=True))


def encode_time(time_frac_year, time_frac_day=0):
    """Quick_Run.py:9-12."""
    return np.array([np.cos(time_frac_year * 2 * np.pi), np.sin(time_frac_year * 2 * np.pi),
                     np.cos(time_frac_day * 2 * np.pi), np.sin(time_frac_day * 2 * np.pi)])


def ray_table_from_P(P, img_shape, downscale, bounds_model, device):
    """The per-image ray cache of mg_Pt_holder.setup_quick_loader (mg_Pt_holder.py:169-187) built on the device: every
    (down-scaled) pixel of an image with the affine-approximated RPC camera `P` (P_img_Pinhole.P after scale_P), its ray
    end points at z = bounds_model[2,1] / [2,0], filtered to the model bounds.
    -> valid_img_pts [n,2] int64 (down-scaled pixel indices), tops [n,3], bots [n,3] float32 (device tensors)."""
    import torch as t
    from . import ops
    H, W = int(img_shape[0]) // downscale, int(img_shape[1]) // downscale
    b = np.asarray(bounds_model, dtype=np.float64)
    tops, bots, good, _ = ops.camera_rays(P, t.device(device), grid=(H, W, downscale), z_top=b[2, 1], z_bot=b[2, 0],
                                          bounds=(b[0, 0], b[0, 1], b[1, 0], b[1, 1]))
    keep = good.nonzero().squeeze(1)
    pts = t.stack([keep // W, keep % W], 1)
    return pts, tops[keep], bots[keep]
