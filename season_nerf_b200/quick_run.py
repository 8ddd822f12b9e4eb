"""Programmatic renderer: drop-in for T_NeRF_Full_2/Quick_Run.py (encode_time :9-12, Quick_Run_Net :61-226).
Engine render convention (SURVEY 8a' item 4): linspace(0,1,S+1)[:-1], no OOB mask, float32 composite on device."""
import numpy as np
import torch as t

from .engine import All_in_One_Eval, create_solor_rays_uniform
from .geometry import encode_time, world_angle_2_local_vec


def _transform_output_dict(output_dict, out_img_size):
    """Quick_Run.py:14-35."""
    out_img = np.zeros([out_img_size, out_img_size, 3])
    mask = np.zeros([out_img_size, out_img_size], dtype=bool)
    XY = output_dict["XY"]
    out_img[XY[:, 0], XY[:, 1]] = output_dict["Rendered_Col"].numpy()
    mask[XY[:, 0], XY[:, 1]] = True
    imgs = {"Col_Img": out_img}
    for key, name in (("Solar_Vis", "Shadow_Mask"), ("Est_Solar_Vis", "Estimated_Shadow_Mask")):
        if key in output_dict.keys():
            sh = np.zeros([out_img_size, out_img_size])
            sh[XY[:, 0], XY[:, 1]] = t.sum(output_dict["PS"] * output_dict[key], 1)[:, 0]
            imgs[name] = sh
    return imgs, mask


def _transform_output_dict_for_dsm(output_dict, out_img_size):
    """Quick_Run.py:37-40 (expected height, 96 samples hard-coded by the reference)."""
    out_img = np.zeros([out_img_size[0], out_img_size[1], 1]) * np.nan
    n = output_dict["PS"].shape[1]
    out_img[output_dict["XY"][:, 0], output_dict["XY"][:, 1]] = np.sum(
        output_dict["PS"].numpy() * np.linspace(1, -1, n).reshape([1, -1, 1]), 1)
    return out_img[:, :, 0]


class Quick_Run_Net():
    def __init__(self, network, args, world_center_LLA, World_2_Local_H, device, max_input_size=50000, use_tqdm=False,
                 use_full_solar=True):
        self.eval_tool = All_in_One_Eval(args, device, 5, False, False, World_2_Local_H, world_center_LLA)
        self.network = network
        self.world_center_LLA = world_center_LLA
        self.W2L_H = World_2_Local_H
        self.n_samples = args.n_samples
        self.n_classes = args.number_low_frequency_cases
        self.use_tqdm = use_tqdm
        self.use_full_solar = use_full_solar
        # the reference chunks by host/GPU memory of 2019 hardware (:72-75); HBM3e takes far larger chunks
        self.max_input_size = 256 if use_full_solar else max(max_input_size // args.n_samples, 4096)

    def _get_input_dict(self, camera_el_az, solar_el_az, time_frac, out_img_size, region):
        """Quick_Run.py:77-109."""
        tup = isinstance(out_img_size, tuple)
        sz = out_img_size if tup else (out_img_size, out_img_size)
        X, Y = np.meshgrid(np.arange(0, sz[0]), np.arange(0, sz[1]), indexing="ij")
        XY = np.stack([X, Y], 2).reshape([-1, 2])
        if not tup:
            mids = np.concatenate([XY * 2. / (out_img_size - 1) - 1, np.zeros([XY.shape[0], 1])], 1)
        else:
            mids = np.concatenate([XY * 2. / (np.array([[sz[0], sz[1]]]) - 1) - 1, np.zeros([XY.shape[0], 1])], 1)
        if region is not None:
            mids[:, 0] = (mids[:, 0] + 1) / 2 * (region[1] - region[0]) + region[0]
            mids[:, 1] = (mids[:, 1] + 1) / 2 * (region[3] - region[2]) + region[2]
        cam = world_angle_2_local_vec(camera_el_az[0], camera_el_az[1], self.world_center_LLA, self.W2L_H)
        tops = mids + cam / cam[2]
        bots = mids - cam / cam[2]
        good = np.all((bots <= 1) * (bots >= -1) * (tops <= 1) * (tops >= -1), 1)
        tops, bots, XY = t.tensor(tops[good]).float(), t.tensor(bots[good]).float(), XY[good]
        if not tup:
            XY[:, 0] = out_img_size - XY[:, 0] - 1
        sv = world_angle_2_local_vec(solar_el_az[0], solar_el_az[1], self.world_center_LLA, self.W2L_H)
        n = XY.shape[0]
        return {"Top": tops, "Bot": bots, "XY": XY, "Sun_Angle": t.tensor(np.stack([sv] * n, 0)).float().reshape(n, 3),
                "Time_Encoded": t.tensor(np.stack([encode_time(time_frac)] * n, 0)).float().reshape(n, 4)}

    def _build_output_dict(self, n, rendered_col_only=False):
        """Quick_Run.py:111-139."""
        S, C = self.n_samples, self.n_classes
        if rendered_col_only:
            return {"Rendered_Col": t.zeros([n, 3])}
        d = {"Rendered_Col": t.zeros([n, 3]), "PE": t.zeros([n, S, 1]), "PV": t.zeros([n, S, 1]),
             "PS": t.zeros([n, S, 1]), "Solar_Vis": t.zeros([n, S, 1]), "Sky_Col": t.zeros([n, S, 3]),
             "Classes": t.zeros([n, S, C]), "Adjust": t.zeros([n, S, 3]), "Col": t.zeros([n, S, 3])}
        if self.use_full_solar:
            d["Est_Solar_Vis"] = t.zeros([n, S, 1])
        return d

    def solar_ray_acc_check(self, n_rays=500, solar_el_and_az=None, H=None, cent=None, solar_vec=None):
        """Quick_Run.py:142-170."""
        test = create_solor_rays_uniform(H, cent)
        if solar_el_and_az is None and solar_vec is None:
            starts, ends, vec, times, az_el = test(n_rays, include_times=True)
            d = {"Top": starts, "Bot": ends, "Sun_Angle": vec, "Time_Encoded": times, "Sun_Angle_Az_El": az_el}
        elif solar_el_and_az is None:
            starts, ends, vec, times = test.create_given_vec(n_rays, solar_vec, True)
            d = {"Top": starts, "Bot": ends, "Sun_Angle": vec, "Time_Encoded": times}
        else:
            raise NotImplementedError("Manual entry of solar rays as az el not yet implemented!")
        with t.no_grad():
            return self.eval_tool.eval_Rho_Only(d, self.network, False)

    def _render(self, input_dict, exact):
        out = self._build_output_dict(input_dict["Top"].shape[0])
        n = input_dict["Top"].shape[0]
        for i in range(0, n, self.max_input_size):
            e = min(i + self.max_input_size, n)
            sub = {k: input_dict[k][i:e] for k in ("Top", "Bot", "Sun_Angle", "Time_Encoded")}
            if exact:
                res = self.eval_tool.eval_exact_solar(sub, self.network, -1, False)
            else:
                res = self.eval_tool.eval(sub, self.network, -1, False)
            for k in out.keys():
                out[k][i:e] = res[k]
        out["XY"] = input_dict["XY"]
        return out

    def render_img(self, camera_el_and_az, solar_el_and_az, time_frac, out_img_size, region=None):
        """Quick_Run.py:173-205."""
        with t.no_grad():
            d = self._get_input_dict(camera_el_and_az, solar_el_and_az, time_frac, out_img_size, region)
            out = self._render(d, self.use_full_solar)
            return _transform_output_dict(out, out_img_size)

    def get_DSM(self, out_img_size, region=None):
        """Quick_Run.py:207-226."""
        with t.no_grad():
            d = self._get_input_dict([90, 0], [90, 0], 0.0, out_img_size, region)
            out = self._render(d, False)
            return _transform_output_dict_for_dsm(out, out_img_size)
