"""Render CLI, drop-in for the reference's main_run_Season_NeRF.py (same flags, same files in --Model_Location:
opts.json, Final_Model.nn, W2C_W2L_H.npy).  Renders on the GPU with the fused sm_100a kernels; --Force_CPU is refused
(there is no CPU path)."""
import argparse
import datetime
import json
from collections import namedtuple

import numpy as np
import torch as t

from season_nerf_b200 import T_NeRF, component_render_by_dir, get_imgs_from_Img_Dict


def get_opts(argv=None):
    """main_run_Season_NeRF.py:10-44."""
    parser = argparse.ArgumentParser()
    parser.add_argument('--Model_Location', type=str, required=True, help="Location of files for input and output.")
    parser.add_argument('--VA', type=float, nargs=2, required=True, help="View elevation and azimuth angles in degrees.")
    parser.add_argument('--SA', type=float, nargs=2, required=True, help="Solar elevation and azimuth angles in degrees.")
    parser.add_argument('--tf', type=str, required=True, help="Month and Day in MM/DD format.")
    parser.add_argument('--Output_Size', type=int, nargs=3, required=False, default=(256, 256, 96),
                        help="Size of output image (n_rows, n_cols, sample per ray).")
    parser.add_argument('--Save_Name', type=str, required=False, help="Save image as Save_Name INSTEAD OF displaying image.")
    parser.add_argument('--ignore_progess', action='store_true', required=False, default=False,
                        help="Do not display rendering progress bars.")
    parser.add_argument('--exact_shadow', action='store_true', required=False, default=False,
                        help="Use exact shadow mask instead of estimated shadow mask.")
    parser.add_argument('--Force_CPU', action='store_true', required=False, default=False,
                        help="Use CPU for rendering, even if GPU is available.")
    return parser.parse_args(argv)


def load_args_from_json(path):
    """misc.py:13-20."""
    with open(path) as f:
        d = json.load(f)
    return namedtuple("args", d.keys())(*d.values())


def load_t_nerf(args, file_loc, model_name="Final_Model.nn"):
    """main_run_Season_NeRF.py:46-50."""
    net = T_NeRF(args.fc_units, args.number_low_frequency_cases)
    net.load_state_dict(t.load(file_loc + "/" + model_name, map_location=t.device("cpu")))
    return net


def load_model(file_loc):
    network_args = load_args_from_json(file_loc + "/opts.json")
    return load_t_nerf(network_args, file_loc), network_args


def parse_time(time_str):
    """main_run_Season_NeRF.py:59-63."""
    ans = datetime.datetime.strptime(time_str, "%m/%d")
    return (ans - datetime.datetime.strptime("01/01", "%m/%d")).days * 1. / 365


def render(args):
    if args.Force_CPU or not t.cuda.is_available():
        raise SystemExit("season_nerf_b200 renders on CUDA only: --Force_CPU / CPU-only hosts are not supported")
    device = t.device("cuda:0")
    the_model, _ = load_model(args.Model_Location)
    W = np.load(args.Model_Location + "/W2C_W2L_H.npy", allow_pickle=True).item()
    the_model = the_model.eval().to(device)
    size = tuple(args.Output_Size)
    raw = component_render_by_dir(the_model, args.VA, args.SA, parse_time(args.tf), size, W2C=W.get("W2C"),
                                  W2L_H=W.get("W2L_H"), include_exact_solar=args.exact_shadow, device=device)
    imgs = get_imgs_from_Img_Dict(raw, size, False)
    key = "Shadow_Adjust_Exact" if args.exact_shadow and "Shadow_Adjust_Exact" in imgs else "Shadow_Adjust"
    # the reference multiplies by Shadow_Adjust even with --exact_shadow (main_run_Season_NeRF.py:90); kept.
    return imgs["Season_Adj_Img"] * imgs["Shadow_Adjust"], imgs, key


def _main():
    args = get_opts()
    out_img, _, _ = render(args)
    if args.Save_Name:
        try:
            from matplotlib import pyplot as plt
            plt.imsave(args.Save_Name, np.clip(np.nan_to_num(out_img), 0, 1))
        except ImportError:
            np.save(args.Save_Name + ".npy", out_img)
    else:
        from matplotlib import pyplot as plt
        plt.imshow(out_img)
        plt.show()


if __name__ == '__main__':
    _main()
