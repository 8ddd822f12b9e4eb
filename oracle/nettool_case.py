"""Seeded inputs of the training-driver parity case (TEST INFRASTRUCTURE, shared by oracle/make_golden_nettool.py, which
runs the UNMODIFIED reference `T_NeRF_Net_Tool` on them, and by tests/test_net_tool_gpu.py, which runs the drop-in).

10 training steps are planned (`max_train_steps=10` -> section 1 = steps 0-1 DSM-guided, section 4 = steps 2-9, save points of
the last section at steps 5, 7, 10: Net_Tool_2.py:23-54); the case runs the first 5: two guided steps, the section switch
with the carried adaptive-loss state, three free steps, then eval_step + eval_img at the first save point."""
from types import SimpleNamespace

import numpy as np
import torch as t

from . import season_oracle as so

N_STEPS_RUN = 5
BATCH = 8


def args(logs_dir):
    return SimpleNamespace(chunk=1024 * 10, n_samples=96, n_importance=0, n_saves=4, max_train_steps=10, batch_size=BATCH,
                           sc_lambda=0.03, logs_dir=logs_dir, exp_name="nettool_case", fc_units=512, number_low_frequency_cases=4,
                           lr=10 ** (-4.86), lr_alpha_scale=1000, Use_MSE_loss=False, jump_start=True, Use_Reg=True,
                           Solar_Type_2=False, Use_Solar=True, use_auto_balance=False, use_HSLuv=False)


def dsms():
    g = t.Generator().manual_seed(71)
    training = (t.rand(16, 16, generator=g) * 1.2 - 0.6).numpy().astype(np.float64)
    gt = training + (t.rand(16, 16, generator=g) * 0.2 - 0.1).numpy()
    gt[3, 5] = np.nan
    return training, gt


def table(n, seed, n_images, img_pts=None):
    """[n, 22] rows in the column layout of mg_run_NeRF.py:122-133"""
    b = so.synthetic_batch(n, seed=seed, n_images=n_images)
    g = t.Generator().manual_seed(seed + 1000)
    if img_pts is None:
        img_pts = t.randint(0, 4, (n, 2), generator=g).float()
    view = t.nn.functional.normalize(b["Top"] - b["Bot"], dim=1)
    w = t.ones(n, 1)
    return t.cat([img_pts, b["Top"], b["Bot"], view, b["Sun_Angle"], b["Time_Encoded"], w, b["GT_Color"]], 1).float()


def train_batches():
    return [table(BATCH, 81 + i, 3) for i in range(N_STEPS_RUN)]


def val_table():
    """two 4x4 validation images, every pixel once -> (table [32,22], img_ids, full_img_size, img_names)"""
    ip = t.stack(t.meshgrid(t.arange(4), t.arange(4), indexing="ij"), -1).reshape(-1, 2).float()
    tabs = [table(16, 91 + i, 1, img_pts=ip) for i in range(2)]
    ids = [0] * 16 + [1] * 16
    return t.cat(tabs, 0), ids, [(4, 4, 3)] * 2, ["val_a", "val_b"]


def seed_step(i):
    """RNG state before step i's get_loss draws (jitter, solar rays, solar jitter: torch global + numpy global)"""
    t.manual_seed(500 + i)
    np.random.seed(500 + i)


class Recorder:
    def __init__(self):
        self.scalars, self.images = [], []

    def add_scalar(self, tag, value, step):
        self.scalars.append((tag, float(value), int(step)))

    def add_image(self, tag, img, step):
        self.images.append((tag, np.asarray(img, dtype=np.float64).copy(), int(step)))
