import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch as t
import bench
import season_nerf_b200 as snb


def make_step(n=4096, **kw):
    dev = t.device("cuda:0")
    args = bench.bench_args()
    W2C = np.array([41.2905, -95.8967, 315.0])
    H = np.eye(4)
    H[0, 0], H[1, 1], H[2, 2] = 2 / 0.0024, 2 / 0.0032, 2 / 70.0
    H[0, 3], H[1, 3], H[2, 3] = -W2C[0] * H[0, 0], -W2C[1] * H[1, 1], -W2C[2] * H[2, 2]
    t.manual_seed(0)
    ts = snb.TrainStep(args, dev, H, W2C, world_size=1, precision="bf16", **kw)
    g = t.Generator().manual_seed(1)
    xy = (t.rand(n, 2, generator=g) * 2 - 1) * 0.8
    dxy = (t.rand(n, 2, generator=g) * 2 - 1) * 0.2
    sun = t.nn.functional.normalize(t.rand(n, 3, generator=g) + t.tensor([0., 0., .5]), dim=1)
    tim = t.rand(n, 4, generator=g)
    batch = {"Top": t.cat([xy, t.ones(n, 1)], 1), "Bot": t.cat([xy + dxy, -t.ones(n, 1)], 1), "Sun_Angle": sun,
             "Time_Encoded": tim, "GT_Color": t.rand(n, 3, generator=g)}
    batch = {k: v.contiguous().to(dev) for k, v in batch.items()}
    np.random.seed(3); t.manual_seed(3)
    st, en, vec, tm, _ = ts.eval_tool.solar_creation_tool(n, include_times=True)
    solar = tuple(x.to(dev) for x in (st, en, vec, tm))
    jit = t.rand(96)
    return ts, batch, solar, jit
