"""Where does the time of one rank's shard of an 8-way sharded 1024^2 render go?  (run on one GPU)"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as t
import season_nerf_b200 as snb
from season_nerf_b200 import fused, ops, render
from season_nerf_b200.engine import sample_ts
from bench_extras import oma_frame

S = 96
dev = t.device("cuda:0")
W2C, H = oma_frame()
t.manual_seed(0)
net = snb.T_NeRF(512, 4).to(dev).eval()
size = (1024, 1024, S)


def wall(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        t.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        t.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3, out


snb.render_image_sharded(net, [80, 0], [45, 135], 184 / 365, (64, 64, S), W2C, H, dev)
ms, (lo, hi, rgb, mask) = wall(lambda: snb.render_shard(net, [80, 0], [45, 135], 184 / 365, size, W2C, H, dev, 0, 8))
print("render_shard (rank 0 of 8, %d rays): %.1f ms" % (hi - lo, ms))
tops, bots = render.view_rays([80, 0], size, W2C, H)
tops, bots = tops[lo:hi].to(dev), bots[lo:hi].to(dev)
ts_ = sample_ts(S, True, True).to(dev)
sun = t.tensor([[0.3, -0.4, 0.866]], device=dev)
with t.no_grad():
    pts, _ = ops.sample_rays(tops, bots, ts_, zero_oob=True)
    p = pts.reshape(-1, 3)
    ms_k, _ = wall(lambda: fused.run(net, p, sun, p.shape[0]))
    print("fused kernel alone on the same points: %.1f ms" % ms_k)
    ms_i, D = wall(lambda: render._internal_render(net, tops, bots, [0.3, -0.4, 0.866], 184 / 365, size, 150000, False, dev))
    print("_internal_render (sampling + kernel + activations into component arrays): %.1f ms" % ms_i)
full = t.rand(1024 * 1024, 3, dtype=t.float64, device=dev)
ms_c, _ = wall(lambda: full.cpu().numpy())
ms_p, _ = wall(lambda: render.device_to_numpy(full, chunk_bytes=8 << 20) if False else full.cpu().numpy())
print("image [1M,3] f64 .cpu().numpy(): %.1f ms" % ms_c)
