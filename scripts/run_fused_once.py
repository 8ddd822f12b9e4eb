"""Launch the fused render kernel a few times on 65536 rays x 96 samples (for ncu captures)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as t
import season_nerf_b200 as snb
from season_nerf_b200 import fused

t.manual_seed(0)
net = snb.T_NeRF(512, 4).cuda().eval()
M = 65536 * 96
pts = (t.rand(M, 3, device="cuda") * 2 - 1)
sun = t.tensor([[0.3, -0.4, 0.866]], device="cuda")
with t.no_grad():
    for _ in range(4):
        s, e = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        s.record()
        fused.run(net, pts, sun, M)
        e.record()
        t.cuda.synchronize()
        ms = s.elapsed_time(e)
        print("fused %.3f ms  %.1f TFLOP/s algorithmic" % (ms, M * 5793792 / ms / 1e9))
