"""Can a persistent tcgen05 GEMM and an HBM-bound element-wise kernel share the SMs?  (two streams vs one)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as t
from season_nerf_b200 import ops
M, N, K = 4096 * 96, 512, 512
X = (t.rand(M, K, device="cuda") * 2 - 1).bfloat16()
W = ((t.rand(N, K, device="cuda") * 2 - 1) * 0.003).bfloat16()
b = t.zeros(N, device="cuda")
Z = t.empty(M, N, device="cuda", dtype=t.bfloat16)
Z2 = t.randn(M, N, device="cuda").bfloat16()
Y2 = t.empty_like(Z2)
a, c = t.ones(N, device="cuda"), t.zeros(N, device="cuda")
s1, s2 = t.cuda.Stream(), t.cuda.Stream()
reps = 10

def run(two):
    t.cuda.synchronize()
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(t.cuda.current_stream()); s2.wait_stream(t.cuda.current_stream())
    for _ in range(reps):
        with t.cuda.stream(s1):
            ops.gemm(X, W, Z, bias=b, alpha=30.0)
        with t.cuda.stream(s2 if two else s1):
            ops.sine_fwd(Z2, a, c, Y2)
    t.cuda.current_stream().wait_stream(s1); t.cuda.current_stream().wait_stream(s2)
    e1.record()
    t.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for _ in range(2):
    print("one stream %.1f us/pair   two streams %.1f us/pair" % (run(False) * 1e3, run(True) * 1e3))
