"""Trunk-shape forward GEMM (M = 4096*96, 512 x 512, bf16): cuBLAS (torch.matmul) beside the CTA-pair kernel without and
with the fused epilogues - separates main-loop / store-path limits from epilogue arithmetic."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as t
from season_nerf_b200 import ops

M, N, K = 4096 * 96, 512, 512
g = t.Generator(device="cuda").manual_seed(0)
X = (t.rand(M, K, device="cuda", generator=g) * 2 - 1).bfloat16()
W = ((t.rand(N, K, device="cuda", generator=g) * 2 - 1) * 0.1 / 30).bfloat16()
b = t.zeros(N, device="cuda")
Z = t.empty(M, N, device="cuda", dtype=t.bfloat16)
Y = t.empty(M, N, device="cuda", dtype=t.bfloat16)
Wt = W.t().contiguous()


def timed(fn, reps=6):
    fn()
    t.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        s, e = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        t.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best * 1e3


fl = 2.0 * M * N * K
for name, fn in (("cublas matmul (X @ W^T -> bf16)", lambda: t.matmul(X, W.t(), out=Z)),
                 ("gemm2 plain store", lambda: ops.gemm(X, W, Z, bias=b, alpha=30.0)),
                 ("gemm2 + BN statistics", lambda: ops.gemm_stats(X, W, Z, bias=b, alpha=30.0)),
                 ("gemm2 + sin (Z and Y out)", lambda: ops.gemm_sine_fwd(X, W, Z, Y, bias=b, alpha=30.0))):
    us = timed(fn)
    print("%-34s %7.1f us  %6.0f TFLOP/s" % (name, us, fl / us / 1e6))
