"""Micro-benchmark of the bf16 tensor-core GEMM at the training shapes (run under gpurun)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as t
from season_nerf_b200 import ops

def bench(name, fn, flops, reps=10):
    for _ in range(3):
        fn()
    t.cuda.synchronize()
    s, e = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    t.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    print("%-34s %8.3f ms  %7.1f TFLOP/s" % (name, ms, flops / ms / 1e9), flush=True)

M = int(sys.argv[1]) if len(sys.argv) > 1 else 393216
for (N, K) in [(512, 512), (256, 512), (512, 576), (256, 256)]:
    A = (t.rand(M, K, device="cuda") - .5).bfloat16()
    W = (t.rand(N, K, device="cuda") - .5).bfloat16()
    b = t.rand(N, device="cuda")
    Z = t.empty(M, N, device="cuda", dtype=t.bfloat16)
    dX = t.empty(M, K, device="cuda", dtype=t.bfloat16)
    dW = t.zeros(N, K, device="cuda")
    fl = 2.0 * M * N * K
    bench("fwd   M=%d N=%d K=%d" % (M, N, K), lambda: ops.gemm(A, W, Z, bias=b, alpha=30.0), fl)
    bench("fwd+stats", lambda: ops.gemm_stats(A, W, Z, bias=b, alpha=30.0), fl)
    bench("dgrad (b_t)", lambda: ops.gemm(Z, W, dX, alpha=30.0, b_t=True), fl)
    bench("dgrad accumulate", lambda: ops.gemm(Z, W, dX, alpha=30.0, b_t=True, accumulate=1), fl)
    bench("wgrad (a_t,b_t,split-K)", lambda: ops.gemm(Z, A, dW, alpha=30.0, accumulate=2, a_t=True, b_t=True), fl)
    ref = t.empty(M, N, device="cuda", dtype=t.bfloat16)
    bench("torch.mm (cuBLAS) fwd", lambda: t.mm(A, W.T, out=ref), fl)
