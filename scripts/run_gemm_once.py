"""Launch each training-path GEMM variant once on the trunk shape (M = 4096*96 rows, 512 x 512) for ncu captures:
forward with fused BatchNorm statistics, forward with fused sin, fused input-gradient (cos + BN-backward sums), weight
gradient (split-K)."""
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
G = t.empty(M, N, device="cuda", dtype=t.bfloat16)
dW = t.zeros(N, K, device="cuda", dtype=t.float32)
ones, zeros = t.ones(N, device="cuda"), t.zeros(N, device="cuda")
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    ev = [t.cuda.Event(enable_timing=True) for _ in range(7)]
    ev[0].record()
    ops.gemm_stats(X, W, Z, bias=b, alpha=30.0)
    ev[1].record()
    ops.gemm_sine_fwd(X, W, Z, Y, bias=b, alpha=30.0)
    ev[2].record()
    ops.gemm_sine_bwd(Y, W, G, Z, ones, zeros, zeros, ones, alpha=30.0)
    ev[3].record()
    ops.gemm(G, X, dW, alpha=30.0, accumulate=2, a_t=True, b_t=True)
    ev[4].record()
    ops.gemm_stats_xf(X, ones, zeros, W, Z, bias=b, alpha=30.0)                 # consumer-side activation, resident A (gemm_tc3.cu)
    ev[5].record()
    ops.gemm_stats_xf(X, ones, zeros, W, Z, bias=b, alpha=30.0, Y=Y)            # + activated operand written back
    ev[6].record()
    t.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(6)]
    fl = 2.0 * M * N * K / 1e9
    print("fwd+stats %.3f ms (%.0f TF)  fwd+sin %.3f ms (%.0f TF)  dgrad+cos+sums %.3f ms (%.0f TF)  wgrad %.3f ms (%.0f TF)  "
          "xf fwd+stats %.3f ms (%.0f TF)  xf+storeY %.3f ms (%.0f TF)"
          % (ms[0], fl / ms[0], ms[1], fl / ms[1], ms[2], fl / ms[2], ms[3], fl / ms[3], ms[4], fl / ms[4], ms[5], fl / ms[5]))
