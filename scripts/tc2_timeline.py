"""In-kernel timeline of the CTA-pair GEMM (gemm_tc2.cu, CTA 0) on the trunk shape: where the producer, the MMA issuer and
the epilogue wait.  usage (gpurun): python scripts/tc2_timeline.py fwd_stats|fwd_sin|dgrad|wgrad"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as t
from season_nerf_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "fwd_sin"
M, N, K = 393216, 512, 512
g = t.Generator(device="cuda").manual_seed(0)
X = (t.rand(M, K, device="cuda", generator=g) * 2 - 1).bfloat16()
W = ((t.rand(N, K, device="cuda", generator=g) * 2 - 1) * 0.1 / 30).bfloat16()
b = t.zeros(N, device="cuda")
Z, Y, G = (t.empty(M, N, device="cuda", dtype=t.bfloat16) for _ in range(3))
dW = t.zeros(N, K, device="cuda", dtype=t.float32)
ones, zeros = t.ones(N, device="cuda"), t.zeros(N, device="cuda")
fn = {"fwd_stats": lambda: ops.gemm_stats(X, W, Z, bias=b, alpha=30.0),
      "fwd_sin": lambda: ops.gemm_sine_fwd(X, W, Z, Y, bias=b, alpha=30.0),
      "dgrad": lambda: ops.gemm_sine_bwd(Y, W, G, Z, ones, zeros, zeros, ones, alpha=30.0),
      "wgrad": lambda: ops.gemm(G, X, dW, alpha=30.0, accumulate=2, a_t=True, b_t=True)}[which]
for _ in range(2):
    fn()
dbg = t.zeros(1024, device="cuda", dtype=t.int64)
os.environ["SNB_TC2_TIMELINE"] = str(dbg.data_ptr())
fn()
t.cuda.synchronize()
del os.environ["SNB_TC2_TIMELINE"]
d = dbg.cpu().tolist()
t0 = min(x for x in d if x > 0)
rel = lambda i: (d[i] - t0) if d[i] else -1
print(which, " all times in SM clocks since the first stamp of CTA 0")
for it in range(12):
    print("item %d: acc free %d" % (it, rel(200 + it * 24 + 23)))
    print("  stage free / load issued:", [rel(it * 8 + kb) for kb in range(8)])
    print("  mma: operands landed    :", [rel(200 + it * 24 + kb * 3) for kb in range(7)])
    print("  mma: issued             :", [rel(200 + it * 24 + kb * 3 + 1) for kb in range(7)])
    print("  epilogue warp 0: accumulator complete %d  done %d" % (rel(600 + it * 2), rel(600 + it * 2 + 1)))
    print("  epilogue done per warp, CTA 0:", [rel(700 + (it * 2) * 8 + w) for w in range(8)], " CTA 1 (its own clock):",
          [d[700 + (it * 2 + 1) * 8 + w] - t0 if d[700 + (it * 2 + 1) * 8 + w] else -1 for w in range(8)])
