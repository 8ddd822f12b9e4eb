"""In-kernel timeline of the resident-A GEMM (gemm_tc3.cu, CTA 0): where the MMA issuer, the transform warps and the
epilogue wait.  usage (gpurun): python scripts/tc3_timeline.py [store_y]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as t
from season_nerf_b200 import ops

store = len(sys.argv) > 1 and sys.argv[1] == "1"
M, N, K = 393216, 512, 512
g = t.Generator(device="cuda").manual_seed(0)
Zp = (t.randn(M, K, device="cuda", generator=g) * 3).bfloat16()
W = ((t.rand(N, K, device="cuda", generator=g) * 2 - 1) * (6 / K) ** 0.5 / 30).bfloat16()
b = t.randn(N, device="cuda", generator=g) * 0.01
a, c = t.rand(K, device="cuda", generator=g) + 0.5, t.randn(K, device="cuda", generator=g)
Z = t.empty(M, N, device="cuda", dtype=t.bfloat16)
Y = t.empty_like(Zp) if store else None
for _ in range(2):
    ops.gemm_stats_xf(Zp, a, c, W, Z, bias=b, alpha=30.0, Y=Y)
dbg = t.zeros(1024, device="cuda", dtype=t.int64)
os.environ["SNB_TC3_TIMELINE"] = str(dbg.data_ptr())
ops.gemm_stats_xf(Zp, a, c, W, Z, bias=b, alpha=30.0, Y=Y)
t.cuda.synchronize()
del os.environ["SNB_TC3_TIMELINE"]
d = dbg.cpu().tolist()
t0 = min(x for x in d if x > 0)
rel = lambda i: (d[i] - t0) if d[i] else -1
print("store_y", store, " all times in SM clocks since the first stamp of CTA 0")
for it in range(6):
    print("tile %d" % it)
    print("  A slot free / load issued :", [rel(700 + it * 8 + kb) for kb in range(8)])
    print("  transform: A landed       :", [rel(400 + (it * 8 + kb) * 2) for kb in range(8)])
    print("  transform: slot done      :", [rel(400 + (it * 8 + kb) * 2 + 1) for kb in range(8)])
    for tn in range(2):
        base = (it * 2 + tn) * 8
        print("  pass %d acc free %d" % (tn, rel(300 + it * 2 + tn)))
        print("    mma: aready  :", [rel((base + kb) * 3) for kb in range(8)])
        print("    mma: bfull   :", [rel((base + kb) * 3 + 1) for kb in range(8)])
        print("    mma: issued  :", [rel((base + kb) * 3 + 2) for kb in range(8)])
        print("    epilogue warp 0: tfull %d  done %d" % (rel(600 + (it * 2 + tn) * 2), rel(600 + (it * 2 + tn) * 2 + 1)))
