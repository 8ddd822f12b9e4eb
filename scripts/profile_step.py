"""Per-kernel time table of one training step (torch.profiler / CUPTI; kernels run back to back, warm caches).
usage (under gpurun): python scripts/profile_step.py [rays] > gpurun_out/step_table.txt"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch as t
import bench
import season_nerf_b200 as snb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = t.device("cuda:0")
args = bench.bench_args()
W2C = np.array([41.2905, -95.8967, 315.0])
H = np.eye(4)
H[0, 0], H[1, 1], H[2, 2] = 2 / 0.0024, 2 / 0.0032, 2 / 70.0
H[0, 3], H[1, 3], H[2, 3] = -W2C[0] * H[0, 0], -W2C[1] * H[1, 1], -W2C[2] * H[2, 2]
t.manual_seed(0)
ts = snb.TrainStep(args, dev, H, W2C, world_size=1, precision="bf16", use_graph=os.environ.get("SNB_PROFILE_GRAPH", "0") == "1")
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
for i in range(3):
    ts.step(batch, i, jitter=jit, solar=solar, solar_jitter=jit)
t.cuda.synchronize()
t0 = time.perf_counter()
for i in range(3):
    ts.step(batch, 3 + i, jitter=jit, solar=solar, solar_jitter=jit)
t.cuda.synchronize()
print("wall ms/step (no profiler): %.2f" % ((time.perf_counter() - t0) / 3 * 1e3))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    ts.step(batch, 10, jitter=jit, solar=solar, solar_jitter=jit)
    t.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == t.autograd.DeviceType.CUDA]
agg = {}
for e in ev:
    k = e.name[:110]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(v[1] for v in agg.values())
span = max(e.time_range.end for e in ev) - min(e.time_range.start for e in ev)
print("first kernel start -> last kernel end: %.3f ms (kernels on parallel streams overlap; gaps = launch / dependency latency)" % (span / 1e3))
print("total kernel time %.3f ms in %d launches" % (tot / 1e3, sum(v[0] for v in agg.values())))
for k, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%9.3f ms %5d %5.1f%%  %s" % (d / 1e3, c, 100 * d / tot, k))

# per-launch durations of the dense kernels, in launch order (forward image pass, solar pass, backward)
print("\nper-launch durations (us) of the tcgen05 GEMM / activation kernels, in launch order:")
evs = sorted((e for e in ev if "snb::" in e.name), key=lambda e: e.time_range.start)
line = []
for e in evs:
    nm = e.name.split("snb::")[1].split("(")[0]
    nm = nm.replace("gemm2_bf16_kernel", "g2").replace("_vec_kernel", "").replace("_kernel", "").replace("__nv_bfloat16", "bf16")
    d = e.device_time if hasattr(e, "device_time") else e.cuda_time
    line.append("%s:%.0f" % (nm, d))
for i in range(0, len(line), 8):
    print("  " + "  ".join(line[i:i + 8]))
