"""Which Python lines launch the small torch kernels of one training step (eager)?  Aggregates CUDA kernels launched by
aten ops by the innermost repo source line on the Python stack."""
import collections
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from scripts_common import make_step
import torch as t
from torch.profiler import profile, ProfilerActivity

ts, batch, solar, jit = make_step()
for i in range(3):
    ts.step(batch, i, jitter=jit, solar=solar, solar_jitter=jit)
t.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], with_stack=True) as prof:
    ts.step(batch, 10, jitter=jit, solar=solar, solar_jitter=jit)
    t.cuda.synchronize()
agg = collections.Counter()
tim = collections.Counter()
for e in prof.events():
    if e.device_type != t.autograd.DeviceType.CPU or not list(e.kernels):
        continue
    if any(list(c.kernels) for c in (e.cpu_children or [])):
        continue                      # count a kernel at the innermost op that owns it
    chain, p_ = [e.name], e.cpu_parent
    while p_ is not None and len(chain) < 5:
        chain.append(p_.name[:40])
        p_ = p_.cpu_parent
    key = " <- ".join(chain)
    agg[key] += len(list(e.kernels))
    tim[key] += sum(k.duration for k in e.kernels)
print("kernels by innermost op <- callers (count, total us):")
for k, c in agg.most_common(70):
    print("%5d %8.1f  %s" % (c, tim[k], k[:150]))
print("total", sum(agg.values()))
