"""Host-side profile of the end-to-end training step (pinned host batch in, loss read back): where the ~0.6 ms between the
device-timed step and the e2e step go.  usage (gpurun): python scripts/e2e_profile.py"""
import cProfile, pstats, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch as t
from scripts_common import make_step
ts, batch, solar, jit = make_step(use_graph=True)
host = {k: v.cpu().pin_memory() for k, v in batch.items()}
for i in range(6):
    ts.step(host, i); float(ts.last_loss)
t.cuda.synchronize()
t0 = time.perf_counter()
for i in range(20):
    ts.step(host, 10 + i); float(ts.last_loss)
t.cuda.synchronize()
print("e2e ms/step %.3f (use_graph=%s)" % ((time.perf_counter() - t0) / 20 * 1e3, ts.use_graph))
pr = cProfile.Profile()
pr.enable()
for i in range(20):
    ts.step(host, 40 + i); float(ts.last_loss)
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print(s.getvalue()[:6000])
