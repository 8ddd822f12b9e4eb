"""torchrun --nproc-per-node N scripts/shard_probe_n.py : phases of the ray-sharded 1024^2 render on every rank
(SNB_SHARD_TIMING=1 prints them) and the max-over-ranks CUDA-event time, twice."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch as t
import torch.distributed as dist
import season_nerf_b200 as snb
from bench_extras import oma_frame

os.environ["SNB_SHARD_TIMING"] = "1"
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = t.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
t.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
S = 96
W2C, H = oma_frame()
t.manual_seed(0)
net = snb.T_NeRF(512, 4).to(dev).eval()
snb.render_image_sharded(net, [80, 0], [45, 135], 184 / 365, (64, 64, S), W2C, H, dev, rank, world)
for rep in range(3):
    dist.barrier()
    t.cuda.synchronize()
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    e0.record()
    img, mask = snb.render_image_sharded(net, [80, 0], [45, 135], 184 / 365, (1024, 1024, S), W2C, H, dev, rank, world)
    e1.record()
    t.cuda.synchronize()
    ms = t.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        sys.stderr.write("== rep %d: %.1f ms max over %d ranks (%.2f M rays/s)\n" % (rep, float(ms), world, 1048576 / float(ms) / 1e3))
dist.barrier()
dist.destroy_process_group()
