"""Time the public render APIs on one GPU: ray-sharded image render and component render + composite (1024^2 and 512^2)."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as t
import season_nerf_b200 as snb
from bench_extras import oma_frame

S = 96
dev = t.device("cuda:0")
W2C, H = oma_frame()
t.manual_seed(0)
net = snb.T_NeRF(512, 4).to(dev).eval()
snb.render_image_sharded(net, [80, 0], [45, 135], 184 / 365, (64, 64, S), W2C, H, dev)
for size in (512, 1024, 1024):
    t.cuda.synchronize()
    t0 = time.perf_counter()
    img, mask = snb.render_image_sharded(net, [80, 0], [45, 135], 184 / 365, (size, size, S), W2C, H, dev)
    t.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("render_image_sharded %4d^2: %7.1f ms  %.3f M rays/s  peak mem %.2f GB" % (size, dt * 1e3, size * size / dt / 1e6,
                                                                                 t.cuda.max_memory_allocated() / 2 ** 30))
for size in (512, 512):
    t.cuda.synchronize()
    t0 = time.perf_counter()
    D = snb.component_render_by_dir(net, [80, 0], [45, 135], 184 / 365, (size, size, S), W2C, H, dev, include_exact_solar=False)
    t.cuda.synchronize()
    t1 = time.perf_counter()
    imgs = snb.get_imgs_from_Img_Dict(D, (size, size, S), False)
    t.cuda.synchronize()
    t2 = time.perf_counter()
    print("component_render_by_dir %d^2: %.1f ms, get_imgs_from_Img_Dict %.1f ms" % (size, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
    del D, imgs
