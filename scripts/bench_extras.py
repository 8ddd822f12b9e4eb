"""Secondary measurements of the hot path on one B200 (imported by bench.py; also runnable alone under gpurun):
  composite   : alpha-compositing forward / backward kernels against the HBM roofline (SURVEY 8d bytes-per-ray model)
  shadow      : BASELINE.json configs[2] - 512x512 novel view with the exact solar-visibility march (sigma-only fused MLP)
  year_sweep  : BASELINE.json configs[4] - 365 time-of-year renders of a 1024x1024 view (one MLP render + fused recombination)
usage: python scripts/bench_extras.py [composite] [shadow[=SIZE]] [year[=SIZE[,T]]]"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

S = 96
RENDER_FLOP_PER_RAY = 556_750_336
MARCH_FLOP_PER_RAY = S * S * 4_061_696   # S^2 sigma-only evaluations


def oma_frame():
    import numpy as np
    W2C = np.array([41.2905, -95.8967, 315.0])
    H = np.eye(4)
    H[0, 0], H[1, 1], H[2, 2] = 2 / 0.0024, 2 / 0.0032, 2 / 70.0
    H[0, 3], H[1, 3], H[2, 3] = -W2C[0] * H[0, 0], -W2C[1] * H[1, 1], -W2C[2] * H[2, 2]
    return W2C, H


def _events(fn, reps):
    import torch as t
    fn()
    t.cuda.synchronize()
    ev = []
    for _ in range(reps):
        s, e = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        ev.append((s, e))
    t.cuda.synchronize()
    return sum(s.elapsed_time(e) for s, e in ev) / reps


def bench_composite(dev, hbm_peak, peak_source, N=1 << 20, reps=5):
    """one warp per ray; inputs (2.4 GB at 1 Mi rays) far exceed the 126 MB L2, so every rep streams from HBM."""
    import torch as t
    from season_nerf_b200 import _lib, ops
    g = t.Generator(device=dev).manual_seed(0)
    rho = t.rand(N, S, device=dev, generator=g) * 4
    deltas = t.full((N, S), 2.0 / S, device=dev)
    col = t.rand(N, S, 3, device=dev, generator=g)
    vis = t.rand(N, S, device=dev, generator=g)
    sky = t.rand(N, 3, device=dev, generator=g)
    d_rend = t.rand(N, 3, device=dev, generator=g)
    albedo, rendered, vsum = t.empty(N, 3, device=dev), t.empty(N, 3, device=dev), t.empty(N, device=dev)
    d_rho, d_col, d_sky = t.empty_like(rho), t.empty_like(col), t.empty_like(sky)
    lib = _lib.load()
    p = lambda x: None if x is None else C.c_void_p(x.data_ptr())
    st = lambda: C.c_void_p(t.cuda.current_stream().cuda_stream)

    def fwd():
        _lib.check(lib.snb_composite_fwd(p(rho), p(deltas), p(col), p(vis), p(sky), 0, N, S, 0, None, None, None, p(albedo),
                                         p(rendered), p(vsum), st()))

    def bwd():
        _lib.check(lib.snb_composite_bwd(p(rho), p(deltas), p(col), p(vis), p(sky), 0, N, S, 0, p(d_rend), None, None, None,
                                         None, p(d_rho), p(d_col), p(d_sky), None, st()))

    ms_f, ms_b = _events(fwd, reps), _events(bwd, reps)
    # bytes the kernel interface moves per ray (fp32): rho, deltas, vis [S], col [S,3], sky [3] in; albedo, rendered [3], vis_sum out
    b_fwd = S * 4 * 6 + 12 + 12 + 12 + 4
    # backward: the same inputs again + d_rendered [3]; d_rho [S], d_col [S,3], d_sky [3] out
    b_bwd = S * 4 * 6 + 12 + 12 + S * 4 * 4 + 12
    def _traffic(key):
        # dram bytes of one launch at 1 Mi rays from the ncu --set full capture (profiles/r02_ncu_traffic.json); null at other sizes
        try:
            import json
            d = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_ncu_traffic.json")))
            return d[key]["dram_bytes"] if N == 1 << 20 else None
        except Exception:
            return None

    mk = lambda ms, b, name: {"bound": "hbm", "kernel": name, "achieved": N * b / (ms * 1e-3) / 1e9, "peak": hbm_peak,
                              "unit": "GB/s", "frac": N * b / (ms * 1e-3) / 1e9 / hbm_peak,
                              "traffic": _traffic("composite_fwd" if "fwd" in name else "composite_bwd"),
                              "bytes_per_ray": b, "rays": N, "ms_per_launch": ms, "rays_per_s": N / (ms * 1e-3),
                              "peak_source": peak_source}
    return {"fwd": mk(ms_f, b_fwd, "composite_fwd_kernel<false,false>"), "bwd": mk(ms_b, b_bwd, "composite_bwd_kernel<false,false,3>")}


def bench_shadow(snb, net, dev, tensor_peak, peak_source, size=512):
    """configs[2]: size x size view, fixed VA/SA, time 07/04, exact shadow march; through the public render API."""
    import torch as t
    W2C, H = oma_frame()
    net.eval()
    t.cuda.synchronize()
    t0 = time.perf_counter()
    D = snb.component_render_by_dir(net, [80, 0], [45, 135], 184 / 365, (size, size, S), W2C, H, dev, include_exact_solar=True)
    imgs = snb.get_imgs_from_Img_Dict(D, (size, size, S), False)
    out = imgs["Season_Adj_Img"] * imgs["Shadow_Adjust_Exact"]
    t.cuda.synchronize()
    dt = time.perf_counter() - t0
    N = size * size
    flop = N * (RENDER_FLOP_PER_RAY + MARCH_FLOP_PER_RAY)
    return {"workload": "%dx%dx%d novel view, VA (80,0), SA (45,135), time 07/04, exact shadow march (BASELINE.json configs[2])" % (size, size, S),
            "rays": N, "sigma_evals": N * S * S, "seconds": dt, "rays_per_s": N / dt, "finite": bool((out == out).all()),
            "roofline": {"bound": "tensor", "kernel": "fused_eval2_kernel (sigma-only program)", "achieved": flop / dt / 1e12,
                         "peak": tensor_peak, "unit": "TFLOP/s", "frac": flop / dt / 1e12 / tensor_peak,
                         "peak_source": peak_source,
                         "note": "algorithmic FLOPs (0.5568 + 37.43 GFLOP/ray) over the WHOLE API call (host geometry, sampling, MLP, "
                                 "march reduction, float64 composite, image D2H)"}}


def bench_year(snb, net, dev, hbm_peak, peak_source, size=1024, T=365):
    """configs[4]: T time-of-year renders of a size x size view: one component render at time 0, the T class vectors, then
    the fused recombination (mg_merge_seasons.py:145-178 factorisation, mg_Img_Eval.py:192-228)."""
    import numpy as np
    import torch as t
    from season_nerf_b200 import ops
    W2C, H = oma_frame()
    net.eval()
    N = size * size
    t.cuda.synchronize()
    t0 = time.perf_counter()
    D = snb.component_render_by_dir(net, [80, 0], [45, 135], 0.0, (size, size, S), W2C, H, dev, include_exact_solar=False)
    times = np.stack([snb.encode_time(k / T) for k in range(T)], 0)
    with t.no_grad():
        cls = net.get_class_only(t.tensor(times, dtype=t.float32, device=dev)).double().cpu().numpy()
    t.cuda.synchronize()
    t1 = time.perf_counter()
    # device-side recombination alone (CUDA events)
    rho, dl, base, adj = [D.dev[k] for k in ("Rho", "Deltas", "Base_Col", "Adjust_col")]
    clsd = t.tensor(cls, dtype=t.float64, device=dev)
    ms_sweep = _events(lambda: ops.year_sweep(rho.reshape(N, S), dl.reshape(N, S), base, adj, clsd), 2)
    t.cuda.synchronize()
    t2 = time.perf_counter()
    imgs = snb.get_imgs_from_Img_Dict_t_step(D, (size, size, S), cls)
    t.cuda.synchronize()
    t3 = time.perf_counter()
    b = S * 4 * (1 + 1 + 3 + 12) + T * 24          # f32 components in, T float64 RGB out per ray
    render_s, api_s = t1 - t0, t3 - t2
    return {"workload": "%d time-of-year renders of a %dx%d view (BASELINE.json configs[4])" % (T, size, size),
            "rays": N, "times": T, "render_s": render_s, "sweep_kernel_ms": ms_sweep, "recombine_api_s": api_s,
            "ray_renders_per_s_device": N * T / (render_s + ms_sweep * 1e-3),
            "ray_renders_per_s_e2e": N * T / (render_s + api_s), "out_shape": list(imgs.shape),
            "roofline": {"bound": "hbm", "kernel": "year_sweep_kernel", "achieved": N * b / (ms_sweep * 1e-3) / 1e9,
                         "peak": hbm_peak, "unit": "GB/s", "frac": N * b / (ms_sweep * 1e-3) / 1e9 / hbm_peak,
                         "bytes_per_ray": b, "peak_source": peak_source}}


if __name__ == "__main__":
    import torch as t
    import season_nerf_b200 as snb
    import bench
    pk = bench._peaks()
    dev = t.device("cuda:0")
    t.manual_seed(0)
    net = snb.T_NeRF(512, 4).to(dev).eval()
    out = {}
    for a in sys.argv[1:] or ["composite", "shadow", "year"]:
        k, _, v = a.partition("=")
        if k == "composite":
            out[k] = bench_composite(dev, pk["hbm_gbs"], pk["source"])
        elif k == "shadow":
            out[k] = bench_shadow(snb, net, dev, pk["bf16_tflops"], pk["source"], int(v) if v else 512)
        elif k == "year":
            a1 = [int(x) for x in v.split(",")] if v else []
            out[k] = bench_year(snb, net, dev, pk["hbm_gbs"], pk["source"], *(a1 or [1024, 365]))
        print(json.dumps({k: out[k]}), flush=True)
