"""CLI render engine: drop-in for T_NeRF_Eval_Utils/mg_Img_Eval.py:17-228 (_internal_render,
component_render_by_P, component_render_by_dir, get_imgs_from_Img_Dict, get_imgs_from_Img_Dict_t_step).

Same arguments and the same dict of per-sample component arrays (float64 numpy on access, as downstream numpy code
expects).  The components stay resident in HBM as float32 (`DeviceImgDict`), so compositing and the year sweep run
on the device without the six D2H copies per chunk of the reference; float64 host arrays are materialised lazily,
only for the keys a caller actually touches.
"""
import numpy as np
import torch as t

from . import ops
from .engine import sample_ts
from .geometry import encode_time, world_angle_2_local_vec

_KEYS = ["World_Points", "Deltas", "Rho", "Base_Col", "Est_Solar_Vis", "Sky_Col", "Output_class", "Adjust_col"]


class DeviceImgDict(dict):
    """dict of float64 numpy arrays backed by float32 device tensors (converted on first access)."""

    def __init__(self, dev_tensors, host_items=None):
        super().__init__()
        self.dev = dict(dev_tensors)
        for k in self.dev:
            dict.__setitem__(self, k, None)
        for k, v in (host_items or {}).items():
            dict.__setitem__(self, k, v)

    def __getitem__(self, k):
        v = dict.__getitem__(self, k)
        if v is None and k in self.dev:
            v = self.dev[k].detach().cpu().numpy().astype(np.float64)
            dict.__setitem__(self, k, v)
        return v

    def get(self, k, default=None):
        return self[k] if k in self else default

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def values(self):
        return [self[k] for k in self.keys()]

    def __setitem__(self, k, v):
        self.dev.pop(k, None)      # a caller-modified array is the truth from now on
        dict.__setitem__(self, k, v)


_STAGE = {}


def device_to_numpy(x, chunk_bytes=128 << 20, threads=4, min_bytes=64 << 20):
    """Large device tensor -> numpy array of the same dtype / shape.  `x.cpu()` of a multi-GB result (the [T,H,W,3] float64
    year sweep is 9.2 GB at 365 x 1024^2) runs at ~2 GB/s: it copies through an internal staging buffer into freshly
    mapped pageable memory, page fault by page fault, on one thread.  Here the device->host copy goes through two cached
    PINNED staging buffers (asynchronous, PCIe rate) and a few host threads move each landed chunk into the destination
    (first touch in parallel) while the next chunk is in flight."""
    x = x.detach().contiguous()
    nbytes = x.numel() * x.element_size()
    if nbytes < min_bytes or not x.is_cuda:          # small results: the plain copy is as fast
        return x.cpu().numpy()
    from concurrent.futures import ThreadPoolExecutor
    key = (x.device.index, chunk_bytes)
    if key not in _STAGE:
        _STAGE[key] = ([t.empty(chunk_bytes, dtype=t.uint8).pin_memory() for _ in range(2)], ThreadPoolExecutor(max_workers=threads))
    stage, pool = _STAGE[key]
    out = np.empty(tuple(x.shape), dtype=t.empty(0, dtype=x.dtype).numpy().dtype)
    dst = out.reshape(-1).view(np.uint8)
    src = x.reshape(-1).view(t.uint8)
    stage_np = [b.numpy() for b in stage]
    stream = t.cuda.current_stream(x.device)
    pending = [[], []]                 # host copies still reading staging buffer 0 / 1
    events = [None, None]
    spans = [(a, min(a + chunk_bytes, nbytes)) for a in range(0, nbytes, chunk_bytes)]

    def drain(j, a, b):
        events[j].synchronize()
        n = b - a
        step = (n + threads - 1) // threads
        step = (step + 4095) // 4096 * 4096
        pending[j] = [pool.submit(np.copyto, dst[a + o:a + min(o + step, n)], stage_np[j][o:min(o + step, n)])
                      for o in range(0, n, step)]

    for i, (a, b) in enumerate(spans):
        j = i & 1
        for f in pending[j]:
            f.result()                 # the buffer's previous contents have left
        stage[j][:b - a].copy_(src[a:b], non_blocking=True)
        events[j] = t.cuda.Event()
        events[j].record(stream)
        if i > 0:
            drain(j ^ 1, *spans[i - 1])
    drain((len(spans) - 1) & 1, *spans[-1])
    for j in (0, 1):
        for f in pending[j]:
            f.result()
    return out


def _points_per_call(the_network):
    return 1 << 22 if getattr(the_network, "_fused_ready", lambda: False)() else 1 << 19


def _internal_render(the_network, tops, bots, a_sun_el_az_vec, a_year_frac, out_img_size, max_batch_size,
                     include_exact_solar, device):
    """mg_Img_Eval.py:17-72.  `max_batch_size` is accepted for compatibility; chunking is by device memory."""
    device = t.device(device)
    if device.type != "cuda":
        raise ops._lib.SeasonNerfCudaError("season_nerf_b200 renders on CUDA only (no CPU fallback)")
    S = out_img_size[2]
    C = the_network.n_classes
    N = tops.shape[0]
    tops = tops.to(device=device, dtype=t.float32)
    bots = bots.to(device=device, dtype=t.float32)
    sun_np = np.asarray(a_sun_el_az_vec, dtype=np.float64)
    sun = t.tensor(sun_np, dtype=t.float32, device=device).reshape(1, 3)           # .float() of the f64 vector (:43)
    tim = t.tensor(encode_time(a_year_frac), dtype=t.float32, device=device).reshape(1, 4)
    ts = sample_ts(S, eval_mode=True, include_end_pt=True).to(device)
    mk = lambda *s: t.empty(*s, device=device, dtype=t.float32)
    out = {"World_Points": mk(N, S, 3), "Deltas": mk(N, S, 1), "Rho": mk(N, S, 1), "Base_Col": mk(N, S, 3),
           "Est_Solar_Vis": mk(N, S, 1), "Sky_Col": mk(N, S, 3), "Output_class": mk(N, S, C), "Adjust_col": mk(N, S, C, 3)}
    if include_exact_solar:
        out["Exact_Solar"] = mk(N, S, 1)
    step = max(1, _points_per_call(the_network) // S)
    step_exact = max(1, _points_per_call(the_network) // (S * S))
    with t.no_grad():
        for i in range(0, N, step):
            e = min(i + step, N)
            n = e - i
            pts, deltas = ops.sample_rays(tops[i:e], bots[i:e], ts, zero_oob=True)           # :40-42
            pos, vis, adj, sky, cl = the_network.forward_rays(pts.reshape(-1, 3), sun, tim, S)
            out["World_Points"][i:e] = pts
            out["Deltas"][i:e] = deltas.unsqueeze(-1)
            out["Rho"][i:e] = the_network.Softplus(pos[:, 0:1]).reshape(n, S, 1)
            out["Base_Col"][i:e] = pos[:, 1:4].reshape(n, S, 3)
            out["Est_Solar_Vis"][i:e] = the_network.Sigmoid(vis).reshape(n, S, 1)
            out["Sky_Col"][i:e] = the_network.Sigmoid(sky).reshape(n, 1, 3)
            out["Output_class"][i:e] = the_network.SoftMax(cl).reshape(n, 1, C)
            out["Adjust_col"][i:e] = adj.reshape(n, S, C, 3)
        if include_exact_solar:                                                               # :57-70
            for i in range(0, N, step_exact):
                e = min(i + step_exact, N)
                nb = out["World_Points"][i:e].reshape(-1, 3)
                nt = ops.solar_tops(nb, sun_np, f64=True)
                _, nd = ops.sample_rays(nt, nb, ts, zero_oob=True, want_pts=False)
                rho = _sigma_on_rays(the_network, nt, nb, ts, S)
                out["Exact_Solar"][i:e] = ops.march_transmittance(rho, nd).reshape(e - i, S, 1)
    return DeviceImgDict(out)


def _sigma_on_rays(the_network, tops, bots, ts, S):
    """softplus(sigma) at the S samples of each (top, bot) ray -> [n, S] (T_NeRF_net_v2.py:169-170)."""
    pts, _ = ops.sample_rays(tops, bots, ts)
    rho_raw = the_network.forward_rays(pts.reshape(-1, 3), None, None, S, mode="sigma")[0]
    return the_network.Softplus(rho_raw).reshape(-1, S).contiguous()


def component_render_by_dir(the_network, view_el_az, sun_el_az, time_frac, out_img_size: tuple, W2C, W2L_H, device,
                            max_batch_size=150000, include_exact_solar=True):
    """mg_Img_Eval.py:96-115."""
    H, W = out_img_size[0], out_img_size[1]
    with t.no_grad():
        XYZ = np.stack(np.meshgrid(np.linspace(1, -1, H), np.linspace(-1, 1, W), indexing="ij"), -1).reshape([-1, 2])
        XYZ = np.concatenate([XYZ, np.zeros([XYZ.shape[0], 1])], 1)
        view_vec = world_angle_2_local_vec(view_el_az[0], view_el_az[1], W2C, W2L_H)
        sun_vec = world_angle_2_local_vec(sun_el_az[0], sun_el_az[1], W2C, W2L_H)
        tops = t.tensor(XYZ + np.expand_dims(view_vec / view_vec[2], 0)).float()
        bots = t.tensor(XYZ - np.expand_dims(view_vec / view_vec[2], 0)).float()
        R = _internal_render(the_network, tops, bots, sun_vec, time_frac, out_img_size, max_batch_size,
                             include_exact_solar, device)
        R["Image_Points"] = np.stack(np.meshgrid(np.arange(H), np.arange(W), indexing="ij"), -1).reshape([-1, 2])
    return R


def component_render_by_P(the_network, a_P_img, out_img_size: tuple, device, max_batch_size=150000,
                          include_exact_solar=True):
    """mg_Img_Eval.py:74-94 (a_P_img: the reference's P_img object: .img, .invert_P, .sun_el_and_az_vec, .get_year_frac)."""
    with t.no_grad():
        XY = np.stack(np.meshgrid(np.linspace(0, a_P_img.img.shape[0] - 1, out_img_size[0]),
                                  np.linspace(0, a_P_img.img.shape[1] - 1, out_img_size[1]), indexing="ij"), -1)
        XY = np.round(XY).astype(int).reshape([-1, 2])
        P = getattr(a_P_img, "P", None)
        if type(a_P_img).__name__ in ("P_img_Pinhole", "P_img_Parallel") and P is not None and np.asarray(P).shape == (3, 4):
            # closed-form inversion of the affine-approximated RPC camera on the device (bit-exact with invert_P,
            # pre_NeRF/P_Img.py:133-147): no per-pixel host arrays, no H2D of the ray endpoints
            tops_d, bots_d, good_d, _ = ops.camera_rays(P, t.device(device), rows=XY[:, 0], cols=XY[:, 1], bounds=(-1, 1, -1, 1))
            keep = good_d.nonzero().squeeze(1)
            tops_t, bots_t = tops_d[keep], bots_d[keep]
            good = good_d.cpu().numpy()
        else:
            x, y, z = a_P_img.invert_P(XY[:, 0], XY[:, 1], 1.)
            tops = np.stack([x, y, np.ones_like(x)], -1)
            x, y, z = a_P_img.invert_P(XY[:, 0], XY[:, 1], -1.)
            bots = np.stack([x, y, -np.ones_like(x)], -1)
            good = (tops[:, 0] >= -1) * (tops[:, 1] <= 1) * (bots[:, 0] >= -1) * (bots[:, 1] <= 1) * \
                   (tops[:, 1] >= -1) * (tops[:, 0] <= 1) * (bots[:, 1] >= -1) * (bots[:, 0] <= 1)
            tops_t, bots_t = t.tensor(tops[good]).float(), t.tensor(bots[good]).float()
        R = _internal_render(the_network, tops_t, bots_t,
                             a_P_img.sun_el_and_az_vec, a_P_img.get_year_frac(), out_img_size, max_batch_size,
                             include_exact_solar, device)
        R["Image_Points_in_GT_Img"] = XY[good]
        R["Image_Points"] = np.stack(np.meshgrid(np.arange(out_img_size[0]), np.arange(out_img_size[1]),
                                                 indexing="ij"), -1).reshape([-1, 2])[good]
    return R


def sig(X):
    return 1 / (1 + np.exp(-X))


def inv_sig(X):
    return -np.log(1 / X - 1)


def _device_components(D, keys):
    """float32 device tensors if the dict is ours and untouched, else float64 uploads of the caller's arrays."""
    if isinstance(D, DeviceImgDict) and all(k in D.dev for k in keys):
        return [D.dev[k] for k in keys]
    if not t.cuda.is_available():
        raise ops._lib.SeasonNerfCudaError("season_nerf_b200 composites on CUDA only (no CPU fallback)")
    return [t.as_tensor(np.ascontiguousarray(D[k]), dtype=t.float64).cuda() for k in keys]


def _scatter(D, size, vals):
    vals = vals.detach().cpu().numpy()
    img = np.zeros([size[0], size[1]] + list(vals.shape[1:])) * np.nan
    img[D["Image_Points"][:, 0], D["Image_Points"][:, 1]] = vals
    return img


def get_imgs_from_Img_Dict(Img_Dict, out_img_size: tuple, use_classic_shadows: bool):
    """mg_Img_Eval.py:123-190: float64 compositing of the cached components into images."""
    has_exact = "Exact_Solar" in Img_Dict.keys()
    keys = ["Rho", "Deltas", "Base_Col", "Est_Solar_Vis", "Adjust_col", "Output_class", "Sky_Col"]
    # ONE fetch for all components (Exact_Solar included): either every array is the resident float32 device tensor or -
    # as soon as the caller replaced any of them - every array is uploaded as float64; never a mix of element types
    comps = _device_components(Img_Dict, keys + (["Exact_Solar"] if has_exact else []))
    rho, dl, base, vis, adj, ocl, skyc = comps[:7]
    ev = comps[7] if has_exact else None
    N, S = rho.shape[0], rho.shape[1]
    Sky_Col = skyc[0, 0].double().cpu().numpy()                                   # ray 0 / sample 0 (:125-126)
    cls = ocl[0, 0].double()
    base_img, season, extreme, raw, raw_e = ops.cli_composite(
        rho.reshape(N, S), dl.reshape(N, S), base, vis.reshape(N, S), adj, cls.contiguous(),
        None if ev is None else ev.reshape(N, S))
    size = out_img_size
    Raw = _scatter(Img_Dict, size, raw)
    Shadow_Mask = sig((Raw - .2) * 30)
    Shadow_Adjust = np.expand_dims(Shadow_Mask, -1) + np.expand_dims(1 - Shadow_Mask, -1) * Sky_Col.reshape([1, 1, 3])
    R = {"Base_Img": _scatter(Img_Dict, size, base_img), "Season_Adj_Img": _scatter(Img_Dict, size, season),
         "Extreme_Imgs": [_scatter(Img_Dict, size, extreme[i]) for i in range(extreme.shape[0])],
         "Shadow_Adjust": Shadow_Adjust, "Shadow_Mask": Shadow_Mask, "Raw_Shadow_Mask": Raw, "Sky_Col": Sky_Col,
         "Time_Class": cls.cpu().numpy()}
    if has_exact:
        Raw_e = _scatter(Img_Dict, size, raw_e)
        Mask_e = sig((Raw_e - .2) * 30)
        R["Shadow_Adjust_Exact"] = np.expand_dims(Mask_e, -1) + np.expand_dims(1 - Mask_e, -1) * Sky_Col.reshape([1, 1, 3])
        R["Shadow_Mask_Exact"] = Mask_e
        R["Raw_Shadow_Mask_Exact"] = Raw_e
    if use_classic_shadows:                                                       # :166-181
        # per-sample shading vis + (1 - vis) * sky inside the colour sum; the image keeps its name but becomes the ratio
        # classic / seasonal colour at every rendered pixel
        (sky_s,) = _device_components(Img_Dict, ["Sky_Col"])
        ip = Img_Dict["Image_Points"]
        args = (rho.reshape(N, S), dl.reshape(N, S), base)
        classic = ops.cli_classic_shadow(*args, vis.reshape(N, S), adj, sky_s, cls.contiguous())
        R["Shadow_Adjust"][ip[:, 0], ip[:, 1]] = (classic / (season + 1e-8)).cpu().numpy()
        if has_exact:
            classic_e = ops.cli_classic_shadow(*args, ev.reshape(N, S), adj, sky_s, cls.contiguous())
            R["Shadow_Adjust_Exact"][ip[:, 0], ip[:, 1]] = (classic_e / (season + 1e-8)).cpu().numpy()
    return R


def _is_raster_grid(ip, H, W):
    """Image_Points of component_render_by_dir: every pixel once, in row-major order."""
    return ip.shape[0] == H * W and bool((ip[:, 0] * W + ip[:, 1] == np.arange(H * W)).all())


def get_imgs_from_Img_Dict_t_step(Img_Dict, out_img_size: tuple, class_vecs_array):
    """mg_Img_Eval.py:192-228: T seasonal recombinations in one fused pass (components read once).  The shadow factor
    (:214-226) rides in the kernel and the images leave the device once, already in [T,H,W,3] order."""
    has_exact = "Exact_Solar" in Img_Dict.keys()
    vkey = "Exact_Solar" if has_exact else "Est_Solar_Vis"
    rho, dl, base, vis, adj, skyc = _device_components(Img_Dict, ["Rho", "Deltas", "Base_Col", vkey, "Adjust_col", "Sky_Col"])
    N, S = rho.shape[0], rho.shape[1]
    H, W = out_img_size[0], out_img_size[1]
    sky0 = skyc[0, 0].double()
    cls = t.as_tensor(np.asarray(class_vecs_array), dtype=t.float64).to(rho.device).contiguous()
    T = cls.shape[0]
    _, _, _, raw, _ = ops.cli_composite(rho.reshape(N, S), dl.reshape(N, S), base, vis.reshape(N, S), adj, cls[0].contiguous())
    mask = t.sigmoid((raw - .2) * 30).unsqueeze(1)                                 # [N,1] float64
    shade = (mask + (1 - mask) * sky0.reshape(1, 3)).contiguous()                  # Shadow_Adjust per ray
    cols = ops.year_sweep(rho.reshape(N, S), dl.reshape(N, S), base, adj, cls, shade=shade)        # [T,N,3]
    ip = np.asarray(Img_Dict["Image_Points"])
    if _is_raster_grid(ip, H, W):
        return device_to_numpy(cols).reshape(T, H, W, 3)
    imgs = t.full((T, H * W, 3), float("nan"), device=cols.device, dtype=t.float64)
    imgs[:, t.as_tensor(ip[:, 0] * W + ip[:, 1], device=cols.device)] = cols
    return device_to_numpy(imgs).reshape(T, H, W, 3)


# ---- ray-sharded rendering over the GPUs of one box (the reference is single-device) ---------------------------
def view_rays(view_el_az, out_img_size, W2C, W2L_H):
    """tops / bots [H*W,3] float32 of component_render_by_dir's pixel grid (mg_Img_Eval.py:98-106)."""
    H, W = out_img_size[0], out_img_size[1]
    XYZ = np.stack(np.meshgrid(np.linspace(1, -1, H), np.linspace(-1, 1, W), indexing="ij"), -1).reshape([-1, 2])
    XYZ = np.concatenate([XYZ, np.zeros([XYZ.shape[0], 1])], 1)
    view_vec = world_angle_2_local_vec(view_el_az[0], view_el_az[1], W2C, W2L_H)
    tops = t.tensor(XYZ + np.expand_dims(view_vec / view_vec[2], 0)).float()
    bots = t.tensor(XYZ - np.expand_dims(view_vec / view_vec[2], 0)).float()
    return tops, bots


def render_shard(the_network, view_el_az, sun_el_az, time_frac, out_img_size, W2C, W2L_H, device, rank=0, world_size=1,
                 include_exact_solar=False, class_vecs=None):
    """This rank's contiguous range of the H*W rays of a novel view, rendered and composited on the device:
    -> (lo, hi, rgb [hi-lo,3] float64, shadow_mask [hi-lo] float64) with rgb = Season_Adj_Img * Shadow_Adjust(_Exact)
    exactly as main_run_Season_NeRF.py:90-92 forms the output image; with `class_vecs` [T,C]: rgb is the [T,hi-lo,3] year
    sweep of get_imgs_from_Img_Dict_t_step.  No communication: rays are independent."""
    from .train import shard_range
    H, W = out_img_size[0], out_img_size[1]
    lo, hi = shard_range(H * W, rank, world_size)
    sun_vec = world_angle_2_local_vec(sun_el_az[0], sun_el_az[1], W2C, W2L_H)
    with t.no_grad():
        # this rank's rows of the pixel grid of mg_Img_Eval.py:98-106, built on the device in float64 from the two 1-D
        # numpy linspaces (same additions as the reference's float64 meshgrid + offset, then .float(): bit-identical)
        view_vec = world_angle_2_local_vec(view_el_az[0], view_el_az[1], W2C, W2L_H)
        off = t.tensor(view_vec / view_vec[2], dtype=t.float64, device=device)
        lin_h = t.tensor(np.linspace(1, -1, H), dtype=t.float64, device=device)
        lin_w = t.tensor(np.linspace(-1, 1, W), dtype=t.float64, device=device)
        # the shard is rendered and composited in blocks of rays: the per-sample component arrays of a block (28 floats per
        # sample point, 11 GB for a whole 1024^2 view) are consumed by the compositing kernels right away and never kept
        block = max(1, _points_per_call(the_network) // out_img_size[2])
        if include_exact_solar:
            block = max(1, min(block, 16384))
        cls_all = None if class_vecs is None else t.as_tensor(np.asarray(class_vecs), dtype=t.float64).to(device).contiguous()
        rgb_parts, mask_parts = [], []
        S = out_img_size[2]
        sun_np = np.asarray(sun_vec, dtype=np.float64)
        sun = t.tensor(sun_np, dtype=t.float32, device=device).reshape(1, 3)
        tim = t.tensor(encode_time(time_frac), dtype=t.float32, device=device).reshape(1, 4)
        ts = sample_ts(S, eval_mode=True, include_end_pt=True).to(device)
        step_exact = max(1, _points_per_call(the_network) // (S * S))
        for b0 in range(lo, max(hi, lo + 1), block):
            b1 = min(b0 + block, hi)
            n = b1 - b0
            if n <= 0:
                rgb_parts.append(t.zeros((0, 3) if cls_all is None else (cls_all.shape[0], 0, 3), dtype=t.float64, device=device))
                mask_parts.append(t.zeros(0, dtype=t.float64, device=device))
                break
            idx = t.arange(b0, b1, device=device)
            xyz = t.stack([lin_h[idx // W], lin_w[idx % W], t.zeros(n, dtype=t.float64, device=device)], 1)
            tops, bots = (xyz + off).float(), (xyz - off).float()
            # the RAW heads of the block are composited right away (ops.render_composite_raw / year_sweep_raw): neither the
            # activated copies nor the per-sample component arrays of _internal_render (28 floats per sample) are formed
            pts, deltas = ops.sample_rays(tops, bots, ts, zero_oob=True)                        # mg_Img_Eval.py:40-42
            pos, vis, adj, sky, cl = the_network.forward_rays(pts.reshape(-1, 3), sun, tim, S)
            sky0 = the_network.Sigmoid(sky[0]).double()       # sun direction and time are those of the whole image: every
            cls0 = the_network.SoftMax(cl[0:1])[0].double().contiguous() if cls_all is None else cls_all[0].contiguous()   # ray = ray 0
            ev = None
            if include_exact_solar:                                                             # mg_Img_Eval.py:57-70
                ev = t.empty(n, S, device=device, dtype=t.float32)
                for i in range(0, n, step_exact):
                    e = min(i + step_exact, n)
                    nb = pts[i:e].reshape(-1, 3)
                    nt = ops.solar_tops(nb, sun_np, f64=True)
                    _, nd = ops.sample_rays(nt, nb, ts, zero_oob=True, want_pts=False)
                    ev[i:e] = ops.march_transmittance(_sigma_on_rays(the_network, nt, nb, ts, S), nd).reshape(e - i, S)
            season, raw, raw_e = ops.render_composite_raw(pos, vis.reshape(-1), adj, deltas, cls0, ev)
            mask = t.sigmoid(((raw_e if include_exact_solar else raw) - .2) * 30)
            shade = (mask.unsqueeze(1) + (1 - mask.unsqueeze(1)) * sky0.reshape(1, 3)).contiguous()
            mask_parts.append(mask)
            if cls_all is None:
                rgb_parts.append(season * shade)
            else:
                rgb_parts.append(ops.year_sweep_raw(pos, deltas, adj, cls_all, shade=shade))
            del pts, pos, vis, adj
        cat_dim = 0 if cls_all is None else 1
        rgb = rgb_parts[0] if len(rgb_parts) == 1 else t.cat(rgb_parts, cat_dim)
        mask = mask_parts[0] if len(mask_parts) == 1 else t.cat(mask_parts, 0)
        return lo, hi, rgb, mask


def render_image_sharded(the_network, view_el_az, sun_el_az, time_frac, out_img_size, W2C, W2L_H, device, rank=0,
                         world_size=1, include_exact_solar=False, dst=0):
    """Ray-sharded novel-view render with the final gather (12 bytes of colour + a mask value per ray).
    dst=r (default 0): rank r returns the full [H,W,3] float64 image and the [H,W] shadow mask, every other rank returns
    (None, None) - ONE gather to one GPU and ONE device->host copy of the image.  dst=None: every rank returns the image
    (all_gather + a host copy per rank: world_size times the PCIe / host traffic; kept for callers that need it).
    world_size 1 needs no process group."""
    from .train import gather_rows
    import os
    import sys
    import time
    timing = os.environ.get("SNB_SHARD_TIMING") == "1"
    H, W = out_img_size[0], out_img_size[1]
    if timing:
        t.cuda.synchronize()
        t0 = time.perf_counter()
    lo, hi, rgb, mask = render_shard(the_network, view_el_az, sun_el_az, time_frac, out_img_size, W2C, W2L_H, device, rank,
                                     world_size, include_exact_solar)
    if timing:
        t.cuda.synchronize()
        t1 = time.perf_counter()
    both = t.cat([rgb, mask.unsqueeze(1)], 1)                   # colour + mask travel together: one gather, one D2H copy
    if world_size > 1:
        both = gather_rows(both, H * W, rank, world_size, dst=dst)
    if timing:
        t.cuda.synchronize()
        t2 = time.perf_counter()
    out = (None, None)
    if both is not None:
        # split on the device (strided host copies of a 33 MB array cost 30 ms) and leave through the pinned staging path
        out = (device_to_numpy(both[:, :3].contiguous(), chunk_bytes=32 << 20, min_bytes=4 << 20).reshape(H, W, 3),
               device_to_numpy(both[:, 3].contiguous(), chunk_bytes=32 << 20, min_bytes=4 << 20).reshape(H, W))
    if timing:
        t3 = time.perf_counter()
        sys.stderr.write("[shard timing] rank %d: render %.1f ms, gather %.1f ms, to host %.1f ms\n"
                         % (rank, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
    return out
