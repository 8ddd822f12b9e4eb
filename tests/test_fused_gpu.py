"""GPU parity of the fused tcgen05 render kernel against (a) the CPU interpretation of the same program (identical
bf16 rounding points), (b) the layer-wise bf16 path and (c) the fp32 oracle network."""
import numpy as np
import pytest
import torch as t

from gpu_util import make_net, maxabs

pytestmark = pytest.mark.gpu


def _pts(M, seed):
    g = t.Generator().manual_seed(seed)
    return t.rand(M, 3, generator=g) * 2 - 1


@pytest.mark.parametrize("M,S", [(128, 1), (96 * 5, 96), (128 * 148 + 37, 1), (96 * 700, 96)])
def test_fused_matches_program_interpreter_and_oracle(params0, M, S):
    from oracle import season_oracle as so
    from season_nerf_b200 import fused, packing2 as packing
    net = make_net(params0, "bf16")
    pts = _pts(M, M)
    nr = (M + S - 1) // S
    sun = t.nn.functional.normalize(t.rand(nr, 3, generator=t.Generator().manual_seed(3)) + .1, dim=1)
    with t.no_grad():
        _, pos4, vis, adj = fused.run(net, pts.cuda(), sun.cuda(), S)
        t.cuda.synchronize()
    sel = t.cat([t.arange(0, min(M, 200)), t.arange(max(M - 150, 0), M)]).unique()
    _, info = packing.build_program(params0)
    sun_pts = sun.repeat_interleave(S, 0)[:M]
    enc = t.cat([so.pe_encode(pts[sel], 10), t.zeros(len(sel), 1)], 1)
    senc = t.cat([so.pe_encode(sun_pts[sel], 4), t.zeros(len(sel), 64 - 27)], 1)
    ref = packing.interpret(info, enc, senc)
    assert maxabs(pos4[sel.cuda()], ref[packing.OUT_POS]) < 1.5e-2
    assert maxabs(vis[sel.cuda()], ref[packing.OUT_VIS][:, 0]) < 1.5e-2
    assert maxabs(adj[sel.cuda()], ref[packing.OUT_ADJ]) < 1.5e-2
    with t.no_grad():
        rho, col, v, sky, cls, a = so._link(params0, pts[sel], sun_pts[sel], t.zeros(len(sel), 4), False)
    assert maxabs(pos4[sel.cuda()], t.cat([rho, col], 1)) < 6e-2
    assert maxabs(vis[sel.cuda()], v[:, 0]) < 6e-2
    assert maxabs(adj[sel.cuda()], a.reshape(-1, 12)) < 6e-2
    assert bool(t.isfinite(pos4).all()) and bool(t.isfinite(adj).all())


def test_fused_sigma_only_and_dispatch(params0):
    from oracle import season_oracle as so
    from season_nerf_b200 import fused
    net = make_net(params0, "bf16")
    pts = _pts(1000, 9)
    with t.no_grad():
        rho = net.forward_Classic_Sigma_Only(pts.cuda())                 # per-point API: layer-wise path
        (raw,) = net.forward_rays(pts.cuda(), None, None, 1, mode="sigma")  # fused sigma-only program
        ref = so.forward_sigma_only(params0, pts)
    assert maxabs(t.nn.functional.softplus(raw), ref) < 3e-2
    assert maxabs(rho, ref) < 3e-2
    # program cache follows parameter updates
    with t.no_grad():
        net.G_NeRF_net.fc10Sigma.bias.add_(1.0)
        (raw2,) = net.forward_rays(pts.cuda(), None, None, 1, mode="sigma")
    assert abs(float((raw2 - raw).mean()) - 1.0) < 1e-3


def test_fused_is_deterministic(params0):
    from season_nerf_b200 import fused
    net = make_net(params0, "bf16")
    pts = _pts(128 * 300, 4).cuda()
    sun = t.tensor([[0.3, -0.4, 0.866]]).cuda()
    with t.no_grad():
        a = fused.run(net, pts, sun, pts.shape[0])
        b = fused.run(net, pts, sun, pts.shape[0])
    for x, y in zip(a[1:], b[1:]):
        assert t.equal(x, y)


def test_ray_sharded_render_equals_unsharded(params0):
    """rays shard with no data-path exchange; shards rendered independently (here: both 'ranks' on one GPU) and
    concatenated must reproduce the single-device image of main_run_Season_NeRF.py:90-92."""
    import season_nerf_b200 as snb
    from oracle import season_oracle as so
    net = make_net(params0, "bf16")
    size = (12, 10, 96)
    dev = t.device("cuda")
    D = snb.component_render_by_dir(net, [80, 0], [45, 135], 184 / 365, size, so.OMA_W2C, so.oma_w2l_h(), dev,
                                    include_exact_solar=True)
    imgs = snb.get_imgs_from_Img_Dict(D, size, False)
    ref = imgs["Season_Adj_Img"] * imgs["Shadow_Adjust_Exact"]
    full, mask = snb.render_image_sharded(net, [80, 0], [45, 135], 184 / 365, size, so.OMA_W2C, so.oma_w2l_h(), dev,
                                          include_exact_solar=True)
    assert maxabs(full, ref) < 1e-6 and maxabs(mask, imgs["Shadow_Mask_Exact"]) < 1e-6
    parts = [snb.render_shard(net, [80, 0], [45, 135], 184 / 365, size, so.OMA_W2C, so.oma_w2l_h(), dev, r, 3,
                              include_exact_solar=True) for r in range(3)]
    assert parts[0][0] == 0 and parts[-1][1] == 120 and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    cat = t.cat([p[2] for p in parts], 0).reshape(12, 10, 3).cpu().numpy()
    assert maxabs(cat, ref) < 1e-6
    # year sweep of the shard (raw heads -> ops.year_sweep_raw) vs the component path (get_imgs_from_Img_Dict_t_step)
    tf = np.array([snb.encode_time(k / 5) for k in range(5)])
    with t.no_grad():
        cv = net.get_class_only(t.tensor(tf, dtype=t.float32, device=dev)).double().cpu().numpy()
    sweep_ref = snb.get_imgs_from_Img_Dict_t_step(D, size, cv)
    _, _, slab, _ = snb.render_shard(net, [80, 0], [45, 135], 184 / 365, size, so.OMA_W2C, so.oma_w2l_h(), dev, 0, 1,
                                     include_exact_solar=True, class_vecs=cv)
    assert maxabs(slab.reshape(5, 12, 10, 3), sweep_ref) < 2e-6


def test_fused_full_size_permutation_and_chunk_invariance(params0):
    """BASELINE.json configs[2] size (512 x 512 rays x 96 samples = 25.2 M points, rendered in the API's 4 M-point calls):
    every point is evaluated independently of its tile, of the CTA pair that renders it and of the call it falls in -
    a permuted batch gives the permuted outputs BIT FOR BIT, two half calls give the whole call, and all outputs are
    finite.  Size-independent properties of the fused kernel where the CPU oracle cannot follow."""
    from season_nerf_b200 import fused
    net = make_net(params0, "bf16")
    M = 1 << 22                                     # one API call of the render path
    g = t.Generator(device="cuda").manual_seed(21)
    sun = t.tensor([[0.3, -0.4, 0.866]], device="cuda")
    total, checksum = 0, 0.0
    with t.no_grad():
        for call in range(6):                       # 6 x 4.19 M = 25.2 M points
            pts = t.rand(M, 3, device="cuda", generator=g) * 2 - 1
            _, pos4, vis, adj = fused.run(net, pts, sun, M)
            assert bool(t.isfinite(pos4).all()) and bool(t.isfinite(vis).all()) and bool(t.isfinite(adj).all())
            total += M
            checksum += float(pos4.double().sum())
            if call == 0:
                perm = t.randperm(M, device="cuda", generator=g)
                _, p2, v2, a2 = fused.run(net, pts[perm], sun, M)
                assert t.equal(p2, pos4[perm]) and t.equal(v2, vis[perm]) and t.equal(a2, adj[perm])
                h = M // 2 + 77                      # ragged split: the second call starts inside a 256-point tile
                _, pa, va, aa = fused.run(net, pts[:h], sun, h)
                _, pb, vb, ab = fused.run(net, pts[h:], sun, M - h)
                assert t.equal(t.cat([pa, pb]), pos4) and t.equal(t.cat([va, vb]), vis) and t.equal(t.cat([aa, ab]), adj)
                rho, _, _, _ = fused.run(net, pts, None, M, sigma_only=True)
                assert float((rho - pos4[:, 0]).abs().max()) == 0.0      # the sigma-only program is the same trunk
    assert total == 6 * M and np.isfinite(checksum)
