"""GPU parity of the drop-in network / engine / CLI render against fixtures produced by the unmodified reference
(tests/golden, oracle/make_golden.py) and against the CPU oracle on the same seeded inputs.
Tolerances (north star): fp32 validation build 1e-4 max-abs on rendered outputs; bf16 production build 1e-2."""
import numpy as np
import pytest
import torch as t

from conftest import load_golden
from gpu_util import T, make_net, maxabs, relerr

pytestmark = pytest.mark.gpu
S = 96
TOL = {"fp32": dict(out=1e-4, rho=2e-3, grad=2e-3), "bf16": dict(out=1e-2, rho=6e-2, grad=8e-2)}


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_network_entry_points_eval(params0, precision):
    g = load_golden("net_eval")
    net = make_net(params0, precision)
    X, sun, Time = T(g["X"]), T(g["sun"]), T(g["Time"])
    tol = TOL[precision]
    with t.no_grad():
        fw = net.forward(X, sun, Time)
        fs = net.forward_seperate(X, sun, Time)
        sol = net.forward_Solar(X, sun, Time)
        sig = net.forward_Classic_Sigma_Only(X)
        cls = net.get_class_only(Time)
        colo = net.G_NeRF_net.forward_color_only(X)
    names = ["rho", "col", "vis", "sky", "cls", "adj"]
    for o, k in zip(fw, names):
        assert maxabs(o, g["fw_" + k]) < (tol["rho"] if k in ("rho", "adj") else tol["out"]), ("fw", k, maxabs(o, g["fw_" + k]))
    for o, k in zip(fs, names):
        assert tuple(o.shape) == g["fs_" + k].shape
        assert maxabs(o, g["fs_" + k]) < (tol["out"] if k in ("vis", "sky", "cls") else tol["rho"]), ("fs", k)
    for o, k in zip(sol, ["rho", "vis", "sky"]):
        assert maxabs(o, g["sol_" + k]) < tol["rho"], ("sol", k)
    assert maxabs(sig, g["sigma_only"]) < tol["rho"]
    assert maxabs(cls, g["class_only"]) < tol["out"]
    assert maxabs(colo, g["color_only"]) < tol["out"]


KEYS = ["Rendered_Col", "PE", "PV", "PS", "Solar_Vis", "Sky_Col", "Classes", "Adjust", "Rho", "Col", "deltas",
        "Albedo_Color"]


def _data(g):
    return {k[3:]: t.tensor(v) for k, v in g.items() if k.startswith("in_")}     # CPU tensors, like the DataLoader rows


def _tool(args=None, use_prior=False, ada=None, n_steps=100):
    import season_nerf_b200 as snb
    from oracle import season_oracle as so
    return snb.All_in_One_Eval(args or so.default_args(), t.device("cuda"), n_steps, use_prior, ada, so.oma_w2l_h(), so.OMA_W2C)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_engine_eval_golden(params0, precision):
    g = load_golden("engine_eval")
    net = make_net(params0, precision)
    with t.no_grad():
        R = _tool().eval(_data(g), net, 0, False)
    assert np.array_equal(R["sample_pts"].cpu().numpy(), g["sample_pts"])
    tol = TOL[precision]
    for k in KEYS:
        assert tuple(R[k].shape) == g[k].shape, k
        lim = tol["rho"] if k in ("Rho", "Adjust") else tol["out"]
        assert maxabs(R[k], g[k]) < lim, (k, maxabs(R[k], g[k]))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_engine_train_forward_batchnorm_golden(params0, precision):
    g = load_golden("engine_train_fwd")
    net = make_net(params0, precision, train=True)
    with t.no_grad():
        R = _tool().eval(_data(g), net, 0, True, jitter=g["jitter"])
    assert np.array_equal(R["sample_pts"].cpu().numpy(), g["sample_pts"])
    tol = TOL[precision]
    for k in KEYS:
        lim = tol["rho"] if k in ("Rho", "Adjust") else tol["out"]
        assert maxabs(R[k], g[k]) < lim, (k, maxabs(R[k], g[k]))
    sd = net.state_dict()
    rt = 1e-4 if precision == "fp32" else 2e-2
    assert relerr(sd["G_NeRF_net.fc2.norm.running_mean"], g["fc2_rm"]) < rt
    assert relerr(sd["G_NeRF_net.fc2.norm.running_var"], g["fc2_rv"]) < rt
    assert relerr(sd["G_NeRF_net.fc9.norm.running_var"], g["fc9_rv"]) < rt
    assert int(sd["G_NeRF_net.fc2.norm.num_batches_tracked"]) == 1


def _ada(use_prior):
    from season_nerf_b200 import AdaptiveLossFunction as mk
    a0 = mk(3, t.float32, "cuda", alpha_hi=2.99, alpha_init=2.0, scale_init=0.03, scale_lo=0.01)
    if not use_prior:
        return a0
    return [a0, mk(1, t.float32, "cuda", alpha_hi=2.99, alpha_init=2.0, scale_init=0.5, scale_lo=0.05)]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name,kw", [("loss_barron", {}), ("loss_mse", {"mse": True}), ("loss_prior", {"use_prior": True}),
                                     ("loss_type2", {"type2": True})])
def test_get_loss_and_gradients_golden(params0, precision, name, kw):
    from oracle import season_oracle as so
    import season_nerf_b200 as snb
    g = load_golden(name)
    use_prior, mse, type2 = kw.get("use_prior", False), kw.get("mse", False), kw.get("type2", False)
    args = so.default_args(Use_MSE_loss=mse, Solar_Type_2=type2)
    net = snb.T_NeRF(512, 4, **({"HM": g["hm"]} if use_prior else {}), precision=precision)
    net.load_state_dict({k: v.clone() for k, v in params0.items()})
    net = net.cuda().train()
    ada = None if mse else _ada(use_prior)
    tool = _tool(args, use_prior, ada)
    solar = tuple(t.tensor(g[k]) for k in ("s_top", "s_bot", "s_sun", "s_time"))
    L = tool.get_loss(_data(g), net, 30, True, jitter=g["jitter"], solar=solar, solar_jitter=g["solar_jitter"])
    tol = TOL[precision]
    assert set(L.keys()) == {k[5:] for k in g if k.startswith("loss_")}
    for k in L:
        ref = float(g["loss_" + k])
        assert abs(float(L[k][0]) - ref) < (3e-3 if precision == "fp32" else 8e-2) * max(abs(ref), 1e-2), (k, float(L[k][0]), ref)
        assert abs(float(L[k][1]) - float(g["w_" + k])) <= 1e-4 * abs(float(g["w_" + k]))
    tot = sum(L[k][0] * L[k][1] for k in L)
    tot.backward()
    norms = dict(zip(g["grad_names"].tolist(), g["grad_norms"].tolist()))
    scale = max(norms.values())
    bad = []
    for k, p in net.named_parameters():
        n = 0.0 if p.grad is None else float(p.grad.norm())
        if abs(n - norms[k]) > tol["grad"] * norms[k] + 1e-4 * scale:
            bad.append((k, n, norms[k]))
    assert not bad, bad
    for k in g:
        if k.startswith("grad_") and k not in ("grad_names", "grad_norms"):
            p = dict(net.named_parameters())[k[5:]]
            if float(np.abs(g[k]).max()) < 1e-4 * scale:
                continue
            assert relerr(p.grad, g[k]) < (5e-3 if precision == "fp32" else 0.15), (k, relerr(p.grad, g[k]))
    for k in ["adjust_rho.weight", "adjust_solar_vis.weight", "adjust_sky_col.weight"]:
        p = dict(net.named_parameters())[k]
        assert p.grad is None or float(p.grad.abs().max()) == 0.0
    sd = net.state_dict()
    assert int(sd["G_NeRF_net.fc2.norm.num_batches_tracked"]) == 2           # image pass + no_grad solar pass
    assert relerr(sd["G_NeRF_net.fc2.norm.running_var"], g["fc2_rv"]) < (1e-4 if precision == "fp32" else 2e-2)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_cli_component_render_and_composite_golden(params0, precision):
    import season_nerf_b200 as snb
    from oracle import season_oracle as so
    g = load_golden("cli_render")
    net = make_net(params0, precision)
    D = snb.component_render_by_dir(net, [80, 0], [45, 135], 184 / 365, (5, 6, S), so.OMA_W2C, so.oma_w2l_h(),
                                    t.device("cuda"), include_exact_solar=False)
    assert D["World_Points"].dtype == np.float64
    assert np.array_equal(D["World_Points"].astype(np.float32), g["d_World_Points"])
    assert np.array_equal(D["Deltas"].astype(np.float32), g["d_Deltas"])
    assert np.array_equal(D["Image_Points"], g["d_Image_Points"])
    tol = TOL[precision]
    for k in ["Rho", "Base_Col", "Est_Solar_Vis", "Sky_Col", "Output_class", "Adjust_col"]:
        assert D[k].shape == g["d_" + k].shape, k
        lim = tol["out"] if k in ("Est_Solar_Vis", "Sky_Col", "Output_class") else tol["rho"]
        assert maxabs(D[k], g["d_" + k]) < lim, (k, maxabs(D[k], g["d_" + k]))
    imgs = snb.get_imgs_from_Img_Dict(D, (5, 6, S), False)
    for k in ["Base_Img", "Season_Adj_Img", "Shadow_Adjust", "Shadow_Mask", "Sky_Col", "Time_Class"]:
        assert maxabs(imgs[k], g[k]) < tol["out"], (k, maxabs(imgs[k], g[k]))
    assert maxabs(np.array(imgs["Extreme_Imgs"]), g["Extreme_Imgs"]) < tol["out"]
    sweep = snb.get_imgs_from_Img_Dict_t_step(D, (5, 6, S), g["class_vecs"].astype(np.float64))
    assert maxabs(sweep, g["sweep"]) < tol["out"]
    # float64 compositing kernels on the REFERENCE's own component arrays: pure arithmetic parity
    Dref = {k[2:]: g[k].astype(np.float64) for k in g if k.startswith("d_") and k != "d_Image_Points"}
    Dref["Image_Points"] = g["d_Image_Points"]
    imgs = snb.get_imgs_from_Img_Dict(Dref, (5, 6, S), False)
    for k in ["Base_Img", "Season_Adj_Img", "Shadow_Adjust", "Shadow_Mask", "Raw_Shadow_Mask"]:
        assert maxabs(imgs[k], g[k]) < 2e-6, (k, maxabs(imgs[k], g[k]))
    sweep = snb.get_imgs_from_Img_Dict_t_step(Dref, (5, 6, S), g["class_vecs"].astype(np.float64))
    assert maxabs(sweep, g["sweep"]) < 2e-6


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_cli_exact_shadow_march_golden(params0, precision):
    import season_nerf_b200 as snb
    from oracle import season_oracle as so
    g = load_golden("cli_render_exact")
    net = make_net(params0, precision)
    D = snb.component_render_by_dir(net, [70, 30], [35, 200], 0.25, (2, 3, S), so.OMA_W2C, so.oma_w2l_h(),
                                    t.device("cuda"), include_exact_solar=True)
    tol = TOL[precision]
    assert maxabs(D["Exact_Solar"], g["d_Exact_Solar"]) < tol["out"]
    imgs = snb.get_imgs_from_Img_Dict(D, (2, 3, S), False)
    for k in ["Season_Adj_Img", "Shadow_Adjust_Exact", "Shadow_Mask_Exact"]:
        assert maxabs(imgs[k], g[k]) < tol["out"], (k, maxabs(imgs[k], g[k]))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_quick_run_and_engine_exact_solar_golden(params0, precision):
    import season_nerf_b200 as snb
    from oracle import season_oracle as so
    g = load_golden("quick_run")
    net = make_net(params0, precision)
    args = so.default_args()
    q = snb.Quick_Run_Net(net, args, so.OMA_W2C, so.oma_w2l_h(), t.device("cuda"), use_full_solar=False)
    img, mask = q.render_img([75, 20], [50, 120], 0.4, 6)
    tol = TOL[precision]
    assert np.array_equal(mask, g["mask"])
    assert maxabs(img["Col_Img"], g["Col_Img"]) < tol["out"]
    assert maxabs(img["Shadow_Mask"], g["Shadow_Mask"]) < tol["out"]
    dsm = q.get_DSM((4, 4))
    assert maxabs(np.nan_to_num(dsm), np.nan_to_num(g["dsm"])) < tol["out"] * 2
    with t.no_grad():
        R = _tool().eval_exact_solar(_data(g), net, -1, False)
    assert maxabs(R["Solar_Vis"], g["ex_Solar_Vis"]) < tol["out"]
    assert maxabs(R["Rendered_Col"], g["ex_Rendered_Col"]) < tol["out"]


def test_component_render_by_P_device_rays_equal_host_rays(params0):
    """component_render_by_P (mg_Img_Eval.py:74-94): the on-device camera-ray path (objects of the reference's
    P_img_Pinhole class) and the host path (any other object with invert_P) give identical rays, masks and components."""
    import season_nerf_b200 as snb
    from oracle import season_oracle as so

    class _Cam:
        def __init__(self):
            P = so.synthetic_camera_P(seed=7)
            self.P = P / P[-1, -1]
            self.img = np.zeros((2048, 2048, 3), dtype=np.uint8)
            self.sun_el_and_az_vec = so.world_angle_2_local_vec(45, 135, so.OMA_W2C, so.oma_w2l_h())

        def invert_P(self, row, col, h=0):
            return so.invert_P(self.P, row, col, h)

        def get_year_frac(self):
            return 0.5

    P_img_Pinhole = type("P_img_Pinhole", (_Cam,), {})
    net = make_net(params0, "bf16")
    size = (9, 8, 24)
    Dd = snb.component_render_by_P(net, P_img_Pinhole(), size, t.device("cuda"), include_exact_solar=False)
    Dh = snb.component_render_by_P(net, _Cam(), size, t.device("cuda"), include_exact_solar=False)
    assert np.array_equal(Dd["Image_Points"], Dh["Image_Points"]) and np.array_equal(Dd["Image_Points_in_GT_Img"], Dh["Image_Points_in_GT_Img"])
    assert Dd["World_Points"].shape[0] < size[0] * size[1]          # some rays fall outside the cube and are dropped
    for k in ("World_Points", "Deltas", "Rho", "Base_Col", "Adjust_col"):
        assert np.array_equal(Dd[k], Dh[k]), k
    pts, tops, bots = snb.ray_table_from_P(P_img_Pinhole().P, (2048, 2048), 64, np.array([[-1., 1.]] * 3), "cuda")
    r, c = pts[:, 0].cpu().numpy() * 64, pts[:, 1].cpu().numpy() * 64
    to, bo, go = so.camera_rays(P_img_Pinhole().P, np.repeat(np.arange(32), 32) * 64, np.tile(np.arange(32), 32) * 64)
    assert pts.shape[0] == int(go.sum()) and np.array_equal(tops.cpu().numpy(), to[go].astype(np.float32))
    assert np.array_equal(bots.cpu().numpy(), bo[go].astype(np.float32))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_gen_results_and_height_map(params0, precision):
    """"next" row 4: dense volume consumers (Eval_funcs.py:268-313) against the unmodified reference's arrays."""
    import season_nerf_b200 as snb
    g = load_golden("gen_results")
    shape, S = tuple(int(v) for v in g["shape"]), int(g["S"])
    net = make_net(params0, precision)
    rho, pe, pv, ps, col = snb.gen_results(net, shape, S, t.device("cuda"), 1000)
    tol = TOL[precision]
    assert rho.dtype == np.float64 and rho.shape == g["rho"].shape and col.shape == g["col"].shape
    assert maxabs(rho, g["rho"]) < tol["rho"] and maxabs(col, g["col"]) < tol["out"]
    for a, k in ((pe, "P_E"), (pv, "P_Vis"), (ps, "P_Surf")):
        assert maxabs(a, g[k]) < tol["out"], k
    hm = snb.height_map(net, shape, S, t.device("cuda"))
    assert maxabs(hm, g["height"]) < (5e-2 if precision == "bf16" else 1e-3)
    hm_m = snb.height_map(net, shape, S, t.device("cuda"), h_range=(280.0, 350.0))
    assert maxabs(hm_m, (hm + 1) / 2 * 70.0 + 280.0) < 1e-9


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_seasonal_alignment_search(params0, precision):
    """"next" row 3: Grad_Descent_Seasonal_Align_v3 (mg_Img_Eval.py:349-414) on the reference's cached components - one
    fused sweep over the 367 candidate times - returns the reference's time of year, class vector and sky colour."""
    import season_nerf_b200 as snb
    from test_oracle_golden import _align_inputs
    g, P, D = _align_inputs(params0)
    net = make_net(P, precision)
    adj, sky, best_t = snb.Grad_Descent_Seasonal_Align_v3(D, g["target"], float(g["t0"]), net, t.device("cuda"))
    assert abs(best_t - float(g["best_t"])) < (1e-7 if precision == "fp32" else 2.5 / 365)      # bf16 class vectors: +-2 days
    tol = 1e-4 if precision == "fp32" else 3e-2
    assert tuple(sky.shape) == (1, 1, 3) and maxabs(sky, g["sky"]) < tol and maxabs(adj, g["adj_vec"]) < tol


def test_classic_shadow_images_and_alignment_golden(params0):
    """use_classic_shadows=True: get_imgs_from_Img_Dict (mg_Img_Eval.py:166-181) on the REFERENCE's component arrays
    (float64 arithmetic parity with the unmodified reference), on device-resident components of our own render, and the
    classic-shadow alignment search (:416-475) - two weighted year-sweep launches - against the reference's result."""
    import season_nerf_b200 as snb
    from oracle import season_oracle as so
    from test_oracle_golden import _align_inputs
    gc = load_golden("cli_classic")
    for fx, size, keys in (("cli_render", (5, 6, S), (("Shadow_Adjust", "sa_classic"),)),
                           ("cli_render_exact", (2, 3, S), (("Shadow_Adjust", "sa_classic_x"), ("Shadow_Adjust_Exact", "sae_classic_x")))):
        g = load_golden(fx)
        Dref = {k[2:]: (g[k].astype(np.float64) if g[k].dtype == np.float32 else g[k]) for k in g if k.startswith("d_")}
        imgs = snb.get_imgs_from_Img_Dict(Dref, size, True)
        for k, gk in keys:
            assert maxabs(imgs[k], gc[gk]) < 2e-6, (fx, k, maxabs(imgs[k], gc[gk]))
        plain = snb.get_imgs_from_Img_Dict(Dref, size, False)
        assert maxabs(plain["Shadow_Adjust"], imgs["Shadow_Adjust"]) > 1e-3
        assert maxabs(plain["Season_Adj_Img"], imgs["Season_Adj_Img"]) == 0
    # device-resident float32 components of our own fp32 render (DeviceImgDict path)
    net = make_net(params0, "fp32")
    D = snb.component_render_by_dir(net, [80, 0], [45, 135], 184 / 365, (5, 6, S), so.OMA_W2C, so.oma_w2l_h(),
                                    t.device("cuda"), include_exact_solar=False)
    imgs = snb.get_imgs_from_Img_Dict(D, (5, 6, S), True)
    assert maxabs(imgs["Shadow_Adjust"], gc["sa_classic"]) < 1e-3
    # alignment
    g, P, Da = _align_inputs(params0)
    net = make_net(P, "fp32")
    adj, sky, best_t = snb.Grad_Descent_Seasonal_Align_v3(Da, g["target"], float(g["t0"]), net, t.device("cuda"),
                                                          use_classic_shadows=True)
    assert abs(best_t - float(gc["best_t"])) < 1e-7
    assert tuple(sky.shape) == (1, 1, 3) and maxabs(sky, gc["sky"]) < 1e-4 and maxabs(adj, gc["adj_vec"]) < 1e-4


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_remaining_entry_points_golden(params0, precision):
    """full_eval (both solar conventions), approx_Solar, forward_full_eval, forward_Position and create_given_vec against
    the unmodified reference (fixture api_extra)."""
    import season_nerf_b200 as snb
    from oracle import season_oracle as so
    g = load_golden("api_extra")
    net = make_net(params0, precision)
    tol = TOL[precision]
    X, Xs, sun, Time = (T(g[k]) for k in ("X", "Xs", "sun", "Time"))
    with t.no_grad():
        ap = net.approx_Solar(X, Xs, Time)
        fe = net.forward_full_eval(X, sun, Time)
        fp = net.G_NeRF_net.forward_Position(X)
    for i, (o, k) in enumerate(zip(ap, ["rho", "rho", "out", "out", "rho"])):
        assert tuple(o.shape) == g["ap_%d" % i].shape and maxabs(o, g["ap_%d" % i]) < tol[k], ("approx_Solar", i, maxabs(o, g["ap_%d" % i]))
    for i, (o, k) in enumerate(zip(fe, ["rho", "rho", "out", "out", "out", "rho"])):
        assert tuple(o.shape) == g["fe_%d" % i].shape and maxabs(o, g["fe_%d" % i]) < tol[k], ("forward_full_eval", i)
    for i, o in enumerate(fp):
        # fp[0] = X_Encode, the 256 hidden activations after nine x30 SIREN layers (not a rendered quantity): in bf16 a
        # single activation may move by up to ~0.06 while the heads computed from all of them stay within tol["rho"]
        lim = tol["rho"] if (i > 0 or precision == "fp32") else 0.12
        assert tuple(o.shape) == g["fp_%d" % i].shape and maxabs(o, g["fp_%d" % i]) < lim, ("forward_Position", i, maxabs(o, g["fp_%d" % i]))
    d = _data(g)
    for tag, classic in (("full", False), ("fullc", True)):
        R = _tool(so.default_args(Solar_Type_2=classic)).full_eval(d, net, 0)
        assert np.array_equal(R["sample_pts"].cpu().numpy(), g[tag + "_sample_pts"])
        assert np.array_equal(R["deltas"].cpu().numpy(), g[tag + "_deltas"])
        for k in ["Rendered_Col", "PE", "PV", "PS", "Solar_Vis", "Sky_Col", "Classes", "Adjust", "Rho", "Col"]:
            assert tuple(R[k].shape) == g[tag + "_" + k].shape, (tag, k, tuple(R[k].shape))
            lim = tol["rho"] if k in ("Rho", "Adjust") else tol["out"]
            assert maxabs(R[k], g[tag + "_" + k]) < lim, (tag, k, maxabs(R[k], g[tag + "_" + k]))
    tool = snb.create_solor_rays_uniform(so.oma_w2l_h(), so.OMA_W2C)
    t.manual_seed(9)
    st, en, sv, tm = tool.create_given_vec(12, g["gv_vec"].astype(np.float64), include_times=True)
    assert np.array_equal(st.numpy(), g["gv_starts"]) and maxabs(en, g["gv_ends"]) < 1e-5
    assert maxabs(sv, g["gv_sun"]) < 1e-7 and maxabs(tm, g["gv_times"]) < 1e-6


PRIOR_KEYS = ["Rendered_Col", "Albedo_Color", "PS", "PV_Supervised", "PE_Supervised", "PS_Supervised", "Rendered_Col_Supervised",
              "PV_Merged", "PE_Merged", "PS_Merged", "Rendered_Col_Merged", "Rho_Merged"]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_engine_eval_prior_golden(params0, precision):
    """DSM-guided section (Eval_Tools_2.py:217-246) vs the unmodified reference: the supervised and merged colours are shaded
    with the Solar_Vis3 of the network's own PS (fixture engine_eval_prior)."""
    import season_nerf_b200 as snb
    g = load_golden("engine_eval_prior")
    tol = TOL[precision]
    for tag, train in (("ev_", False), ("tr_", True)):
        net = snb.T_NeRF(512, 4, HM=g["hm"], precision=precision)
        net.load_state_dict({k: v.clone() for k, v in params0.items()})
        net = net.cuda().train(train)
        with t.no_grad():
            R = _tool(use_prior=True).eval(_data(g), net, 30, train, jitter=g["jitter"] if train else None)
        for k in PRIOR_KEYS:
            lim = tol["rho"] if k == "Rho_Merged" else tol["out"]
            assert maxabs(R[k], g[tag + k]) < lim, (tag, k, maxabs(R[k], g[tag + k]))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_get_loss_prior_mse_golden(params0, precision):
    """--Use_MSE_loss + prior: Rendered_Col_Merged is the training target, so the shading fix reaches the gradients."""
    from oracle import season_oracle as so
    import season_nerf_b200 as snb
    g = load_golden("loss_prior_mse")
    net = snb.T_NeRF(512, 4, HM=g["hm"], precision=precision)
    net.load_state_dict({k: v.clone() for k, v in params0.items()})
    net = net.cuda().train()
    tool = _tool(so.default_args(Use_MSE_loss=True), True, None)
    solar = tuple(t.tensor(g[k]) for k in ("s_top", "s_bot", "s_sun", "s_time"))
    L = tool.get_loss(_data(g), net, 30, True, jitter=g["jitter"], solar=solar, solar_jitter=g["solar_jitter"])
    assert set(L.keys()) == {k[5:] for k in g if k.startswith("loss_")}
    for k in L:
        ref = float(g["loss_" + k])
        assert abs(float(L[k][0]) - ref) < (3e-3 if precision == "fp32" else 8e-2) * max(abs(ref), 1e-2), (k, float(L[k][0]), ref)
    sum(L[k][0] * L[k][1] for k in L).backward()
    norms = dict(zip(g["grad_names"].tolist(), g["grad_norms"].tolist()))
    scale = max(norms.values())
    tol = TOL[precision]
    bad = []
    for k, p in net.named_parameters():
        n = 0.0 if p.grad is None else float(p.grad.norm())
        if abs(n - norms[k]) > tol["grad"] * norms[k] + 1e-4 * scale:
            bad.append((k, n, norms[k]))
    assert not bad, bad
    for k in g:
        if k.startswith("grad_") and k not in ("grad_names", "grad_norms"):
            p = dict(net.named_parameters())[k[5:]]
            if float(np.abs(g[k]).max()) < 1e-4 * scale:
                continue
            assert relerr(p.grad, g[k]) < (5e-3 if precision == "fp32" else 0.15), (k, relerr(p.grad, g[k]))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("type2", [False, True])
def test_get_loss_fused_heads_path_equals_general_path(params0, precision, type2):
    """get_loss through the fused raw-head kernels (engine.FAST_LOSS, the default) vs the general dictionary path
    (eval / eval_Rho_Only + torch activations): same losses, same gradients up to float32 summation order"""
    from oracle import season_oracle as so
    import season_nerf_b200 as snb
    from season_nerf_b200 import engine
    n = 64
    batch = so.synthetic_batch(n, seed=5, n_images=7)
    st, en, vec, tm, _ = so.create_solar_rays_uniform(n, so.OMA_W2C, so.oma_w2l_h(), np.random.RandomState(6), t.Generator().manual_seed(6))
    jit = t.rand(S, generator=t.Generator().manual_seed(8))
    res = {}
    for fast in (True, False):
        engine.FAST_LOSS = fast
        try:
            net = make_net(params0, precision, train=True)
            ada = _ada(False)
            tool = _tool(so.default_args(Solar_Type_2=type2), False, ada)
            L = tool.get_loss(batch, net, 30, True, jitter=jit, solar=(st, en, vec, tm), solar_jitter=jit)
            sum(L[k][0] * L[k][1] for k in L).backward()
            res[fast] = ({k: float(L[k][0]) for k in L}, {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None},
                         ada.latent_scale.grad.clone())
        finally:
            engine.FAST_LOSS = True
    lf, gf, af = res[True]
    lg, gg, ag = res[False]
    assert lf.keys() == lg.keys() and gf.keys() == gg.keys()
    for k in lf:
        assert abs(lf[k] - lg[k]) <= (2e-5 if precision == "fp32" else 2e-3) * max(abs(lg[k]), 1e-2), (k, lf[k], lg[k])
    scale = max(float(v.norm()) for v in gg.values())
    worst = max(float((gf[k] - gg[k]).norm()) / max(float(gg[k].norm()), 1e-4 * scale) for k in gg)
    assert worst < (2e-4 if precision == "fp32" else 5e-2), worst
    assert relerr(af, ag) < (1e-4 if precision == "fp32" else 2e-2)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_volume_consumers_golden(params0, precision):
    """eval_shadow_data (mg_Shadow_Eval.py:72-104) and eval_HM to the scores before alignment (Eval_funcs.py:298-395) on the
    device vs the unmodified reference (fixture evals)."""
    import season_nerf_b200 as snb
    from oracle import season_oracle as so
    g = load_golden("evals")
    net = make_net(params0, precision)
    tol = TOL[precision]
    ve, vs, sk = snb.eval_shadow_data(net, g["angles"], g["ground_points"], 96, so.OMA_W2C, so.oma_w2l_h(), 15000, t.device("cuda"))
    assert ve.shape == g["vis_exact"].shape and vs.shape == g["vis_est"].shape and sk.shape == g["sky_col"].shape and ve.dtype == np.float64
    assert maxabs(ve, g["vis_exact"]) < tol["out"] and maxabs(vs, g["vis_est"]) < tol["out"] and maxabs(sk, g["sky_col"]) < tol["rho"]
    Imgs, scores, conf_stats = snb.eval_HM(net, g["hm_GT"], (300., 370.), 96, t.device("cuda"), 5000)
    np.testing.assert_allclose(Imgs["GT"], g["hm_GT_m"], rtol=1e-12, equal_nan=True)
    m = ~np.isnan(g["hm_est_no_shift"])
    # heights in metres over a 70 m range: 1e-4 / 1e-2 of the normalised cube = 3.5 mm / 35 cm
    assert np.abs(Imgs["Est_HM_no_Shift"][m] - g["hm_est_no_shift"][m]).max() < 35 * tol["out"]
    for k, v in zip(g["before_keys"].tolist(), g["before_vals"].tolist()):
        assert abs(scores[k] - v) <= (1e-3 if precision == "fp32" else 5e-2) * max(abs(v), 1.0), (k, scores[k], v)
    if precision == "fp32":
        assert abs(conf_stats[0] - float(g["conf_mean"])) < 1e-6 and abs(conf_stats[1] - float(g["conf_median"])) < 1e-6
    else:       # one sample of a 96-sample column is 0.73 m of the 70 m range
        assert abs(conf_stats[0] - float(g["conf_mean"])) < 1.5


@pytest.mark.gpu
@pytest.mark.parametrize("type2", [False, True])
def test_get_loss_fused_tail_equals_torch_terms(params0, type2):
    """get_loss with the fused loss kernels (csrc/loss.cu, the default) against get_loss with the torch arithmetic
    (engine.FUSED_TAIL = False) on the fp32 validation build: same dictionary (keys, order, terms, weights), same gradients
    of the network and of the adaptive-loss parameters; and a caller that sums term * weight itself - the reference's
    train_step (mg_run_NeRF.py:299-318) - gets the gradient of the fused total"""
    import season_nerf_b200 as snb
    from season_nerf_b200 import engine
    from oracle import season_oracle as so
    from gpu_util import make_net
    args = so.default_args()
    args.Solar_Type_2 = type2
    n = 256
    batch = so.synthetic_batch(n, seed=2)
    st, en, vec, tm, _ = so.create_solar_rays_uniform(n, so.OMA_W2C, so.oma_w2l_h(), np.random.RandomState(5), t.Generator().manual_seed(5))
    jit = t.rand(96, generator=t.Generator().manual_seed(6))
    res = {}
    for fused in (True, False):
        engine.FUSED_TAIL = fused
        try:
            net = make_net(params0, "fp32", train=True)
            ada = snb.AdaptiveLossFunction(3, t.float32, "cuda", alpha_hi=2.99, alpha_init=1.8, scale_init=0.03, scale_lo=0.01)
            tool = snb.All_in_One_Eval(args, t.device("cuda"), 100, False, ada, so.oma_w2l_h(), so.OMA_W2C)
            L = tool.get_loss(batch, net, 30, True, jitter=jit, solar=(st, en, vec, tm), solar_jitter=jit)
            assert (tool._fused_total is not None) == fused
            fused_total = tool._fused_total
            tot = sum(L[k][0] * L[k][1] for k in L)
            if fused:
                assert abs(float(fused_total) - float(tot)) <= 1e-6 * abs(float(tot))
            tot.backward()
            res[fused] = ([(k, float(v[0]), float(v[1])) for k, v in L.items()],
                          {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None},
                          (ada.latent_alpha.grad.clone(), ada.latent_scale.grad.clone()))
        finally:
            engine.FUSED_TAIL = True
    (la, ga, aa), (lb, gb, ab) = res[True], res[False]
    assert [k for k, _, _ in la] == [k for k, _, _ in lb]
    for (k, v1, w1), (_, v0, w0) in zip(la, lb):
        assert abs(v1 - v0) <= 5e-6 * max(abs(v0), 1e-3) and abs(w1 - w0) <= 1e-6 * abs(w0), (k, v1, v0, w1, w0)
    assert set(ga) == set(gb)
    scale = max(float(g.norm()) for g in gb.values())
    for k in gb:
        assert float((ga[k] - gb[k]).norm()) <= 2e-5 * float(gb[k].norm()) + 1e-7 * scale, k
    for x, y in zip(aa, ab):
        assert float((x - y).abs().max()) <= 5e-5 * float(y.abs().max()) + 1e-9
