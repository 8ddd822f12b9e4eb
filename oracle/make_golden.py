"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported
from /root/reference with stubbed optional deps) on seeded synthetic inputs.

Run in the build container only:  python -m oracle.make_golden
The fixtures pin the oracle (tests/test_oracle_golden.py) and, through it and
directly, the CUDA path (tests/test_*_gpu.py).  Weights are NOT stored: they
are ``season_oracle.init_params(seed, perturb_bn=True)`` (torch CPU generator,
deterministic) loaded into the reference module with load_state_dict.
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch as t

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import season_oracle as so           # noqa: E402
from oracle import barron_loss                    # noqa: E402
from oracle.ref_import import import_reference    # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
S = 96


def np32(x):
    if isinstance(x, t.Tensor):
        x = x.detach().cpu().numpy()
    x = np.asarray(x)
    return x.astype(np.float32) if x.dtype == np.float64 else x


def save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **{k: np32(v) for k, v in arrs.items()})
    print("wrote", name, {k: np.asarray(v).shape for k, v in arrs.items() if np.asarray(v).size > 1} and "")


def ref_net(ref, params, hm=None, train=False):
    net = ref.T_NeRF(512, 4) if hm is None else ref.T_NeRF(512, 4, HM=hm)
    net.load_state_dict({k: v.clone() for k, v in params.items()}, strict=True)
    net.train(train)
    return net


def rays(n, seed):
    return so.synthetic_batch(n, seed=seed, n_images=5)


def main():
    ref = import_reference()
    t.set_num_threads(8)
    P0 = so.init_params(seed=0, perturb_bn=True)
    args = so.default_args()

    # ---- A: network entry points, eval mode ------------------------------------
    g = t.Generator().manual_seed(11)
    M = 192
    X = t.rand(M, 3, generator=g) * 2 - 1
    sun = t.nn.functional.normalize(t.rand(M, 3, generator=g) + t.tensor([-.5, -.5, .2]), dim=1)
    f = t.rand(M, generator=g)
    Time = t.stack([t.cos(2 * np.pi * f), t.sin(2 * np.pi * f), t.ones(M), t.zeros(M)], 1)
    net = ref_net(ref, P0)
    with t.no_grad():
        fw = net.forward(X, sun, Time)
        fs = net.forward_seperate(X, sun, Time)
        fsol = net.forward_Solar(X, sun, Time)
        sig = net.forward_Classic_Sigma_Only(X)
        cls = net.get_class_only(Time)
        colo = net.G_NeRF_net.forward_color_only(X)
        pe10 = net.G_NeRF_net.PE_encoder(X)
        pe4 = net.G_NeRF_net.PE_encoder_solar(sun)
        pe2 = net.Time_Enocder(Time[:, 0:2])
    save("net_eval", X=X, sun=sun, Time=Time,
         fw_rho=fw[0], fw_col=fw[1], fw_vis=fw[2], fw_sky=fw[3], fw_cls=fw[4], fw_adj=fw[5],
         fs_rho=fs[0], fs_col=fs[1], fs_vis=fs[2], fs_sky=fs[3], fs_cls=fs[4], fs_adj=fs[5],
         sol_rho=fsol[0], sol_vis=fsol[1], sol_sky=fsol[2], sigma_only=sig, class_only=cls, color_only=colo,
         pe10=pe10, pe4=pe4, pe2=pe2)

    # ---- B: sampling ------------------------------------------------------------
    d = rays(16, 21)
    top, bot = d["Top"].clone(), d["Bot"].clone()
    bot[3] = t.tensor([1.4, -0.2, -1.0])       # leaves the cube
    top[5] = t.tensor([-1.3, 0.9, 1.0])
    p0, d0 = ref.misc.sample_pt_coarse(top, bot, S, True)
    p1, d1 = ref.misc.sample_pt_coarse(top, bot, S, True, include_end_pt=True)
    t.manual_seed(5)
    jit = t.rand(S)
    t.manual_seed(5)
    p2, d2 = ref.misc.sample_pt_coarse(top, bot, S, False)
    bad = ref.misc.zero_invalid_pts()(p1)
    rho = t.rand(16, S, 1, generator=g) * 3
    pv = ref.get_PV(rho, d1)
    save("sampling", top=top, bot=bot, jitter=jit, pts_eval=p0, del_eval=d0, pts_end=p1, del_end=d1,
         pts_train=p2, del_train=d2, bad=bad.numpy(), rho=rho, pv=pv)

    # ---- C: engine eval (eval mode, then train mode with batch-stat BN) ---------
    d = rays(6, 31)
    tool = ref.All_in_One_Eval(args, t.device("cpu"), 100, False, None, so.oma_w2l_h(), so.OMA_W2C)
    with t.no_grad():
        R = tool.eval(d, net, 0, False)
    keys = ["Rendered_Col", "PE", "PV", "PS", "Solar_Vis", "Sky_Col", "Classes", "Adjust", "Rho", "Col", "deltas",
            "sample_pts", "Albedo_Color"]
    save("engine_eval", **{"in_" + k: v for k, v in d.items()}, **{k: R[k] for k in keys})

    netT = ref_net(ref, P0, train=True)
    t.manual_seed(6)
    jit = t.rand(S)
    t.manual_seed(6)
    with t.no_grad():
        R = tool.eval(d, netT, 0, True)
    sd = netT.state_dict()
    save("engine_train_fwd", **{"in_" + k: v for k, v in d.items()}, jitter=jit, **{k: R[k] for k in keys},
         fc2_rm=sd["G_NeRF_net.fc2.norm.running_mean"], fc2_rv=sd["G_NeRF_net.fc2.norm.running_var"],
         fc9_rm=sd["G_NeRF_net.fc9.norm.running_mean"], fc9_rv=sd["G_NeRF_net.fc9.norm.running_var"])

    # ---- D: get_loss + backward ------------------------------------------------
    def loss_case(name, n, seed, use_prior=False, mse=False, type2=False):
        a = so.default_args(Use_MSE_loss=mse, Solar_Type_2=type2)
        hm = None
        if use_prior:
            hm = (t.rand(16, 16, generator=t.Generator().manual_seed(seed)) * 1.2 - 0.6).numpy()
        nt = ref_net(ref, P0, hm=hm, train=True)
        if mse:
            ada = None
        elif use_prior:
            ada = [barron_loss.AdaptiveLossFunction(3, t.float32, "cpu", alpha_hi=2.99, alpha_init=2.0,
                                                    scale_init=0.03, scale_lo=0.01),
                   barron_loss.AdaptiveLossFunction(1, t.float32, "cpu", alpha_hi=2.99, alpha_init=2.0,
                                                    scale_init=0.5, scale_lo=0.05)]
        else:
            ada = barron_loss.AdaptiveLossFunction(3, t.float32, "cpu", alpha_hi=2.99, alpha_init=2.0,
                                                   scale_init=0.03, scale_lo=0.01)
        tl = ref.All_in_One_Eval(a, t.device("cpu"), 100, use_prior, ada, so.oma_w2l_h(), so.OMA_W2C)
        dd = rays(n, seed)
        # replay the reference's RNG draws in order to capture the injected inputs
        t.manual_seed(seed)
        np.random.seed(seed)
        jit = t.rand(S)
        az_el = np.random.random(n * 2).reshape([n, 2]) * np.array([[360, 89]]) + np.array([[-180, 1]])
        sx, sy = t.rand(n), t.rand(n)
        fr = t.rand(n, 2) * 2 * np.pi
        sjit = t.rand(S)
        t.manual_seed(seed)
        np.random.seed(seed)
        L = tl.get_loss(dd, nt, 30, True)
        tot = 0
        for k in L:
            tot = tot + L[k][0] * L[k][1]
        tot.backward()
        # the solar rays the reference drew
        vec = np.array([ref.world_angle_2_local_vec(az_el[i][1], az_el[i][0], so.OMA_W2C, so.oma_w2l_h())
                        for i in range(n)])
        starts = t.ones(n, 3)
        starts[:, 0] = 2 * sx - 1
        starts[:, 1] = 2 * sy - 1
        ends = (starts - 2 * (vec / vec[:, 2::])).float()
        stimes = t.stack([t.cos(fr[:, 0]), t.sin(fr[:, 0]), t.cos(fr[:, 1]), t.sin(fr[:, 1])], 1)
        out = {"in_" + k: v for k, v in dd.items()}
        out.update(jitter=jit, solar_jitter=sjit, s_top=starts, s_bot=ends, s_sun=t.tensor(vec).float(),
                   s_time=stimes, az_el=az_el, total=tot.detach())
        if hm is not None:
            out["hm"] = hm
        for k in L:
            out["loss_" + k] = t.as_tensor(L[k][0]).detach()
            out["w_" + k] = np.float32(float(L[k][1]))
        small = ["G_NeRF_net.fc10Sigma.weight", "G_NeRF_net.fc10Col.weight", "G_NeRF_net.fc_solar_4.weight",
                 "get_class_layer.weight", "adjust_col.weight", "G_NeRF_net.fc_sky_color_2.weight",
                 "G_NeRF_net.fc2.norm.weight", "G_NeRF_net.fc2.norm.bias", "G_NeRF_net.fc1.linear.bias",
                 "time_layer_1.linear.weight", "G_NeRF_net.fc9.linear.bias"]
        names, norms = [], []
        for k, prm in nt.named_parameters():
            names.append(k)
            norms.append(0.0 if prm.grad is None else float(prm.grad.norm()))
            if k in small:
                out["grad_" + k] = prm.grad
        out["grad_names"] = np.array(names)
        out["grad_norms"] = np.array(norms, dtype=np.float64)
        if ada is not None:
            adas = ada if isinstance(ada, list) else [ada]
            for i, a_ in enumerate(adas):
                out[f"ada{i}_galpha"] = a_.latent_alpha.grad
                out[f"ada{i}_gscale"] = a_.latent_scale.grad
        sdd = nt.state_dict()
        out["fc2_rm"] = sdd["G_NeRF_net.fc2.norm.running_mean"]
        out["fc2_rv"] = sdd["G_NeRF_net.fc2.norm.running_var"]
        out["fc2_nbt"] = sdd["G_NeRF_net.fc2.norm.num_batches_tracked"]
        save(name, **out)

    loss_case("loss_barron", 8, 41)
    loss_case("loss_mse", 8, 42, mse=True)
    loss_case("loss_prior", 8, 43, use_prior=True)
    loss_case("loss_type2", 8, 44, type2=True)

    # ---- E: CLI component render + float64 compositing --------------------------
    W2L = so.oma_w2l_h()
    size = (5, 6, S)
    Dd = ref.mg_Img_Eval.component_render_by_dir(net, [80, 0], [45, 135], 184 / 365, size, so.OMA_W2C, W2L,
                                                 t.device("cpu"), include_exact_solar=False)
    imgs = ref.mg_Img_Eval.get_imgs_from_Img_Dict(Dd, size, False)
    tf = np.array([so.encode_time(k / 7) for k in range(7)])
    with t.no_grad():
        cv = net.get_class_only(t.tensor(tf).float()).numpy()
    sweep = ref.mg_Img_Eval.get_imgs_from_Img_Dict_t_step(Dd, size, cv)
    save("cli_render", **{"d_" + k: v for k, v in Dd.items()},
         Base_Img=imgs["Base_Img"], Season_Adj_Img=imgs["Season_Adj_Img"], Shadow_Adjust=imgs["Shadow_Adjust"],
         Shadow_Mask=imgs["Shadow_Mask"], Raw_Shadow_Mask=imgs["Raw_Shadow_Mask"], Sky_Col=imgs["Sky_Col"],
         Time_Class=imgs["Time_Class"], Extreme_Imgs=np.array(imgs["Extreme_Imgs"]), class_vecs=cv, sweep=sweep)

    size = (2, 3, S)
    Dd = ref.mg_Img_Eval.component_render_by_dir(net, [70, 30], [35, 200], 0.25, size, so.OMA_W2C, W2L,
                                                 t.device("cpu"), include_exact_solar=True)
    imgs = ref.mg_Img_Eval.get_imgs_from_Img_Dict(Dd, size, False)
    save("cli_render_exact", **{"d_" + k: v for k, v in Dd.items()},
         Season_Adj_Img=imgs["Season_Adj_Img"], Shadow_Adjust_Exact=imgs["Shadow_Adjust_Exact"],
         Shadow_Mask_Exact=imgs["Shadow_Mask_Exact"], Raw_Shadow_Mask_Exact=imgs["Raw_Shadow_Mask_Exact"])

    # ---- F: Quick_Run (engine render convention) + engine exact solar -----------
    qargs = SimpleNamespace(**vars(args))
    q = ref.Quick_Run_Net(net, qargs, so.OMA_W2C, W2L, t.device("cpu"), use_full_solar=False)
    img, mask = q.render_img([75, 20], [50, 120], 0.4, 6)
    dsm = q.get_DSM((4, 4))
    d2 = rays(2, 51)
    with t.no_grad():
        Rx = tool.eval_exact_solar(d2, net, -1, False)
    save("quick_run", Col_Img=img["Col_Img"], Shadow_Mask=img["Shadow_Mask"], mask=mask, dsm=dsm,
         **{"in_" + k: v for k, v in d2.items()}, ex_Rendered_Col=Rx["Rendered_Col"], ex_Solar_Vis=Rx["Solar_Vis"],
         ex_Est_Solar_Vis=Rx["Est_Solar_Vis"])

    # ---- G: geometry ------------------------------------------------------------
    ang = np.array([[80, 0], [45, 135], [20, 270], [89, 10], [1, -170]], dtype=np.float64)
    vecs = np.array([ref.world_angle_2_local_vec(a[0], a[1], so.OMA_W2C, W2L) for a in ang])
    np.savez_compressed(os.path.join(OUT, "geometry.npz"), ang=ang, vecs=vecs, W2C=so.OMA_W2C, W2L_H=W2L)
    print("done")


if __name__ == "__main__":
    main()
