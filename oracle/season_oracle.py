"""CPU oracle for the Season-NeRF render / train hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
module: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may.  It is a *restatement* (own
code, torch-CPU float32 tensor arithmetic, no autograd tricks) of the
reference algorithm; every function cites the reference file:line it follows
(paths relative to the upstream repo root).

Parity pinning: the upstream repo holds no tests or golden vectors
(SURVEY.md section 4), so this oracle is pinned against OUTPUTS OF THE
REFERENCE ITSELF: ``oracle/make_golden.py`` imports the unmodified reference
(with its absent optional dependencies stubbed) in the build container, runs
it on seeded synthetic inputs and stores small fixtures under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file against
those fixtures.  The one exception is the Barron adaptive loss
(``oracle/barron_loss.py``): the third-party package is absent, so that term
is "parity unpinned" (restated from the paper).

Model weights are a plain ``dict[str, torch.Tensor]`` carrying exactly the 94
``state_dict`` keys of the reference ``T_NeRF`` module (T_NeRF_net_v2.py:20-68,
G_NeRF.py:6-71), so weights move freely between the reference, this oracle and
the CUDA product.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch as t

PI_HALF_F32 = float(np.float32(np.pi / 2))
OMEGA_0 = 30.0
BN_EPS = 1e-5
BN_MOMENTUM = 0.01

TRUNK = ["fc1", "fc2", "fc3", "fc4", "fc5", "fc6", "fc7", "fc8", "fc9"]


# --------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------
def _sine_init(g, out_f, in_f, is_first):
    """misc.py:178-186 (SineLayer.init_weights) + nn.Linear default bias."""
    if is_first:
        bound = 1.0 / in_f
    else:
        bound = math.sqrt(6.0 / in_f) / OMEGA_0
    w = (t.rand(out_f, in_f, generator=g) * 2 - 1) * bound
    bb = 1.0 / math.sqrt(in_f)
    b = (t.rand(out_f, generator=g) * 2 - 1) * bb
    return w, b


def _linear_init(g, out_f, in_f):
    bb = 1.0 / math.sqrt(in_f)
    w = (t.rand(out_f, in_f, generator=g) * 2 - 1) * bb
    b = (t.rand(out_f, generator=g) * 2 - 1) * bb
    return w, b


def layer_table(lw=512, n_classes=4):
    """(key prefix, out, in, kind) for every layer of T_NeRF.
    kind: 'first' | 'sine' | 'sine_bn' | 'linear'.  G_NeRF.py:41-63,
    T_NeRF_net_v2.py:36-52."""
    lw2, lw4 = max(lw // 2, 1), max(lw // 4, 1)
    g = "G_NeRF_net."
    tab = [
        (g + "fc1", lw, 63, "first"),
        (g + "fc2", lw, lw, "sine_bn"), (g + "fc3", lw, lw, "sine_bn"), (g + "fc4", lw, lw, "sine_bn"),
        (g + "fc5", lw, lw + 63, "sine_bn"),
        (g + "fc6", lw, lw, "sine_bn"), (g + "fc7", lw, lw, "sine_bn"), (g + "fc8", lw, lw, "sine_bn"),
        (g + "fc9", lw2, lw, "sine_bn"),
        (g + "fc10Col", 3, lw2, "linear"), (g + "fc10Sigma", 1, lw2, "linear"),
        (g + "fc_solar_1", lw2, lw2 + 27, "first"), (g + "fc_solar_2", lw2, lw2, "sine"),
        (g + "fc_solar_3", lw2, lw2, "sine"), (g + "fc_solar_4", 1, lw2, "linear"),
        (g + "fc_sky_color_1", lw4, 27, "first"), (g + "fc_sky_color_2", 3, lw4, "linear"),
        ("time_layer_1", lw, 10, "first"), ("time_layer_2", lw, lw, "sine"),
        ("get_class_layer", n_classes, lw, "linear"),
        ("adjust_layer_1", lw, lw2, "sine"), ("adjust_layer_2", lw, lw, "sine"),
        ("adjust_layer_3", lw, lw, "sine"),
        ("adjust_col", n_classes * 3, lw, "linear"), ("adjust_rho", n_classes, lw, "linear"),
        ("adjust_solar_vis", n_classes, lw, "linear"), ("adjust_sky_col", n_classes * 3, lw, "linear"),
    ]
    return tab


def init_params(seed=0, lw=512, n_classes=4, perturb_bn=False):
    """Random-init weights with the reference's distributions (not its RNG
    stream); same 94 keys/shapes as T_NeRF(lw, n_classes).state_dict()."""
    g = t.Generator().manual_seed(seed)
    p = {}
    for name, o, i, kind in layer_table(lw, n_classes):
        if kind == "linear":
            w, b = _linear_init(g, o, i)
            p[name + ".weight"], p[name + ".bias"] = w, b
        else:
            w, b = _sine_init(g, o, i, kind == "first")
            p[name + ".linear.weight"], p[name + ".linear.bias"] = w, b
            if kind == "sine_bn":
                p[name + ".norm.weight"] = t.ones(o)
                p[name + ".norm.bias"] = t.zeros(o)
                p[name + ".norm.running_mean"] = t.zeros(o)
                p[name + ".norm.running_var"] = t.ones(o)
                p[name + ".norm.num_batches_tracked"] = t.tensor(0, dtype=t.long)
                if perturb_bn:
                    p[name + ".norm.weight"] = 1 + 0.2 * (t.rand(o, generator=g) - 0.5)
                    p[name + ".norm.bias"] = 0.2 * (t.rand(o, generator=g) - 0.5)
                    p[name + ".norm.running_mean"] = 0.5 * (t.rand(o, generator=g) - 0.5)
                    p[name + ".norm.running_var"] = 0.5 + t.rand(o, generator=g)
    return p


# --------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------
def pe_encode(X, n, extended=True):
    """misc.py:105-139 PE_Encode.  out = [X | per dim d: cos(k_j x_d) j<n,
    sin(k_j x_d) j<n], k_j = 2^j * fl32(pi/2)."""
    k = (2 ** t.arange(0, n)).to(t.float32) * PI_HALF_F32  # misc.py:109 (int64 * python float -> f32)
    arg = X.unsqueeze(2) * k.reshape(1, 1, n)               # [M, D, n]   fl32(k_j * x_d)
    enc = t.stack([t.cos(arg), t.sin(arg)], 2)              # [M, D, 2, n]  misc.py:130-132
    enc = enc.reshape(X.shape[0], -1)
    if extended:
        enc = t.cat([X, enc], 1)                            # misc.py:118-120
    return enc


def batch_norm(x, p, name, training):
    """nn.BatchNorm1d(momentum=0.01, eps=1e-5), misc.py:169-170.  Train mode
    normalises with biased batch variance and updates running stats with the
    unbiased one (torch semantics), also under no_grad (G_NeRF.py:141-145)."""
    w, b = p[name + ".weight"], p[name + ".bias"]
    if training:
        mean = x.mean(0)
        var = x.var(0, unbiased=False)
        with t.no_grad():
            n = x.shape[0]
            p[name + ".running_mean"] = (1 - BN_MOMENTUM) * p[name + ".running_mean"] + BN_MOMENTUM * mean.detach()
            p[name + ".running_var"] = (1 - BN_MOMENTUM) * p[name + ".running_var"] + \
                BN_MOMENTUM * var.detach() * (n / max(n - 1, 1))
            p[name + ".num_batches_tracked"] = p[name + ".num_batches_tracked"] + 1
    else:
        mean, var = p[name + ".running_mean"], p[name + ".running_var"]
    return (x - mean) / t.sqrt(var + BN_EPS) * w + b


def sine_layer(x, p, name, training):
    """misc.py:188-189: sin(norm(omega_0 * linear(x)))."""
    z = OMEGA_0 * t.nn.functional.linear(x, p[name + ".linear.weight"], p[name + ".linear.bias"])
    if (name + ".norm.weight") in p:
        z = batch_norm(z, p, name + ".norm", training)
    return t.sin(z)


def linear(x, p, name):
    return t.nn.functional.linear(x, p[name + ".weight"], p[name + ".bias"])


# --------------------------------------------------------------------------
# network (G_NeRF.py / T_NeRF_net_v2.py)
# --------------------------------------------------------------------------
def encode_x(p, X, training):
    """G_NeRF.py:80-91 _encode_X."""
    g = "G_NeRF_net."
    enc = pe_encode(X, 10)
    h = sine_layer(enc, p, g + "fc1", training)
    h = sine_layer(h, p, g + "fc2", training)
    h = sine_layer(h, p, g + "fc3", training)
    h = sine_layer(h, p, g + "fc4", training)
    h = sine_layer(t.cat([h, enc], 1), p, g + "fc5", training)
    h = sine_layer(h, p, g + "fc6", training)
    h = sine_layer(h, p, g + "fc7", training)
    h = sine_layer(h, p, g + "fc8", training)
    return sine_layer(h, p, g + "fc9", training)


def forward_position(p, X, training):
    """G_NeRF.py:93-98."""
    x1 = encode_x(p, X, training)
    return x1, linear(x1, p, "G_NeRF_net.fc10Sigma"), linear(x1, p, "G_NeRF_net.fc10Col")


def forward_solar_heads(p, x_enc, sun, training):
    """G_NeRF.py:100-122 forward_Solar (ignore_hue False)."""
    g = "G_NeRF_net."
    se = pe_encode(sun, 4)
    a = sine_layer(t.cat([x_enc, se], 1), p, g + "fc_solar_1", training)
    a = sine_layer(a, p, g + "fc_solar_2", training)
    a = sine_layer(a, p, g + "fc_solar_3", training)
    vis = linear(a, p, g + "fc_solar_4")
    sky = linear(sine_layer(se, p, g + "fc_sky_color_1", training), p, g + "fc_sky_color_2")
    return vis, sky


def time_classes(p, Time, training=False):
    """T_NeRF_net_v2.py:72-73,77-78,160-163 (only Time[:,0:2] is used)."""
    te = pe_encode(Time[:, 0:2], 2)
    h = sine_layer(sine_layer(te, p, "time_layer_1", training), p, "time_layer_2", training)
    return t.softmax(linear(h, p, "get_class_layer"), 1)


def adjust_branch(p, x_enc, n_classes, training):
    """T_NeRF_net_v2.py:80-87."""
    y = sine_layer(x_enc, p, "adjust_layer_1", training)
    y = sine_layer(y, p, "adjust_layer_2", training)
    y = sine_layer(y, p, "adjust_layer_3", training)
    return linear(y, p, "adjust_col").reshape(x_enc.shape[0], n_classes, -1)


def _link(p, X, sun, Time, training):
    n_classes = p["get_class_layer.weight"].shape[0]
    x_enc, rho, col = forward_position(p, X, training)        # G_NeRF.py:130-133
    vis, sky = forward_solar_heads(p, x_enc, sun, training)
    cls = time_classes(p, Time, training)
    adj = adjust_branch(p, x_enc, n_classes, training)
    return rho, col, vis, sky, cls, adj


def forward(p, X, sun, Time, training=False):
    """T_NeRF.forward, T_NeRF_net_v2.py:75-105."""
    rho, col, vis, sky, cls, adj = _link(p, X, sun, Time, training)
    adjust_col = t.sum(adj * cls.unsqueeze(2), 1)
    return (t.nn.functional.softplus(rho), t.sigmoid(col + adjust_col), t.sigmoid(vis),
            t.sigmoid(sky), cls, adjust_col)


def forward_seperate(p, X, sun, Time, training=False):
    """T_NeRF.forward_seperate, T_NeRF_net_v2.py:131-151 (raw col, unmixed Adj)."""
    rho, col, vis, sky, cls, adj = _link(p, X, sun, Time, training)
    return t.nn.functional.softplus(rho), col, t.sigmoid(vis), t.sigmoid(sky), cls, adj


forward_full_eval = forward_seperate  # T_NeRF_net_v2.py:184-204 returns the same tuple


def approx_solar(p, X, X_solar, Time, training=False):
    """T_NeRF.approx_Solar, T_NeRF_net_v2.py:107-128: one trunk pass over [X; X_solar], adjust branch on the X part."""
    n = X.shape[0]
    n_classes = p["get_class_layer.weight"].shape[0]
    x_enc, rho, col = forward_position(p, t.cat([X, X_solar], 0), training)
    cls = time_classes(p, Time, training)
    adj = adjust_branch(p, x_enc[0:n], n_classes, training)
    adjust_col = t.sum(adj * cls.unsqueeze(2), 1)
    rho = t.nn.functional.softplus(rho)
    return rho[0:n], rho[n:], t.sigmoid(col[0:n] + adjust_col), cls, adjust_col


def forward_solar(p, X, sun, Time=None, training=False):
    """T_NeRF.forward_Solar, T_NeRF_net_v2.py:154-157 + G_NeRF.py:141-145:
    trunk + sigma head under no_grad, solar/sky heads with grad, RAW sky."""
    with t.no_grad():
        x_enc, rho, _ = forward_position(p, X, training)
    vis, sky = forward_solar_heads(p, x_enc, sun, training)
    return t.nn.functional.softplus(rho), t.sigmoid(vis), sky


def forward_sigma_only(p, X, training=False):
    """T_NeRF.forward_Classic_Sigma_Only, T_NeRF_net_v2.py:169-170; G_NeRF.py:74-77."""
    return t.nn.functional.softplus(linear(encode_x(p, X, training), p, "G_NeRF_net.fc10Sigma"))


def forward_color_only(p, X, training=False):
    """G_NeRF.py:154-157."""
    return t.sigmoid(linear(encode_x(p, X, training), p, "G_NeRF_net.fc10Col"))


def supervised_sample(hm, world_pts, delta):
    """T_NeRF.Supervised_Sample, T_NeRF_net_v2.py:175-181."""
    hm = t.as_tensor(hm)
    hm_const = t.tensor(hm.shape).reshape(1, 2) - 1
    xy = ((world_pts[:, 0:2] + 1) / 2 * hm_const).long()
    prob = (hm[xy[:, 0], xy[:, 1]] >= world_pts[:, 2]).float()
    prob[prob > .99] = 0.99
    return -t.log(1 - prob.unsqueeze(1)) / delta


# --------------------------------------------------------------------------
# sampling / transmittance (misc.py, Eval_Tools_2.py)
# --------------------------------------------------------------------------
def sample_ts(n, eval_mode, include_end_pt=False, jitter=None):
    """misc.py:236-241.  ``jitter`` replaces t.rand(n): ONE vector per call
    shared by every ray (SURVEY 8a' item 1)."""
    if include_end_pt is False or eval_mode is False:
        ts = t.linspace(0, 1, n + 1)[0:-1].clone()
    else:
        ts = t.linspace(0, 1, n)
    if eval_mode is False:
        if jitter is None:
            jitter = t.rand(n)
        ts = ts + 1 / n * jitter
    return ts


def sample_pt_coarse(tops, bots, n, eval_mode, include_end_pt=False, jitter=None):
    """misc.py:234-247."""
    ts = sample_ts(n, eval_mode, include_end_pt, jitter).reshape(1, -1, 1)
    deltas = t.sqrt(t.sum((tops - bots) ** 2, 1)) / n
    pts = tops.unsqueeze(1) * (1 - ts) + bots.unsqueeze(1) * ts
    deltas = deltas.reshape(-1, 1, 1) * t.ones(deltas.shape[0], n, 1, dtype=deltas.dtype)
    return pts, deltas


def invalid_pts(Xs):
    """misc.py:255-261 zero_invalid_pts.__call__: True where a point leaves [-1,1]^3."""
    return ~((Xs <= 1).all(-1) & (Xs >= -1).all(-1))


def get_PV(Rhos, Deltas):
    """Eval_Tools_2.py:13-16: exclusive-prefix transmittance."""
    Y = t.cat([t.zeros(Rhos.shape[0], 1, 1, dtype=Rhos.dtype), Rhos * Deltas], 1)
    return t.exp(-t.cumsum(Y, 1))[:, 0:-1]


# --------------------------------------------------------------------------
# geometry helpers (all_NeRF/mg_unit_converter.py)
# --------------------------------------------------------------------------
def world_angle_2_local_vec(world_el, world_az, world_center, W2L_H):
    """all_NeRF/mg_unit_converter.py:5-9,29-34,59-68 (float64 numpy)."""
    Y = np.cos(np.deg2rad(world_az))
    X = np.sin(np.deg2rad(world_az))
    Z = np.tan(np.deg2rad(world_el)) * np.sqrt(X ** 2 + Y ** 2)
    norm = np.sqrt(X ** 2 + Y ** 2 + Z ** 2) / 1000
    X, Y, Z = X / norm, Y / norm, Z / norm
    R = 6378.137
    dLat = Y / (1000. * R)
    dLon = X / (1000. * R * np.cos(np.deg2rad(world_center[0])))
    lla = np.array([world_center[0] + np.rad2deg(dLat), world_center[1] + np.rad2deg(dLon), world_center[2] + Z])
    temp = (np.asarray(W2L_H) @ np.array([[lla[0], lla[1], lla[2], 1]]).T)[0:-1]
    return (temp / np.sqrt(np.sum(temp ** 2))).T[0]


def encode_time(time_frac_year, time_frac_day=0):
    """T_NeRF_Full_2/Quick_Run.py:9-12."""
    return np.array([np.cos(time_frac_year * 2 * np.pi), np.sin(time_frac_year * 2 * np.pi),
                     np.cos(time_frac_day * 2 * np.pi), np.sin(time_frac_day * 2 * np.pi)])


def create_solar_rays_uniform(n, WC, W2L_H, np_rng=None, torch_gen=None):
    """Eval_Tools_2.py:72-108 create_solor_rays_uniform.__call__(n, True).
    The reference draws from the global numpy / torch RNGs; the oracle takes
    explicit generators (inputs are injected in parity tests anyway)."""
    np_rng = np_rng or np.random
    az_el = np_rng.random(n * 2).reshape([n, 2]) * np.array([[360, 89]]) + np.array([[-180, 1]])
    vec = np.array([world_angle_2_local_vec(az_el[i][1], az_el[i][0], WC, W2L_H) for i in range(n)])
    delta = 2 * (vec / vec[:, 2::])
    starts = t.ones(n, 3)
    starts[:, 0] = 2 * t.rand(n, generator=torch_gen) - 1
    starts[:, 1] = 2 * t.rand(n, generator=torch_gen) - 1
    ends = (starts - delta).float()
    vec = t.tensor(vec).float()
    f = t.rand(n, 2, generator=torch_gen) * 2 * np.pi
    times = t.stack([t.cos(f[:, 0]), t.sin(f[:, 0]), t.cos(f[:, 1]), t.sin(f[:, 1])], 1)
    return starts, ends, vec, times, az_el


def create_solar_rays_given_vec(n, solar_angle_vec, torch_gen=None):
    """create_solor_rays_uniform.create_given_vec(n, vec, include_times=True), Eval_Tools_2.py:50-70."""
    delta = 2 * (solar_angle_vec / solar_angle_vec[2::])
    starts = t.ones(n, 3)
    starts[:, 0] = 2 * t.rand(n, generator=torch_gen) - 1
    starts[:, 1] = 2 * t.rand(n, generator=torch_gen) - 1
    ends = (starts - np.expand_dims(delta, 0)).float()
    vec = t.stack([t.tensor(solar_angle_vec).float()] * n, 0)
    f = t.rand(n, 2, generator=torch_gen) * 2 * np.pi
    times = t.stack([t.cos(f[:, 0]), t.sin(f[:, 0]), t.cos(f[:, 1]), t.sin(f[:, 1])], 1)
    return starts, ends, vec, times


# --------------------------------------------------------------------------
# render / loss engine (T_NeRF_Full_2/Eval_Tools_2.py)
# --------------------------------------------------------------------------
def default_args(**kw):
    a = dict(n_samples=96, Use_Reg=True, Solar_Type_2=False, Use_MSE_loss=False, sc_lambda=0.03,
             Use_Solar=True, number_low_frequency_cases=4)
    a.update(kw)
    return SimpleNamespace(**a)


def engine_eval(args, data, p, current_step, train_mode, jitter=None, use_prior=False, n_steps=1, hm=None):
    """All_in_One_Eval.eval, Eval_Tools_2.py:165-252."""
    S = args.n_samples
    Xs, deltas = sample_pt_coarse(data["Top"], data["Bot"], S, not train_mode, jitter=jitter)
    N = Xs.shape[0]
    sun = (t.ones_like(Xs) * data["Sun_Angle"].unsqueeze(1)).reshape(-1, 3)
    tim = (t.ones(N, S, 4) * data["Time_Encoded"].unsqueeze(1)).reshape(-1, 4)
    Rho, Col, Vis, Sky, Cls, Adj = forward(p, Xs.reshape(-1, 3), sun, tim, training=train_mode)
    Col, Rho, Vis = Col.reshape(N, S, -1), Rho.reshape(N, S, 1), Vis.reshape(N, S, 1)
    Sky, Cls, Adj = Sky.reshape(N, S, -1), Cls.reshape(N, S, -1), Adj.reshape(N, S, -1)
    PV = get_PV(Rho, deltas)
    PE = 1 - t.exp(-Rho * deltas)
    PS = PV * PE
    Albedo = t.sum(PS * Col, 1)

    def shade(ps, albedo):
        if args.Solar_Type_2:
            return t.sum(ps * Col * (Vis + (1 - Vis) * Sky), 1)
        return albedo * (SV3 + (1 - SV3) * t.mean(Sky, 1))

    SV3 = None if args.Solar_Type_2 else t.sigmoid((t.sum(Vis.detach() * PS, 1) - .2) * 30)
    Rendered = shade(PS, Albedo)
    R = {"Rendered_Col": Rendered, "PE": PE, "PV": PV, "PS": PS, "Solar_Vis": Vis, "Sky_Col": Sky,
         "Classes": Cls, "Adjust": Adj, "Rho": Rho, "Col": Col, "Col_Adj": -1, "deltas": deltas,
         "sample_pts": Xs, "Albedo_Color": Albedo}
    if use_prior:
        trust = current_step / n_steps
        Rho_S = supervised_sample(hm, Xs.reshape(-1, 3), deltas.reshape(-1, 1)).reshape(N, S, 1)
        PV_S = get_PV(Rho_S, deltas)
        PE_S = 1 - t.exp(-Rho_S * deltas)
        PS_S = PV_S * PE_S
        Rend_S = shade(PS_S, t.sum(PS_S * Col, 1))
        Rho_M = Rho * trust + Rho_S * (1 - trust)
        PV_M = get_PV(Rho_M, deltas)
        PE_M = 1 - t.exp(-Rho_M * deltas)
        PS_M = PV_M * PE_M
        Albedo = t.sum(PS_M * Col, 1)
        Rend_M = shade(PS_M, Albedo)
        R.update({"PV_Supervised": PV_S, "PE_Supervised": PE_S, "PS_Supervised": PS_S,
                  "Rendered_Col_Supervised": Rend_S, "PV_Merged": PV_M, "PE_Merged": PE_M,
                  "PS_Merged": PS_M, "Rendered_Col_Merged": Rend_M, "Rho_Merged": Rho_M,
                  "Albedo_Color": Albedo})
    return R


def engine_full_eval(args, data, p):
    """All_in_One_Eval.full_eval, Eval_Tools_2.py:127-163: eval-mode sampling, `forward` (activated colour), and the
    colour passes through Sigmoid a second time inside the compositing (:155,158)."""
    S = args.n_samples
    Xs, deltas = sample_pt_coarse(data["Top"], data["Bot"], S, True)
    N = Xs.shape[0]
    sun = (t.ones_like(Xs) * data["Sun_Angle"].unsqueeze(1)).reshape(-1, 3)
    tim = (t.ones(N, S, 4) * data["Time_Encoded"].unsqueeze(1)).reshape(-1, 4)
    Rho, Col, Vis, Sky, Cls, Adj = forward(p, Xs.reshape(-1, 3), sun, tim)
    Col, Rho, Vis = Col.reshape(N, S, -1), Rho.reshape(N, S, 1), Vis.reshape(N, S, 1)
    Sky, Cls, Adj = Sky.reshape(N, S, -1), Cls.reshape(N, S, -1), Adj.reshape(N, S, -1)
    PV = get_PV(Rho, deltas)
    PE = 1 - t.exp(-Rho * deltas)
    PS = PV * PE
    if args.Solar_Type_2:
        Rendered = t.sum(PS * t.sigmoid(Col) * (Vis + (1 - Vis) * Sky), 1)
    else:
        SV3 = t.sigmoid((t.sum(Vis.detach() * PS, 1) - .2) * 30)
        Rendered = t.sum(PS * t.sigmoid(Col), 1) * (SV3 + (1 - SV3) * t.mean(Sky, 1))
    return {"Rendered_Col": Rendered, "PE": PE, "PV": PV, "PS": PS, "Solar_Vis": Vis, "Sky_Col": Sky, "Classes": Cls,
            "Adjust": Adj, "Rho": Rho, "Col": Col, "deltas": deltas, "sample_pts": Xs}


def eval_rho_only(args, data, p, train_mode, current_step=0, jitter=None, use_prior=False, n_steps=1, hm=None):
    """All_in_One_Eval.eval_Rho_Only, Eval_Tools_2.py:297-337."""
    S = args.n_samples
    Xs, deltas = sample_pt_coarse(data["Top"], data["Bot"], S, not train_mode, include_end_pt=True, jitter=jitter)
    N = Xs.shape[0]
    sun = (t.ones_like(Xs) * data["Sun_Angle"].unsqueeze(1)).reshape(-1, 3)
    Rho, Vis, Sky = forward_solar(p, Xs.reshape(-1, 3), sun, None, training=train_mode)
    Rho, Vis, Sky = Rho.reshape(N, S, 1), Vis.reshape(N, S, 1), Sky.reshape(N, S, -1)
    if use_prior:
        trust = current_step / n_steps
        Xs2, d2 = Xs.reshape(-1, 3), deltas.reshape(-1, 1)
        good = t.all((Xs2 <= 1.) * (Xs2 >= -1.), 1)
        Rho_S = Rho.reshape(-1, 1).detach().clone()
        Rho_S[good] = supervised_sample(hm, Xs2[good], d2[good])
        Rho_M = Rho * trust + Rho_S.reshape(N, S, 1) * (1 - trust)
        return {"PE": 1 - t.exp(-Rho_M * deltas), "PV_Exact": get_PV(Rho_M, deltas), "Solar_Vis": Vis, "Sky_Col": Sky}
    return {"PE": 1 - t.exp(-Rho * deltas), "PV_Exact": get_PV(Rho, deltas), "Solar_Vis": Vis, "Sky_Col": Sky}


def get_exact_solar(args, world_pts, sun_angle, p):
    """All_in_One_Eval._get_exact_solar, Eval_Tools_2.py:255-269."""
    sun_ext = t.stack([sun_angle] * world_pts.shape[0], 0)
    K = (1 - world_pts[:, 2]) / sun_angle[2]
    tops = world_pts + K.unsqueeze(1) * sun_ext
    r = eval_rho_only(args, {"Top": tops, "Bot": world_pts, "Sun_Angle": sun_ext}, p, False)
    return r["PV_Exact"][:, -1], r["Solar_Vis"][:, -1]


def eval_exact_solar(args, data, p, current_step=-1):
    """All_in_One_Eval.eval_exact_solar, Eval_Tools_2.py:273-295 (eval mode)."""
    R = engine_eval(args, data, p, current_step, False)
    R["Est_Solar_Vis"] = R["Solar_Vis"].clone()
    for i in range(data["Sun_Angle"].shape[0]):
        exact, _ = get_exact_solar(args, R["sample_pts"][i], data["Sun_Angle"][i], p)
        R["Solar_Vis"][i] = exact
    R["Col_Adj"] = (R["Solar_Vis"] + (1 - R["Solar_Vis"]) * R["Sky_Col"]) * R["Col"]
    if args.Solar_Type_2:
        R["Rendered_Col"] = t.sum(R["PS"] * R["Col"] * (R["Solar_Vis"] + (1 - R["Solar_Vis"]) * R["Sky_Col"]), 1)
    else:
        sv3 = t.sigmoid((t.sum(R["Solar_Vis"] * R["PS"], 1) - .2) * 30)
        R["Rendered_Col"] = t.sum(R["PS"] * R["Col"], 1) * (sv3 + (1 - sv3) * t.mean(R["Sky_Col"], 1))
    return R


def get_loss(args, data, p, current_step, train_mode, ada_loss, jitter=None, solar=None, solar_jitter=None,
             use_prior=False, n_steps=1, hm=None):
    """All_in_One_Eval.get_loss, Eval_Tools_2.py:340-459.  ``solar`` =
    (starts, ends, sun_vec, times) replaces the random solar-ray draw (:350);
    ``jitter`` / ``solar_jitter`` replace the two t.rand(S) draws."""
    Loss = {}
    w_sc = args.sc_lambda
    out = engine_eval(args, data, p, current_step, train_mode, jitter, use_prior, n_steps, hm)
    if args.Use_Solar:
        starts, ends, svec, stime = solar
        sol = eval_rho_only(args, {"Top": starts, "Bot": ends, "Sun_Angle": svec, "Time_Encoded": stime}, p,
                            train_mode, current_step, solar_jitter, use_prior, n_steps, hm)
        Loss["Solar_Correction"] = [t.mean(t.sum((sol["Solar_Vis"] - sol["PV_Exact"].detach()) ** 2, 1)), w_sc]
        absorb = t.mean(1 - t.sum(sol["PE"].detach() * sol["PV_Exact"].detach() * sol["Solar_Vis"], 1))
        Loss["Solar_Correction_2"] = [absorb if args.Solar_Type_2 else absorb.detach(), w_sc]
        if not args.Solar_Type_2:
            sk_alb, _ = t.min(out["Albedo_Color"], 0)
            sel = sk_alb[sk_alb < .2]
            alb_loss = t.sum((1. - sel / .2) ** 2) / out["Albedo_Color"].shape[0] if sel.shape[0] > 0 else t.tensor(0.0)
            sk = (out["Sky_Col"] - .5) / .5
            pos = sk[sk > 0]
            if pos.shape[0] > 0:
                sk_loss = t.sum(pos ** 2) / float(np.prod(sk.shape))
                if use_prior:
                    sk_loss = sk_loss.detach()
            else:
                sk_loss = t.tensor(0.)
            Loss["Sky_Color_Var"] = [sk_loss, w_sc]
            Loss["Albedo_Color"] = [alb_loss, w_sc]
    gt = data["GT_Color"]
    mse = t.nn.functional.mse_loss
    if args.Use_MSE_loss:
        key = "Rendered_Col_Merged" if (use_prior and train_mode) else "Rendered_Col"
        Loss["Color"] = [mse(out[key], gt), 1.0]
        if use_prior:
            Loss["Alpha_Adjust"] = [mse(out["PE"], out["PE_Supervised"].detach()), 1.]
    else:
        diff = out["Rendered_Col"] - gt
        if use_prior:
            a0, a1 = ada_loss
            adiff = (out["PE"] - out["PE_Supervised"].detach()).reshape(-1, 1)
            Loss["Alpha_Adjust_ada"] = [t.mean(a1.lossfun(adiff)), 1.]
            Loss["Color_ada"] = [t.mean(a0.lossfun(diff)), 1.0]
            Loss["Color_alpha"] = [t.mean(a0.alpha().detach()), 1.]
            Loss["Color_width"] = [t.mean(a0.scale().detach()), 1.]
            Loss["Alpha_Adjust"] = [mse(out["PE"], out["PE_Supervised"].detach()), 1.]
            scale = t.mean(a0.scale().detach()) ** 2
            Loss["Solar_Correction"][1] = Loss["Solar_Correction"][1] / scale
            Loss["Solar_Correction_2"][1] = Loss["Solar_Correction_2"][1] / scale
            Loss["Alpha_alpha"] = [t.mean(a1.alpha().detach()), 1.]
            Loss["Alpha_width"] = [t.mean(a1.scale().detach()), 1.]
        else:
            Loss["Color_ada"] = [t.mean(ada_loss.lossfun(diff)), 1.0]
            Loss["Color_alpha"] = [t.mean(ada_loss.alpha().detach()), 1.]
            Loss["Color_width"] = [t.mean(ada_loss.scale().detach()), 1.]
            scale = t.mean(ada_loss.scale().detach()) ** 2
            Loss["Solar_Correction"][1] = Loss["Solar_Correction"][1] / scale
            Loss["Solar_Correction_2"][1] = Loss["Solar_Correction_2"][1] / scale
        with t.no_grad():
            key = "Rendered_Col_Merged" if (use_prior and train_mode) else "Rendered_Col"
            Loss["Color"] = [mse(out[key], gt).detach(), 1.0]
    return Loss, out


def total_loss(Loss):
    """mg_run_NeRF.py:293-306: sum of value * weight."""
    tot = 0
    for k in Loss:
        tot = tot + Loss[k][0] * Loss[k][1]
    return tot


# --------------------------------------------------------------------------
# CLI render path (T_NeRF_Eval_Utils/mg_Img_Eval.py)
# --------------------------------------------------------------------------
def internal_render(p, tops, bots, sun_vec, year_frac, out_img_size, max_batch_size=150000, include_exact_solar=False):
    """_internal_render, mg_Img_Eval.py:17-72 (float64 result arrays)."""
    C = p["get_class_layer.weight"].shape[0]
    S = out_img_size[2]
    N = tops.shape[0]
    keys = ["World_Points", "Deltas", "Rho", "Base_Col", "Est_Solar_Vis", "Sky_Col", "Output_class"]
    last = [3, 1, 1, 3, 1, 3, C]
    R = {k: np.zeros([N, S, d]) for k, d in zip(keys, last)}
    R["Adjust_col"] = np.zeros([N, S, C, 3])
    if include_exact_solar:
        R["Exact_Solar"] = np.zeros([N, S, 1])
    step = max_batch_size // S if not include_exact_solar else max_batch_size // (S ** 2)
    sun_vec = np.asarray(sun_vec)
    sun_t = t.tensor(sun_vec).float()
    time_t = t.tensor(encode_time(year_frac)).float()
    with t.no_grad():
        for i in range(0, N, step):
            e = min(i + step, N)
            pts, deltas = sample_pt_coarse(tops[i:e], bots[i:e], S, eval_mode=True, include_end_pt=True)
            deltas[invalid_pts(pts)] = 0.
            M = pts.shape[0] * S
            outs = forward_seperate(p, pts.reshape(-1, 3), sun_t.reshape(1, 3).expand(M, 3),
                                    time_t.reshape(1, 4).expand(M, 4))
            for k, o in zip(keys[2:7], outs[:5]):
                R[k][i:e] = o.reshape(pts.shape[0], S, -1).numpy()
            R["Adjust_col"][i:e] = outs[5].reshape(pts.shape[0], S, C, -1).numpy()
            R["World_Points"][i:e] = pts.numpy()
            R["Deltas"][i:e] = deltas.numpy()
            if include_exact_solar:
                nb = pts.reshape(-1, 3)
                Sf = (1. - nb[:, 2]) / sun_vec[2]                                    # mg_Img_Eval.py:58-60
                nt = (nb + Sf.reshape(-1, 1) * sun_vec.reshape(1, -1)).float()       # float64 -> float32
                npts, nd = sample_pt_coarse(nt, nb, S, eval_mode=True, include_end_pt=True)
                nd[invalid_pts(npts)] = 0.
                rhos = forward_sigma_only(p, npts.reshape(-1, 3)).reshape(nt.shape[0], S, 1)
                pv = t.exp(-t.sum((rhos * nd)[:, 0:-1, :], 1)).reshape(pts.shape[0], S, 1)
                R["Exact_Solar"][i:e] = pv.numpy()
    return R


def component_render_by_dir(p, view_el_az, sun_el_az, time_frac, out_img_size, W2C, W2L_H,
                            max_batch_size=150000, include_exact_solar=True):
    """mg_Img_Eval.py:96-115."""
    H, W = out_img_size[0], out_img_size[1]
    XYZ = np.stack(np.meshgrid(np.linspace(1, -1, H), np.linspace(-1, 1, W), indexing="ij"), -1).reshape([-1, 2])
    XYZ = np.concatenate([XYZ, np.zeros([XYZ.shape[0], 1])], 1)
    vv = world_angle_2_local_vec(view_el_az[0], view_el_az[1], W2C, W2L_H)
    sv = world_angle_2_local_vec(sun_el_az[0], sun_el_az[1], W2C, W2L_H)
    tops = t.tensor(XYZ + np.expand_dims(vv / vv[2], 0)).float()
    bots = t.tensor(XYZ - np.expand_dims(vv / vv[2], 0)).float()
    R = internal_render(p, tops, bots, sv, time_frac, out_img_size, max_batch_size, include_exact_solar)
    R["Image_Points"] = np.stack(np.meshgrid(np.arange(H), np.arange(W), indexing="ij"), -1).reshape([-1, 2])
    return R


def _sig(X):
    return 1 / (1 + np.exp(-X))


def _ps_f64(D):
    pv = get_PV(t.tensor(D["Rho"]), t.tensor(D["Deltas"])).numpy()
    return pv * (1 - np.exp(-D["Rho"] * D["Deltas"]))


def _scatter(D, size, vals):
    img = np.zeros([size[0], size[1]] + list(vals.shape[1:])) * np.nan
    img[D["Image_Points"][:, 0], D["Image_Points"][:, 1]] = vals
    return img


def get_imgs_from_img_dict(D, size, use_classic_shadows=False):
    """get_imgs_from_Img_Dict, mg_Img_Eval.py:123-190 (float64 numpy), both shadow conventions (the CLI takes
    use_classic_shadows=False; True = :166-181)."""
    sky = D["Sky_Col"][0, 0]
    cls = D["Output_class"][0, 0]
    PS = _ps_f64(D)
    base = _scatter(D, size, np.sum(PS * _sig(D["Base_Col"]), 1))
    raw = _scatter(D, size, np.sum(PS * D["Est_Solar_Vis"], 1)[:, 0])
    mask = _sig((raw - .2) * 30)
    adj = np.expand_dims(mask, -1) + np.expand_dims(1 - mask, -1) * sky.reshape(1, 1, 3)
    mixed = (np.expand_dims(D["Output_class"], 2) @ D["Adjust_col"])[:, :, 0, :]
    season = _scatter(D, size, np.sum(PS * _sig(D["Base_Col"] + mixed), 1))
    extreme = [_scatter(D, size, np.sum(PS * _sig(D["Base_Col"] + D["Adjust_col"][:, :, i]), 1))
               for i in range(D["Adjust_col"].shape[2])]
    R = {"Base_Img": base, "Season_Adj_Img": season, "Extreme_Imgs": extreme, "Shadow_Adjust": adj,
         "Shadow_Mask": mask, "Raw_Shadow_Mask": raw, "Sky_Col": sky, "Time_Class": cls}
    if "Exact_Solar" in D:
        raw_e = _scatter(D, size, np.sum(PS * D["Exact_Solar"], 1)[:, 0])
        mask_e = _sig((raw_e - .2) * 30)
        R["Shadow_Adjust_Exact"] = np.expand_dims(mask_e, -1) + np.expand_dims(1 - mask_e, -1) * sky.reshape(1, 1, 3)
        R["Shadow_Mask_Exact"] = mask_e
        R["Raw_Shadow_Mask_Exact"] = raw_e
    if use_classic_shadows:                                                    # mg_Img_Eval.py:166-181
        ip = D["Image_Points"]
        cols_season = np.sum(PS * _sig(D["Base_Col"] + mixed), 1)
        for key, name in (("Est_Solar_Vis", "Shadow_Adjust"), ("Exact_Solar", "Shadow_Adjust_Exact")):
            if key not in D:
                continue
            shadow_term = D[key] + (1 - D[key]) * D["Sky_Col"]
            classic = np.sum(PS * (_sig(D["Base_Col"] + mixed) * shadow_term), 1)
            R[name][ip[:, 0], ip[:, 1]] = classic / (cols_season + 1e-8)
    return R


def get_imgs_from_img_dict_t_step(D, size, class_vecs):
    """get_imgs_from_Img_Dict_t_step, mg_Img_Eval.py:192-228."""
    sky = D["Sky_Col"][0, 0]
    PS = _ps_f64(D)
    key = "Exact_Solar" if "Exact_Solar" in D else "Est_Solar_Vis"
    raw = _scatter(D, size, np.sum(PS * D[key], 1)[:, 0])
    mask = _sig((raw - .2) * 30)
    adj = np.expand_dims(mask, -1) + np.expand_dims(1 - mask, -1) * sky.reshape(1, 1, 3)
    imgs = []
    for i in range(class_vecs.shape[0]):
        mixed = (class_vecs[i].reshape(1, 1, -1) @ D["Adjust_col"])[:, :, 0, :]
        imgs.append(_scatter(D, size, np.sum(PS * _sig(D["Base_Col"] + mixed), 1)) * adj)
    return np.array(imgs)


def seasonal_align_v3_classic(p, D, target_img, t0):
    """T_NeRF_Eval_Utils/mg_Img_Eval.py:416-475 (_grad_descent_v3_classic_shadows): per-sample shading inside the sum."""
    ts = t.tensor([t0] + list(np.linspace(0, 1, 366))).float()
    ts_scaled = t.stack([t.cos(ts * 2 * np.pi), t.sin(ts * 2 * np.pi), t.cos(ts * 2 * np.pi), t.sin(ts * 2 * np.pi)], 1)
    with t.no_grad():
        tv = time_classes(p, ts_scaled)
        ip = np.asarray(D["Image_Points_in_GT_Img"])
        GT = t.tensor(np.asarray(target_img)[ip[:, 0], ip[:, 1]]).float()
        rho, dl = np.asarray(D["Rho"], dtype=np.float64), np.asarray(D["Deltas"], dtype=np.float64)
        PS = t.tensor(get_PV(t.tensor(rho), t.tensor(dl)).numpy() * (1 - np.exp(-rho * dl))).float()
        Base, Adj = t.tensor(np.asarray(D["Base_Col"])).float(), t.tensor(np.asarray(D["Adjust_col"])).float()
        SV = t.tensor(np.asarray(D["Est_Solar_Vis"])).float()
        scores, skies = np.ones(ts.shape[0]), np.zeros([ts.shape[0], 3])
        for i in range(ts.shape[0]):
            col = t.sigmoid(Base + t.sum(Adj * tv[i].reshape([1, 1, -1, 1]), 2))
            Y = GT - t.sum(PS * col * SV, 1)
            X = t.sum(PS * col * (1 - SV), 1)
            good = t.sum(X * X, 0) > 0
            sky = t.clamp((1 / t.sum(X * X, 0)[good]) * t.sum(X * Y, 0)[good], 0, 1)
            R = t.sum(PS * col * (SV + (1 - SV) * sky), 1)
            scores[i] = float(t.mean((R - GT) ** 2))
            skies[i] = sky.numpy()
        best = int(np.argmin(scores))
    return tv[best], t.tensor(skies[best]).reshape([1, 1, 3]).float(), ts[best].item(), scores


# --------------------------------------------------------------------------
# synthetic OMA_281-shaped inputs
# --------------------------------------------------------------------------
OMA_W2C = np.array([41.2905, -95.8967, 315.0])


def seasonal_align_v3(p, D, target_img, t0):
    """T_NeRF_Eval_Utils/mg_Img_Eval.py:354-414 (_grad_descent_v3): 367 candidate times, closed-form sky colour, MSE."""
    ts = t.tensor([t0] + list(np.linspace(0, 1, 366))).float()
    ts_scaled = t.stack([t.cos(ts * 2 * np.pi), t.sin(ts * 2 * np.pi), t.cos(ts * 2 * np.pi), t.sin(ts * 2 * np.pi)], 1)
    with t.no_grad():
        tv = time_classes(p, ts_scaled)
        ip = np.asarray(D["Image_Points_in_GT_Img"])
        GT = t.tensor(np.asarray(target_img)[ip[:, 0], ip[:, 1]]).float()
        rho, dl = np.asarray(D["Rho"], dtype=np.float64), np.asarray(D["Deltas"], dtype=np.float64)
        PS = get_PV(t.tensor(rho), t.tensor(dl)).numpy() * (1 - np.exp(-rho * dl))
        PS = t.tensor(PS).float()
        Base, Adj = t.tensor(np.asarray(D["Base_Col"])).float(), t.tensor(np.asarray(D["Adjust_col"])).float()
        SV = t.sigmoid((t.sum(PS * t.tensor(np.asarray(D["Est_Solar_Vis"])).float(), 1) - .2) * 30)
        good = (SV < .99)[:, 0]
        scores, skies = np.ones(ts.shape[0]), np.zeros([ts.shape[0], 3])
        for i in range(ts.shape[0]):
            A = t.sum(PS * t.sigmoid(Base + t.sum(Adj * tv[i].reshape([1, 1, -1, 1]), 2)), 1)
            Y = GT[good] - A[good] * SV[good]
            X = (1 - SV[good]) * A[good]
            sky = t.clamp((1 / t.sum(X * X, 0)) * t.sum(X * Y, 0), 0, 1)
            R = A * (SV + (1 - SV) * sky)
            scores[i] = float(t.mean((R - GT) ** 2))
            skies[i] = sky.numpy()
        best = int(np.argmin(scores))
    return tv[best], t.tensor(skies[best]).reshape([1, 1, 3]).float(), ts[best].item(), scores


def gen_results(p, img_shape, n_samples):
    """T_NeRF_Eval_Utils/Eval_funcs.py:268-296 (dense sigma / colour volume and its vertical compositing, float64 numpy)."""
    H, W, S = img_shape[0], img_shape[1], n_samples
    XYZ = np.stack(np.meshgrid(np.arange(0, H), np.arange(0, W), np.arange(S), indexing="ij"), -1).reshape([-1, 3])
    s = np.array([1 / H, 1 / W, 1 / S])
    xyz = XYZ * s * 2 - 1
    xyz[:, 2] *= -1
    X = t.tensor(xyz).float()
    with t.no_grad():
        Rho = forward_sigma_only(p, X, training=False)
        Col = forward_color_only(p, X, training=False)
    all_Rhos = Rho.double().numpy().reshape(H, W, S, 1)
    all_Cols = Col.double().numpy().reshape(H, W, S, 3)
    delta = 2 / S
    P_E = 1 - np.exp(-all_Rhos * delta)[:, :, :, 0]
    P_Vis = np.exp(-np.cumsum(np.concatenate([np.zeros([H, W, 1, 1]), all_Rhos * delta], 2), 2)[:, :, 0:-1])[:, :, :, 0]
    return all_Rhos, P_E, P_Vis, P_E * P_Vis, all_Cols


def height_map(p, shape, n_samples):
    """Eval_funcs.py:298-313: expected surface height in the normalised cube."""
    _, _, _, P_Surf, _ = gen_results(p, shape, n_samples)
    return np.sum(P_Surf * np.linspace(1, -1, n_samples).reshape([1, 1, -1]), 2) / np.sum(P_Surf, 2)


def synthetic_camera_P(seed=7):
    """A 3x4 projection shaped like the reference's affine-approximated RPC camera after scale_P (pre_NeRF/P_Img.py:168-201):
    pixel (row, col) of a 2048 x 2048 image from normalised scene coordinates in [-1,1]^3, off-nadir by a few degrees,
    with a small perspective row (the least-squares fit of compute_Approx_RPC, P_Img.py:331-371, is not exactly affine)."""
    g = np.random.RandomState(seed)
    P = np.array([[1180.0, 35.0, -140.0, 1023.5],
                  [-28.0, 1175.0, 95.0, 1023.5],
                  [0.0, 0.0, 0.0, 1.0]])
    P[0:2, 0:3] += g.uniform(-5, 5, (2, 3))
    P[2, 0:3] = g.uniform(-2e-3, 2e-3, 3)
    return P


def invert_P(P, row, col, h=0):
    """P_img_Pinhole.invert_P, pre_NeRF/P_Img.py:133-147 (float64 numpy, same evaluation order)."""
    row, col = np.asarray(row), np.asarray(col)
    A = P[1, 2] * h + P[1, 3] - P[2, 2] * h * col - P[2, 3] * col
    B = P[0, 2] * h + P[0, 3] - P[2, 2] * h * row - P[2, 3] * row
    P11mP31x = P[0, 0] - P[2, 0] * row
    P22mP32y = P[1, 1] - P[2, 1] * col
    P12mP32x = P[0, 1] - P[2, 1] * row
    P21mP31y = P[1, 0] - P[2, 0] * col
    den = P11mP31x * P22mP32y - P12mP32x * P21mP31y
    x = (P12mP32x * A - P22mP32y * B) / den
    y = (-P11mP31x * A + P21mP31y * B) / den
    return x, y, h


def camera_rays(P, rows, cols, z_top=1.0, z_bot=-1.0, bounds=(-1.0, 1.0, -1.0, 1.0)):
    """tops / bots (float64 [n,3]) and the inside-bounds mask of mg_Img_Eval.py:78-84 / mg_Pt_holder.py:176-187."""
    xt, yt, _ = invert_P(P, rows, cols, z_top)
    xb, yb, _ = invert_P(P, rows, cols, z_bot)
    tops = np.stack([xt, yt, np.full_like(xt, z_top)], -1)
    bots = np.stack([xb, yb, np.full_like(xb, z_bot)], -1)
    x0, x1, y0, y1 = bounds
    good = (xt <= x1) & (x0 <= xt) & (yt <= y1) & (y0 <= yt) & (xb <= x1) & (x0 <= xb) & (yb <= y1) & (y0 <= yb)
    return tops, bots, good


def oma_w2l_h():
    """diag scale mapping a ~0.0024 deg x 0.0032 deg x 70 m box to [-1,1]^3 (pre_NeRF/P_Img.py:168-176)."""
    H = np.eye(4)
    H[0, 0], H[1, 1], H[2, 2] = 2 / 0.0024, 2 / 0.0032, 2 / 70.0
    H[0, 3], H[1, 3], H[2, 3] = -OMA_W2C[0] * H[0, 0], -OMA_W2C[1] * H[1, 1], -OMA_W2C[2] * H[2, 2]
    return H


def synthetic_batch(n_rays, seed=1, n_images=41):
    """Training batch shaped like mg_run_NeRF.py:122-133 rows."""
    g = t.Generator().manual_seed(seed)
    xy = (t.rand(n_rays, 2, generator=g) * 2 - 1) * 0.8
    dxy = (t.rand(n_rays, 2, generator=g) * 2 - 1) * 0.2
    top = t.cat([xy, t.ones(n_rays, 1)], 1)
    bot = t.cat([xy + dxy, -t.ones(n_rays, 1)], 1)
    el = t.deg2rad(20 + 50 * t.rand(n_images, generator=g))
    az = 2 * math.pi * t.rand(n_images, generator=g)
    sun_img = t.stack([t.cos(el) * t.sin(az), t.cos(el) * t.cos(az), t.sin(el)], 1)
    f = t.rand(n_images, generator=g)
    time_img = t.stack([t.cos(2 * math.pi * f), t.sin(2 * math.pi * f),
                        t.full_like(f, math.cos(2 * math.pi * 0.70)), t.full_like(f, math.sin(2 * math.pi * 0.70))], 1)
    img = t.randint(0, n_images, (n_rays,), generator=g)
    return {"Top": top, "Bot": bot, "Sun_Angle": sun_img[img], "Time_Encoded": time_img[img],
            "GT_Color": t.rand(n_rays, 3, generator=g)}


# --------------------------------------------------------------------------
# sigma-only volume consumers (T_NeRF_Eval_Utils/mg_Shadow_Eval.py, Eval_funcs.py)
def eval_shadow_data(p, shadow_angles, ground_points, Z_points, WC, W2L_H):
    """eval_shadow_data, mg_Shadow_Eval.py:72-104 -> Vis_Exact [A,G,Z,1], Vis_Est [A,G,Z,1], Sky_Col [A,3] (float64)."""
    A, G = shadow_angles.shape[0], ground_points.shape[0]
    vec0 = np.array([world_angle_2_local_vec(shadow_angles[i, 0], shadow_angles[i, 1], WC, W2L_H) for i in range(A)])
    vec = vec0 / vec0[:, -1::]
    gp3 = np.expand_dims(np.concatenate([ground_points, np.zeros([G, 1])], 1), 0)
    tops = t.tensor(gp3 + np.expand_dims(vec, 1)).float()
    bots = t.tensor(gp3 - np.expand_dims(vec, 1)).float()
    ve, vs, sk = np.zeros([A, G, Z_points, 1]), np.zeros([A, G, Z_points, 1]), np.zeros([A, 3])
    with t.no_grad():
        for i in range(A):
            Xs, deltas = sample_pt_coarse(tops[i], bots[i], Z_points, eval_mode=True)
            deltas[invalid_pts(Xs)] = 0.
            sv = t.tensor(vec0[i]).float().reshape(1, 3).expand(G * Z_points, 3)
            rho, vis, sky = forward_solar(p, Xs.reshape(-1, 3), sv, None)
            rho = rho.reshape(G, Z_points, 1)
            ve[i] = get_PV(rho, deltas).numpy()
            vs[i] = vis.reshape(G, Z_points, 1).numpy()
            sk[i] = sky.reshape(G, Z_points, -1)[0, 0].numpy()
    return ve, vs, sk


def eval_hm_head(p, GT, h_range, n_samples):
    """eval_HM up to the scores before alignment, Eval_funcs.py:298-395 (the confidence loop :315-330 verbatim in structure)
    -> GT in metres, shifted estimated height map, conf_range [H,W,3], scores dict."""
    rho, pe, pv, ps, _ = gen_results(p, GT.shape, n_samples)
    est = np.sum(ps * np.linspace(1, -1, n_samples).reshape([1, 1, -1]), 2) / np.sum(ps, 2)
    pdf = ps / np.sum(ps, 2, keepdims=True)
    conf = np.zeros([pdf.shape[0], pdf.shape[1], 3])
    for i in range(conf.shape[0]):
        for j in range(conf.shape[1]):
            z0 = int(np.argmax(pdf[i, j]))
            z1 = z0 + 1
            value = pdf[i, j, z0]
            while value < .67 and (z0 != 0 or z1 != pdf.shape[2]):
                z0, z1 = max(0, z0 - 1), min(z1 + 1, pdf.shape[2])
                value = np.sum(pdf[i, j, z0:z1])
            conf[i, j, 0], conf[i, j, 1] = z0, z1
    conf[:, :, 2] = (conf[:, :, 1] - conf[:, :, 0]) / pdf.shape[2] * (h_range[1] - h_range[0])
    h0, h1 = h_range
    est = (est + 1) / 2 * (h1 - h0) + h0
    GTm = (GT + 1) / 2 * (h1 - h0) + h0
    est = est + np.nanmean((GTm - est).ravel())
    diff = est - GTm
    diff = np.ravel(diff[diff == diff])
    scores = {"MAE": np.mean(np.abs(diff)), "RMSE": np.sqrt(np.mean(diff ** 2)),
              "Acc_1_m": np.sum(np.abs(diff) <= 1) / diff.shape[0], "Median": np.median(np.abs(diff))}
    return GTm, est, conf, scores
