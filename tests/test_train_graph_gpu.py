"""The CUDA-graph training step (train.TrainStep(use_graph=True)) must be the eager step, replayed: same losses step by
step and the same weights / BatchNorm running statistics after a few steps (identical kernels and inputs; the only
non-determinism is the order of the split-K fp32 atomics of the weight-gradient GEMMs)."""
import types

import numpy as np
import pytest
import torch as t

from gpu_util import maxabs, relerr

pytestmark = pytest.mark.gpu
S = 96


def _args():
    return types.SimpleNamespace(n_samples=S, Use_Reg=True, Solar_Type_2=False, Use_MSE_loss=False, sc_lambda=0.03,
                                 Use_Solar=True, number_low_frequency_cases=4, fc_units=512, lr=10 ** -4.86,
                                 lr_alpha_scale=1000.0, max_train_steps=1000)


def _run(use_graph, n_steps, n=192, micro_batch=None):
    import season_nerf_b200 as snb
    from oracle import season_oracle as so
    dev = t.device("cuda")
    t.manual_seed(0)
    ts = snb.TrainStep(_args(), dev, so.oma_w2l_h(), so.OMA_W2C, use_graph=use_graph, graph_warmup=1, micro_batch=micro_batch)
    batch = so.synthetic_batch(n, seed=1, n_images=5)
    losses = []
    for i in range(n_steps):
        rs = np.random.RandomState(10 + i)
        g = t.Generator().manual_seed(10 + i)
        st, en, vec, tm, _ = so.create_solar_rays_uniform(n, so.OMA_W2C, so.oma_w2l_h(), rs, g)
        jit = t.rand(S, generator=g)
        L = ts.step(batch, i, jitter=jit, solar=(st, en, vec, tm), solar_jitter=jit)
        losses.append({k: float(v[0]) for k, v in L.items()} | {"total": float(ts.last_loss)})
    sd = {k: v.detach().float().cpu().clone() for k, v in ts.network.state_dict().items()}
    lr = float(ts.optim.param_groups[0]["lr"])
    return losses, sd, lr, ts


@pytest.mark.parametrize("micro_batch", [None, 64])
def test_graph_step_matches_eager(micro_batch):
    """micro_batch=64: 192 rays as 3 chunks whose gradients accumulate before one optimiser step (eager loop vs the
    accumulate-in-place graph replayed 3 times)."""
    n_steps = 5
    le, sde, lre, _ = _run(False, n_steps, micro_batch=micro_batch)
    lg, sdg, lrg, tsg = _run(True, n_steps, micro_batch=micro_batch)
    assert tsg._graphs and tsg.launches_replayed > 0, "the graph path did not run"
    assert abs(lre - lrg) < 1e-12 * max(1.0, abs(lre)) + 1e-9
    for a, b in zip(le, lg):
        for k in a:
            assert abs(a[k] - b[k]) <= 2e-3 * max(abs(a[k]), 1e-3), (k, a[k], b[k])
    for k in sde:
        if "num_batches_tracked" in k:
            assert t.equal(sde[k], sdg[k]), k
        elif sde[k].numel() > 1 and float(sde[k].norm()) > 0:
            # Adam moves every parameter by about lr per step whatever the gradient's magnitude: parameters that start at
            # zero (BatchNorm biases) are compared on that scale, the others relative to their norm
            assert relerr(sdg[k], sde[k]) < 2e-3 or maxabs(sdg[k], sde[k]) < n_steps * lre, (k, relerr(sdg[k], sde[k]))


def test_graph_step_accepts_host_batches_and_changes_inputs():
    """static buffers are refreshed before every replay: two different batches give two different losses, and the same
    batch replayed with the same draws gives (almost) the same loss as the first time modulo the weight update."""
    import season_nerf_b200 as snb
    from oracle import season_oracle as so
    dev = t.device("cuda")
    t.manual_seed(0)
    ts = snb.TrainStep(_args(), dev, so.oma_w2l_h(), so.OMA_W2C, use_graph=True, graph_warmup=1)
    b1 = so.synthetic_batch(128, seed=1, n_images=5)
    b2 = so.synthetic_batch(128, seed=2, n_images=5)
    np.random.seed(0)
    t.manual_seed(1)
    vals = []
    for i, b in enumerate([b1, b1, b2, b1, b2]):
        ts.step(b, i)
        vals.append(float(ts.last_loss))
    assert all(np.isfinite(v) for v in vals)
    assert abs(vals[2] - vals[3]) > 1e-6


@pytest.mark.parametrize("use_graph", [False, True])
def test_checkpoint_resume_continues_the_same_trajectory(use_graph):
    """TrainStep.state_dict(): network + both Adam optimisers + both OneCycle schedulers + adaptive-loss parameters; a run
    resumed from it produces the losses of the uninterrupted run."""
    import season_nerf_b200 as snb
    from oracle import season_oracle as so
    dev = t.device("cuda")
    batch = so.synthetic_batch(128, seed=1, n_images=5)

    def draws(i):
        rs, g = np.random.RandomState(50 + i), t.Generator().manual_seed(50 + i)
        st, en, vec, tm, _ = so.create_solar_rays_uniform(128, so.OMA_W2C, so.oma_w2l_h(), rs, g)
        jit = t.rand(S, generator=g)
        return dict(jitter=jit, solar=(st, en, vec, tm), solar_jitter=jit)

    t.manual_seed(0)
    a = snb.TrainStep(_args(), dev, so.oma_w2l_h(), so.OMA_W2C, use_graph=use_graph, graph_warmup=1)
    for i in range(3):
        a.step(batch, i, **draws(i))
    ck = {k: (v if not isinstance(v, dict) else v) for k, v in a.state_dict().items()}
    import copy
    ck = copy.deepcopy(ck)
    ref = []
    for i in range(3, 6):
        a.step(batch, i, **draws(i))
        ref.append(float(a.last_loss))
    t.manual_seed(123)                       # different initial weights: everything must come from the checkpoint
    b = snb.TrainStep(_args(), dev, so.oma_w2l_h(), so.OMA_W2C, use_graph=use_graph, graph_warmup=1)
    b.load_state_dict(ck)
    got = []
    for i in range(3, 6):
        b.step(batch, i, **draws(i))
        got.append(float(b.last_loss))
    for x, y in zip(ref, got):
        assert abs(x - y) <= 2e-3 * max(abs(x), 1e-3), (ref, got)


def test_graph_step_with_prior_dsm_matches_eager():
    """DSM-guided section (use_prior, Net_Tool_2.py:23-33,83-86): the trust factor step/n_steps changes every step and
    reaches the captured graph through a device scalar."""
    import season_nerf_b200 as snb
    from oracle import season_oracle as so
    dev = t.device("cuda")
    hm = (np.random.RandomState(5).rand(64, 64) * 1.2 - 0.6).astype(np.float32)
    batch = so.synthetic_batch(128, seed=1, n_images=5)

    def run(use_graph):
        t.manual_seed(0)
        ts = snb.TrainStep(_args(), dev, so.oma_w2l_h(), so.OMA_W2C, training_DSM=hm, use_prior=True, total_steps=20,
                           use_graph=use_graph, graph_warmup=1)
        out = []
        for i in range(5):
            rs, g = np.random.RandomState(70 + i), t.Generator().manual_seed(70 + i)
            st, en, vec, tm, _ = so.create_solar_rays_uniform(128, so.OMA_W2C, so.oma_w2l_h(), rs, g)
            jit = t.rand(S, generator=g)
            ts.step(batch, i, jitter=jit, solar=(st, en, vec, tm), solar_jitter=jit)
            out.append(float(ts.last_loss))
        return out, ts

    le, _ = run(False)
    lg, tsg = run(True)
    assert tsg._graphs and tsg.launches_replayed > 0
    assert len(set(round(x, 6) for x in le)) > 1                # the loss really changes with the trust factor / weights
    for a, b in zip(le, lg):
        assert abs(a - b) <= 2e-3 * max(abs(a), 1e-3), (le, lg)


@pytest.mark.parametrize("use_graph,micro_batch", [(False, None), (True, None), (True, 64)])
def test_device_solar_rng_steps_are_reproducible(use_graph, micro_batch):
    """TrainStep(solar_rng='device') (the default): the random solar rays of every step are drawn with torch's CUDA
    generator and built by the solar_rays kernel - nothing on the host.  Same seeds -> same loss trajectory; a different
    CUDA seed -> different solar rays -> a different solar loss; the batch may still arrive from the host."""
    import season_nerf_b200 as snb
    from oracle import season_oracle as so

    def run(cuda_seed):
        t.manual_seed(0)
        t.cuda.manual_seed(cuda_seed)
        ts = snb.TrainStep(_args(), t.device("cuda"), so.oma_w2l_h(), so.OMA_W2C, use_graph=use_graph, graph_warmup=1,
                           micro_batch=micro_batch)
        assert ts.solar_rng == "device" and ts.eval_tool.solar_on_device
        batch = so.synthetic_batch(192, seed=1, n_images=5)                    # host tensors
        out = []
        for i in range(4):
            jit = t.rand(S, generator=t.Generator().manual_seed(20 + i))
            L = ts.step(batch, i, jitter=jit, solar_jitter=jit)
            out.append((float(L["Solar_Correction"][0]), float(ts.last_loss)))
        return out

    a, b, c = run(11), run(11), run(12)
    assert all(np.isfinite(v) for pair in a for v in pair)
    for (sa, ta), (sb, tb) in zip(a, b):
        assert abs(sa - sb) <= 2e-3 * abs(sa) + 1e-6 and abs(ta - tb) <= 2e-3 * abs(ta) + 1e-6
    assert any(abs(x[0] - y[0]) > 1e-4 * abs(x[0]) for x, y in zip(a, c))


def test_host_solar_rng_still_follows_the_reference_streams():
    """solar_rng='host': numpy / CPU-torch global streams in the reference's order (Eval_Tools_2.py:72-108) - the same
    seeds give the rays the oracle draws"""
    import season_nerf_b200 as snb
    from oracle import season_oracle as so
    ts = snb.TrainStep(_args(), t.device("cuda"), so.oma_w2l_h(), so.OMA_W2C, solar_rng="host")
    assert not ts.eval_tool.solar_on_device
    np.random.seed(4)
    t.manual_seed(4)
    _, solar, _ = ts._draw_inputs(16, {"jitter": t.zeros(S), "solar_jitter": t.zeros(S)})
    ref = so.create_solar_rays_uniform(16, so.OMA_W2C, so.oma_w2l_h(), np.random.RandomState(4), t.Generator().manual_seed(4))
    for x, y in zip(solar, ref[:4]):
        assert maxabs(x, y) < 1e-6


def test_step_scalars_one_copy_and_tensorboard_tags():
    """TrainStep.log_scalars: the reference's TensorBoard tags (mg_run_NeRF.py:301-308,325) from one device->host copy"""
    import season_nerf_b200 as snb
    from oracle import season_oracle as so

    class Writer:
        def __init__(self):
            self.rows = []

        def add_scalar(self, tag, value, step):
            self.rows.append((tag, float(value), step))

    t.manual_seed(0)
    ts = snb.TrainStep(_args(), t.device("cuda"), so.oma_w2l_h(), so.OMA_W2C, use_graph=True, graph_warmup=1)
    batch = so.synthetic_batch(128, seed=1, n_images=5)
    w = Writer()
    for i in range(3):
        L = ts.step(batch, i)
        sc = ts.log_scalars(w, i)
        assert abs(sc["total"] - float(ts.last_loss)) < 1e-6 * max(1.0, abs(sc["total"]))
        for k in L:
            assert abs(sc[k] - float(L[k][0])) <= 1e-6 * max(1.0, abs(sc[k])), k
    tags = {r[0] for r in w.rows}
    assert {"Training/Color_ada", "Training/Solar_Correction", "Training/Albedo_Color", "LR/Learning_Rate"} <= tags
    assert all(np.isfinite(r[1]) for r in w.rows) and {r[2] for r in w.rows} == {0, 1, 2}
