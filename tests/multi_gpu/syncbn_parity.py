"""torchrun --nproc-per-node 2 tests/multi_gpu/syncbn_parity.py [fp32|bf16] [graph]

Data-parallel parity of the training step: W ranks with B/W rays each and TrainStep(sync_bn=True) must take
the step ONE device takes on the B rays (the reference's semantics: BatchNorm over the whole batch, Albedo_Color minimum
over the whole batch) - same loss terms, same weights and BatchNorm running statistics after a few steps.  Rank 0 also
runs the single-device step and compares; without sync_bn the two trajectories differ (checked too).  Prints one JSON
line; exit code 0 = parity."""
import json
import os
import sys
import types

import numpy as np
import torch as t
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
S = 96


def _args():
    return types.SimpleNamespace(n_samples=S, Use_Reg=True, Solar_Type_2=False, Use_MSE_loss=False, sc_lambda=0.03,
                                 Use_Solar=True, number_low_frequency_cases=4, fc_units=512, lr=10 ** -4.86,
                                 lr_alpha_scale=1000.0, max_train_steps=1000)


def run(world, rank, sync_bn, precision, use_graph, n_total, n_steps, dev):
    import season_nerf_b200 as snb
    from oracle import season_oracle as so
    t.manual_seed(0)                                                   # identical initial weights everywhere
    ts = snb.TrainStep(_args(), dev, so.oma_w2l_h(), so.OMA_W2C, world_size=world, precision=precision, use_graph=use_graph,
                       graph_warmup=1, sync_bn=sync_bn)
    batch = so.synthetic_batch(n_total, seed=1, n_images=5)
    n = n_total // world
    lo, hi = rank * n, (rank + 1) * n
    losses = []
    for i in range(n_steps):
        rs, g = np.random.RandomState(10 + i), t.Generator().manual_seed(10 + i)
        st, en, vec, tm, _ = so.create_solar_rays_uniform(n_total, so.OMA_W2C, so.oma_w2l_h(), rs, g)
        jit = t.rand(S, generator=g)
        L = ts.step({k: v[lo:hi] for k, v in batch.items()}, i, jitter=jit, solar=tuple(x[lo:hi] for x in (st, en, vec, tm)),
                    solar_jitter=jit)
        vals = t.stack([t.as_tensor(L[k][0], device=dev).detach().float().reshape(()) * L[k][1] for k in sorted(L)])
        if world > 1:                                                  # the step's loss = mean of the ranks' losses
            dist.all_reduce(vals)
            vals /= world
        losses.append(dict(zip(sorted(L), vals.tolist())))
    sd = {k: v.detach().double().cpu() for k, v in ts.network.state_dict().items()}
    return losses, sd


def compare(a, b, adam_floor):
    """-> (max relative loss-term error, max relative weight error).  Adam moves every parameter by about lr per step
    whatever the size of its gradient, so two runs whose gradients agree to rounding still differ by up to steps * lr on
    parameters with near-zero gradients: tensors whose largest difference is below that floor count as equal."""
    la, sa = a
    lb, sb = b
    loss_err = max(abs(x[k] - y[k]) / max(abs(y[k]), 1e-3) for x, y in zip(la, lb) for k in x)
    w_err = 0.0
    for k in sa:
        if "num_batches_tracked" in k or sa[k].numel() < 2:
            continue
        diff = sa[k] - sb[k]
        if float(diff.abs().max()) <= adam_floor:
            continue
        ref = float(sb[k].norm())
        if ref > 0:
            w_err = max(w_err, float(diff.norm()) / ref)
    return loss_err, w_err


def main():
    precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    use_graph = "graph" in sys.argv[2:]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = t.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    t.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    n_total, n_steps = 256, 3
    dp_sync = run(world, rank, True, precision, use_graph, n_total, n_steps, dev)
    dp_plain = run(world, rank, False, precision, False, n_total, n_steps, dev)
    dist.barrier()
    rc = 0
    if rank == 0:
        single = run(1, 0, False, precision, False, n_total, n_steps, dev)
        floor = n_steps * _args().lr
        le, we = compare(dp_sync, single, floor)
        le_p, we_p = compare(dp_plain, single, floor)
        tol_l, tol_w = (2e-3, 2e-3) if precision == "fp32" else (5e-2, 5e-2)
        ok = le < tol_l and we < tol_w and (we_p > 3 * we or le_p > 3 * le)
        print(json.dumps({"precision": precision, "graph": use_graph, "world": world, "rays_total": n_total, "steps": n_steps,
                          "sync_bn_vs_single": {"loss_rel_err": le, "weight_rel_err": we},
                          "plain_dp_vs_single": {"loss_rel_err": le_p, "weight_rel_err": we_p}, "ok": bool(ok)}))
        rc = 0 if ok else 1
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(rc)


if __name__ == "__main__":
    main()
