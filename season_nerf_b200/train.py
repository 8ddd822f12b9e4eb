"""Training step: drop-in for mg_run_NeRF.py:288-326 (Net_tool.train_step) and the optimiser / scheduler set-up of
T_NeRF_Full_2/Net_Tool_2.py:63-129 (T_NeRF_Net_Tool.reset_eval), plus the ONLY parallelism of the build: ray-sharded
data parallel over the GPUs of one box with one flat NCCL all-reduce of the gradients per step.
The reference's DataLoader / TensorBoard / checkpoint glue stays out of scope; `step()` takes the batch dict."""
import os

import torch as t
import torch.distributed as dist

from .adaptive_loss import AdaptiveLossFunction
from .engine import All_in_One_Eval, sample_ts
from .network import T_NeRF

FUSED_ADAM = os.environ.get("SNB_FUSED_ADAM", "1") != "0"


def flat_allreduce_mean_(tensors, world_size, flat=None):
    """One flat fp32 bucket: pack (one concatenation kernel) -> all_reduce(SUM) -> /world_size -> unpack (one multi-tensor
    copy), in place.  Works for NCCL (CUDA, capturable in a CUDA graph) and gloo (CPU, used by the host-logic tests).
    Returns the bucket."""
    if not tensors:
        return flat
    flat = t.cat([x.reshape(-1).float() for x in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(world_size)
    parts = flat.split([x.numel() for x in tensors])
    t._foreach_copy_(list(tensors), [v.view_as(x) for v, x in zip(parts, tensors)])
    return flat


def shard_range(n, rank, world_size):
    """contiguous ray range [lo, hi) of `rank` (render sharding; remainder rays go to the first ranks)."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(local, n_total, rank, world_size, dst=None):
    """Final gather of a ray-sharded render: per-rank row blocks of unequal length -> [n_total, ...].
    dst=None: every rank receives the rows (all_gather).  dst=r: only rank r receives them (`gather`; the other ranks
    return None) - the image leaves the box once, from one GPU, instead of world_size redundant device->host copies.
    One collective on equal, padded chunks; the valid rows are sliced out afterwards."""
    sizes = [shard_range(n_total, r, world_size)[1] - shard_range(n_total, r, world_size)[0] for r in range(world_size)]
    mx = max(sizes)
    tail = tuple(local.shape[1:])
    if local.shape[0] == mx:
        pad = local.contiguous()
    else:
        pad = t.zeros((mx,) + tail, dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
    if dst is not None:
        outs = [t.empty_like(pad) for _ in range(world_size)] if rank == dst else None
        dist.gather(pad, outs, dst=dst)
        if rank != dst:
            return None
        return t.cat([o[:s_] for o, s_ in zip(outs, sizes)], 0)
    flat = t.empty((world_size * mx,) + tail, dtype=local.dtype, device=local.device)
    try:
        dist.all_gather_into_tensor(flat, pad)
    except (RuntimeError, NotImplementedError, AttributeError):          # backend without the flat variant
        outs = [t.empty_like(pad) for _ in range(world_size)]
        dist.all_gather(outs, pad)
        flat = t.cat(outs, 0)
    if all(s_ == mx for s_ in sizes):
        return flat
    return t.cat([flat[r * mx:r * mx + s_] for r, s_ in enumerate(sizes)], 0)


class TrainStep:
    """One learning-mode section of T_NeRF_Net_Tool (learning_mode 2..4: free training, use_prior False) or the
    DSM-guided section (use_prior True, learning_mode 1 with jump_start)."""

    def __init__(self, args, device, H, WC, network=None, training_DSM=None, use_prior=False, total_steps=None,
                 world_size=1, precision="bf16", use_graph=False, graph_warmup=2, micro_batch=None, solar_rng="device",
                 sync_bn=False, trust_steps=None, ada_init=None, base_solar_vecs=None):
        """use_graph: after `graph_warmup` eager steps at a given batch size, the whole step (sampling, network forward,
        losses, backward and - single GPU - both Adam updates) is captured once into a CUDA graph and replayed; inputs live
        in static device buffers refreshed before each replay.  The arithmetic and kernel sequence are those of the eager
        step; only the per-step launches stop costing host time.  The DSM-guided section (use_prior) is captured too: its
        trust factor step / n_steps lives in a device scalar refreshed before each replay.
        solar_rng: "device" (default) draws the random solar rays of a step with torch's CUDA generator and builds them in
        one kernel (Eval_Tools_2.py:72-108 without its 89 us / ray host loop and without the H2D copy); "host" draws them
        from the reference's numpy / CPU-torch streams in the reference's order.  Injected rays (`solar=`) bypass both.
        sync_bn (world_size > 1): BatchNorm batch statistics - forward sums and the two backward sums of every layer - and
        the batch minimum of the Albedo_Color term are all-reduced over the ranks, so that N ranks with B/N rays each take
        the step the reference takes on one device with B rays (SURVEY 8e caveats 1-2); without it each rank normalises
        with its own shard (DistributedDataParallel semantics).  Ranks must hold equal numbers of rays.
        trust_steps: n_steps of the section's All_in_One_Eval (the prior's trust factor is step / n_steps,
        Net_Tool_2.py:88-90 passes the section END); default total_steps.  ada_init = (alpha, scale): start values of the
        colour loss carried over from the previous section (Net_Tool_2.py:70-78)."""
        if solar_rng not in ("device", "host"):
            raise ValueError("solar_rng must be 'device' or 'host'")
        self.solar_rng = solar_rng
        self.args, self.device = args, t.device(device)
        self.world_size = world_size
        self.use_graph = bool(use_graph) and self.device.type == "cuda"
        self.use_prior = use_prior
        self.graph_warmup = graph_warmup
        self.micro_batch = micro_batch
        self._graphs = {}
        self._eager_calls = {}
        self.launches_replayed = 0     # kernels of this library launched through graph replays (not seen by snb_launch_count)
        self.network = network if network is not None else T_NeRF(
            args.fc_units, n_classes=args.number_low_frequency_cases,
            **({} if training_DSM is None else {"HM": training_DSM}), precision=precision).to(self.device)
        self.network.train()
        total_steps = total_steps or args.max_train_steps
        a_init, s_init = ada_init if ada_init is not None else (2.0, .03)
        mk = lambda d, ai, si, sl: AdaptiveLossFunction(d, t.float32, self.device, alpha_hi=2.99, alpha_init=ai,
                                                        scale_init=si, scale_lo=sl)
        if args.Use_MSE_loss:
            ada, ada_params = None, []
        elif use_prior:                                                            # Net_Tool_2.py:69,82
            ada = [mk(3, a_init, s_init, 0.01), mk(1, 2.0, 0.5, 0.05)]
            ada_params = list(ada[0].parameters()) + list(ada[1].parameters())
        else:                                                                      # Net_Tool_2.py:78
            ada = mk(3, a_init, s_init, 0.01)
            ada_params = list(ada.parameters())
        self.eval_tool = All_in_One_Eval(args, self.device, trust_steps or total_steps, use_prior, ada, H, WC, base_solar_vecs)
        self.eval_tool.solar_on_device = solar_rng == "device"
        self.sync_bn = bool(sync_bn) and world_size > 1
        if self.sync_bn:
            if micro_batch:
                raise ValueError("sync_bn and micro_batch do not combine: a micro-batch is its own BatchNorm batch")
            self.network._sync_bn = world_size
            self.eval_tool.sync_world = world_size
        self.params = [p for p in self.network.parameters()]
        self.ada_params = ada_params
        gk = {}
        lr1, lr2 = args.lr, args.lr * args.lr_alpha_scale
        if self.use_graph:       # graph-capturable Adam: step counters and learning rates are device tensors
            # fused: one multi-tensor kernel per optimiser instead of ~12 foreach launches over the 70 parameter tensors
            # (0.4 ms of a 15.7 ms step); same fp32 arithmetic as the reference's default Adam (Net_Tool_2.py:110-119)
            gk = dict(capturable=True, fused=FUSED_ADAM)
            lr1, lr2 = t.tensor(lr1, device=self.device), t.tensor(lr2, device=self.device)
        self.optim = t.optim.Adam(self.params, lr=lr1, **gk)                       # Net_Tool_2.py:110-119
        self.optim2 = t.optim.Adam(ada_params, lr=lr2, **gk) if ada_params else None
        oc = dict(total_steps=total_steps, base_momentum=0.85, max_momentum=0.95, cycle_momentum=False)
        self.sched = t.optim.lr_scheduler.OneCycleLR(self.optim, max_lr=args.lr, **oc)
        self.sched2 = t.optim.lr_scheduler.OneCycleLR(self.optim2, max_lr=args.lr * args.lr_alpha_scale, **oc) \
            if self.optim2 is not None else None
        self.capture_optim = True       # the optimiser updates ride in the captured graph (capturable Adam)
        self._flat = None
        self.last_loss = None
        self._versioned = [v for v in self.network.state_dict(keep_vars=True).values()] + list(ada_params)

    @classmethod
    def adopt(cls, args, device, network, eval_tool, optim, optim2=None, sched=None, sched2=None, use_graph=True,
              graph_warmup=2, world_size=1):
        """A step around objects somebody else built - the way the reference's own `T_NeRF_Net_Tool.reset_eval`
        (Net_Tool_2.py:63-129) creates `eval_tool`, `optim`, `optim2`, `sched`, `sched2` itself and then calls
        `train_step`.  Forward + backward are still captured in a CUDA graph; the adopted optimisers are ordinary
        (non-capturable) torch optimisers, so their updates run eagerly after each replay."""
        self = cls.__new__(cls)
        self.solar_rng = "device" if getattr(eval_tool, "solar_on_device", False) else "host"
        self.args, self.device = args, t.device(device)
        self.world_size = world_size
        self.use_graph = bool(use_graph) and self.device.type == "cuda"
        self.use_prior = bool(eval_tool.use_prior)
        self.graph_warmup = max(1, graph_warmup)
        self.micro_batch = None
        self._graphs, self._eager_calls = {}, {}
        self.launches_replayed = 0
        self.network = network
        self.eval_tool = eval_tool
        self.sync_bn = False
        ada = eval_tool.ada_loss
        ada = [] if ada is None else (list(ada) if isinstance(ada, (list, tuple)) else [ada])
        self.params = [p for p in network.parameters()]
        self.ada_params = [p for a in ada for p in a.parameters()]
        self.optim, self.optim2, self.sched, self.sched2 = optim, optim2, sched, sched2
        self.capture_optim = False
        self._flat = None
        self.last_loss = None
        self._versioned = [v for v in network.state_dict(keep_vars=True).values()] + list(self.ada_params)
        return self

    # ---- checkpoint / resume (the reference saves only network.state_dict(), mg_run_NeRF.py:225: no resume path) --------
    def state_dict(self):
        """everything a resumed run needs: network (the reference's 94 keys), both optimisers, both schedulers and the
        adaptive-loss parameters"""
        ada = self.eval_tool.ada_loss
        ada = [] if ada is None else (list(ada) if isinstance(ada, (list, tuple)) else [ada])
        return {"network": self.network.state_dict(), "optim": self.optim.state_dict(),
                "optim2": None if self.optim2 is None else self.optim2.state_dict(),
                "sched": self.sched.state_dict(), "sched2": None if self.sched2 is None else self.sched2.state_dict(),
                "ada_loss": [a.state_dict() for a in ada]}

    def load_state_dict(self, sd):
        self.network.load_state_dict(sd["network"])
        self.optim.load_state_dict(sd["optim"])
        self.sched.load_state_dict(sd["sched"])
        if self.optim2 is not None and sd.get("optim2") is not None:
            self.optim2.load_state_dict(sd["optim2"])
            self.sched2.load_state_dict(sd["sched2"])
        ada = self.eval_tool.ada_loss
        ada = [] if ada is None else (list(ada) if isinstance(ada, (list, tuple)) else [ada])
        for a, s_ in zip(ada, sd.get("ada_loss", [])):
            a.load_state_dict(s_)
        self._graphs.clear()            # captured graphs baked the old optimiser state tensors
        self._eager_calls.clear()

    def _allreduce_grads(self):
        """one flat fp32 bucket (3.19 M network gradients + the adaptive-loss scalars), NCCL sum -> mean"""
        grads = [p.grad for p in self.params + self.ada_params if p.grad is not None]
        self._flat = flat_allreduce_mean_(grads, self.world_size, self._flat)

    # ---- one step -----------------------------------------------------------------------------------------
    _BATCH_KEYS = ("Top", "Bot", "Sun_Angle", "Time_Encoded", "GT_Color")

    def _draw_inputs(self, n, inject, solar_in_graph=False):
        """host-side random draws of one step, in the reference's order (Eval_Tools_2.py:165-170 jitter, :350 solar rays,
        :300-301 solar jitter) -> ts_img [S], solar 4-tuple, ts_solar [S] (CPU tensors unless injected on the device)."""
        S = self.args.n_samples
        ts_img = sample_ts(S, False, False, inject.get("jitter"))
        solar = inject.get("solar")
        if solar is None and self.args.Use_Solar and not solar_in_graph:
            if self.solar_rng == "device":
                solar = self.eval_tool.solar_creation_tool.on_device(n, self.device, include_times=True)
            else:
                solar = self.eval_tool.solar_creation_tool(n, include_times=True)[:4]
        ts_sol = sample_ts(S, False, True, inject.get("solar_jitter")) if self.args.Use_Solar else None
        return ts_img, solar, ts_sol

    def _chunks(self, n):
        """micro-batches of one step: [(lo, hi)] - equal sizes (the captured graph has one shape)"""
        mb = self.micro_batch if (self.micro_batch and self.micro_batch < n) else n
        if n % mb:
            raise ValueError("batch of %d rays is not a multiple of micro_batch=%d" % (n, mb))
        return [(lo, lo + mb) for lo in range(0, n, mb)]

    def _static(self, data_dict, mb):
        dev = self.device
        f32 = dict(device=dev, dtype=t.float32)
        S = self.args.n_samples
        st = {"batch": {k: t.empty((mb,) + tuple(data_dict[k].shape[1:]), **f32) for k in self._BATCH_KEYS},
              "ts_img": t.empty(S, **f32), "ts_sol": t.empty(S, **f32),
              "solar": tuple(t.empty(mb, w, **f32) for w in (3, 3, 3, 4)) if self.args.Use_Solar else None}
        return st

    def _fill_static(self, st, data_dict, ts_img, solar, ts_sol, lo, hi):
        for k in self._BATCH_KEYS:
            st["batch"][k].copy_(data_dict[k][lo:hi], non_blocking=True)
        st["ts_img"].copy_(ts_img, non_blocking=True)
        if st["solar"] is not None:
            st["ts_sol"].copy_(ts_sol, non_blocking=True)
            if solar is not None:            # None: the graph draws its own solar rays on the device
                for d, s_ in zip(st["solar"], solar):
                    d.copy_(s_[lo:hi], non_blocking=True)

    def _fwd_bwd(self, batch, current_step, scale, **kw):
        """loss of one (micro-)batch and its backward; gradients accumulate into .grad (scale = 1 / number of chunks)"""
        loss = self.eval_tool.get_loss(batch, self.network, current_step, train_mode=True, **kw)
        total = getattr(self.eval_tool, "_fused_total", None)        # the weighted sum, already formed by the loss kernel
        self.eval_tool._fused_total = None
        if total is None:
            total = 0
            for k in loss.keys():
                total = total + loss[k][0] * loss[k][1]
        (total * scale if scale != 1.0 else total).backward()
        # the step has consumed the autograd graph: hand back plain values (a caller that kept graph-attached losses alive
        # would also keep this iteration's AccumulateGrad nodes alive, which breaks a later CUDA-graph capture)
        return {k: [v[0].detach() if isinstance(v[0], t.Tensor) else v[0], v[1]] for k, v in loss.items()}, total.detach()

    def _optim_step(self):
        self.optim.step()
        if self.optim2 is not None:
            self.optim2.step()
        # torch's fused Adam (torch._fused_adam_) rewrites the parameters WITHOUT bumping their version counters (measured:
        # scripts/adam_probe.py): every cache derived from them - staged bf16 weights, packed render program - would go stale
        if self.use_graph and FUSED_ADAM and not t.cuda.is_current_stream_capturing():
            t.autograd.graph.increment_version(self._versioned)

    def _zero_grads(self, to_none):
        self.optim.zero_grad(set_to_none=to_none)
        if self.optim2 is not None:
            self.optim2.zero_grad(set_to_none=to_none)

    @staticmethod
    def _merge(losses, totals):
        k = len(losses)
        if k == 1:
            return losses[0], totals[0]
        out = {}
        for name in losses[0]:
            v0 = losses[0][name][0]
            out[name] = [sum(l[name][0] for l in losses) / k if isinstance(v0, t.Tensor) else v0, losses[0][name][1]]
        return out, sum(totals) / k

    def _step_graphed(self, data_dict, current_step, inject):
        n = data_dict["Top"].shape[0]
        chunks = self._chunks(n)
        k = len(chunks)
        mb = chunks[0][1] - chunks[0][0]
        # device-drawn solar rays are drawn INSIDE the graph (torch's CUDA generator is graph-safe: every replay continues the
        # stream): the eager draw + 4 copies into static buffers in front of every replay cost 0.15 ms of host time while the
        # GPU idled.  Injected rays (tests, the device-timed bench loop) keep their static buffers - a graph of its own.
        solar_in_graph = bool(self.args.Use_Solar and self.solar_rng == "device" and inject.get("solar") is None)
        ts_img, solar, ts_sol = self._draw_inputs(n, inject, solar_in_graph)
        st = self._graphs.get((n, mb, solar_in_graph))
        if st is None:
            st = self._graphs[(n, mb, solar_in_graph)] = self._static(data_dict, mb)
        if self.use_prior:
            if getattr(self.eval_tool, "trust_tensor", None) is None:
                self.eval_tool.trust_tensor = t.zeros((), device=self.device, dtype=t.float32)
            self.eval_tool.trust_tensor.fill_(current_step / self.eval_tool.n_steps)
        # one chunk: the gradient all-reduce (NCCL, capturable) and the optimiser updates ride in the same graph
        fused_opt = k == 1 and self.capture_optim
        if k > 1 and self.graph_warmup < 1:
            raise ValueError("micro-batched graph capture needs graph_warmup >= 1 (the eager step creates the .grad tensors)")
        if "g_fb" not in st:
            # record: fwd + bwd of one (micro-)batch.  One chunk: gradients are (re)created by the graph itself.  Several
            # chunks: the graph accumulates into the existing .grad tensors, zeroed before the first chunk of every step.
            self._fill_static(st, data_dict, ts_img, solar, ts_sol, *chunks[0])
            self._zero_grads(to_none=(k == 1))
            t.cuda.synchronize()
            from . import _lib
            l0 = _lib.launch_count()
            g_fb = t.cuda.CUDAGraph()
            from . import network as _nw
            _nw._capture_epoch[0] += 1          # weights staged by earlier passes / captures are not reused inside this one
            with t.cuda.graph(g_fb):
                st["loss"], st["total"] = self._fwd_bwd(st["batch"], current_step, 1.0 / k,
                                                        solar=None if solar_in_graph else st["solar"],
                                                        ts=st["ts_img"], solar_ts=st["ts_sol"])
                if fused_opt:
                    if self.world_size > 1:
                        self._allreduce_grads()
                    self._optim_step()
            st["g_fb"] = g_fb
            st["launches"] = _lib.launch_count() - l0
            # the gradient tensors the graph writes: a caller's `optim.zero_grad()` (set_to_none) drops them from the
            # parameters between steps; they are re-attached after every replay
            st["grads"] = [(p_, p_.grad) for p_ in self.params + self.ada_params if p_.grad is not None]
            if not fused_opt and self.capture_optim:
                g_opt = t.cuda.CUDAGraph()
                with t.cuda.graph(g_opt, pool=g_fb.pool()):
                    self._optim_step()
                st["g_opt"] = g_opt
        for p_, g_ in st["grads"]:
            p_.grad = g_
        if k > 1:
            self._zero_grads(to_none=False)
        losses, totals = [], []
        for lo, hi in chunks:
            self._fill_static(st, data_dict, ts_img, solar, ts_sol, lo, hi)
            st["g_fb"].replay()
            self.launches_replayed += st["launches"]
            if k > 1:
                losses.append({kk: [v[0].clone() if isinstance(v[0], t.Tensor) else v[0], v[1]] for kk, v in st["loss"].items()})
                totals.append(st["total"].clone())
        # a replay rewrites parameters and BatchNorm buffers behind autograd's back: bump their version counters so that
        # every derived cache (packed render program, staged bf16 weights) sees the change
        t.autograd.graph.increment_version(self._versioned)
        if not fused_opt:
            if self.world_size > 1:
                self._allreduce_grads()
            if self.capture_optim:
                st["g_opt"].replay()
                t.autograd.graph.increment_version(self._versioned)      # the replayed optimiser rewrote the parameters
            else:
                self._optim_step()
        if self.sched is not None:
            self.sched.step()
        if self.sched2 is not None:
            self.sched2.step()
        loss, self.last_loss = (st["loss"], st["total"]) if k == 1 else self._merge(losses, totals)
        self._last_loss_dict = loss
        return loss

    def scalars(self, loss=None):
        """{name: float} of a step's loss terms + total + learning rate from ONE device->host copy (the reference pays one
        `.item()` synchronisation per term and step, mg_run_NeRF.py:299-308)."""
        loss = loss if loss is not None else self._last_loss_dict
        names = [k for k in loss if isinstance(loss[k][0], t.Tensor)]
        vals = t.stack([loss[k][0].detach().float().reshape(()) for k in names] + [self.last_loss.detach().float().reshape(())])
        host = vals.cpu().tolist()
        out = dict(zip(names, host[:-1]))
        out.update({k: float(loss[k][0]) for k in loss if k not in out})
        out["total"] = host[-1]
        out["Learning_Rate"] = float(self.sched.get_last_lr()[0])
        return out

    def log_scalars(self, writer, current_step, loss=None):
        """TensorBoard tags of mg_run_NeRF.py:301-308,325 ('Training/<term>', 'LR/Learning_Rate') for any object with
        `add_scalar(tag, value, step)`."""
        sc = self.scalars(loss)
        for k, v in sc.items():
            if k not in ("total", "Learning_Rate"):
                writer.add_scalar("Training/" + k, v, current_step)
        writer.add_scalar("LR/Learning_Rate", sc["Learning_Rate"], current_step)
        return sc

    def step(self, data_dict, current_step, **inject):
        """mg_run_NeRF.py:288-326 without the per-term TensorBoard .item() syncs; returns the loss dict.
        With `micro_batch`, the batch is processed in equal chunks whose gradients accumulate before ONE optimiser step
        (each chunk is its own BatchNorm batch and its own Albedo_Color minimum; one jitter vector per step)."""
        n = data_dict["Top"].shape[0]
        if self.use_graph:
            if self._eager_calls.get(n, 0) >= self.graph_warmup:
                return self._step_graphed(data_dict, current_step, inject)
            self._eager_calls[n] = self._eager_calls.get(n, 0) + 1
        chunks = self._chunks(n)
        self._zero_grads(to_none=True)
        if self.use_prior and getattr(self.eval_tool, "trust_tensor", None) is not None:
            self.eval_tool.trust_tensor.fill_(current_step / self.eval_tool.n_steps)
        if len(chunks) == 1:
            loss, total = self._fwd_bwd(data_dict, current_step, 1.0, **inject)
        else:
            ts_img, solar, ts_sol = self._draw_inputs(n, inject)
            dev = self.device
            ts_img, ts_sol = ts_img.to(dev), (ts_sol.to(dev) if ts_sol is not None else None)
            losses, totals = [], []
            for lo, hi in chunks:
                l, tt = self._fwd_bwd({kk: data_dict[kk][lo:hi] for kk in self._BATCH_KEYS}, current_step, 1.0 / len(chunks),
                                      solar=None if solar is None else tuple(x[lo:hi] for x in solar), ts=ts_img, solar_ts=ts_sol)
                losses.append(l)
                totals.append(tt)
            loss, total = self._merge(losses, totals)
        if self.world_size > 1:
            self._allreduce_grads()
        self._optim_step()
        if self.sched is not None:
            self.sched.step()
        if self.sched2 is not None:
            self.sched2.step()
        self.last_loss = total
        self._last_loss_dict = loss
        return loss
