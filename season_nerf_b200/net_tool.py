"""Training drivers: drop-in for `Net_tool` (mg_run_NeRF.py:42-335) and `T_NeRF_Net_Tool` (T_NeRF_Full_2/Net_Tool_2.py:11-145).

`main.py`'s loop (`net_tool = T_NeRF_Net_Tool(args, training_DSM, GT_DSM, device, H, WC); for i in range(n): net_tool.step()`,
main.py:85-104) runs unchanged on top of `train.TrainStep`: every learning-mode section owns one TrainStep (fresh Adam x 2 and
OneCycleLR x 2 over the section's steps, Net_Tool_2.py:110-129), so the reference entry point reaches the CUDA-graph step,
the device-built solar rays and the one-copy scalar logging.  What the tool keeps from the reference, line by line:

  * the section schedule: `ps = [0.2, 0, 0, 0.8]` of `max_train_steps`, `section_starts / section_Ends / Section_Steps`,
    `sub_section_outputs` from `misc.get_output_loc_lin_first` (Net_Tool_2.py:23-54); `learning_mode =
    sum(step >= section_starts)` (:134), i.e. mode 1 (DSM-guided when `jump_start`) then mode 4;
  * `reset_eval` (:63-129): mode 1 starts the colour loss at alpha 2 / scale 0.03 (+ the 1-d loss of the prior term when
    `jump_start`); later modes start from the mean alpha / scale the previous section reached (:70-78, with the reference's
    fall-back to the defaults when the previous loss object cannot be indexed); the section's eval tool gets
    `n_steps = section_Ends[mode-1]` (the prior's trust denominator) and the schedulers `Section_Steps[mode-1]`;
  * `step` (:133-145): train step, then `eval_step` + `eval_img` at the save points of the section;
  * `Net_tool.get_Dist / data_to_dict / get_data / train_step / eval_step / eval_img / get_num_epochs / get_output_loc*`
    with the reference's TensorBoard tags.

Data: the reference builds `pt_loader` datasets from prepared files (NN_loaders/mg_Color_Loader.py, GDAL side: out of
scope).  Here `train_data` / `val_data` are `ColorTable`s (or any object with the same attributes: `all_data [n,22]`,
`img_ids`, `full_img_size`, `img_names`, `solar_vecs`); batches come from a device-resident shuffled `data.RayTable`
instead of a 4-worker DataLoader.  Without them the constructor tries the reference's own `build_data_loaders(args)`.
"""
import numpy as np
import torch as t

from . import ops
from .adaptive_loss import AdaptiveLossFunction  # noqa: F401  (re-exported: Net_Tool_2.py:8 imports it from the package)
from .data import RayTable, data_to_dict
from .engine import All_in_One_Eval, sample_pt_coarse  # noqa: F401
from .network import T_NeRF
from .train import TrainStep


def get_output_loc(n_steps, n_outputs):
    """misc.py:35-42 / mg_run_NeRF.py:309-316."""
    if n_outputs > 0:
        alpha = np.log(n_steps) / np.log(n_outputs)
        ans = (np.arange(1, n_outputs + 1) ** alpha).astype(int)
        ans[-1] = n_steps
    else:
        ans = np.array([n_steps])
    return ans


def get_output_loc_lin_first(n_steps, n_outputs, min_gap):
    """misc.py:45-53 (the variant T_NeRF_Net_Tool uses for its per-section save points)."""
    if n_outputs * min_gap >= n_steps:
        ans = np.linspace(1, n_steps, n_outputs + 1, dtype=int)[1::]
    else:
        ans = get_output_loc(n_steps, n_outputs)
        lin = np.arange(1, n_outputs + 1) * min_gap
        ans = np.maximum(ans, lin)
    return ans


def section_schedule(n_steps, n_saves):
    """Net_Tool_2.py:15-54: the four learning-mode sections (20 % DSM-guided, then - two empty sections later - mode 4 for the
    rest) -> section_starts, section_Ends, Section_Steps, sub_section_outputs (the save points of each section)."""
    ps = [0.2, 0.0, 0.0]
    ps.append(1 - np.sum(ps))
    p1 = int(ps[0] * n_steps)
    p2 = int(ps[1] * n_steps)
    p3 = int(ps[2] * n_steps)
    p4 = n_steps - p3 - p2 - p1
    pi = [p1, p2, p3, p4]
    section_starts = np.array([0, p1, p1 + p2, p1 + p2 + p3])
    section_Ends = np.array([p1, p1 + p2, p1 + p2 + p3, n_steps])
    Section_Steps = []
    for i in range(section_starts.shape[0] - 1):
        Section_Steps.append(int(section_starts[i + 1] - section_starts[i]))
    Section_Steps.append(int(n_steps - section_starts[-1]))
    sub_section_outputs = []
    for i in range(section_starts.shape[0]):
        output_points = get_output_loc_lin_first(pi[i], int(n_saves * ps[i]), min_gap=1000)
        sub_section_outputs.append(section_starts[i] + output_points)
    sub_section_outputs[-1][-1] = n_steps
    return section_starts, section_Ends, Section_Steps, sub_section_outputs


class ColorTable:
    """The attributes of NN_loaders.mg_Color_Loader.pt_loader (:40-105) that the training loop reads, over an in-memory
    [n,22] ray table: `all_data`, `img_ids`, `full_img_size`, `img_names`, `solar_vecs`, `get_id`, `[j]`, `len`."""

    def __init__(self, all_data, img_ids=None, full_img_size=None, img_names=None, solar_vecs=None):
        self.all_data = t.as_tensor(all_data).float()
        n = self.all_data.shape[0]
        self.img_ids = [0] * n if img_ids is None else [int(i) for i in img_ids]
        n_img = max(self.img_ids) + 1 if n else 0
        if full_img_size is None:
            h = int(self.all_data[:, 0].max()) + 1 if n else 1
            w = int(self.all_data[:, 1].max()) + 1 if n else 1
            full_img_size = [(h, w, 3)] * n_img
        self.full_img_size = list(full_img_size)
        self.img_names = list(img_names) if img_names is not None else ["img_%d" % i for i in range(n_img)]
        self.solar_vecs = solar_vecs

    def __len__(self):
        return self.all_data.shape[0]

    def __getitem__(self, item):
        return self.all_data[item]

    def get_id(self, item):
        return self.img_ids[item]


class _NullWriter:
    """stands in for torch.utils.tensorboard.SummaryWriter when tensorboard is not installed; records the scalars"""

    def __init__(self):
        self.scalars, self.images = [], []

    def add_scalar(self, tag, value, step):
        self.scalars.append((tag, float(value), int(step)))

    def add_image(self, tag, img, step):
        self.images.append((tag, np.asarray(img).shape, int(step)))


def _make_writer(args):
    try:
        from torch.utils.tensorboard import SummaryWriter
        return SummaryWriter(args.logs_dir, comment=getattr(args, "exp_name", ""))
    except Exception:
        return _NullWriter()


class Net_tool():
    """mg_run_NeRF.py:42-335."""

    def __init__(self, args, device, training_DSM, GT_DSM, init_network=True, has_weight_term=False, train_data=None,
                 val_data=None, writer=None, precision="bf16", use_graph=True, world_size=1, rank=0, seed=0):
        self.args = args
        self.num_rays_to_eval = args.chunk
        self.num_course_samples = args.n_samples
        self.num_fine_samples = args.n_importance
        self.n_saves = args.n_saves
        self.n_steps = args.max_train_steps
        self.batch_size = args.batch_size
        self.n_DSM_samples = args.n_samples + args.n_importance
        self.has_weight_term = has_weight_term
        self.use_solar = args.sc_lambda > 0
        self._step_count = 0
        self.save_points = self.get_output_loc_lin_first(args.max_train_steps, args.n_saves, 500)
        self.device = t.device(device)
        if self.device.type != "cuda":
            raise ops._lib.SeasonNerfCudaError("season_nerf_b200.Net_tool trains on CUDA only (no CPU fallback)")
        self.precision, self.use_graph, self.world_size, self.rank = precision, use_graph, world_size, rank
        self.training_DSM, self.GT_DSM = np.asarray(training_DSM), np.asarray(GT_DSM)
        # mg_run_NeRF.py:59-74 builds dense float64 occupancy volumes [H,W,n] = (DSM >= h_k) + DSM*0 on the host; only their
        # definition is needed: the two height maps and the n levels live on the device (get_Dist evaluates the comparison)
        self._dsm_dev = {True: t.tensor(self.GT_DSM, dtype=t.float64, device=self.device),
                         False: t.tensor(self.training_DSM, dtype=t.float64, device=self.device)}
        self._levels = t.tensor(np.linspace(-1, 1, self.n_DSM_samples), dtype=t.float64, device=self.device)
        if train_data is None or val_data is None:
            try:
                from mg_run_NeRF import build_data_loaders          # the reference's file-based loaders (needs its data)
            except Exception as e:
                raise RuntimeError("season_nerf_b200.Net_tool needs train_data / val_data (ColorTable or [n,22] ray tables): the "
                                   "reference's build_data_loaders is not importable here (%s)" % e)
            self.train_data, self.val_data = build_data_loaders(args)
        else:
            wrap = lambda d: d if isinstance(d, dict) else {"Color_Loader": d if hasattr(d, "all_data") else ColorTable(d)}
            self.train_data, self.val_data = wrap(train_data), wrap(val_data)
        self._tables = {}
        for name, dd, sd in (("train", self.train_data, seed), ("val", self.val_data, seed + 1)):
            table = dd["Color_Loader"].all_data
            if world_size > 1 and name == "train":
                table = RayTable.shard(table, rank, world_size)
            self._tables[name] = RayTable(table, args.batch_size, self.device, seed=sd + 7919 * rank)
        self.writer = writer if writer is not None else _make_writer(args)
        self.GT_Cache = None
        self.network = None
        self.eval_tool = None
        self.optim = self.optim2 = self.sched = self.sched2 = None
        self._ts = None
        self.log_every = 1

    # ---- DSM distances (mg_run_NeRF.py:99-120) ---------------------------------------------------------------------
    def _scale_to_DSM(self, pts, use_GT):
        d = self.GT_DSM if use_GT else self.training_DSM
        K = t.tensor([d.shape[0] - 1, d.shape[1] - 1, self.n_DSM_samples - 1], device=pts.device).reshape([1, 1, 3])
        return ((pts + 1) / 2 * K).type(t.long)

    def _surf_loc(self, pts, delta, use_GT):
        """expected distance to the first occupied DSM cell along each ray (:103-106 / :108-113), float64 like the reference;
        out-of-cube sample points are clamped into the grid (the reference's host indexing would wrap or raise there)"""
        dsm = self._dsm_dev[use_GT]
        n = self.n_DSM_samples
        idx = self._scale_to_DSM(pts, use_GT)
        ix = idx[..., 0].clamp(0, dsm.shape[0] - 1)
        iy = idx[..., 1].clamp(0, dsm.shape[1] - 1)
        iz = idx[..., 2].clamp(0, n - 1)
        h = dsm[ix, iy]                                               # [N, n]
        PE = ((h >= self._levels[iz]).double() + h * 0).unsqueeze(-1)    # + h*0 keeps the NaN cells of the DSM (:63-64)
        ones = t.ones([PE.shape[0], 1, 1], dtype=t.float64, device=PE.device)
        prob = PE * t.cumprod(t.cat([ones, 1 - PE], 1), 1)[:, 0:-1]
        return t.sum(prob * t.cumsum(delta, 1), 1) / t.sum(prob, 1)

    def get_Dist(self, top, bot):
        """-> Surf_Loc_GT, Surf_Loc_Prior [N,1] float64 (device tensors; the reference returns CPU tensors)"""
        with t.no_grad():
            pts, delta = sample_pt_coarse(top, bot, self.n_DSM_samples, eval_mode=True, device=self.device)
            return self._surf_loc(pts, delta, True), self._surf_loc(pts, delta, False)

    def data_to_dict(self, data):
        return data_to_dict(data)

    # ---- data (mg_run_NeRF.py:229-264) -----------------------------------------------------------------------------
    def get_data(self, eval_mode=False):
        data_dict, _ = self._tables["val" if eval_mode else "train"].next_batch()
        if eval_mode:       # the training step never reads the distances: they are looked up where they are consumed
            data_dict["Dist_to_Surf_GT"], data_dict["Dist_to_Surf_Prior"] = self.get_Dist(data_dict["Top"], data_dict["Bot"])
        return data_dict

    # ---- one step (mg_run_NeRF.py:135-146) -------------------------------------------------------------------------
    def step(self):
        data_dict = self.get_data(eval_mode=False)
        self.train_step(data_dict, self._step_count)
        self._step_count += 1
        if self._step_count in self.save_points or self._step_count == 1:
            print("Evaluating step", self._step_count)
            data_dict = self.get_data(eval_mode=True)
            self.eval_step(data_dict, self._step_count - 1)
            self.eval_img(self._step_count - 1)

    def _bind(self, ts):
        """expose the section's TrainStep under the attribute names of the reference tool"""
        self._ts = ts
        self.eval_tool = ts.eval_tool
        self.optim, self.optim2, self.sched, self.sched2 = ts.optim, ts.optim2, ts.sched, ts.sched2

    def train_step(self, data_dict, current_step, **inject):
        """mg_run_NeRF.py:288-326: zero_grad, get_loss, sum of value * weight, backward, Adam x 2, OneCycleLR x 2, scalars.
        The scalars of a step leave the device in ONE copy (TrainStep.log_scalars) instead of one .item() per term."""
        ts = self._ts
        if ts is None or ts.eval_tool is not self.eval_tool or ts.optim is not self.optim or ts.network is not self.network:
            # a subclass set up `eval_tool / optim / optim2 / sched / sched2` itself, the way the reference's own
            # T_NeRF_Net_Tool.reset_eval does (Net_Tool_2.py:63-129): adopt them - forward + backward still replay from a
            # CUDA graph, the adopted (non-capturable) optimisers step eagerly
            if self.eval_tool is None or self.optim is None or self.network is None:
                raise RuntimeError("no training section is set up: call reset_eval() (T_NeRF_Net_Tool.step does)")
            self.eval_tool.solar_on_device = getattr(self, "solar_rng", "device") == "device"
            ts = self._ts = TrainStep.adopt(self.args, self.device, self.network, self.eval_tool, self.optim, self.optim2,
                                            self.sched, self.sched2, use_graph=self.use_graph, world_size=self.world_size)
        loss = ts.step(data_dict, current_step, **inject)
        if self.log_every and current_step % self.log_every == 0:
            ts.log_scalars(self.writer, current_step, loss)
        return loss

    def eval_step(self, data_dict, current_step):
        """mg_run_NeRF.py:328-339."""
        with t.no_grad():
            self.network.eval()
            loss = self.eval_tool.get_loss(data_dict, self.network, current_step, train_mode=False)
            self.network.train()
            names = [k for k in loss if isinstance(loss[k][0], t.Tensor)]
            vals = t.stack([loss[k][0].detach().float().reshape(()) for k in names]).cpu().tolist() if names else []
            for k, v in zip(names, vals):
                self.writer.add_scalar("Testing/" + k, v, current_step)
        return dict(zip(names, vals))

    def eval_img(self, step_count, save=True):
        """mg_run_NeRF.py:148-227: render every validation ray (eval mode), scatter colour / expected height / height error
        into per-image rasters on the device, log them and the two summary scalars, save `Model_<step>.nn`.
        -> dict with the rasters (numpy) and the scalars (the reference returns nothing)."""
        val = self.val_data["Color_Loader"]
        self.network.eval()
        dev = self.device
        with t.no_grad():
            n_pts = len(val)
            BS = max(1, self.args.chunk // (self.args.n_samples + self.args.n_importance))
            sizes = val.full_img_size
            n_img = len(sizes)
            imgs = t.zeros([n_img] + list(sizes[0]), dtype=t.float64, device=dev)
            hm = t.zeros([n_img] + list(sizes[0][0:2]), dtype=t.float64, device=dev)
            mae = t.zeros_like(hm)
            update_GT_cache = self.GT_Cache is None
            if update_GT_cache:
                self.GT_Cache = np.zeros(tuple(imgs.shape))
                self.GT_Cache_n = np.zeros(n_img)
                gt_dev = t.zeros_like(imgs)
            ids_all = t.as_tensor(np.asarray(val.img_ids), device=dev, dtype=t.long)
            table = val.all_data.to(dev)
            BS = max(BS, 4096)      # chunking is by device memory, not by args.chunk (results do not depend on it: eval-mode BatchNorm)
            for i in range(0, n_pts, BS):
                e = min(n_pts, i + BS)
                data_dict = self.data_to_dict(table[i:e])
                d_gt, _ = self.get_Dist(data_dict["Top"], data_dict["Bot"])
                R = self.eval_tool.eval(data_dict, self.network, current_step=self.args.max_train_steps, train_mode=False)
                P_Surf, deltas, sample_pts, rendered = R["PS"], R["deltas"], R["sample_pts"], R["Rendered_Col"]
                loc = t.sum(P_Surf * sample_pts, 1) / (t.sum(P_Surf, 1) + 1e-8)
                dist = t.sum(t.cumsum(deltas, 1) * P_Surf, 1) / t.sum(P_Surf, 1)
                a_mae = t.abs(d_gt - dist)
                ip = data_dict["Img_Pt"].int().long()
                ids = ids_all[i:e]
                imgs[ids, ip[:, 0], ip[:, 1]] = rendered.double()
                hm[ids, ip[:, 0], ip[:, 1]] = loc[:, 2].double()
                mae[ids, ip[:, 0], ip[:, 1]] = a_mae[:, 0]
                if update_GT_cache:
                    gt_dev[ids, ip[:, 0], ip[:, 1]] = data_dict["GT_Color"].double()
            out_val_images, out_val_hm, out_val_MAE = imgs.cpu().numpy(), ((hm + 1) / 2).cpu().numpy(), mae.cpu().numpy()
            if update_GT_cache:
                self.GT_Cache = gt_dev.cpu().numpy()
            img_error, mean_h_err = 0, None
            for i in range(n_img):
                self.writer.add_image("HM/Img_" + val.img_names[i], np.expand_dims(out_val_hm[i], 0), step_count)
                if i != n_img - 1:
                    out_img = np.moveaxis(np.concatenate([self.GT_Cache[i], out_val_images[i]], 1), -1, 0)
                    if update_GT_cache:
                        self.GT_Cache_n[i] = np.sum(np.any(self.GT_Cache[i] != 0, 2)) * 3
                    img_error += np.sum(np.log(1 / 2 * (self.GT_Cache[i] - out_val_images[i]) ** 2 + 1)) / self.GT_Cache_n[i]
                else:
                    out_img = np.moveaxis(out_val_images[i], -1, 0)
                    a = out_val_MAE[i]
                    mean_h_err = float(np.mean(a[a == a]))
                    self.writer.add_scalar("Testing/Mean_Height_Error", mean_h_err, step_count)
                if getattr(self.args, "use_HSLuv", False):
                    import hsluv
                    out_img = out_img * np.array([360., 100, 100]).reshape([3, 1, 1])
                    for x_idx in range(out_img.shape[1]):
                        for y_idx in range(out_img.shape[2]):
                            out_img[:, x_idx, y_idx] = hsluv.hsluv_to_rgb(out_img[:, x_idx, y_idx])
                self.writer.add_image("Col/Img_" + val.img_names[i], out_img, step_count)
        img_error = img_error / (n_img - 1) if n_img > 1 else float("nan")
        self.writer.add_scalar("Testing/Overall_Cauchy_Color_Error", img_error, step_count)
        if save:
            t.save(self.network.state_dict(), self.args.logs_dir + "/Model_" + str(step_count) + ".nn")
        self.network.train()
        return {"images": out_val_images, "height": out_val_hm, "height_error": out_val_MAE, "Mean_Height_Error": mean_h_err,
                "Overall_Cauchy_Color_Error": img_error}

    def get_num_epochs(self):
        return (self.n_steps * self.batch_size) / len(self.train_data["Color_Loader"])

    def get_output_loc(self, n_steps, n_outputs):
        return get_output_loc(n_steps, n_outputs)

    def get_output_loc_lin_first(self, n_steps, n_outputs, min_gap):
        """mg_run_NeRF.py:318-326 (NOT misc.get_output_loc_lin_first: this one starts its linear ramp at 1)."""
        if n_outputs * min_gap >= n_steps:
            ans = np.linspace(1, n_steps, n_outputs, dtype=int)
        else:
            ans = self.get_output_loc(n_steps, n_outputs)
            lin = np.arange(n_outputs) * min_gap + 1
            ans = np.maximum(ans, lin)
        return ans


class T_NeRF_Net_Tool(Net_tool):
    """T_NeRF_Full_2/Net_Tool_2.py:11-145."""

    def __init__(self, args, training_DSM, GT_DSM, device, H, WC, network=None, **kw):
        super(T_NeRF_Net_Tool, self).__init__(args, device, training_DSM, GT_DSM, init_network=False, has_weight_term=True, **kw)
        self.section_starts, self.section_Ends, self.Section_Steps, self.sub_section_outputs = section_schedule(
            args.max_train_steps, args.n_saves)
        self.learning_mode = -1
        self.network = network if network is not None else T_NeRF(
            args.fc_units, n_classes=args.number_low_frequency_cases, HM=training_DSM, precision=self.precision).to(self.device)
        self.lr = args.lr
        self.H = H
        self.WC = WC

    def reset_eval(self):
        """Net_Tool_2.py:63-129: the eval tool, optimisers and schedulers of the section `self.learning_mode` (1..4)."""
        mode = int(self.learning_mode)
        if mode not in (1, 2, 3, 4):
            raise ValueError("Error: Invalid learning mode %r (5, 'Seasonal Learning with Outliers', is not implemented by the "
                             "reference either)" % (self.learning_mode,))
        scale_init = .03
        ada_init = None
        if not self.args.Use_MSE_loss and mode != 1:
            try:                                                                   # :70-78
                prev = self.eval_tool.ada_loss[0]
                ada_init = (t.mean(prev.alpha()).item(), t.mean(prev.scale()).item())
            except Exception:
                print("WARNING: Unable to load alpha and scale start, using default")
                ada_init = (2.0, scale_init)
        print({1: "Guided Classic Learning", 2: "Classic Learning", 3: "Classic and Seasonal Learning",
               4: "Classic and Seasonal Learning with Outliers"}[mode])
        use_prior = mode == 1 and bool(self.args.jump_start)
        solar_vecs = getattr(self.train_data["Color_Loader"], "solar_vecs", None)
        ts = TrainStep(self.args, self.device, self.H, self.WC, network=self.network, use_prior=use_prior,
                       total_steps=self.Section_Steps[mode - 1], trust_steps=int(self.section_Ends[mode - 1]),
                       world_size=self.world_size, precision=self.precision, use_graph=self.use_graph, ada_init=ada_init,
                       base_solar_vecs=solar_vecs, solar_rng=getattr(self, "solar_rng", "device"))
        self._bind(ts)

    def step(self):
        """Net_Tool_2.py:133-145."""
        mode = np.sum(self._step_count >= self.section_starts)
        if mode != self.learning_mode:
            self.learning_mode = mode
            self.reset_eval()
        data_dict = self.get_data(eval_mode=False)
        self.train_step(data_dict, self._step_count)
        self._step_count += 1
        if self._step_count in self.sub_section_outputs[mode - 1]:
            print("Evaluating step", self._step_count)
            data_dict = self.get_data(eval_mode=True)
            self.eval_step(data_dict, self._step_count - 1)
            self.eval_img(self._step_count - 1)
