"""Golden fixture of the training driver from the UNMODIFIED reference: `T_NeRF_Net_Tool` (T_NeRF_Full_2/Net_Tool_2.py:11-145)
on its base `Net_tool` (mg_run_NeRF.py:42-335) runs `step()` five times across the section switch on the seeded case of
oracle/nettool_case.py.  Only the data layer is replaced (the reference builds its loaders from prepared image files):
`mg_run_NeRF.build_data_loaders` returns in-memory datasets with the attributes of NN_loaders.pt_loader, the instance's
`get_data` hands out the case's batches (and re-seeds the global RNGs so that the draws of every step are reproducible),
the TensorBoard writer is a recorder.  Everything else - reset_eval, train_step, eval_step, eval_img, get_Dist, the
optimisers and schedulers - is the reference's own code.
Run in the build container: python -m oracle.make_golden_nettool"""
import os
import sys
import tempfile

import numpy as np
import torch as t

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import barron_loss                     # noqa: E402
from oracle import nettool_case as case            # noqa: E402
from oracle import season_oracle as so             # noqa: E402
from oracle.make_golden import save                # noqa: E402
from oracle.ref_import import import_reference     # noqa: E402


class FakeLoader(t.utils.data.Dataset):
    def __init__(self, all_data, img_ids, full_img_size, img_names):
        self.all_data, self.img_ids, self.full_img_size, self.img_names = all_data, img_ids, full_img_size, img_names
        self.solar_vecs = None

    def __len__(self):
        return self.all_data.shape[0]

    def __getitem__(self, item):
        return self.all_data[item]

    def get_id(self, item):
        return self.img_ids[item]


def main():
    import_reference()
    t.set_num_threads(8)
    import mg_run_NeRF
    import robust_loss_pytorch
    robust_loss_pytorch.AdaptiveLossFunction = barron_loss.AdaptiveLossFunction      # the absent third-party package
    from T_NeRF_Full_2 import Net_Tool_2
    Net_Tool_2.AdaptiveLossFunction = barron_loss.AdaptiveLossFunction
    vt, vids, vsize, vnames = case.val_table()
    batches = case.train_batches()
    train_ds = FakeLoader(t.cat(batches, 0), [0] * (len(batches) * case.BATCH), [(4, 4, 3)], ["train"])
    val_ds = FakeLoader(vt, vids, vsize, vnames)
    mg_run_NeRF.build_data_loaders = lambda args: ({"Color_Loader": train_ds, "SC_Loader": FakeLoader(t.zeros(4, 2), [0] * 4, [], [])},
                                                   {"Color_Loader": val_ds, "SC_Loader": FakeLoader(t.zeros(4, 2), [0] * 4, [], [])})
    logs = tempfile.mkdtemp()
    a = case.args(logs)
    training_DSM, GT_DSM = case.dsms()
    tool = Net_Tool_2.T_NeRF_Net_Tool(a, training_DSM, GT_DSM, t.device("cpu"), so.oma_w2l_h(), so.OMA_W2C)
    tool.network.load_state_dict({k: v.clone() for k, v in so.init_params(seed=0, perturb_bn=True).items()}, strict=True)
    rec = case.Recorder()
    tool.writer = rec
    state = {"i": 0}

    def get_data(eval_mode=False):
        if eval_mode:
            d = tool.data_to_dict(vt[:case.BATCH].clone())
        else:
            d = tool.data_to_dict(batches[state["i"]].clone())
            state["i"] += 1
        d["Dist_to_Surf_GT"], d["Dist_to_Surf_Prior"] = tool.get_Dist(d["Top"], d["Bot"])
        case.seed_step(tool._step_count + (100 if eval_mode else 0))
        return d

    tool.get_data = get_data
    out = {"section_starts": tool.section_starts, "section_Ends": tool.section_Ends, "Section_Steps": np.array(tool.Section_Steps),
           "save_points": tool.save_points}
    for i, o in enumerate(tool.sub_section_outputs):
        out["sub_section_outputs_%d" % i] = np.asarray(o)
    dg, dp = tool.get_Dist(vt[:, 2:5], vt[:, 5:8])
    out["dist_gt"], out["dist_prior"] = dg.numpy(), dp.numpy()            # float64, NaN where the GT DSM has none
    modes, ada_state = [], []
    for i in range(case.N_STEPS_RUN):
        tool.step()
        modes.append(int(tool.learning_mode))
        al = tool.eval_tool.ada_loss
        a0 = al[0] if isinstance(al, list) else al
        ada_state.append([float(t.mean(a0.alpha())), float(t.mean(a0.scale()))])
    out["modes"] = np.array(modes)
    out["ada_state"] = np.array(ada_state, dtype=np.float64)
    out["lr_last"] = np.float64(tool.sched.get_last_lr()[0])
    tags = sorted({s[0] for s in rec.scalars})
    out["scalar_tags"] = np.array(tags)
    for tag in tags:
        out["sc_" + tag.replace("/", "__")] = np.array([[s[2], s[1]] for s in rec.scalars if s[0] == tag], dtype=np.float64)
    out["image_tags"] = np.array([im[0] for im in rec.images])
    for j, im in enumerate(rec.images):
        out["im_%d" % j] = im[1]
    sd = tool.network.state_dict()
    P0 = so.init_params(seed=0, perturb_bn=True)
    names, dnorm = [], []
    for k, v in sd.items():
        if v.is_floating_point():
            names.append(k)
            dnorm.append(float((v - P0[k]).norm()))
    out["w_names"], out["w_delta_norms"] = np.array(names), np.array(dnorm, dtype=np.float64)
    for k in ["G_NeRF_net.fc10Sigma.weight", "G_NeRF_net.fc10Col.weight", "G_NeRF_net.fc2.norm.weight", "G_NeRF_net.fc9.linear.bias",
              "adjust_col.weight", "get_class_layer.weight", "G_NeRF_net.fc2.norm.running_mean", "G_NeRF_net.fc2.norm.running_var",
              "G_NeRF_net.fc_solar_4.weight", "G_NeRF_net.fc6.linear.weight"]:
        out["w_" + k] = sd[k]
    assert os.path.exists(os.path.join(logs, "Model_4.nn"))
    save("net_tool", **out)
    print("modes", modes, "ada", ada_state, "tags", tags)


if __name__ == "__main__":
    main()
