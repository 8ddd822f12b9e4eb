import numpy as np
import torch as t


def T(x, dev="cuda"):
    return t.tensor(np.asarray(x)).to(dev)


def maxabs(a, b):
    a = a.detach().float().cpu().numpy() if isinstance(a, t.Tensor) else np.asarray(a, dtype=np.float64)
    b = b.detach().float().cpu().numpy() if isinstance(b, t.Tensor) else np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))) if a.size else 0.0


def relerr(a, b):
    a = a.detach().double().cpu().numpy() if isinstance(a, t.Tensor) else np.asarray(a, dtype=np.float64)
    b = b.detach().double().cpu().numpy() if isinstance(b, t.Tensor) else np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def make_net(params, precision, train=False):
    import season_nerf_b200 as snb
    net = snb.T_NeRF(512, 4, precision=precision)
    net.load_state_dict({k: v.clone() for k, v in params.items()}, strict=True)
    net = net.cuda()
    net.train(train)
    return net
