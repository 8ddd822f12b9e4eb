"""Oracle restatement of ``robust_loss_pytorch.AdaptiveLossFunction``.

TEST INFRASTRUCTURE ONLY (see oracle/season_oracle.py header).

PARITY UNPINNED: the reference installs this third-party package with an
unpinned ``pip install git+https://github.com/jonbarron/robust_loss_pytorch``
(reference README.md:26); it is absent from /root/reference and from this
image, and no reference test touches it.  This file restates the published
algorithm (Barron, "A General and Adaptive Robust Loss Function", CVPR 2019,
eqs. 8, 13-16) and is anchored only on the reference's call sites:
construction Net_Tool_2.py:69,78,82, use Eval_Tools_2.py:426-443.

Differences from the package that are visible only in d/d(latent_alpha): the
package interpolates a pre-computed cubic spline table of log Z(alpha)
(``partition_spline.npz``); here log Z(alpha) is evaluated by Gauss-Legendre
quadrature of exp(-rho(x, alpha, 1)) under x = tan(theta).  The network
gradients do not depend on log Z.
"""
from __future__ import annotations

import numpy as np
import torch as t

_EPS = float(np.finfo(np.float32).eps)


def general_lossfun(x, alpha, scale):
    """rho(x, alpha, c) of eq. 8 with the package's numerically safe special cases."""
    sq = (x / scale) ** 2
    loss_two = 0.5 * sq
    loss_zero = t.log1p(t.clamp(0.5 * sq, max=3e37))
    loss_neginf = -t.expm1(-0.5 * sq)
    loss_posinf = t.expm1(t.clamp(0.5 * sq, max=87.5))
    beta_safe = t.clamp(t.abs(alpha - 2.), min=_EPS)
    alpha_safe = t.where(alpha >= 0, t.ones_like(alpha), -t.ones_like(alpha)) * t.clamp(t.abs(alpha), min=_EPS)
    loss_otherwise = (beta_safe / alpha_safe) * (t.pow(sq / beta_safe + 1., 0.5 * alpha) - 1.)
    return t.where(alpha == -float("inf"), loss_neginf,
           t.where(alpha == 0, loss_zero,
           t.where(alpha == 2, loss_two,
           t.where(alpha == float("inf"), loss_posinf, loss_otherwise))))


_GL_NODES, _GL_WEIGHTS = np.polynomial.legendre.leggauss(1024)


def log_base_partition_function(alpha):
    """log Z(alpha) = log int exp(-rho(x, alpha, 1)) dx, by quadrature (x = tan(theta))."""
    th = t.tensor(_GL_NODES * (np.pi / 2), dtype=t.float64)
    w = t.tensor(_GL_WEIGHTS * (np.pi / 2), dtype=t.float64)
    x = t.tan(th)
    a = alpha.to(t.float64).reshape(-1, 1)
    rho = general_lossfun(x.reshape(1, -1), a, t.ones_like(a))
    integrand = t.exp(-rho) / t.cos(th).reshape(1, -1) ** 2
    return t.log(t.sum(integrand * w.reshape(1, -1), 1)).reshape(alpha.shape).to(alpha.dtype)


def _inv_softplus(y):
    return np.log(np.expm1(y))


class AdaptiveLossFunction(t.nn.Module):
    """Same constructor / methods the reference calls (Net_Tool_2.py:69-82,
    Eval_Tools_2.py:426-443): lossfun(x[N,d]) -> NLL[N,d], alpha(), scale()."""

    def __init__(self, num_dims, float_dtype=t.float32, device="cpu", alpha_lo=0.001, alpha_hi=1.999,
                 alpha_init=None, scale_lo=1e-5, scale_init=1.0):
        super().__init__()
        self.num_dims, self.alpha_lo, self.alpha_hi = num_dims, alpha_lo, alpha_hi
        self.scale_lo, self.scale_init = scale_lo, scale_init
        if alpha_init is None:
            alpha_init = (alpha_lo + alpha_hi) / 2.
        q = (alpha_init - alpha_lo) / (alpha_hi - alpha_lo)
        latent_alpha_init = float(np.log(q) - np.log1p(-q))            # inv_affine_sigmoid
        self.latent_alpha = t.nn.Parameter(t.full((1, num_dims), latent_alpha_init, dtype=float_dtype, device=device))
        self.latent_scale = t.nn.Parameter(t.zeros((1, num_dims), dtype=float_dtype, device=device))

    def alpha(self):
        return t.sigmoid(self.latent_alpha) * (self.alpha_hi - self.alpha_lo) + self.alpha_lo

    def scale(self):
        shift = float(_inv_softplus(1.0))
        return (self.scale_init - self.scale_lo) * t.nn.functional.softplus(self.latent_scale + shift) + self.scale_lo

    def lossfun(self, x):
        alpha, scale = self.alpha(), self.scale()
        return general_lossfun(x, alpha, scale) + t.log(scale) + log_base_partition_function(alpha)
