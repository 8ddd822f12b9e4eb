"""Barron's adaptive robust loss, the product-side stand-in for the third-party `robust_loss_pytorch` package the
reference imports (Net_Tool_2.py:8,69-82; used at Eval_Tools_2.py:426-443).  The package is not vendored by the
reference and is absent offline, so `AdaptiveLossFunction` here follows the paper (Barron, CVPR 2019) with the
package's parameterisation: alpha = sigmoid(latent)*(hi-lo)+lo, scale = (init-lo)*softplus(latent+softplus^-1(1))+lo,
NLL = rho(x, alpha, scale) + log(scale) + log Z(alpha).  A user who has the real package can pass its object
instead: the engine only calls .lossfun(), .alpha(), .scale(), .parameters().  Operates on [N,3] / [N*S,1]
residuals: small tensors, plain torch ops on the current device.
"""
import numpy as np
import torch as t

_EPS = float(np.finfo(np.float32).eps)


def lossfun(x, alpha, scale):
    sq = (x / scale) ** 2
    b = t.clamp(t.abs(alpha - 2.), min=_EPS)
    a = t.where(alpha >= 0, t.ones_like(alpha), -t.ones_like(alpha)) * t.clamp(t.abs(alpha), min=_EPS)
    general = (b / a) * (t.pow(sq / b + 1., 0.5 * alpha) - 1.)
    return t.where(alpha == 2, 0.5 * sq, t.where(alpha == 0, t.log1p(0.5 * sq), general))


class AdaptiveLossFunction(t.nn.Module):
    _nodes, _weights = np.polynomial.legendre.leggauss(768)

    def __init__(self, num_dims, float_dtype=t.float32, device="cpu", alpha_lo=0.001, alpha_hi=1.999, alpha_init=None,
                 scale_lo=1e-5, scale_init=1.0):
        super().__init__()
        self.alpha_lo, self.alpha_hi, self.scale_lo, self.scale_init = alpha_lo, alpha_hi, scale_lo, scale_init
        if alpha_init is None:
            alpha_init = (alpha_lo + alpha_hi) / 2.
        q = (alpha_init - alpha_lo) / (alpha_hi - alpha_lo)
        self.latent_alpha = t.nn.Parameter(t.full((1, num_dims), float(np.log(q) - np.log1p(-q)), dtype=float_dtype, device=device))
        self.latent_scale = t.nn.Parameter(t.zeros((1, num_dims), dtype=float_dtype, device=device))
        self.register_buffer("_th", t.tensor(self._nodes * (np.pi / 2), dtype=t.float64, device=device), persistent=False)
        self.register_buffer("_w", t.tensor(self._weights * (np.pi / 2), dtype=t.float64, device=device), persistent=False)

    def alpha(self):
        return t.sigmoid(self.latent_alpha) * (self.alpha_hi - self.alpha_lo) + self.alpha_lo

    def scale(self):
        shift = float(np.log(np.expm1(1.0)))
        return (self.scale_init - self.scale_lo) * t.nn.functional.softplus(self.latent_scale + shift) + self.scale_lo

    def log_partition(self, alpha):
        """log of int exp(-rho(x, alpha, 1)) dx by Gauss-Legendre quadrature under x = tan(theta)."""
        a = alpha.to(t.float64).reshape(-1, 1)
        x = t.tan(self._th).reshape(1, -1)
        rho = lossfun(x, a, t.ones_like(a))
        integrand = t.exp(-rho) / t.cos(self._th).reshape(1, -1) ** 2
        return t.log(t.sum(integrand * self._w.reshape(1, -1), 1)).reshape(alpha.shape).to(alpha.dtype)

    def lossfun(self, x):
        alpha, scale = self.alpha(), self.scale()
        return lossfun(x, alpha, scale) + t.log(scale) + self.log_partition(alpha)
