"""Regularisation-weight (tau) update of the sampler: the scalar recursion that stays on the host.

The reference re-estimates the spatial-prior weight by projected gradient ascent on the marginal likelihood
(``EBayesMMLELogRate``, sampler/utils/mml.py:123-179, 240-456), fed once per iteration with the *unweighted* spatial
potential g(Theta) (``Sampler.sample_regu_hyperparams``, sampler/abstract_sampler.py:172-210).  g is a sum over the whole map:
on the GPU path one reduction kernel (``bt_spatial_prior_unweighted``) plus, when the map is sharded, one all-reduce.  What is
left is this recursion on one scalar, written here as a single ``update`` (the reference spreads it over an abstract base and
three helpers); the public surface a caller touches is the reference's: class name, constructor arguments, ``update(theta,
iteration, gX)``, and the running-average attributes ``mean_theta`` / ``mean_theta_old`` / ``sum_weights``.  Pinned against the
unmodified reference class by tests/golden/mml_golden.json.
"""
from __future__ import annotations

import math
from typing import Union

import numpy as np


def _is_count_or_inf(x) -> bool:
    return isinstance(x, (int, np.integer)) or (isinstance(x, (float, np.floating)) and math.isinf(x) and x > 0)


class EBayesMMLELogRate:
    """tau_{t+1} = clip( tau_t * exp(delta_t (dim / homogeneity - g(Theta_t) tau_t)) ), delta_t = (scale / dim) (t - N0 + 1)^-exponent,
    clipped to [vmin, vmax] in log scale; a weighted running mean of the iterates is kept alongside (weights 1 up to N1,
    delta_t afterwards)."""

    def __init__(self, scale: float, N0: Union[int, float], N1: Union[int, float], dim: int, vmin: float = 1e-8,
                 vmax: float = 1e8, homogeneity: float = 1.0, exponent: float = 0.8):
        for name, val in (("N0", N0), ("N1", N1)):
            assert _is_count_or_inf(val), f"{name} should be a finite int or +np.inf, but it is {val}"
        self.scale, self.dim, self.homogeneity, self.exponent = scale, dim, homogeneity, exponent
        self.N0, self.N1 = N0, N1
        self.vmin, self.vmax = math.log(vmin), math.log(vmax)      # bounds of log tau (mml.py:445-454)
        self.mean_theta = self.mean_theta_old = self.sum_weights = 0.0

    def update(self, theta: float, iteration: int, gX: float) -> float:
        if iteration + 1 <= self.N0:                               # warm-up: the weight is left alone (mml.py:373-404)
            return theta
        delta = (self.scale / self.dim) / (iteration - self.N0 + 1) ** self.exponent          # mml.py:323-337
        log_tau = math.log(theta) + delta * (self.dim / self.homogeneity - gX * theta)         # pgd_rate_log, mml.py:123-179
        theta = float(math.exp(min(max(log_tau, self.vmin), self.vmax)))
        w = delta if iteration > self.N1 else 1                    # running mean of the iterates (mml.py:339-358)
        self.mean_theta_old = self.mean_theta
        total = self.sum_weights + w
        self.mean_theta = (self.sum_weights * self.mean_theta + w * theta) / total
        self.sum_weights = total
        return theta
