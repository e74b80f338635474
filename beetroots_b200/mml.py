"""Regularisation-weight (tau) update of the sampler: host mirror of the scalar optimiser.

The reference re-estimates the spatial-prior weight by projected gradient ascent on the marginal likelihood
(``EBayesMMLELogRate``, sampler/utils/mml.py:123-179, 240-456), fed once per iteration with the *unweighted* spatial
potential g(Theta) (``Sampler.sample_regu_hyperparams``, sampler/abstract_sampler.py:172-210).  g is a sum over the
whole map -- on the GPU path it is one reduction kernel (``bt_spatial_prior_unweighted``) plus, when the map is sharded,
one all-reduce; the scalar recursion below is all that is left on the host.  Same class name, constructor arguments,
attributes (``mean_theta``, ``sum_weights`` ...) and ``update`` signature as the reference.
"""
from __future__ import annotations

from typing import Union

import numpy as np


class EBayesMMLELogRate:
    def __init__(self, scale: float, N0: Union[int, float], N1: Union[int, float], dim: int, vmin: float = 1e-8,
                 vmax: float = 1e8, homogeneity: float = 1.0, exponent: float = 0.8):
        assert isinstance(N0, (int, np.integer)) or np.isposinf(N0), f"N0 should be a finite int or +np.inf, but it is {N0}"
        assert isinstance(N1, (int, np.integer)) or np.isposinf(N1), f"N1 should be a finite int or +np.inf, but it is {N1}"
        self.scale = scale
        self.N0 = N0
        self.N1 = N1
        self.dim = dim
        self.vmin = np.log(vmin)            # the log-rate variant clips in log scale (mml.py:445-454)
        self.vmax = np.log(vmax)
        self.homogeneity = homogeneity
        self.exponent = exponent
        self._c = scale / dim
        self.mean_theta = 0.0
        self.mean_theta_old = 0.0
        self.sum_weights = 0.0

    def _compute_stepsize(self, iteration: int) -> float:
        return self._c / ((iteration - self.N0 + 1) ** self.exponent)                      # mml.py:323-337

    def _update_mean_theta(self, theta: float, iteration: int, stepsize: float) -> None:  # mml.py:339-358
        self.mean_theta_old = self.mean_theta
        step = stepsize if iteration > self.N1 else 1
        self.mean_theta = self.sum_weights * self.mean_theta + step * theta
        self.sum_weights += step
        self.mean_theta /= self.sum_weights

    @staticmethod
    def _update_step(theta, stepsize, vmin, vmax, dim, gX, homogeneity) -> float:          # pgd_rate_log, mml.py:123-179
        log_theta = np.log(theta)
        return float(np.exp(np.maximum(np.minimum(log_theta + stepsize * (dim / homogeneity - gX * theta), vmax), vmin)))

    def update(self, theta: float, iteration: int, gX: float) -> float:                    # mml.py:373-404
        if iteration + 1 > self.N0:
            stepsize = self._compute_stepsize(iteration)
            theta = self._update_step(theta, stepsize, self.vmin, self.vmax, self.dim, gX, self.homogeneity)
            self._update_mean_theta(theta, iteration, stepsize)
        return theta
