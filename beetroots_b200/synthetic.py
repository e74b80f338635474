"""Synthetic inversion problems of the shapes named in BASELINE.json (SURVEY.md §8d recipe).

Mirrors the reference toy-case setup (simulations/astro/observation/abstract_toy_case.py:68-77
and data/toycases/nn_N30_fixed_angle.yaml): a smooth true map Theta* in scaled space,
observations ``y = max(omega, eps_m f(Theta*) + eps_a)`` with eps_a ~ N(0, sigma_a),
eps_m ~ logN(-sigma_m^2/2, sigma_m), sigma_a = 1.38715e-10, sigma_m = ln 1.1, omega = 3 sigma_a,
tau = [10, 1, 2, 4], 4-neighbour L2 prior, box = scaled images of
[0.1,10] x [1e5,1e9] x [1,1e5] x [1,40], Delta = 0.1, angle fixed at 0.

``f(Theta*)`` is supplied by the caller (the GPU forward map in bench.py, the oracle in CPU
tests): this module does no forward-map arithmetic.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

from .modelling import (L22DiscreteGradSpatialPrior, MixingModelsLikelihood, NeuralNetworkApprox,
                        Posterior, SmoothIndicatorPrior)
from .network import DenseMLPWeights, synthetic_weights

LOG10 = float(np.log(10.0))
# data/models/meudon_pdr_model_dense/scaler.pickle (StandardScaler on log10 P, log10 radm, log10 Avmax, angle)
SCALER_MEAN = np.array([0.0, 6.9990482, 2.4990482, 0.80041319, 30.0])
SCALER_STD = np.array([1.0 / LOG10, 1.24061804, 1.55073016, 0.49778857, 20.0])
IS_LOG = np.array([True, True, True, True, False])
LOWER_LIN = np.array([1.0e-1, 1.0e5, 1.0e0, 1.0e0, 0.0])
UPPER_LIN = np.array([1.0e1, 1.0e9, 1.0e5, 4.0e1, 60.0])


def lin_to_scaled(theta_lin: np.ndarray) -> np.ndarray:
    """space_transform/transform.py:31-50 for the 5 astro parameters."""
    x = np.where(IS_LOG, np.log10(np.where(IS_LOG, theta_lin, 1.0)), theta_lin)
    return (x - SCALER_MEAN) / SCALER_STD


@dataclass
class SyntheticProblem:
    posterior: Posterior
    theta_true: np.ndarray        # (N, D_sampling) scaled
    theta_init: np.ndarray        # (N, D_sampling) scaled
    H: int
    W: int
    sampler_kwargs: dict
    a0: Optional[np.ndarray] = None      # (L,) centres / radii (log10) of the likelihood's mixing window,
    a1: Optional[np.ndarray] = None      # the columns a0_best / a1_best of the reference's best_params.csv


def smooth_theta_map(H: int, W: int, seed: int = 1, amplitude: float = 1.1) -> np.ndarray:
    """Low-frequency sinusoid mixture in scaled space, (H*W, 4), row-major (X slow, Y fast)."""
    rng = np.random.default_rng(seed)
    xx, yy = np.meshgrid(np.arange(H) / max(H, 2), np.arange(W) / max(W, 2), indexing="ij")
    out = np.zeros((H * W, 4))
    for d in range(4):
        acc = np.zeros((H, W))
        for _ in range(3):
            fx, fy = rng.uniform(0.3, 2.0, size=2)
            ph = rng.uniform(0, 2 * np.pi)
            acc += rng.uniform(0.3, 1.0) * np.sin(2 * np.pi * (fx * xx + fy * yy) + ph)
        out[:, d] = amplitude * acc.reshape(-1) / np.abs(acc).max()
    return out


def make_problem(
    H: int,
    W: int,
    log_f_of_theta: Callable[[NeuralNetworkApprox, np.ndarray], np.ndarray],
    L: int = 10,
    k_mtm: int = 50,
    seed_obs: int = 2023,
    seed_net: int = 0,
    seed_map: int = 1,
    sigma_a: float = 1.38715e-10,
    sigma_m: float = float(np.log(1.1)),
    omega_factor: float = 3.0,
    use_next_nearest_neighbors: bool = False,
    net: Optional[DenseMLPWeights] = None,
    censor_boost: float = 1.0,
    init_jitter: float = 0.05,
) -> SyntheticProblem:
    """Build the synthetic cube.  ``censor_boost`` > 1 raises sigma_a (hence omega) so that a
    sizeable fraction of (n, l) is censored (config c3)."""
    N, D = H * W, 4
    net = synthetic_weights(seed=seed_net, n_out=L) if net is None else net
    fixed = {"kappa": None, "P": None, "radm": None, "Avmax": None,
             "angle": float(lin_to_scaled(np.array([1.0, 1e5, 1.0, 1.0, 0.0]))[4])}
    fm = NeuralNetworkApprox(net, fixed)
    theta_true = smooth_theta_map(H, W, seed_map)
    log_f = np.asarray(log_f_of_theta(fm, theta_true), dtype=np.float64)       # natural log, (N, L)
    f = np.exp(log_f)
    sa = sigma_a * censor_boost
    rng = np.random.default_rng(seed_obs)
    eps_a = rng.normal(0.0, sa, size=(N, L))
    eps_m = rng.lognormal(-(sigma_m ** 2) / 2, sigma_m, size=(N, L))
    omega = omega_factor * sa
    y = np.maximum(omega, eps_m * f + eps_a)
    rl = np.random.default_rng(seed_obs + 1)
    a0 = rl.uniform(-8.9, -8.4, size=L) + np.log10(censor_boost)                # range of data/toycases/best_params.csv
    a1 = rl.uniform(0.15, 0.85, size=L)
    lik = MixingModelsLikelihood(fm, D, L, N, y, sa, sigma_m, omega, a0=a0[None, :], a1=a1[None, :])
    X = np.repeat(np.arange(H), W)
    Y = np.tile(np.arange(W), H)
    prior_s = L22DiscreteGradSpatialPrior(D, N, X, Y, None, use_next_nearest_neighbors,
                                          np.array([10.0, 1.0, 2.0, 4.0]))
    lo = lin_to_scaled(LOWER_LIN)[:4]
    hi = lin_to_scaled(UPPER_LIN)[:4]
    prior_i = SmoothIndicatorPrior(D, N, 0.1, lo, hi)
    post = Posterior(D, L, N, lik, prior_s, prior_i)
    theta_init = theta_true + init_jitter * np.random.default_rng(seed_obs + 2).standard_normal((N, D))
    kw = dict(initial_step_size=1e-4, extreme_grad=1e-5, history_weight=0.99,
              selection_probas=[0.5, 0.5], k_mtm=k_mtm)
    return SyntheticProblem(post, theta_true, theta_init, H, W, kw, a0, a1)
