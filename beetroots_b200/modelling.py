"""Host-side mirror of the reference modelling interface for the sampler hot path.

Same class names, constructor-argument meaning and attribute names as the reference
(/root/reference/beetroots/modelling/**), but *arithmetic-free*: these objects only hold the
inputs; every evaluation happens in the CUDA library (beetroots_b200/csrc) behind
``B200Sampler`` / ``DevicePosterior``.  ``DevicePosterior.from_posterior`` reads attributes by
name, so an unmodified reference ``Posterior`` (with its ``MixingModelsLikelihood``,
``L22DiscreteGradSpatialPrior``, ``SmoothIndicatorPrior``, ``NeuralNetworkApprox``) can be
passed instead of these mirrors -- that is the drop-in boundary.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from .network import DenseMLPWeights

LOGE_10 = float(np.log(10.0))


class NeuralNetworkApprox:
    """Mirror of forward_maps/neural_network_approx.py:12-105 (attributes only).

    ``ln f = LOGE_10 * NN(theta_full[1:]) + theta_full[0]`` where theta_full combines the fixed
    values and the sampled columns (neural_network_approx.py:186-200).
    """

    LOGE_10 = LOGE_10

    def __init__(self, network: DenseMLPWeights, dict_fixed_values_scaled: Dict[str, Optional[float]]):
        self.network = network
        self.D_no_kappa = network.n_in
        self.D = network.n_in + 1
        self.L = network.n_out
        # forward_maps/abstract_base.py:44-73
        vals = list(dict_fixed_values_scaled.values())
        assert len(vals) == self.D, (len(vals), self.D)
        self.dict_fixed_values_scaled = dict_fixed_values_scaled
        self.list_indices_to_sample = [d for d, v in enumerate(vals) if v is None]
        self.list_indices_fixed = [d for d, v in enumerate(vals) if v is not None]
        self.D_sampling = len(self.list_indices_to_sample)
        self.arr_fixed_values = np.array([0.0 if v is None else float(v) for v in vals])
        self.list_indices_to_sample_for_nn = [d - 1 for d in self.list_indices_to_sample if d >= 1]


class MixingModelsLikelihood:
    """Mirror of likelihoods/approx_censored_add_mult.py:95-266 (attributes only)."""

    def __init__(self, forward_map, D: int, L: int, N: int, y, sigma_a, sigma_m, omega,
                 log_fm1=None, log_fp1=None, a0=None, a1=None):
        y = np.asarray(y, dtype=np.float64)
        if y.shape != (N, L):
            raise ValueError(f"y must have the shape (N, L) = ({N}, {L}) elements")
        self.forward_map, self.D, self.L, self.N = forward_map, D, L, N
        self.y = y
        self.log_y = np.log(y)
        self.sigma_a = self._full(sigma_a, N, L)
        self.sigma_m = self._full(sigma_m, N, L)
        self.omega = self._full(omega, N, L)
        if log_fm1 is None:
            # approx_censored_add_mult.py:225-265: centre a0, radius a1, in log10 -> natural log
            a0 = np.broadcast_to(np.asarray(a0, dtype=np.float64), (N, L))
            a1 = np.broadcast_to(np.asarray(a1, dtype=np.float64), (N, L))
            log_fm1, log_fp1 = a0 * LOGE_10 - a1 * LOGE_10, a0 * LOGE_10 + a1 * LOGE_10
        self.log_fm1 = self._full(log_fm1, N, L)
        self.log_fp1 = self._full(log_fp1, N, L)

    @staticmethod
    def _full(v, N, L):
        v = np.asarray(v, dtype=np.float64)
        if v.ndim == 0:
            return np.full((N, L), float(v))
        assert v.shape == (N, L), v.shape
        return np.ascontiguousarray(v)


def saver_nll_utils(likelihood, log_f: np.ndarray) -> dict:
    """The per-(pixel, line) entries of ``nll_utils`` that the reference saver stores next to the chain
    (sampler/saver/my_saver.py:57-61, 128-130: every key without ``nll_`` / ``grad`` / ``hess_diag``), evaluated on the host in
    float64 from ln f of the saved iterate -- saver glue, NOT the hot path (the device fuses these quantities into the
    likelihood kernels and never stores them).  Formulas: likelihoods/approx_censored_add_mult.py:1041-1066 and :341-352;
    ``likelihood`` is the reference object or its mirror (same attribute names)."""
    N, L = log_f.shape
    full = lambda a: np.broadcast_to(np.asarray(a, dtype=np.float64), (N, L))      # noqa: E731
    y, sa, sm, om = full(likelihood.y), full(likelihood.sigma_a), full(likelihood.sigma_m), full(likelihood.omega)
    fm1, fp1 = full(likelihood.log_fm1), full(likelihood.log_fp1)
    sm2 = sm ** 2
    s_a = sa * np.sqrt(np.expm1(sm2) * np.exp(2.0 * (log_f - np.log(sa))) + 1.0)
    m_m = -0.5 * (sm2 + np.log1p(np.exp(2.0 * (np.log(sa) - log_f - sm2 / 2.0))))
    s_m2 = -2.0 * m_m
    u = np.clip((log_f - fm1) / (fp1 - fm1), 0.0, 1.0)
    lam = np.where(log_f <= fm1, 1.0, np.where(log_f >= fp1, 0.0, ((-6.0 * u + 15.0) * u - 10.0) * u ** 3 + 1.0))
    return {"censored_mask": (y <= om) * 1, "sigma_a": sa * 1, "sigma_m": sm * 1, "m_a": np.zeros((N, L)), "s_a": s_a,
            "s_a2": s_a ** 2, "m_m": m_m, "s_m2": s_m2, "s_m": np.sqrt(s_m2), "lambda_": lam}


class SmoothIndicatorPrior:
    """Mirror of priors/smooth_indicator_prior.py:113-166 (attributes only)."""

    def __init__(self, D: int, N: int, indicator_margin_scale: float, lower_bounds, upper_bounds,
                 list_idx_sampling: Optional[Sequence[int]] = None):
        lower_bounds = np.asarray(lower_bounds, dtype=np.float64)
        upper_bounds = np.asarray(upper_bounds, dtype=np.float64)
        if list_idx_sampling is not None:
            lower_bounds = lower_bounds[list(list_idx_sampling)]
            upper_bounds = upper_bounds[list(list_idx_sampling)]
        assert lower_bounds.size == D and upper_bounds.size == D
        self.D, self.N = D, N
        self.indicator_margin_scale = float(indicator_margin_scale)
        self.lower_bounds, self.upper_bounds = lower_bounds, upper_bounds


def build_list_edges(X: np.ndarray, Y: np.ndarray, idx: np.ndarray,
                     use_next_nearest_neighbors: bool) -> np.ndarray:
    """Edge list of the pixel graph; same offsets and ordering as
    priors/abstract_spatial_prior.py:13-55 ((1,0),(0,1)[,(1,1),(-1,1)], pixels in input order),
    built with a dictionary lookup instead of pandas."""
    pos = {(int(x), int(y)): int(i) for x, y, i in zip(X, Y, idx)}
    offsets = [(1, 0), (0, 1), (1, 1), (-1, 1)] if use_next_nearest_neighbors else [(1, 0), (0, 1)]
    edges = []
    for x, y, i in zip(X, Y, idx):
        for dx, dy in offsets:
            j = pos.get((int(x) + dx, int(y) + dy))
            if j is not None:
                edges.append((int(i), j))
    return np.asarray(edges, dtype=np.int64).reshape(-1, 2)


def build_sites(X, Y, idx, use_next_nearest_neighbors: bool) -> Dict[int, np.ndarray]:
    """Checkerboard colour classes, priors/abstract_spatial_prior.py:125-153."""
    X, Y, idx = np.asarray(X), np.asarray(Y), np.asarray(idx)
    colour = 2 * (X % 2) + Y % 2 if use_next_nearest_neighbors else (X + Y) % 2
    n_sites = 4 if use_next_nearest_neighbors else 2
    return {k: np.sort(idx[colour == k]) for k in range(n_sites)}


def build_neighbor_table(list_edges: np.ndarray, N: int, max_degree: int = 8) -> np.ndarray:
    """(N, max_degree) int32 neighbour table, -1 padded.

    Neighbour order = order of appearance in ``list_edges`` (what
    sampler/utils/utils.py:100-125 ``get_neighboring_pixels`` returns), which fixes the meaning
    of the neighbour-subset bit masks used by the MTM proposal on both the GPU and the oracle.
    """
    nbr = -np.ones((N, max_degree), dtype=np.int32)
    e = np.asarray(list_edges, dtype=np.int64).reshape(-1, 2)
    if e.size == 0:
        return nbr
    # directed copies interleaved in edge order: (a->b) then (b->a) for every edge
    src = e.reshape(-1)                       # a0,b0,a1,b1,...
    dst = e[:, ::-1].reshape(-1)              # b0,a0,b1,a1,...
    order = np.argsort(src, kind="stable")    # stable: keeps edge-list order inside each pixel
    src_s, dst_s = src[order], dst[order]
    start = np.searchsorted(src_s, src_s, side="left")
    slot = np.arange(src_s.size) - start
    assert slot.max() < max_degree, "pixel degree exceeds max_degree"
    nbr[src_s, slot] = dst_s
    return nbr


class L22DiscreteGradSpatialPrior:
    """Mirror of priors/l22_discrete_grad_prior.py:83-156 + abstract_spatial_prior.py:58-123."""

    def __init__(self, D: int, N: int, X, Y, idx=None, use_next_nearest_neighbors: bool = False,
                 initial_regu_weights=None):
        X, Y = np.asarray(X, dtype=np.int64), np.asarray(Y, dtype=np.int64)
        idx = np.arange(N, dtype=np.int64) if idx is None else np.asarray(idx, dtype=np.int64)
        assert X.shape == (N,) and Y.shape == (N,) and idx.shape == (N,)
        self.D, self.N = D, N
        self.X, self.Y, self.idx = X, Y, idx
        self.use_next_nearest_neighbors = bool(use_next_nearest_neighbors)
        self.list_edges = build_list_edges(X, Y, idx, self.use_next_nearest_neighbors)
        self.dict_sites = build_sites(X, Y, idx, self.use_next_nearest_neighbors)
        w = np.ones(D) if initial_regu_weights is None else np.asarray(initial_regu_weights, dtype=np.float64)
        assert w.shape == (D,)
        self.initial_weights = w.copy()
        self.weights = w.copy()


class Posterior:
    """Mirror of modelling/posterior.py:6-57 (attributes only)."""

    __slots__ = ("D", "L", "N", "likelihood", "prior_spatial", "prior_indicator", "dict_sites")

    def __init__(self, D: int, L: int, N: int, likelihood, prior_spatial=None, prior_indicator=None,
                 separable: bool = True, dict_sites: Optional[Dict[int, np.ndarray]] = None):
        self.D, self.L, self.N = D, L, N
        self.likelihood = likelihood
        self.prior_spatial = prior_spatial
        self.prior_indicator = prior_indicator
        if dict_sites is not None:
            self.dict_sites = dict_sites
        elif prior_spatial is not None:
            self.dict_sites = prior_spatial.dict_sites
        elif separable:
            self.dict_sites = {0: np.arange(N)}
        else:
            self.dict_sites = {n: np.array([n]) for n in range(N)}
