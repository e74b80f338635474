"""Oracle (test infrastructure): float64 restatement of Posterior.compute_all and its parts.

Follows, formula by formula, the reference functions on the hot path (SURVEY.md §8a A1-A10,
Appendix B).  Inputs are attribute containers with the reference attribute names (either the
mirrors in beetroots_b200/modelling.py or unmodified reference objects).
The only structural change: the O(N*E) boolean-mask loops of the priors
(l22_discrete_grad_prior.py:9-80) are evaluated through a neighbour table (same sums).
"""
from __future__ import annotations

import numpy as np
from scipy.special import log_ndtr

from . import mlp as _mlp

LOGE_10 = float(np.log(10.0))
_HALF_LOG_2PI = 0.5 * np.log(2.0 * np.pi)


# ----------------------------------------------------------------------------- forward map
def forward_map_compute_all(fm, Theta: np.ndarray, compute_derivatives: bool) -> dict:
    """neural_network_approx.py:152-305 (lin+log, no 2nd order).

    Returns log_f_Theta (M,L), f_Theta (M,L) and, with derivatives, grad_log_f_Theta /
    grad_f_Theta (M, D_sampling, L).
    """
    M = Theta.shape[0]
    Theta_combined = np.zeros((M, fm.D)) + fm.arr_fixed_values[None, :]          # :189-192
    for i, idx in enumerate(fm.list_indices_to_sample):
        Theta_combined[:, idx] += Theta[:, i]
    y, jac = _mlp.mlp_forward(fm.network, Theta_combined[:, 1:], with_jac=compute_derivatives)
    log_f = Theta_combined[:, 0][:, None] + y * LOGE_10                          # :198-200
    out = {"log_f_Theta": log_f, "f_Theta": np.exp(log_f)}
    if compute_derivatives:
        g_torch = np.transpose(jac, (0, 2, 1))                                   # (M, n_in, L) :208-214
        nn_cols = [d - 1 for d in fm.list_indices_to_sample if d >= 1]
        grad_log_f = np.zeros((M, fm.D_sampling, log_f.shape[1]))
        if 0 in fm.list_indices_to_sample:                                       # :216-228
            grad_log_f[:, 0, :] += 1.0
            grad_log_f[:, 1:, :] += g_torch[:, nn_cols, :] * LOGE_10
        else:
            grad_log_f[:, :, :] += g_torch[:, nn_cols, :] * LOGE_10
        out["grad_log_f_Theta"] = grad_log_f
        out["grad_f_Theta"] = grad_log_f * out["f_Theta"][:, None, :]           # :283-289
    return out


# ----------------------------------------------------------------------------- likelihood
def _norm_pdf_cdf_ratio(x):
    """likelihoods/utils.py:19-58."""
    return np.exp(-(x ** 2) / 2 - _HALF_LOG_2PI - log_ndtr(x))


def _P(u):      # approx_censored_add_mult.py:193  poly1d([-6, 15, -10, 0, 0, 1])
    return ((((-6.0 * u + 15.0) * u - 10.0) * u) * u) * u + 1.0


def _dP(u):
    return (((-30.0 * u + 60.0) * u - 30.0) * u) * u


def likelihood_eval(lik, fm_evals: dict, idx=None, compute_derivatives: bool = True) -> dict:
    """MixingModelsLikelihood.evaluate_all_nll_utils + neglog_pdf + gradient_neglog_pdf
    (approx_censored_add_mult.py:982-1246, 544-572, 633-743), literally, including the quirks
    listed in SURVEY.md §8a (y in the omega slot of nll_ac, ``y == omega`` in the gradient
    branch, sign of m_m in gradient_cost_mu, m_a == 0).

    ``idx`` = pixel index of every row group when evaluating candidates (rows are
    n_pix * k consecutive rows per pixel, :1003-1039); None = one row per pixel.
    Returns nll_full (M,L), nll_pix (M,), and grad (M,D) when derivatives are requested.
    """
    lf = fm_evals["log_f_Theta"]
    f = fm_evals["f_Theta"]
    M, L = lf.shape
    if idx is None:
        rows = slice(None)
    else:
        k = M // idx.size
        assert k * idx.size == M
        rows = np.repeat(idx, k)
    sigma_a, sigma_m = lik.sigma_a[rows], lik.sigma_m[rows]
    y, omega = lik.y[rows], lik.omega[rows]
    log_fm1, log_fp1 = lik.log_fm1[rows], lik.log_fp1[rows]
    with np.errstate(all="ignore"):
        # :999-1000 full maps use the stored self.log_y (which the model check leaves stale, my_sampler.py:176-177);
        # :1036 candidate rows recompute it
        log_y = lik.log_y[rows] if idx is None else np.log(y)
        censored_mask = y <= omega                                              # :1042
        sigma_m2 = sigma_m ** 2
        log_combination = np.log(sigma_a) - lf - sigma_m2 / 2                   # :1047-1049
        E = np.exp(-2 * log_combination)                                        # :1050
        exp_sm2 = np.exp(sigma_m2)
        m_a = np.zeros_like(y)                                                  # :1055
        s_a = sigma_a * np.sqrt((exp_sm2 - 1) * np.exp(2 * (lf - np.log(sigma_a))) + 1)  # :1056-1060
        s_a2 = s_a ** 2
        m_m = -0.5 * (sigma_m2 + np.log(1 + 1 / E))                             # :1063
        s_m2 = -2 * m_m
        s_m = np.sqrt(s_m2)
        # mixing weight :341-352
        u = (lf - log_fm1) / (log_fp1 - log_fm1)
        lam = np.where(lf <= log_fm1, 1.0, np.where(lf >= log_fp1, 0.0, _P(u)))
        nll_au = 0.5 * ((y - f - m_a) / s_a) ** 2 + np.log(s_a)                 # :14-19
        nll_ac = np.nan_to_num(-log_ndtr((y - f - m_a) / s_a))                  # :574-585 with y (:1212)
        nll_mu = 0.5 * ((log_y - lf - m_m) / s_m) ** 2 + np.log(s_m)            # :22-33
        nll_mc = np.nan_to_num(-log_ndtr((np.log(omega) - lf - m_m) / s_m))     # :604-616
        nll_full = lam * np.where(censored_mask, nll_ac, nll_au) + (1 - lam) * np.where(
            censored_mask, nll_mc, nll_mu)                                      # :553-557
        nll_full = np.nan_to_num(nll_full)
        nll_full = nll_full - (np.log(sigma_a) + np.log(sigma_m))               # :560
    out = {"nll_full": nll_full, "nll_pix": nll_full.sum(axis=1)}
    if not compute_derivatives:
        return out
    assert idx is None, "reference computes derivatives on full maps only"
    glf = fm_evals["grad_log_f_Theta"]                                          # (M, D, L)
    gf = fm_evals["grad_f_Theta"]
    e = lambda a: a[:, None, :]                                                 # noqa: E731
    with np.errstate(all="ignore"):
        grad_m_a = np.zeros_like(gf)                                            # :1073
        grad_s_a2 = 2 * e(f * (exp_sm2 - 1)) * gf                               # :1078-1082
        grad_s_a = e(1 / (2 * s_a)) * grad_s_a2                                 # :1095-1097
        grad_m_m = e(1 / (1 + E) / f) * gf                                      # :1107-1111
        grad_s_m2 = -2 * grad_m_m
        grad_s_m = e(1 / (2 * s_m)) * grad_s_m2                                 # :1132-1136
        grad_lam = np.where(e((lf <= log_fm1) | (lf >= log_fp1)), 0.0,
                            glf / e(log_fp1 - log_fm1) * e(_dP(u)))             # :385-405
        # gradient_cost_au :36-59
        r = f - m_a - y
        g_au = 0.5 * grad_s_a2 / e(s_a2) + 0.5 * e(r / s_a2 ** 2) * (
            2 * e(s_a2) * (gf + grad_m_a) - e(r) * grad_s_a2)
        # gradient_cost_mu :62-92  (note lf - m_m - log_y)
        rm = lf - m_m - lik.log_y
        g_mu = 0.5 * grad_s_m2 / e(s_m2) + np.where(
            e(np.isnan(lik.log_y)) | np.zeros_like(glf, dtype=bool), 0.0,
            0.5 * e(rm / s_m2 ** 2) * (2 * e(s_m2) * (glf + grad_m_m) - e(rm) * grad_s_m2))
        # gradient_neglog_pdf_ac :676-692
        t_a = f + m_a - lik.omega
        g_ac = e(_norm_pdf_cdf_ratio(-t_a / s_a)) * e(1 / s_a2) * (
            (gf + grad_m_a) * e(s_a) - e(t_a) * grad_s_a)
        # gradient_neglog_pdf_mc :709-728
        t_m = lf + m_m - np.log(lik.omega)
        g_mc = e(_norm_pdf_cdf_ratio(-t_m / s_m)) * e(1 / s_m2) * (
            (glf + grad_m_m) * e(s_m) - e(t_m) * grad_s_m)
        grad_ = np.where(                                                       # :654-667
            e(lik.y == lik.omega),
            e(lam) * g_ac + grad_lam * e(nll_ac) + e(1 - lam) * g_mc - grad_lam * e(nll_mc),
            e(lam) * g_au + grad_lam * e(nll_au) + e(1 - lam) * g_mu - grad_lam * e(nll_mu),
        )
        grad_ = np.nan_to_num(grad_)                                            # :673
    out["grad"] = grad_.sum(axis=2)                                             # :674
    return out


# ----------------------------------------------------------------------------- priors
def spatial_sums(Theta: np.ndarray, nbr: np.ndarray, idx_pix=None, values=None):
    """Per-pixel sums over the neighbour set V_n (l22_discrete_grad_prior.py:9-80).

    Theta (N, D).  ``values`` (n_pix, ..., D) optional replacement values for the centre pixel
    (MTM candidates); defaults to Theta[idx_pix].
    Returns (hadamard, laplacian): sum_i (v - theta_i)^2 and sum_i (v - theta_i), shaped like values.
    """
    if idx_pix is None:
        idx_pix = np.arange(Theta.shape[0])
    centre = Theta[idx_pix] if values is None else values
    had = np.zeros_like(centre)
    lap = np.zeros_like(centre)
    extra = (None,) * (centre.ndim - 2)
    for j in range(nbr.shape[1]):
        col = nbr[idx_pix, j]
        ok = col >= 0
        th = Theta[np.where(ok, col, 0)]                                        # (n_pix, D)
        th = th[(slice(None),) + extra + (slice(None),)]
        diff = (centre - th) * ok[(slice(None),) + (None,) * (centre.ndim - 1)]
        had += diff ** 2
        lap += diff
    return had, lap


def indicator_penalty(Theta, prior):
    """smooth_indicator_prior.py:35-54 (penalty_one_pix): (..., D) -> (...)"""
    lo, hi, delta = prior.lower_bounds, prior.upper_bounds, prior.indicator_margin_scale
    d = np.where(Theta > hi, (Theta - hi) / delta, np.where(Theta < lo, (lo - Theta) / delta, 0.0))
    return (d ** 4).sum(axis=-1)


def indicator_gradient(Theta, prior):
    """smooth_indicator_prior.py:57-82 (gradient_penalty)."""
    lo, hi, delta = prior.lower_bounds, prior.upper_bounds, prior.indicator_margin_scale
    d = np.where(Theta > hi, Theta - hi, np.where(Theta < lo, Theta - lo, 0.0))
    return 4 / delta ** 4 * d ** 3


# ----------------------------------------------------------------------------- posterior
def compute_all(post, Theta: np.ndarray, nbr=None) -> dict:
    """Posterior.compute_all (posterior.py:332-446) without 2nd-order terms."""
    assert np.sum(np.isnan(Theta)) == 0
    lik = post.likelihood
    fm = forward_map_compute_all(lik.forward_map, Theta, True)
    le = likelihood_eval(lik, fm, None, True)
    N = Theta.shape[0]
    ind = indicator_penalty(Theta, post.prior_indicator) if post.prior_indicator is not None else np.zeros(N)
    grad = le["grad"].copy()
    if post.prior_spatial is not None:
        had, lap = spatial_sums(Theta, nbr)
        w = post.prior_spatial.weights
        chrom = (w * had).sum(axis=1)                                           # l22:107,125-128
        glob = 0.5 * chrom
        grad += w[None, :] * 2 * lap                                            # l22:137-139
    else:
        chrom = np.zeros(N)
        glob = np.zeros(N)
    if post.prior_indicator is not None:
        grad += indicator_gradient(Theta, post.prior_indicator)
    grad = np.nan_to_num(grad)                                                  # posterior.py:240
    obj_c = le["nll_pix"] + chrom + ind                                         # :402-411
    obj_g = le["nll_pix"] + glob + ind
    return {
        "Theta": Theta, "forward_map_evals": fm, "nll_full": le["nll_full"], "nll_pix": le["nll_pix"],
        "grad_lik": le["grad"],
        "objective_pix_chromatic": obj_c, "objective_pix_global": obj_g,
        "objective_chromatic": obj_c.sum(), "objective_global": obj_g.sum(), "grad": grad,
    }
