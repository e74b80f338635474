"""DevicePosterior: the reference ``Posterior`` resident on one B200, behind the C ABI.

Reads everything it needs *from the reference objects by attribute name* (SURVEY.md §7 step 2):
``posterior.likelihood.{y, sigma_a, sigma_m, omega, log_fm1, log_fp1, forward_map}``,
``posterior.prior_spatial.{weights, initial_weights, list_edges, dict_sites,
use_next_nearest_neighbors}``, ``posterior.prior_indicator.{lower_bounds, upper_bounds,
indicator_margin_scale}``, ``forward_map.{network, arr_fixed_values, list_indices_to_sample}``
and uploads them once.  ``Posterior`` has ``__slots__`` (modelling/posterior.py:7-15), so the
device handle lives here and not on the reference object.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import _lib as L
from .modelling import build_neighbor_table
from .network import DenseMLPWeights

MLP_FP32 = 0
MLP_BF16_TC = 1


def _network_weights(network) -> DenseMLPWeights:
    """Accept our container or an nnbma-like torch module (PolynomialNetwork -> DenselyConnected)."""
    if isinstance(network, DenseMLPWeights):
        return network
    if hasattr(network, "to_dense_mlp_weights"):
        return network.to_dense_mlp_weights()
    # best-effort extraction from an nnbma PolynomialNetwork: poly.{means,stds}, subnetwork.layers.N[.0].{weight,bias},
    # subnetwork.output_layer.{weight,bias} (key names of data/models/*/parameters.pth / state_dict.pth)
    from .network import poly_exponents
    sd = {k: v.detach().cpu().double().numpy() for k, v in network.state_dict().items()}
    means = sd["poly.means"]
    stds = sd["poly.stds"]
    order = int(getattr(network, "order", 3))
    n_in = int(network.input_features)
    ws, bs, i = [], [], 0
    while True:
        for key in (f"subnetwork.layers.{i}.weight", f"subnetwork.layers.{i}.0.weight"):
            if key in sd:
                ws.append(sd[key])
                bs.append(sd[key.replace("weight", "bias")])
                break
        else:
            break
        i += 1
    w_out, b_out = sd["subnetwork.output_layer.weight"], sd["subnetwork.output_layer.bias"]
    sub = getattr(getattr(network, "subnetwork", network), "current_output_subset_indices", None)
    if sub is not None:
        w_out, b_out = w_out[list(sub)], b_out[list(sub)]
    net = DenseMLPWeights(n_in, order, poly_exponents(n_in, order), means, stds, ws, bs, w_out, b_out)
    net.validate()
    return net


def host_inputs_from_posterior(posterior, dict_sites=None) -> dict:
    """Everything the device needs, read from a (reference or mirror) ``Posterior`` by attribute name -- pure host
    code, no CUDA.  This is the drop-in boundary: tests/test_reference_dropin.py runs it on UNMODIFIED reference
    objects (build container) and on the mirrors and checks that both give the same arrays."""
    lik = posterior.likelihood
    fm = lik.forward_map
    N, D, Lines = int(posterior.N), int(posterior.D), int(posterior.L)
    ps, pi = posterior.prior_spatial, posterior.prior_indicator
    if ps is not None:
        nnn = bool(getattr(ps, "use_next_nearest_neighbors", True))
        max_degree = 8 if nnn else 4
        nbr = build_neighbor_table(ps.list_edges, N, max_degree)
    else:
        max_degree, nbr = 1, -np.ones((N, 1), dtype=np.int32)
    net = _network_weights(fm.network)
    assert net.n_out == Lines, "restrict the network to the fitted lines first"
    obs = {k: L.as_f64(np.broadcast_to(np.asarray(getattr(lik, k), dtype=np.float64), (N, Lines)))
           for k in ("y", "sigma_a", "sigma_m", "omega", "log_fm1", "log_fp1")}
    sites = dict_sites if dict_sites is not None else posterior.dict_sites
    return {
        "N": N, "D": D, "L": Lines, "max_degree": max_degree, "nbr": L.as_i32(nbr), "net": net,
        "fixed": L.as_f64(fm.arr_fixed_values),
        "sampled_idx": L.as_i32(np.array(fm.list_indices_to_sample, dtype=np.int32)),
        "obs": obs,
        "site_keys": list(sites.keys()),
        "site_pixels": [L.as_i32(np.asarray(sites[k])) for k in sites.keys()],
        "tau": None if ps is None else L.as_f64(ps.weights), "tau_init": None if ps is None else L.as_f64(ps.initial_weights),
        "lower": None if pi is None else L.as_f64(pi.lower_bounds), "upper": None if pi is None else L.as_f64(pi.upper_bounds),
        "delta": None if pi is None else float(pi.indicator_margin_scale),
    }


class DevicePosterior:
    def __init__(self, ctx, N, D, n_lines, n_sites, site_keys, site_pixels, max_degree):
        self._ctx = ctx
        self.N, self.D, self.L = N, D, n_lines
        self.n_sites = n_sites
        self.site_keys = site_keys
        self.site_pixels = site_pixels      # list of np arrays (local pixel ids)
        self.max_degree = max_degree
        self.has_spatial = self.has_indicator = True
        self._lib = L.load()

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_posterior(cls, posterior, device: int = 0, global_id: Optional[np.ndarray] = None,
                       dict_sites: Optional[Dict[int, np.ndarray]] = None, mlp_mode: int = MLP_FP32,
                       stream: int = 0) -> "DevicePosterior":
        lib = L.load()
        hi = host_inputs_from_posterior(posterior, dict_sites)
        N, D, Lines, max_degree, nbr = hi["N"], hi["D"], hi["L"], hi["max_degree"], hi["nbr"]
        ps, pi = posterior.prior_spatial, posterior.prior_indicator
        ctx = C.c_void_p()
        L.check(lib.bt_ctx_create(C.byref(ctx), device, N, D, Lines, max_degree), "bt_ctx_create")
        self = cls(ctx, N, D, Lines, 0, [], [], max_degree)
        try:
            if stream:
                L.check(lib.bt_set_stream(ctx, stream), "bt_set_stream")
            net = hi["net"]
            nb = len(net.weights)
            Ws = [L.as_f64(w) for w in net.weights]
            Bs = [L.as_f64(b) for b in net.biases]
            WP = (L.c_double_p * max(nb, 1))(*[L.dptr(w) for w in Ws])
            BP = (L.c_double_p * max(nb, 1))(*[L.dptr(b) for b in Bs])
            exps = L.as_i32(net.exponents)
            means, stds = L.as_f64(net.poly_means), L.as_f64(net.poly_stds)
            bnew = L.as_i32(np.array(net.block_new, dtype=np.int32))
            w_out, b_out = L.as_f64(net.w_out), L.as_f64(net.b_out)
            L.check(lib.bt_upload_network(ctx, net.n_in, net.order, net.n_feat0, L.dptr(exps, C.c_int32), L.dptr(means),
                                          L.dptr(stds), nb, L.dptr(bnew, C.c_int32), WP, BP, L.dptr(w_out),
                                          L.dptr(b_out), net.n_out), "bt_upload_network")
            fixed, idx = hi["fixed"], hi["sampled_idx"]
            L.check(lib.bt_upload_forward_map(ctx, fixed.size, L.dptr(fixed), idx.size, L.dptr(idx, C.c_int32)),
                    "bt_upload_forward_map")
            arrs = [hi["obs"][k] for k in ("y", "sigma_a", "sigma_m", "omega", "log_fm1", "log_fp1")]
            L.check(lib.bt_upload_observations(ctx, *[L.dptr(a) for a in arrs]), "bt_upload_observations")
            self._upload_priors(ps, pi)
            self.site_keys = hi["site_keys"]
            self.site_pixels = hi["site_pixels"]
            self.n_sites = len(self.site_keys)
            sizes = L.as_i32(np.array([p.size for p in self.site_pixels], dtype=np.int32))
            pix = L.as_i32(np.concatenate(self.site_pixels)) if self.n_sites else np.zeros(0, np.int32)
            gid = None if global_id is None else np.ascontiguousarray(global_id, dtype=np.int64)
            L.check(lib.bt_upload_graph(ctx, L.dptr(L.as_i32(nbr), C.c_int32), self.n_sites, L.dptr(sizes, C.c_int32),
                                        L.dptr(pix, C.c_int32), L.dptr(gid, C.c_int64)), "bt_upload_graph")
            if mlp_mode != MLP_FP32:
                L.check(lib.bt_set_mlp_mode(ctx, mlp_mode), "bt_set_mlp_mode")
        except Exception:
            lib.bt_ctx_destroy(ctx)
            self._ctx = None
            raise
        return self

    def _upload_priors(self, ps, pi):
        D = self.D
        tau = L.as_f64(ps.weights) if ps is not None else np.zeros(D)
        tau0 = L.as_f64(ps.initial_weights) if ps is not None else np.zeros(D)
        lo = L.as_f64(pi.lower_bounds) if pi is not None else np.zeros(D)
        hi = L.as_f64(pi.upper_bounds) if pi is not None else np.zeros(D)
        delta = float(pi.indicator_margin_scale) if pi is not None else 1.0
        self.has_spatial, self.has_indicator = ps is not None, pi is not None
        L.check(self._lib.bt_upload_priors(self._ctx, int(ps is not None), L.dptr(tau), L.dptr(tau0), int(pi is not None),
                                           L.dptr(lo), L.dptr(hi), delta), "bt_upload_priors")

    def update_spatial_weights(self, ps, pi):
        """posterior.prior_spatial.weights changed (sampler/my_sampler.py:415)."""
        self._upload_priors(ps, pi)

    def close(self):
        if self._ctx is not None:
            self._lib.bt_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ state
    def set_mlp_mode(self, mode: int):
        L.check(self._lib.bt_set_mlp_mode(self._ctx, mode), "bt_set_mlp_mode")

    def set_state(self, Theta=None, v=None, j=None):
        Theta, v, j = L.as_f64(Theta), L.as_f64(v), L.as_i32(j)
        for a in (Theta, v, j):
            assert a is None or a.size == self.N * self.D
        L.check(self._lib.bt_set_state(self._ctx, L.dptr(Theta), L.dptr(v), L.dptr(j, C.c_int32)), "bt_set_state")

    def get_state(self, want_v=True, want_j=True, reuse=False):
        """(Theta, v, j) of the resident state as float64 / int32 host arrays.  ``reuse``: return views of buffers owned by
        this object, overwritten by the next call (the sampler's per-iteration hand-over to the saver, which copies: a fresh
        33 MB array at c4 costs ~8 ms of page faults on first touch)."""
        if reuse:
            if getattr(self, "_hb", None) is None:
                self._hb = (np.empty((self.N, self.D)), np.empty((self.N, self.D)), np.empty((self.N, self.D), dtype=np.int32))
            Theta, v, j = self._hb[0], (self._hb[1] if want_v else None), (self._hb[2] if want_j else None)
            L.check(self._lib.bt_get_state(self._ctx, L.dptr(Theta), L.dptr(v), L.dptr(j, C.c_int32)), "bt_get_state")
            return Theta, v, j
        Theta = np.empty((self.N, self.D))
        v = np.empty((self.N, self.D)) if want_v else None
        j = np.empty((self.N, self.D), dtype=np.int32) if want_j else None
        L.check(self._lib.bt_get_state(self._ctx, L.dptr(Theta), L.dptr(v), L.dptr(j, C.c_int32)), "bt_get_state")
        return Theta, v, j

    def compute_all(self, Theta=None) -> dict:
        """Posterior.compute_all (modelling/posterior.py:332-446), evaluated on the GPU."""
        if Theta is not None:
            assert np.sum(np.isnan(Theta)) == 0, np.sum(np.isnan(Theta))    # posterior.py:360
            self.set_state(Theta)
        L.check(self._lib.bt_compute_all(self._ctx), "bt_compute_all")
        return self.get_current()

    def refresh(self):
        L.check(self._lib.bt_compute_all(self._ctx), "bt_compute_all")

    def init_v_from_grad(self):
        L.check(self._lib.bt_init_v_from_grad(self._ctx), "bt_init_v_from_grad")

    def get_current(self, with_forward_map=True) -> dict:
        N, D, Ln = self.N, self.D, self.L
        log_f = np.empty((N, Ln)) if with_forward_map else None
        nll = np.empty(N)
        gl = np.empty((N, D))
        oc, og, g = np.empty(N), np.empty(N), np.empty((N, D))
        L.check(self._lib.bt_get_current(self._ctx, L.dptr(log_f), L.dptr(nll), L.dptr(gl), L.dptr(oc), L.dptr(og),
                                         L.dptr(g)), "bt_get_current")
        Theta, _, _ = self.get_state(False, False)
        fm = {"log_f_Theta": log_f, "f_Theta": np.exp(log_f)} if with_forward_map else {}
        return {
            "Theta": Theta, "forward_map_evals": fm,
            "nll_utils": {"nll": np.float64(nll.sum())},
            "nll_pix": nll, "grad_lik": gl,
            "objective_pix_chromatic": oc, "objective_pix_global": og,
            "objective_chromatic": float(oc.sum()), "objective_global": float(og.sum()), "grad": g,
        }

    def get_log_f(self) -> np.ndarray:
        """ln f(Theta) of the resident state, (N, L), from the per-pixel cache."""
        log_f = np.empty((self.N, self.L))
        L.check(self._lib.bt_get_current(self._ctx, L.dptr(log_f), None, None, None, None, None), "bt_get_current")
        return log_f

    def objective_terms(self, owned=None) -> dict:
        """dict_objective of Posterior.compute_all_for_saver (posterior.py:270-330)."""
        out = np.zeros(2 + 2 * self.D)
        ow = None if owned is None else np.ascontiguousarray(owned, dtype=np.uint8)
        L.check(self._lib.bt_objective_terms(self._ctx, L.dptr(ow, C.c_uint8), L.dptr(out)), "bt_objective_terms")
        return self._dict_objective(out)

    def _dict_objective(self, out) -> dict:
        """Same keys and types as the reference (posterior.py:300-330): numpy scalars (``MySaver.initialize_memory`` reads
        ``v.shape`` of every entry, my_saver.py:57-58), prior entries only when the prior exists."""
        D = self.D
        d = {"nll": np.float64(out[0])}
        if self.has_spatial:
            d["nl_prior_spatial"] = np.array(out[1:1 + D], dtype=np.float64)
        if self.has_indicator:
            d["nl_prior_indicator"] = np.array(out[1 + D:1 + 2 * D], dtype=np.float64)
        d["objective"] = np.float64(out[1 + 2 * D])
        return d

    def spatial_prior_unweighted(self, owned=None) -> np.ndarray:
        """prior_spatial.neglog_pdf(Theta, idx_pix=arange(N), with_weights=False) -> (D,)
        (l22_discrete_grad_prior.py:95-128), at the resident Theta."""
        out = np.zeros(self.D)
        ow = None if owned is None else np.ascontiguousarray(owned, dtype=np.uint8)
        L.check(self._lib.bt_spatial_prior_unweighted(self._ctx, L.dptr(ow, C.c_uint8), L.dptr(out)),
                "bt_spatial_prior_unweighted")
        return out

    # ------------------------------------------------------------------ on-device chain store
    def chain_reserve(self, capacity: int):
        L.check(self._lib.bt_chain_reserve(self._ctx, int(capacity)), "bt_chain_reserve")

    def chain_push(self):
        L.check(self._lib.bt_chain_push(self._ctx), "bt_chain_push")

    def chain_length(self) -> int:
        return int(self._lib.bt_chain_length(self._ctx))

    def chain_get(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        """Stored iterates [first, first + count) as (count, N, D) float64 -- the saver's ``list_Theta``."""
        count = self.chain_length() - first if count is None else count
        out = np.empty((count, self.N, self.D))
        L.check(self._lib.bt_chain_get(self._ctx, first, count, L.dptr(out)), "bt_chain_get")
        return out

    def chain_stats(self, first: int = 0, count: Optional[int] = None, percentiles=(), want_var: bool = False) -> dict:
        """Per (pixel, parameter) summaries of the stored chain, computed on the device: ``mean`` (N, D) in scaled space
        and, per percentile, the bracketing order statistics ``lo`` / ``hi`` (n_q, N, D) with weight ``frac`` (n_q)."""
        count = self.chain_length() - first if count is None else count
        q = L.as_f64(np.asarray(percentiles, dtype=np.float64).reshape(-1))
        nq = q.size
        mean = np.empty((self.N, self.D))
        var = np.empty((self.N, self.D)) if want_var else None
        lo, hi, frac = np.empty((nq, self.N, self.D)), np.empty((nq, self.N, self.D)), np.empty(nq)
        L.check(self._lib.bt_chain_stats(self._ctx, first, count, nq, L.dptr(q) if nq else None, L.dptr(mean), L.dptr(var),
                                         L.dptr(lo) if nq else None, L.dptr(hi) if nq else None,
                                         L.dptr(frac) if nq else None), "bt_chain_stats")
        return {"mean": mean, "var": var, "percentiles": q, "lo": lo, "hi": hi, "frac": frac, "count": count}

    def forward_map(self, Theta: np.ndarray, compute_derivatives: bool = False) -> dict:
        """ForwardMap.compute_all on arbitrary rows (neural_network_approx.py:152-305)."""
        Theta = L.as_f64(Theta)
        M = Theta.shape[0]
        lf = np.empty((M, self.L))
        glf = np.empty((M, self.D, self.L)) if compute_derivatives else None
        L.check(self._lib.bt_forward_map(self._ctx, M, L.dptr(Theta), L.dptr(lf), L.dptr(glf)), "bt_forward_map")
        out = {"log_f_Theta": lf, "f_Theta": np.exp(lf)}
        if compute_derivatives:
            out["grad_log_f_Theta"] = glf
            out["grad_f_Theta"] = glf * out["f_Theta"][:, None, :]
        return out

    # ------------------------------------------------------------------ kernels
    def begin_iteration(self):
        L.check(self._lib.bt_begin_iteration(self._ctx), "bt_begin_iteration")

    def pmala_site(self, site, eps, eta, alpha, seed, t, z=None, u=None):
        z, u = L.as_f32(z), L.as_f32(u)
        L.check(self._lib.bt_pmala_site(self._ctx, site, eps, eta, alpha, seed, t, L.dptr(z, C.c_float),
                                        L.dptr(u, C.c_float)), "bt_pmala_site")

    def mtm_site(self, site, K, seed, t, cand=None, u_sel=None, u_acc=None):
        cand, u_sel, u_acc = L.as_f32(cand), L.as_f32(u_sel), L.as_f32(u_acc)
        L.check(self._lib.bt_mtm_site(self._ctx, site, K, seed, t, L.dptr(cand, C.c_float), L.dptr(u_sel, C.c_float),
                                      L.dptr(u_acc, C.c_float)), "bt_mtm_site")

    def mtm_finish(self, alpha):
        L.check(self._lib.bt_mtm_finish(self._ctx, alpha), "bt_mtm_finish")

    def step_stats(self):
        a, lp = np.empty(self.N), np.empty(self.N)
        L.check(self._lib.bt_get_step_stats(self._ctx, L.dptr(a), L.dptr(lp)), "bt_get_step_stats")
        return a, lp

    def reduce_step_stats(self, owned=None):
        out = np.zeros(2)
        ow = None if owned is None else np.ascontiguousarray(owned, dtype=np.uint8)
        L.check(self._lib.bt_reduce_step_stats(self._ctx, L.dptr(ow, C.c_uint8), L.dptr(out)), "bt_reduce_step_stats")
        return out

    def debug_philox_pmala(self, site, seed, t):
        z = np.zeros((self.N, self.D), dtype=np.float32)
        u = np.zeros(self.N, dtype=np.float32)
        L.check(self._lib.bt_debug_philox_pmala(self._ctx, site, seed, t, L.dptr(z, C.c_float), L.dptr(u, C.c_float)),
                "bt_debug_philox_pmala")
        return z, u

    def debug_philox_mtm(self, site, K, seed, t):
        cand = np.zeros((self.N, K, self.D), dtype=np.float32)
        us = np.zeros(self.N, dtype=np.float32)
        ua = np.zeros(self.N, dtype=np.float32)
        L.check(self._lib.bt_debug_philox_mtm(self._ctx, site, K, seed, t, L.dptr(cand, C.c_float), L.dptr(us, C.c_float),
                                              L.dptr(ua, C.c_float)), "bt_debug_philox_mtm")
        return cand, us, ua

    # ------------------------------------------------------------------ online model checking
    def model_check_reset(self):
        L.check(self._lib.bt_model_check_reset(self._ctx), "bt_model_check_reset")

    def model_check_update(self, seed, t, za=None, zm=None):
        """MySampler._update_model_check_values, sampling branch (sampler/my_sampler.py:158-214)."""
        za, zm = L.as_f32(za), L.as_f32(zm)
        L.check(self._lib.bt_model_check_update(self._ctx, seed, t, L.dptr(za, C.c_float), L.dptr(zm, C.c_float)),
                "bt_model_check_update")

    def model_check_get(self, finalize=True):
        clppd, pvy, pvl = np.empty((self.N, self.L)), np.empty((self.N, self.L)), np.empty(self.N)
        L.check(self._lib.bt_model_check_get(self._ctx, L.dptr(clppd), L.dptr(pvy), L.dptr(pvl), int(finalize)),
                "bt_model_check_get")
        return clppd, pvy, pvl

    def model_check_snapshot_clppd(self):
        """optimisation branch of _update_model_check_values (sampler/my_sampler.py:208-212)."""
        L.check(self._lib.bt_model_check_snapshot_clppd(self._ctx), "bt_model_check_snapshot_clppd")

    def model_check_update_pvalues(self, seed, t, za=None, zm=None):
        """one replication of _finalize_model_check_values, optimisation branch (sampler/my_sampler.py:222-256)."""
        za, zm = L.as_f32(za), L.as_f32(zm)
        L.check(self._lib.bt_model_check_update_pvalues(self._ctx, seed, t, L.dptr(za, C.c_float), L.dptr(zm, C.c_float)),
                "bt_model_check_update_pvalues")

    # ------------------------------------------------------------------ optimisation mode (is_stochastic=False)
    def optim_pmala_step(self, eps, eta, alpha, seed, t, z=None):
        """generate_new_sample_pmala_rmsprop, optimisation branch (sampler/my_sampler.py:908-964) -> (accept, proba)."""
        obj_cur = self.objective_terms()["objective"]
        L.check(self._lib.bt_optim_begin(self._ctx), "bt_optim_begin")
        z = L.as_f32(z)
        accept = False
        for with_noise in (0, 1):
            L.check(self._lib.bt_optim_pmala_propose(self._ctx, eps, eta, with_noise, seed, t, L.dptr(z, C.c_float)),
                    "bt_optim_pmala_propose")
            self.refresh()
            if self.objective_terms()["objective"] < obj_cur:
                accept = True
                break
        L.check(self._lib.bt_optim_update_v(self._ctx, alpha), "bt_optim_update_v")
        L.check(self._lib.bt_optim_end(self._ctx, int(accept)), "bt_optim_end")
        return accept, (1 if accept else 0)

    def mtm_optim_weights(self, site, K, seed, t, cand=None, owned=None) -> np.ndarray:
        cand = L.as_f32(cand)
        ow = None if owned is None else np.ascontiguousarray(owned, dtype=np.uint8)
        out = np.zeros(K)
        L.check(self._lib.bt_mtm_optim_weights(self._ctx, site, K, seed, t, L.dptr(cand, C.c_float), L.dptr(ow, C.c_uint8),
                                               L.dptr(out)), "bt_mtm_optim_weights")
        return out

    def mtm_optim_select(self, site, K, k_star):
        L.check(self._lib.bt_mtm_optim_select(self._ctx, site, K, int(k_star)), "bt_mtm_optim_select")

    def mtm_optim_site(self, site, K, seed, t, cand=None):
        """body of the site loop of generate_new_sample_mtm, optimisation branch (sampler/my_sampler.py:1006-1086)."""
        colsum = self.mtm_optim_weights(site, K, seed, t, cand)
        self.mtm_optim_select(site, K, int(np.argmin(colsum)))

    def synchronize(self):
        L.check(self._lib.bt_synchronize(self._ctx), "bt_synchronize")

    def kernel_launches(self) -> int:
        return int(self._lib.bt_kernel_launches(self._ctx))

    def mlp_kernel_variant(self) -> int:
        """Tensor-core forward-map kernel in use for this network (3 = pair kernel, 1 = single-CTA kernel, 0 = none)."""
        return int(self._lib.bt_mlp_kernel_variant(self._ctx))

    def mlp_profile(self, enable=True, reset=False):
        ms, rows, launches = C.c_double(), C.c_int64(), C.c_int64()
        L.check(self._lib.bt_mlp_profile(self._ctx, int(enable), int(reset), C.byref(ms), C.byref(rows),
                                         C.byref(launches)), "bt_mlp_profile")
        return ms.value, rows.value, launches.value

    KERNEL_PROFILE_NAMES = ("pmala_propose", "pmala_accept", "mtm_gen", "mtm_nl", "mtm_select", "mtm_finish")

    def kernel_profile(self, enable=True, reset=False) -> dict:
        """Device time (ms), algorithmic HBM bytes and launch count of the memory-bound kernels of the step since the last
        reset (``bt_kernel_profile``): {name: (ms, bytes, launches)}."""
        n = len(self.KERNEL_PROFILE_NAMES)
        ms, by, ln = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.int64)
        L.check(self._lib.bt_kernel_profile(self._ctx, int(enable), int(reset), L.dptr(ms), L.dptr(by), L.dptr(ln, C.c_int64)),
                "bt_kernel_profile")
        return {k: (float(ms[i]), float(by[i]), int(ln[i])) for i, k in enumerate(self.KERNEL_PROFILE_NAMES)}

    def reserve_scratch(self, max_rows: int, max_rows_with_jacobian: int = 0):
        """Allocate the step's scratch buffers up front (``bt_reserve_scratch``): no allocation inside a step afterwards."""
        L.check(self._lib.bt_reserve_scratch(self._ctx, int(max_rows), int(max_rows_with_jacobian)), "bt_reserve_scratch")

    # raw pointer plumbing for halo exchange / e2e bench (device or pinned-host addresses as ints)
    def set_theta_f32_ptr(self, host_ptr: int):
        L.check(self._lib.bt_set_theta_f32(self._ctx, C.c_void_p(host_ptr)), "bt_set_theta_f32")

    def get_theta_f32_ptr(self, host_ptr: int):
        L.check(self._lib.bt_get_theta_f32(self._ctx, C.c_void_p(host_ptr)), "bt_get_theta_f32")

    def gather_theta_rows(self, pix_dev_ptr: int, n: int, out_dev_ptr: int):
        L.check(self._lib.bt_gather_theta_rows(self._ctx, C.c_void_p(pix_dev_ptr), n, C.c_void_p(out_dev_ptr)),
                "bt_gather_theta_rows")

    def scatter_theta_rows(self, pix_dev_ptr: int, n: int, in_dev_ptr: int):
        L.check(self._lib.bt_scatter_theta_rows(self._ctx, C.c_void_p(pix_dev_ptr), n, C.c_void_p(in_dev_ptr)),
                "bt_scatter_theta_rows")


def gpu_log_f(fm, Theta: np.ndarray, device: int = 0, mlp_mode: int = MLP_FP32) -> np.ndarray:
    """ln f(Theta) of a (mirror or reference) NeuralNetworkApprox evaluated on the GPU, without a Posterior:
    used to synthesise observations (simulations/astro/observation/abstract_toy_case.py:68-77)."""
    from .modelling import MixingModelsLikelihood, Posterior
    Theta = np.ascontiguousarray(Theta, dtype=np.float64)
    M = Theta.shape[0]
    Ln = int(fm.L)
    lik = MixingModelsLikelihood(fm, Theta.shape[1], Ln, M, np.ones((M, Ln)), 1.0, 0.1, 0.5,
                                 a0=np.zeros((1, Ln)), a1=np.ones((1, Ln)))
    dev = DevicePosterior.from_posterior(Posterior(Theta.shape[1], Ln, M, lik, None, None), device=device,
                                         mlp_mode=mlp_mode)
    try:
        return dev.forward_map(Theta)["log_f_Theta"]
    finally:
        dev.close()
