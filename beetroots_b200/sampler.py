"""B200Sampler: drop-in for the reference ``MySampler`` (sampler/my_sampler.py:19-1202).

Same constructor, same ``sample(...)`` signature and Saver protocol; the two transition kernels
(``generate_new_sample_mtm``, ``generate_new_sample_pmala_rmsprop``) run on the GPU through the
C ABI.  CUDA state is created lazily inside ``sample`` so the object survives the
``copy.deepcopy(sampler_)`` of RunMCMC (inversion/run/run_mcmc.py:164-167); do not use it from a
forked worker (can_run_in_parallel must be False, as the reference already requires for GPU
forward maps, simulations/astro/toy_case/toy_case_nn.py:162,172).

Noise: the kernel type is drawn from ``self.rng`` exactly like the reference (:430-435); all
per-pixel draws come from a counter-based Philox stream keyed on (seed, t, site, global pixel
id), with the 64-bit seed drawn once from ``self.rng`` -- so a chain is reproducible from
``rng`` and independent of how the map is sharded over GPUs.
"""
from __future__ import annotations

from sys import byteorder
from typing import List, Optional, Union

import numpy as np

from .device import MLP_FP32, DevicePosterior
from .mml import EBayesMMLELogRate


class MySamplerParams:
    """Mirror of sampler/utils/my_sampler_params.py:9-74 (same asserts)."""

    __slots__ = ("initial_step_size", "extreme_grad", "history_weight", "selection_probas", "k_mtm",
                 "is_stochastic", "compute_correction_term")

    def __init__(self, initial_step_size: float, extreme_grad: float, history_weight: float,
                 selection_probas: Union[np.ndarray, List[float]], k_mtm: int, is_stochastic: bool = True,
                 compute_correction_term: bool = True) -> None:
        assert initial_step_size > 0
        assert extreme_grad > 0
        assert 0 < history_weight < 1
        assert isinstance(k_mtm, int) and k_mtm >= 1
        if isinstance(selection_probas, list):
            selection_probas = np.array(selection_probas)
        assert np.all(0 <= selection_probas)
        assert selection_probas.sum() == 1
        assert isinstance(is_stochastic, bool)
        assert isinstance(compute_correction_term, bool)
        self.initial_step_size = initial_step_size
        self.extreme_grad = extreme_grad
        self.history_weight = history_weight
        self.selection_probas = selection_probas
        self.k_mtm = k_mtm
        self.is_stochastic = is_stochastic
        self.compute_correction_term = compute_correction_term


class B200Sampler:
    ESS_OPTIM = 1_000                                                              # sampler/abstract_sampler.py:15

    def __init__(self, my_sampler_params, D: int, L: int, N: int,
                 rng: np.random.Generator = None, device: int = 0, mlp_mode: int = MLP_FP32,
                 device_chain: int = 0, device_chain_thin: int = 1, saver_nll_utils: Optional[bool] = None):
        p = my_sampler_params
        self.eps0 = p.initial_step_size
        self.lambda_ = p.extreme_grad
        self.alpha = p.history_weight
        self.k_mtm = p.k_mtm
        self.selection_probas = np.asarray(p.selection_probas, dtype=np.float64)
        assert np.sum(self.selection_probas) == 1, f"{self.selection_probas} should sum to 1"
        self.stochastic = p.is_stochastic
        self.compute_correction_term = p.compute_correction_term
        if self.compute_correction_term:
            # NOT BUILT (second half): the bias-correction term of the preconditioned proposal needs the
            # diagonal of the Hessian of the negative log-posterior (my_sampler.py:771-793, 828-840) -- a second-order
            # forward-mode pass through the network (neural_network_approx.py:237-265) and the reference's hess_diag formulas
            # (approx_censored_add_mult.py:745-980).  In the reference it only runs WITHOUT a spatial prior (posterior.py:259
            # calls prior_spatial.hess_diag_neglog_pdf without idx_pix).  MySamplerParams defaults to True like the
            # reference's, so a drop-in caller must pass compute_correction_term=False explicitly.
            raise NotImplementedError("B200Sampler does not implement the Hessian-diagonal correction term of PMALA "
                                      "(compute_correction_term=True); construct MySamplerParams(..., "
                                      "compute_correction_term=False)")
        self.D, self.L, self.N = D, L, N
        self.rng = np.random.default_rng(42) if rng is None else rng
        self.device = device
        self.mlp_mode = mlp_mode
        self.v = np.zeros((N * D,))
        self.j_t = np.zeros((N * D,))
        self.current = {}
        self._philox = None                # key of the per-pixel Philox streams of the current sample() call
        self._user_seed = None             # a key fixed by the caller (tests, benchmarks): ``sampler.philox_seed = k``
        # on-device chain store: keep up to ``device_chain`` post-burn-in iterates (every
        # ``device_chain_thin``-th) in HBM and summarise them there, see posterior_summary()
        assert device_chain >= 0 and device_chain_thin >= 1
        self.device_chain = int(device_chain)
        self.device_chain_thin = int(device_chain_thin)
        # the reference hands the saver the per-(pixel, line) likelihood auxiliaries (s_a, m_m, lambda_, ...: my_saver.py:57-61);
        # here they are re-evaluated on the host from ln f of the saved iterate.  None: only for maps up to 10^6 (pixel, line)
        # pairs (ten float64 arrays of that size per saved iterate -- 0.8 GB at c4 -- are rarely wanted)
        self.saver_nll_utils = saver_nll_utils
        self._dev: Optional[DevicePosterior] = None
        self._dev_owner = None

    # ------------------------------------------------------------------ reference API
    def get_rng_state(self):
        """sampler/abstract_sampler.py:48-73 (numpy>=2-safe)."""
        st = self.rng.bit_generator.state["state"]
        return (np.array(bytearray(st["state"].to_bytes(32, byteorder))),
                np.array(bytearray(st["inc"].to_bytes(32, byteorder))))

    def set_rng_state(self, state_bytes, inc_bytes):
        """sampler/abstract_sampler.py:75-93."""
        state = self.rng.bit_generator.state
        state["state"]["state"] = int.from_bytes(state_bytes, byteorder)
        state["state"]["inc"] = int.from_bytes(inc_bytes, byteorder)
        self.rng.bit_generator.state = state

    def generate_random_start_Theta(self, posterior):
        """sampler/abstract_sampler.py:95-121."""
        if posterior.prior_indicator is not None:
            lo, hi = posterior.prior_indicator.lower_bounds, posterior.prior_indicator.upper_bounds
            return self.rng.uniform(0, 1, size=(posterior.N, posterior.D)) * (hi - lo)[None, :] + lo[None, :]
        return self.rng.standard_normal(size=(posterior.N, posterior.D))

    def generate_random_start_Theta_1pix(self, Theta, posterior, idx_pix):
        """The MTM proposals are drawn inside the MTM kernel on the device
        (sampler/my_sampler.py:103-156 -> csrc/sampler_kernels.cuh mtm_gen_kernel)."""
        raise NotImplementedError("proposals are generated on the device; see DevicePosterior.debug_philox_mtm")

    @property
    def philox_seed(self):
        return self._philox

    @philox_seed.setter
    def philox_seed(self, seed):
        """Fix the key of the per-pixel Philox streams (None: a fresh one is drawn from ``self.rng`` at every ``sample()``)."""
        self._user_seed = None if seed is None else int(seed)
        self._philox = self._user_seed

    # ------------------------------------------------------------------ device plumbing
    def attach(self, posterior, dev: Optional[DevicePosterior] = None):
        if dev is not None:
            self._dev, self._dev_owner = dev, posterior
        elif self._dev is None or self._dev_owner is not posterior:
            self._dev = DevicePosterior.from_posterior(posterior, device=self.device, mlp_mode=self.mlp_mode)
            self._dev_owner = posterior
        return self._dev

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, val in self.__dict__.items():
            setattr(new, k, None if k in ("_dev", "_dev_owner") else copy.deepcopy(val, memo))
        return new

    def _pull_state(self, with_current=True):
        dev = self._dev
        Theta, v, j = dev.get_state()
        cur = dev.get_current() if with_current else None      # (sharded: every rank takes part in the gathers)
        if Theta is None:                                  # sharded run, not the root rank
            return
        self.v = v.reshape(-1)
        self.j_t = j.reshape(-1).astype(np.float64)
        self.current = cur if with_current else {"Theta": Theta}

    def _pull_for_saver(self, saver, dict_objective):
        """What the Saver protocol stores of an iterate (sampler/saver/my_saver.py:46-83, 85-144): Theta, v and -- only if the
        saver keeps forward-map evaluations -- ln f / f.  j, the per-pixel objectives and the gradients stay on the device
        (the reference recomputes them by ``compute_all_for_saver``; nothing reads them between two iterations)."""
        dev = self._dev
        try:
            Theta, v, _ = dev.get_state(want_v=True, want_j=False, reuse=True)     # (host buffers reused: the saver copies)
        except TypeError:                                  # sharded / stand-in devices without the option
            Theta, v, _ = dev.get_state(want_v=True, want_j=False)
        if Theta is None:                                  # sharded run, not the root rank
            want_aux = self.saver_nll_utils if self.saver_nll_utils is not None else self.N * self.L <= 1_000_000
            if getattr(saver, "save_forward_map_evals", True) or want_aux:
                dev.get_log_f()                            # take part in the gather
            return False
        self.v = v.reshape(-1)
        want_aux = self.saver_nll_utils if self.saver_nll_utils is not None else self.N * self.L <= 1_000_000
        fm, aux = {}, {}
        if getattr(saver, "save_forward_map_evals", True) or want_aux:
            log_f = dev.get_log_f()
            if getattr(saver, "save_forward_map_evals", True):
                fm = {"log_f_Theta": log_f, "f_Theta": np.exp(log_f)}
            if want_aux:
                from .modelling import saver_nll_utils
                aux = saver_nll_utils(self._dev_owner.likelihood, log_f)
        self.current = {"Theta": Theta, "forward_map_evals": fm,
                        "nll_utils": {"nll": np.float64(dict_objective["nll"]), **aux}}
        return True

    # ------------------------------------------------------------------ kernels
    def generate_new_sample_pmala_rmsprop(self, t: int, posterior, noise: Optional[dict] = None):
        """sampler/my_sampler.py:718-906.  ``noise[s] = {"z": (N, D), "u": (N,)}`` injects the draws."""
        dev = self.attach(posterior)
        if not self.stochastic:                                                    # :908-964, noise = {"z": (N, D)}
            return dev.optim_pmala_step(self.eps0, self.lambda_, self.alpha, self.philox_seed or 0, t,
                                        None if noise is None else noise.get("z"))
        dev.begin_iteration()
        for s in range(dev.n_sites):
            nz = noise[dev.site_keys[s]] if noise is not None else {}
            dev.pmala_site(s, self.eps0, self.lambda_, self.alpha, self.philox_seed or 0, t, nz.get("z"), nz.get("u"))
        acc, lpa = dev.reduce_step_stats()
        return acc / self.N, lpa / self.N

    def generate_new_sample_mtm(self, t: int, posterior, noise: Optional[dict] = None):
        """sampler/my_sampler.py:966-1202.  ``noise[s] = {"cand": (N, K, D), "u_sel": (N,), "u_acc": (N,)}``."""
        dev = self.attach(posterior)
        dev.begin_iteration()
        for s in range(dev.n_sites):
            nz = noise[dev.site_keys[s]] if noise is not None else {}
            if self.stochastic:
                dev.mtm_site(s, self.k_mtm, self.philox_seed or 0, t, nz.get("cand"), nz.get("u_sel"), nz.get("u_acc"))
            else:                                                                  # :1050-1086
                dev.mtm_optim_site(s, self.k_mtm, self.philox_seed or 0, t, nz.get("cand"))
        dev.mtm_finish(self.alpha)
        acc, lpa = dev.reduce_step_stats()
        if not self.stochastic:
            return acc / self.N, acc / self.N                                      # :1199-1200
        return acc / self.N, lpa / self.N

    def sample_regu_hyperparams(self, posterior, regu_weights_optimizer, t: int, current_Theta=None) -> np.ndarray:
        """sampler/abstract_sampler.py:172-210; the potential is reduced on the device at the resident Theta
        (``current_Theta`` is accepted for signature compatibility and ignored)."""
        assert self.N > 1
        assert posterior.prior_spatial is not None
        dev = self.attach(posterior)
        tau_tm1 = np.asarray(posterior.prior_spatial.weights, dtype=np.float64) * 1
        neglog_prior = dev.spatial_prior_unweighted()                              # (D,)
        new_tau = regu_weights_optimizer.update(tau_tm1.mean(), t, neglog_prior.sum())
        return new_tau * np.ones((self.D,))

    def posterior_summary(self, list_CI=(68, 90, 95, 99), from_scaled_to_lin=None, first: int = 0,
                          count: Optional[int] = None) -> dict:
        """MMSE and credibility intervals of the chain kept on the device, as the reference's results extraction
        computes them from the HDF5 file (inversion/results/utils/mc.py:190-215): mean in scaled space (mapped to
        linear space if ``from_scaled_to_lin`` is given), ``np.percentile`` (linear interpolation) of the chain in
        linear space for the lower / upper percentile of every credibility level.  Keys follow the reference's csv
        columns (``Theta_MMSE``, ``per_2p5`` ...)."""
        assert self._dev is not None and self.device_chain, "no chain was kept on the device"
        per = sorted({(100 - ci) / 2 for ci in list_CI} | {100 - (100 - ci) / 2 for ci in list_CI})
        st = self._dev.chain_stats(first, count, per)
        if st["mean"] is None:                             # sharded run, not the root rank (it took part in the gathers)
            return {}
        f = (lambda a: a) if from_scaled_to_lin is None else (lambda a: from_scaled_to_lin(a.reshape(-1, self.D)).reshape(a.shape))
        out = {"Theta_MMSE_scaled": st["mean"], "Theta_MMSE": f(st["mean"]), "count": st["count"]}
        for q, p in enumerate(per):
            lo, hi, g = f(st["lo"][q]), f(st["hi"][q]), st["frac"][q]
            out[f"per_{p:.1f}".replace(".", "p")] = lo + g * (hi - lo)
        return out

    # ------------------------------------------------------------------ main loop
    def sample(self, posterior, saver, max_iter: int, Theta_0: Optional[np.ndarray] = None,
               disable_progress_bar: bool = False, regu_spatial_N0: Union[int, float] = np.inf,
               regu_spatial_scale: float = 1.0, regu_spatial_vmin: float = 1e-8, regu_spatial_vmax: float = 1e8,
               T_BI: int = 0) -> None:
        """sampler/my_sampler.py:269-559."""
        additional_sampling_log = {}
        if Theta_0 is None:
            print("starting from a random point")
            Theta_0 = self.generate_random_start_Theta(posterior)
        assert Theta_0 is not None
        assert Theta_0.shape == (self.N, self.D)
        dev = self.attach(posterior)
        # Sharded over several GPUs (parallel.ShardedDevicePosterior) every rank runs this loop; only the root talks to the
        # saver (one writer per results file), the other ranks take part in the collectives behind the pulls.
        is_root = getattr(dev, "rank", 0) == 0
        sharded = getattr(dev, "world", 1) > 1
        # a fresh Philox key per call (a second sample() on the same object continues the host stream like the reference's
        # rng does); recorded with the results so that a run can be reproduced from its saved state
        self._philox = int(self.rng.integers(0, 2 ** 63 - 1)) if self._user_seed is None else int(self._user_seed)
        if sharded:
            self._philox = int(dev.broadcast_ints([self._philox])[0])
        if hasattr(dev, "reserve_scratch"):
            n_max = max([p.size for p in dev.site_pixels] + [1])
            dev.reserve_scratch(n_max * (self.k_mtm + 1), max(n_max, self.N if not sharded else dev.local.N))
        self.current = dev.compute_all(Theta_0)                                    # :319-323
        if is_root:
            assert np.isnan(self.current["objective_global"]) == 0                # :325-326
            assert np.sum(np.isnan(self.current["grad"])) == 0
        dev.init_v_from_grad()                                                     # :337, :354
        if is_root:
            self.v = self.current["grad"].flatten() ** 2
            assert np.sum(np.isnan(self.v)) == 0.0
            assert np.sum(np.isinf(self.v)) == 0.0
        self.j_t = np.zeros((self.N * self.D,))
        dev.model_check_reset()                                                    # :385-401
        best_objective = np.inf                                                    # used only if not self.stochastic
        regu_weights_optimizer = EBayesMMLELogRate(scale=regu_spatial_scale, N0=regu_spatial_N0, N1=+np.inf,
                                                   dim=self.D * self.N, vmin=regu_spatial_vmin, vmax=regu_spatial_vmax,
                                                   homogeneity=2.0, exponent=0.8)           # :369-379
        optimize_regu_weights = regu_weights_optimizer.N0 < np.inf
        if self.device_chain:
            dev.chain_reserve(self.device_chain)

        try:
            from tqdm.auto import tqdm
            it = tqdm(range(1, max_iter + 1), disable=disable_progress_bar or not is_root)
        except Exception:                                                          # pragma: no cover
            it = range(1, max_iter + 1)
        for t in it:
            if optimize_regu_weights and (self.N > 1):                              # :403-427
                assert posterior.prior_spatial is not None
                if t >= regu_weights_optimizer.N0:
                    tau_t = self.sample_regu_hyperparams(posterior, regu_weights_optimizer, t)
                    posterior.prior_spatial.weights = tau_t * 1
                    # the device assembles prior terms from the resident Theta on the fly: nothing to recompute (:419-425)
                    dev.update_spatial_weights(posterior.prior_spatial, posterior.prior_indicator)
                additional_sampling_log["tau"] = posterior.prior_spatial.weights * 1
            type_t = int(np.argmax(self.rng.multinomial(1, pvals=self.selection_probas)))   # :430-435
            # the saver's decisions depend on t and on its own state only, so they are taken before the kernel runs; sharded,
            # the root's kernel choice and decisions are the ones every rank follows (one small broadcast per iteration)
            first = is_root and saver.memory == {}
            need = is_root and (first or saver.check_need_to_update_memory(t))      # :456, :501
            if sharded:
                type_t, need = (int(x) for x in dev.broadcast_ints([type_t, int(need)]))
            if type_t == 0:
                accepted_t, log_proba_accept_t = self.generate_new_sample_mtm(t, posterior)
            else:
                accepted_t, log_proba_accept_t = self.generate_new_sample_pmala_rmsprop(t, posterior)

            if need:
                dict_objective = dev.objective_terms()                             # posterior.py:270-330
                have = self._pull_for_saver(saver, dict_objective)
                if t > T_BI:                                                       # :468-476 / :514-521
                    if self.stochastic:
                        dev.model_check_update(self.philox_seed, t)
                    elif dict_objective["objective"] < best_objective:            # :208-212
                        best_objective = dict_objective["objective"] * 1
                        dev.model_check_snapshot_clppd()
                if have:
                    additional_sampling_log["v"] = self.v.reshape((self.N, self.D))     # (the saver copies it into its batch)
                    additional_sampling_log["type_t"] = type_t
                    additional_sampling_log["accepted_t"] = accepted_t
                    additional_sampling_log["log_proba_accept_t"] = log_proba_accept_t
                    kw = dict(Theta=self.current["Theta"], forward_map_evals=self.current["forward_map_evals"],
                              nll_utils=self.current["nll_utils"], dict_objective=dict_objective,
                              additional_sampling_log=additional_sampling_log)
                    if first:
                        saver.initialize_memory(max_iter, t, **kw)                 # :478-486
                    rng_state_array, rng_inc_array = self.get_rng_state()
                    saver.update_memory(t, rng_state_array=rng_state_array, rng_inc_array=rng_inc_array, **kw)
            if is_root and saver.check_need_to_save(t):                            # :539-541
                saver.save_to_file()
            if (self.device_chain and t > T_BI and t % self.device_chain_thin == 0
                    and dev.chain_length() < self.device_chain):
                dev.chain_push()
        self._pull_state(with_current=True)
        if not self.stochastic:                                                    # :222-256: p-values at the estimate
            for rep in range(self.ESS_OPTIM):
                dev.model_check_update_pvalues(self.philox_seed, max_iter + 1 + rep)
        clppd, p_values_y, p_values_llh = dev.model_check_get(finalize=True)       # :544-549
        if is_root:
            saver.save_additional(list_arrays=[clppd, p_values_y, p_values_llh, np.array([self.philox_seed], dtype=np.uint64)],
                                  list_names=["clppd", "p-values-y", "p-values-llh", "philox_seed"])   # :551-558
        return
