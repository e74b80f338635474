"""Oracle (test infrastructure): float64 restatement of the two MySampler transition kernels.

``pmala_step`` follows MySampler.generate_new_sample_pmala_rmsprop, sampling branch
(/root/reference/beetroots/sampler/my_sampler.py:735-906); ``mtm_step`` follows
generate_new_sample_mtm, sampling branch (:987-1202) with the proposal / proposal-density
helpers of sampler/utils/utils.py:274-495.  The random draws are *injected* (noise tape), in
the order the reference consumes them, so a step can be compared bit-for-decision with the
reference and with the CUDA path:

  PMALA, per site : z (n_pix, D) standard normal (:767), u (n_pix,) uniform (:873)
  MTM,   per site : cand (n_pix, K, D) proposals (:1013), u_sel (n_pix,) selection uniform
                    (:1128, ``Generator.choice`` == searchsorted(cdf, u, 'right')),
                    u_acc (n_pix,) acceptance uniform (:1161)

The structure (full-map ``compute_all`` after every site, (N, K+1, D) candidate tensor) is the
reference's own, so that timing this port is a fair CPU baseline.
"""
from __future__ import annotations

import numpy as np

from . import posterior as P


class OracleSampler:
    """State + kernels.  ``params`` needs eps0, lambda_, alpha, k_mtm (MySampler.__init__ :50-66)."""

    def __init__(self, post, nbr, eps0, lambda_, alpha, k_mtm):
        self.post, self.nbr = post, nbr
        self.eps0, self.lambda_, self.alpha, self.k_mtm = eps0, lambda_, alpha, k_mtm
        self.N, self.D = post.N, post.D
        self.current = None
        self.v = None
        self.j_t = None

    def init(self, Theta_0):
        """my_sampler.py:319-354."""
        self.current = P.compute_all(self.post, Theta_0 * 1, self.nbr)
        self.v = self.current["grad"] ** 2
        self.j_t = np.zeros((self.N, self.D))

    def sites(self):
        return list(self.post.dict_sites.keys())

    # ------------------------------------------------------------------ model checking :158-214, :260-265
    def model_check_reset(self):
        N, L = self.N, self.post.L
        self.mc = {"clppd": np.zeros((N, L)), "p_values_y": np.zeros((N, L)), "p_values_llh": np.zeros(N), "count": 0}

    def model_check_update(self, za, zm, y_rep=None):
        """_update_model_check_values, sampling branch, with the replicated observation of
        MixingModelsLikelihood.sample_observation_model (approx_censored_add_mult.py:268-280) driven by the
        standard normals za (additive) and zm (multiplicative)."""
        import copy
        lik = self.post.likelihood
        fm = self.current["forward_map_evals"]
        nll_full = self.current["nll_full"]
        c = self.mc["count"]
        self.mc["clppd"] = self.mc["clppd"] * (c / (c + 1)) + np.exp(-nll_full) / (c + 1)              # :169-170
        eps_a = lik.sigma_a * za                                                                        # rng.normal(0, sigma_a)
        eps_m = np.exp(-(lik.sigma_m ** 2) / 2 + lik.sigma_m * zm)                                      # rng.lognormal
        if y_rep is None:
            y_rep = np.maximum(lik.omega, eps_m * fm["f_Theta"] + eps_a)                                # :279
        lik_rep = copy.copy(lik)
        lik_rep.y = y_rep            # NB the reference does NOT refresh log_y (:176-177): the multiplicative branch of
        #                              the replicated likelihood keeps ln(y) of the ORIGINAL observation -- reproduced
        nll_rep = P.likelihood_eval(lik_rep, fm, None, False)["nll_full"]                               # :181-191
        self.mc["p_values_y"] = self.mc["p_values_y"] * (c / (c + 1)) + (y_rep <= lik.y) / (c + 1)      # :194-195
        self.mc["p_values_llh"] = self.mc["p_values_llh"] * (c / (c + 1)) + (nll_rep.sum(1) >= nll_full.sum(1)) / (c + 1)
        self.mc["count"] = c + 1

    def model_check_get(self):
        py = self.mc["p_values_y"]
        return self.mc["clppd"], np.where(py > 0.5, 1 - py, py), self.mc["p_values_llh"]               # :260-265

    # ------------------------------------------------------------------ PMALA :735-906
    def pmala_step(self, noise: dict):
        post = self.post
        accept_total = np.zeros(self.N)
        log_pa_total = np.zeros(self.N)
        list_idx = self.sites()
        objective_type = "objective_pix_chromatic" if len(list_idx) > 1 else "objective_pix_global"  # :745-750
        for s in list_idx:
            idx_pix = post.dict_sites[s]
            new_Theta = self.current["Theta"] * 1
            grad_t = self.current["grad"][idx_pix]
            v_cur = self.v[idx_pix]
            G = 1 / (self.lambda_ + np.sqrt(v_cur))                               # :761
            z = noise[s]["z"] * np.sqrt(self.eps0 * G)                            # :767-768
            mu = new_Theta[idx_pix] - self.eps0 / 2 * G * grad_t                   # :796-800 (no correction)
            cand = mu + z
            log_q_cand = -0.5 * np.log(G).sum(1) - 1 / (2 * self.eps0) * ((cand - mu) ** 2 / G).sum(1)  # :804-808
            cand_full = new_Theta * 1
            cand_full[idx_pix] = cand
            cand_all = P.compute_all(post, cand_full, self.nbr)                   # :817-821
            grad_c = cand_all["grad"][idx_pix]
            v_cand = self.alpha * v_cur + (1 - self.alpha) * grad_c ** 2           # :823-825
            Gc = 1 / (self.lambda_ + np.sqrt(v_cand))
            mu_c = cand - self.eps0 / 2 * Gc * grad_c                              # :841-845
            log_q_cur = -0.5 * np.log(Gc).sum(1) - 1 / (2 * self.eps0) * (
                (new_Theta[idx_pix] - mu_c) ** 2 / Gc).sum(1)                      # :847-851
            log_pa = (-cand_all[objective_type][idx_pix] + self.current[objective_type][idx_pix]
                      + log_q_cur - log_q_cand)                                    # :857-870
            with np.errstate(divide="ignore"):
                log_u = np.log(noise[s]["u"])
            acc = log_u < log_pa                                                   # :873-874
            new_Theta[idx_pix] = np.where(acc[:, None], cand, new_Theta[idx_pix])
            accept_total[idx_pix] = acc
            log_pa_total[idx_pix] = log_pa
            self.v[idx_pix] = v_cand                                               # :886-888 (always)
            self.j_t[idx_pix] = np.where(acc[:, None], 0.0, self.j_t[idx_pix] + 1)  # :890-896
            if acc.max() > 0:
                self.current = P.compute_all(post, new_Theta, self.nbr)            # :898-903
        return accept_total, log_pa_total

    # ------------------------------------------------------------------ MTM :987-1202
    def proposal_log_density(self, Theta, idx_pix, cand_pix):
        """compute_nlpdf_spatial_proposal (utils.py:419-495): for every non-empty subset S of the
        neighbours (size m, mean mu_S): sum_d log sum_S sqrt(tau_d) m exp(-(x_d-mu_Sd)^2 tau_d m^2).
        cand_pix (n_pix, K+1, D) -> (n_pix, K+1).  tau = min(initial_weights, weights) (:1097-1100)."""
        ps = self.post.prior_spatial
        tau = np.minimum(ps.initial_weights, ps.weights)
        out = np.zeros(cand_pix.shape[:2])
        nb = self.nbr[idx_pix]
        deg = (nb >= 0).sum(1)
        for d_ in np.unique(deg):
            sel = np.where(deg == d_)[0]
            if d_ == 0:
                # empty sum -> log(0) = -inf in the reference (logsumexp over an empty array)
                out[sel] = -np.inf
                continue
            th = Theta[nb[sel][:, :d_]]                                           # (n, deg, D)
            x = cand_pix[sel]                                                     # (n, K+1, D)
            terms = []
            wts = []
            for mask in range(1, 2 ** d_):
                bits = [(mask >> b) & 1 for b in range(d_)]
                m = sum(bits)
                mu = (th * np.asarray(bits)[None, :, None]).sum(1) / m            # (n, D)
                w = np.sqrt(tau) * m                                              # (D,)
                terms.append(-((x - mu[:, None, :]) ** 2) * w ** 2)               # (n, K+1, D)
                wts.append(w)
            terms = np.stack(terms, 0)                                            # (S, n, K+1, D)
            wts = np.stack(wts, 0)[:, None, None, :]
            mx = terms.max(0)
            out[sel] = (np.log((wts * np.exp(terms - mx)).sum(0)) + mx).sum(-1)
        return out

    def mtm_step(self, noise: dict):
        post = self.post
        K = self.k_mtm
        new_Theta = self.current["Theta"] * 1
        accept_total = np.zeros(self.N)
        log_rg_total = np.zeros(self.N)
        list_idx = self.sites()
        chromatic = len(list_idx) > 1
        for s in list_idx:
            idx_pix = post.dict_sites[s]
            n_pix = idx_pix.size
            cand_pix = np.zeros((n_pix, K + 1, self.D))                           # rows of :1011-1015
            cand_pix[:, :K] = noise[s]["cand"]
            cand_pix[:, K] = new_Theta[idx_pix]
            # posterior.mtm_neglog_pdf_priors :59-111
            nl = np.zeros((n_pix, K + 1))
            if post.prior_spatial is not None:
                had, _ = P.spatial_sums(new_Theta, self.nbr, idx_pix, cand_pix)
                factor = 1.0 if (chromatic or n_pix < self.N) else 0.5            # l22:105-107
                nl += factor * (post.prior_spatial.weights * had).sum(-1)
            if post.prior_indicator is not None:
                nl += P.indicator_penalty(cand_pix, post.prior_indicator)
            # likelihood.neglog_pdf_candidates abstract_likelihood.py:73-111
            fm = P.forward_map_compute_all(post.likelihood.forward_map,
                                           cand_pix.reshape(n_pix * (K + 1), self.D), False)
            nl_lik = P.likelihood_eval(post.likelihood, fm, idx_pix, False)["nll_pix"]
            nl += nl_lik.reshape(n_pix, K + 1)                                    # :1044-1046
            if post.prior_spatial is not None:
                nl += self.proposal_log_density(new_Theta, idx_pix, cand_pix)     # :1093-1105
            with np.errstate(all="ignore"):
                nl = nl - nl.min(axis=1, keepdims=True)                           # :1107-1110
                pdf = np.exp(-nl)
                log_num = np.log(pdf[:, :K].sum(1))                               # :1114
                w = np.exp(-nl[:, :K] - (-nl[:, :K]).max(1, keepdims=True))       # softmax :1123
                w = w / w.sum(1, keepdims=True)
                cdf = np.cumsum(w, axis=1)
                cdf = cdf / cdf[:, -1:]
                idx_ch = (cdf <= noise[s]["u_sel"][:, None]).sum(1)               # searchsorted right :1127-1131
                idx_ch = np.minimum(idx_ch, K - 1)
                ar = np.arange(n_pix)
                chall = cand_pix[ar, idx_ch]
                log_den = np.log(pdf.sum(1) - np.exp(-nl[ar, idx_ch]))            # :1142-1144
                log_rg = log_num - log_den
                log_rg = np.where(np.isfinite(log_rg), log_rg, 1e-15)             # :1156-1159
                log_u = np.log(noise[s]["u_acc"])
            acc = log_u < log_rg                                                  # :1161-1162
            new_Theta[idx_pix] = np.where(acc[:, None], chall, cand_pix[:, K])
            accept_total[idx_pix] = acc
            log_rg_total[idx_pix] = log_rg
            self.j_t[idx_pix] = np.where(acc[:, None], 0.0, self.j_t[idx_pix])    # :1174-1178
        if accept_total.max() > 0:                                                # :1182-1197
            self.current = P.compute_all(post, new_Theta, self.nbr)
            self.v = np.where(accept_total[:, None] > 0,
                              self.alpha * self.v + (1 - self.alpha) * self.current["grad"] ** 2, self.v)
        return accept_total, log_rg_total

    # ------------------------------------------------------------------ optimisation mode (is_stochastic=False)
    def pmala_optim_step(self, z):
        """generate_new_sample_pmala_rmsprop, optimisation branch (my_sampler.py:908-964).  z (N, D) standard normals."""
        post = self.post
        key = "objective_global"
        grad_t = self.current["grad"]
        G = 1 / (self.lambda_ + np.sqrt(self.v))                                   # :915
        z_t = z * np.sqrt(self.eps0 * G)                                           # :922-923
        mu = self.current["Theta"] - self.eps0 / 2 * G * grad_t                    # :926-928
        cand_all = P.compute_all(post, mu, self.nbr)
        if cand_all[key] < self.current[key]:                                      # :935-938
            self.current = cand_all
            accept = True
        else:
            cand_all = P.compute_all(post, mu + z_t, self.nbr)                     # :940-945
            accept = bool(cand_all[key] < self.current[key])
            if accept:
                self.current = cand_all
        self.v = self.alpha * self.v + (1 - self.alpha) * cand_all["grad"] ** 2    # :952-953 (last evaluated candidate)
        return accept, (1 if accept else 0)

    def mtm_optim_step(self, noise: dict):
        """generate_new_sample_mtm, optimisation branch (my_sampler.py:1006-1086, 1182-1200).  noise[s] = {"cand"}."""
        post = self.post
        K = self.k_mtm
        new_Theta = self.current["Theta"] * 1
        accept_total = np.zeros(self.N)
        list_idx = self.sites()
        chromatic = len(list_idx) > 1
        for s in list_idx:
            idx_pix = post.dict_sites[s]
            n_pix = idx_pix.size
            cand_pix = np.zeros((n_pix, K + 1, self.D))
            cand_pix[:, :K] = noise[s]["cand"]
            cand_pix[:, K] = new_Theta[idx_pix]
            nl = np.zeros((n_pix, K + 1))
            if post.prior_spatial is not None:
                had, _ = P.spatial_sums(new_Theta, self.nbr, idx_pix, cand_pix)
                factor = 1.0 if (chromatic or n_pix < self.N) else 0.5
                nl += factor * (post.prior_spatial.weights * had).sum(-1)
            if post.prior_indicator is not None:
                nl += P.indicator_penalty(cand_pix, post.prior_indicator)
            fm = P.forward_map_compute_all(post.likelihood.forward_map, cand_pix.reshape(n_pix * (K + 1), self.D), False)
            nl += P.likelihood_eval(post.likelihood, fm, idx_pix, False)["nll_pix"].reshape(n_pix, K + 1)
            k_star = int(np.argmin(nl[:, :-1].sum(axis=0)))                        # :1051-1053 (one index per class)
            nc, n0 = nl[:, k_star], nl[:, -1]
            acc = (nc < n0) & np.isfinite(nc) & np.isfinite(n0)                    # :1073-1077
            new_Theta[idx_pix] = np.where(acc[:, None], cand_pix[:, k_star], cand_pix[:, K])
            accept_total[idx_pix] = acc
        if accept_total.max() > 0:                                                 # :1182-1197
            self.current = P.compute_all(post, new_Theta, self.nbr)
            self.v = np.where(accept_total[:, None] > 0,
                              self.alpha * self.v + (1 - self.alpha) * self.current["grad"] ** 2, self.v)
        return accept_total
