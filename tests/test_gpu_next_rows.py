"""GPU: the SURVEY §8f "next" rows built on top of the sampler step -- on-device chain store + posterior summaries
(rank 1) and the regularisation-weight update (rank 3) -- through the C ABI, against numpy / the oracle / reference
goldens.  Order statistics and stored iterates are bit-exact (float32 moves, no arithmetic); means and the spatial
potential are float32 sums compared at rtol 1e-6 / 1e-5."""
import json
import os

import numpy as np
import pytest

from _cases import GOLDEN, load_golden, problem_from_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib(built_lib):
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    from beetroots_b200 import _lib
    return _lib.load()


def _synthetic(H, W, K, nnn=False):
    from beetroots_b200.synthetic import make_problem
    from oracle import posterior as OP
    return make_problem(H, W, lambda fm, th: OP.forward_map_compute_all(fm, th, False)["log_f_Theta"], L=10, k_mtm=K,
                        use_next_nearest_neighbors=nnn)


def _np_percentile_from_brackets(st):
    return st["lo"] + st["frac"][:, None, None] * (st["hi"] - st["lo"])


@pytest.mark.parametrize("T,N", [(1, 7), (2, 33), (37, 300), (401, 130), (900, 70), (1700, 40), (2500, 33)])
def test_chain_store_and_summaries_vs_numpy(lib, T, N):
    """Synthetic iterates pushed through bt_set_state / bt_chain_push: stored chain bit-identical to what was pushed;
    mean / variance / percentiles equal numpy's on the same float32 chain (all three shared-memory tile widths and the
    global-scratch path for long chains; ties, constants and single-iterate chains included)."""
    from beetroots_b200 import DevicePosterior
    prob = _synthetic(1, N, 4) if N < 64 else _synthetic(N // 10, 10, 4)
    post = prob.posterior
    rng = np.random.default_rng(T * 1000 + N)
    dev = DevicePosterior.from_posterior(post)
    try:
        dev.chain_reserve(T + 3)
        chain = rng.standard_normal((T + 3, post.N, post.D)).astype(np.float32)
        chain[:, 0, :] = 1.25                                        # constant series
        chain[:, 1, 0] = np.round(chain[:, 1, 0])                    # many ties
        if T > 2:
            chain[:, 2, 1] = np.sort(chain[:, 2, 1])[::-1]           # already sorted, descending
        for t in range(T + 3):
            dev.set_state(chain[t].astype(np.float64))
            dev.chain_push()
        assert dev.chain_length() == T + 3
        with pytest.raises(Exception, match="full"):
            dev.chain_push()
        got = dev.chain_get()
        assert np.array_equal(got, chain.astype(np.float64))
        per = [0.5, 2.5, 5, 16, 50, 84, 95, 97.5, 99.5, 0, 100]
        first = 2                                                    # burn-in of two stored iterates
        st = dev.chain_stats(first, T, per, want_var=True)
        ref = chain[first:first + T].astype(np.float64)
        np.testing.assert_allclose(st["mean"], ref.mean(axis=0), rtol=0, atol=2e-7 * (1 + np.abs(ref).max()))
        np.testing.assert_allclose(st["var"], ref.var(axis=0), rtol=2e-6, atol=1e-7)
        srt = np.sort(ref, axis=0)
        pos = np.array(per) / 100 * (T - 1)
        lo = np.floor(pos).astype(int)
        hi = np.minimum(lo + 1, T - 1)
        assert np.array_equal(st["lo"], srt[lo]) and np.array_equal(st["hi"], srt[hi])       # exact order statistics
        np.testing.assert_allclose(st["frac"], pos - lo, rtol=0, atol=1e-12)
        np.testing.assert_allclose(_np_percentile_from_brackets(st), np.percentile(ref, per, axis=0), rtol=1e-12, atol=1e-12)
        with pytest.raises(Exception, match="outside the stored chain"):
            dev.chain_stats(first, T + 2, per)
        dev.chain_reserve(0)
        with pytest.raises(Exception, match="no chain"):
            dev.chain_push()
    finally:
        dev.close()


@pytest.mark.parametrize("name", ["ref_n4_6x5", "ref_n8_5x6_censored"])
def test_spatial_potential_vs_reference_golden(lib, name):
    """bt_spatial_prior_unweighted at the golden states == reference L22 neglog_pdf(with_weights=False)
    (tests/golden/mml_golden.json, made by make_golden_mml.py)."""
    from beetroots_b200 import DevicePosterior
    with open(os.path.join(GOLDEN, "mml_golden.json")) as f:
        ref = json.load(f)[f"{name}|scale=1.0|N0=1"]["gX_per_dim"]
    g = load_golden(name)
    post, _ = problem_from_golden(g)
    states = [g["theta0"]] + [g[f"t{t}_Theta"] for t in range(1, len(g["kernels"]) + 1)]
    dev = DevicePosterior.from_posterior(post)
    try:
        for th, r in zip(states, ref):
            dev.set_state(th)
            np.testing.assert_allclose(dev.spatial_prior_unweighted(), r, rtol=1e-5)
            # weights do not enter, whatever they are
            post.prior_spatial.weights = post.prior_spatial.weights * 3.0
            dev.update_spatial_weights(post.prior_spatial, post.prior_indicator)
            np.testing.assert_allclose(dev.spatial_prior_unweighted(), r, rtol=1e-5)
    finally:
        dev.close()


class _MemSaver:
    def __init__(self):
        self.memory, self.tau, self.Theta, self.additional = {}, [], [], None

    def initialize_memory(self, T_MC, t, **kw):
        self.memory = {"started": True}

    def update_memory(self, t, Theta, additional_sampling_log, **kw):
        self.Theta.append(Theta.copy())
        self.tau.append(np.array(additional_sampling_log["tau"]))

    def check_need_to_update_memory(self, t):
        return True

    def check_need_to_save(self, t):
        return False

    def save_to_file(self):
        pass

    def save_additional(self, list_arrays, list_names):
        self.additional = dict(zip(list_names, list_arrays))


def test_sample_with_regu_weight_update_and_device_chain(lib):
    """``sample(..., regu_spatial_N0=3)``: the logged tau follows the reference recursion (mirror optimiser fed with the
    float64 oracle potential of the saved iterates), prior_spatial.weights is updated in place like the reference
    does, the step keeps using the new weights (objective terms agree with the oracle at the final tau); the chain
    kept on the device equals the iterates handed to the saver and its summaries equal numpy's."""
    from beetroots_b200 import B200Sampler, MySamplerParams
    from beetroots_b200.mml import EBayesMMLELogRate
    from beetroots_b200.modelling import build_neighbor_table
    from oracle import posterior as OP
    prob = _synthetic(12, 9, 5)
    post = prob.posterior
    nbr = build_neighbor_table(post.prior_spatial.list_edges, post.N, 4)
    tau0 = post.prior_spatial.weights.copy()
    T, N0, T_BI = 14, 3, 4
    smp = B200Sampler(MySamplerParams(1e-4, 1e-5, 0.99, [0.5, 0.5], 5, True, False), post.D, post.L, post.N,
                      rng=np.random.default_rng(11), device_chain=T, device_chain_thin=2)
    sv = _MemSaver()
    smp.sample(post, sv, T, Theta_0=prob.theta_init, disable_progress_bar=True, regu_spatial_N0=N0,
               regu_spatial_scale=1.0, T_BI=T_BI)
    assert len(sv.tau) == T
    opt = EBayesMMLELogRate(scale=1.0, N0=N0, N1=+np.inf, dim=post.D * post.N, vmin=1e-8, vmax=1e8, homogeneity=2.0,
                            exponent=0.8)
    tau = float(tau0.mean())
    prev = prob.theta_init
    for t in range(1, T + 1):
        if t >= N0:
            had, _ = OP.spatial_sums(prev, nbr)
            tau = opt.update(tau, t, float(0.5 * had.sum()))
            expect = tau * np.ones(post.D)
        else:
            expect = tau0
        np.testing.assert_allclose(sv.tau[t - 1], expect, rtol=1e-4)
        prev = sv.Theta[t - 1]
    assert np.abs(sv.tau[-1] - tau0).max() > 1e-3 * tau0.max(), "tau did not move: the test would be vacuous"
    np.testing.assert_allclose(post.prior_spatial.weights, sv.tau[-1], rtol=0, atol=0)
    # the device keeps sampling with the updated weights
    cur = OP.compute_all(post, sv.Theta[-1], nbr)
    dobj = smp._dev.objective_terms()
    np.testing.assert_allclose(dobj["nl_prior_spatial"].sum(), (cur["objective_pix_global"] - cur["nll_pix"]
                               - OP.indicator_penalty(sv.Theta[-1], post.prior_indicator)).sum(), rtol=1e-4)
    # on-device chain: iterates t = 6, 8, ..., 14 (t > T_BI, every second one)
    kept = [t for t in range(1, T + 1) if t > T_BI and t % 2 == 0]
    chain = smp._dev.chain_get()
    assert chain.shape[0] == len(kept)
    for k, t in enumerate(kept):
        np.testing.assert_allclose(chain[k], sv.Theta[t - 1], rtol=0, atol=0)
    summ = smp.posterior_summary(list_CI=[68, 95], from_scaled_to_lin=lambda a: 10.0 ** a)
    lin = 10.0 ** chain
    np.testing.assert_allclose(summ["Theta_MMSE_scaled"], chain.mean(axis=0), rtol=0, atol=1e-6)
    np.testing.assert_allclose(summ["Theta_MMSE"], 10.0 ** summ["Theta_MMSE_scaled"], rtol=1e-12)
    assert set(summ) == {"Theta_MMSE_scaled", "Theta_MMSE", "count", "per_2p5", "per_16p0", "per_84p0", "per_97p5"}
    for key, p in (("per_2p5", 2.5), ("per_16p0", 16.0), ("per_84p0", 84.0), ("per_97p5", 97.5)):
        np.testing.assert_allclose(summ[key], np.percentile(lin, p, axis=0), rtol=1e-12)


# ---- MTM proposals from the smooth indicator prior when there is no spatial prior -------------
def _no_spatial(H, W, K):
    from beetroots_b200 import modelling as M
    prob = _synthetic(H, W, K)
    p = prob.posterior
    return prob, M.Posterior(p.D, p.L, p.N, p.likelihood, None, p.prior_indicator)      # separable: one site, all pixels


def test_smooth_indicator_proposals_distribution(lib):
    """Device draws (Philox) of sampler/utils/utils.py:50-97: with probability Z_u / (1 + Z_u) uniform in the box,
    otherwise a generalised-Gaussian (power 4, scale delta) tail beyond the nearer bound, sides equally likely."""
    from scipy import stats
    from scipy.special import gamma
    from beetroots_b200 import DevicePosterior
    K = 384
    prob, post = _no_spatial(8, 8, K)
    pi = post.prior_indicator
    lo, hi, delta = pi.lower_bounds, pi.upper_bounds, pi.indicator_margin_scale
    dev = DevicePosterior.from_posterior(post)
    try:
        assert dev.n_sites == 1
        dev.compute_all(prob.theta_init)
        cand, us, ua = dev.debug_philox_mtm(0, K, 2024, 3)
        cand2, _, _ = dev.debug_philox_mtm(0, K, 2024, 4)
        assert not np.array_equal(cand, cand2)                                   # a fresh draw every iteration
        x = cand.reshape(-1, post.D).astype(np.float64)
        n = x.shape[0]
        for d in range(post.D):
            zu = 2 * (hi[d] - lo[d]) / (delta * gamma(0.25))
            p_out = 1 / (1 + zu)
            out = (x[:, d] < lo[d]) | (x[:, d] > hi[d])
            assert abs(out.mean() - p_out) < 5 * np.sqrt(p_out * (1 - p_out) / n), (d, out.mean(), p_out)
            inside = x[~out, d]
            assert stats.kstest((inside - lo[d]) / (hi[d] - lo[d]), "uniform").pvalue > 1e-4
            excess = np.where(x[out, d] > hi[d], x[out, d] - hi[d], x[out, d] - lo[d])
            assert excess.size > 200
            assert stats.kstest(excess, stats.gennorm(beta=4.0, scale=delta).cdf).pvalue > 1e-4
            assert abs((excess > 0).mean() - 0.5) < 5 * 0.5 / np.sqrt(excess.size)
        assert np.all((us > 0) & (us < 1)) and np.all((ua > 0) & (ua < 1))
    finally:
        dev.close()


def test_mtm_without_spatial_prior_vs_oracle(lib):
    """MTM on a separable posterior (likelihood + indicator, my_sampler.py:135-143 and posterior.py:105-110): the
    Philox path equals the injected path bit for bit, and both follow the float64 oracle on the same candidates
    (decisions identical up to near-ties, accepted states exact: they are copies of candidates)."""
    from beetroots_b200 import DevicePosterior
    from oracle.sampler import OracleSampler
    K, seed = 12, 77
    prob, post = _no_spatial(9, 7, K)
    devs = [DevicePosterior.from_posterior(post) for _ in range(2)]
    try:
        nbr = -np.ones((post.N, 1), dtype=np.int32)
        orc = OracleSampler(post, nbr, 1e-4, 1e-5, 0.99, K)
        pi = post.prior_indicator
        # start far from the mode (uniform in the box), otherwise no uniform proposal ever beats the current state
        start = pi.lower_bounds + (pi.upper_bounds - pi.lower_bounds) * np.random.default_rng(5).uniform(size=(post.N, post.D))
        start = start.astype(np.float32).astype(np.float64)
        orc.init(start)
        for dev in devs:
            dev.compute_all(start)
            dev.init_v_from_grad()
        n_flip = 0
        for t in range(1, 5):
            obj_scale = np.abs(orc.current["objective_pix_chromatic"])
            cand, us, ua = devs[0].debug_philox_mtm(0, K, seed, t)
            for dev, inj in zip(devs, (False, True)):
                dev.begin_iteration()
                if inj:
                    dev.mtm_site(0, K, seed, t, cand, us, ua)
                else:
                    dev.mtm_site(0, K, seed, t)
                dev.mtm_finish(0.99)
            th0, v0, _ = devs[0].get_state()
            th1, v1, _ = devs[1].get_state()
            assert np.array_equal(th0, th1) and np.array_equal(v0, v1)
            o_acc, o_lrg = orc.mtm_step({0: {"cand": cand.astype(np.float64), "u_sel": us.astype(np.float64),
                                             "u_acc": ua.astype(np.float64)}})
            acc, lrg = devs[0].step_stats()
            flips = np.where((acc > 0) != (o_acc > 0))[0]
            for n in flips:
                assert abs(lrg[n] - o_lrg[n]) < 1e-3 * (1 + abs(o_lrg[n])) + 1e-5 * obj_scale[n], (t, n, lrg[n], o_lrg[n])
            n_flip += flips.size
            same = (acc > 0) == (o_acc > 0)
            # which candidate is selected is itself a float comparison: allow rare selection near-ties
            bad = np.abs(th0 - orc.current["Theta"]).max(axis=1) > 1e-6
            assert (bad & same).sum() <= 1, (t, np.where(bad & same)[0])
            ok = same & ~bad
            # the reference (and the oracle) form the denominator as (sum of all weights) - w_selected in float64
            # (my_sampler.py:1142-1144): when the selected weight dominates, that difference carries an absolute error
            # of one ulp of the total, i.e. 2.2e-16 e^{log_rg} relative to itself (0.15 at log_rg = 35, certain accept
            # either way); the device sums the other weights directly
            tol = 2e-3 * (1 + np.abs(o_lrg)) + 1e-5 * obj_scale + 2.3e-16 * np.exp(np.minimum(np.abs(o_lrg), 40.0))
            viol = np.where(ok & ~(np.abs(lrg - o_lrg) <= tol))[0]
            assert viol.size == 0, (t, viol, lrg[viol], o_lrg[viol], obj_scale[viol])
            assert acc.sum() > 0
            if flips.size or (bad & same).any():
                from oracle import posterior as OP
                orc.current = OP.compute_all(post, th0, nbr)
                orc.v = v0.copy()
        assert n_flip <= 2
    finally:
        for dev in devs:
            dev.close()


# ---- optimisation mode (is_stochastic=False) -------------------------------------
@pytest.mark.parametrize("name", ["ref_n4_6x5", "ref_n8_5x6_censored"])
def test_optimisation_mode_vs_reference_golden(lib, name):
    """Device PMALA / MTM optimisation steps with the reference's recorded noise (tests/golden/*_optim.npz):
    accept decisions identical, states within rtol 1e-5, v within 2e-5 of its scale, objective within 1e-6."""
    from beetroots_b200 import B200Sampler, MySamplerParams
    g, go = load_golden(name), load_golden(name + "_optim")
    post, _ = problem_from_golden(g)
    K = int(g["K"])
    smp = B200Sampler(MySamplerParams(1e-4, 1e-5, 0.99, [0.5, 0.5], K, False, False), post.D, post.L, post.N)
    dev = smp.attach(post)
    dev.compute_all(g["theta0"])
    dev.init_v_from_grad()
    for t, kind in enumerate(go["kernels"], start=1):
        if kind == 1:
            acc, proba = smp.generate_new_sample_pmala_rmsprop(t, post, noise={"z": go[f"t{t}_z"]})
            assert float(acc) == float(go[f"t{t}_accept"]) and proba == int(acc)
        else:
            full = {}
            for s, idx in post.dict_sites.items():
                c = np.zeros((post.N, K, post.D))
                c[idx] = go[f"t{t}_s{s}_cand"]
                full[s] = {"cand": c}
            acc_mean, acc_mean2 = smp.generate_new_sample_mtm(t, post, noise=full)
            a, _ = dev.step_stats()
            assert np.array_equal(a, go[f"t{t}_accept"]) and acc_mean == acc_mean2 == a.mean()
        th, v, _ = dev.get_state()
        np.testing.assert_allclose(th, go[f"t{t}_Theta"], rtol=1e-5, atol=2e-6)
        assert np.max(np.abs(v - go[f"t{t}_v"])) <= 2e-5 * np.abs(go[f"t{t}_v"]).max()
        obj = dev.objective_terms()["objective"]
        assert abs(obj - float(go[f"t{t}_objective"])) < 1e-6 * abs(float(go[f"t{t}_objective"]))
    dev.close()


def test_optimisation_pmala_retry_and_reject_vs_oracle(lib):
    """Step sizes large enough that the plain gradient step overshoots: the noisy retry (my_sampler.py:939-950) and the
    rejection (state restored, v still updated with the rejected candidate's gradient, :952-953) against the oracle."""
    from beetroots_b200 import DevicePosterior
    from beetroots_b200.modelling import build_neighbor_table
    from oracle.sampler import OracleSampler
    prob = _synthetic(9, 8, 4)
    post = prob.posterior
    nbr = build_neighbor_table(post.prior_spatial.list_edges, post.N, 4)
    seen = set()
    for eps in (1e-4, 3e-2, 1.0):
        dev = DevicePosterior.from_posterior(post)
        orc = OracleSampler(post, nbr, eps, 1e-5, 0.99, 4)
        orc.init(prob.theta_init)
        dev.compute_all(prob.theta_init)
        dev.init_v_from_grad()
        rng = np.random.default_rng(int(eps * 1e6))
        for t in range(1, 5):
            z = rng.standard_normal((post.N, post.D)).astype(np.float32)
            before = orc.current["Theta"].copy()
            step_std = np.sqrt(eps / (1e-5 + np.sqrt(orc.v)))
            o_acc, _ = orc.pmala_optim_step(z.astype(np.float64))
            acc, _ = dev.optim_pmala_step(eps, 1e-5, 0.99, 0, t, z)
            assert bool(acc) == bool(o_acc), (eps, t)
            th, v, _ = dev.get_state()
            if not acc:
                assert np.array_equal(th.astype(np.float32), before.astype(np.float32))      # restored exactly
            # same float32 bound as the sampling-mode chain test: rounding of nearly cancelling gradients is amplified by
            # the preconditioner, at most 0.5 % of the local step scale sqrt(eps G)
            assert np.all(np.abs(th - orc.current["Theta"]) <= 1e-5 * np.abs(th) + 2e-6 + 5e-3 * step_std)
            assert np.max(np.abs(v - orc.v)) <= 1e-4 * np.abs(orc.v).max()
            cur = dev.get_current()
            np.testing.assert_allclose(cur["objective_global"], orc.current["objective_global"], rtol=1e-6)
            seen.add(bool(acc))
        dev.close()
    assert seen == {True, False}, "both the accepted and the rejected branch must be exercised"


def test_sample_in_optimisation_mode(lib):
    """Full ``sample`` loop with is_stochastic=False: the objective never increases, clppd is exp(-nll) of the best
    saved iterate (my_sampler.py:208-212), p-values come from ESS_OPTIM replications at the final iterate (:222-256)."""
    from beetroots_b200 import B200Sampler, MySamplerParams
    prob = _synthetic(8, 8, 6)
    post = prob.posterior
    smp = B200Sampler(MySamplerParams(1e-4, 1e-5, 0.99, [0.3, 0.7], 6, False, False), post.D, post.L, post.N,
                      rng=np.random.default_rng(5))
    smp.ESS_OPTIM = 40
    objs = []

    class Saver(_MemSaver):
        def update_memory(self, t, Theta, dict_objective, **kw):
            objs.append(dict_objective["objective"])
            self.Theta.append(Theta.copy())

    sv = Saver()
    smp.sample(post, sv, 12, Theta_0=prob.theta_init, disable_progress_bar=True)
    assert len(objs) == 12 and np.all(np.diff(objs) <= 1e-6 * np.abs(objs[0])), objs
    assert objs[-1] < objs[0]
    clppd, pvy, pvl = sv.additional["clppd"], sv.additional["p-values-y"], sv.additional["p-values-llh"]
    from oracle import posterior as OP
    from beetroots_b200.modelling import build_neighbor_table
    fm = OP.forward_map_compute_all(post.likelihood.forward_map, sv.Theta[int(np.argmin(objs))], False)
    nll_full = OP.likelihood_eval(post.likelihood, fm, None, False)["nll_full"]
    np.testing.assert_allclose(clppd, np.exp(-nll_full), rtol=2e-4, atol=1e-30)
    assert np.all((pvy >= 0) & (pvy <= 0.5)) and np.all((pvl >= 0) & (pvl <= 1))
    assert np.allclose(pvy * 40, np.round(pvy * 40), atol=1e-3) and pvl.std() > 0


def test_kernel_profile_and_scratch_reservation(built_lib):
    """bt_reserve_scratch / bt_kernel_profile (the measurement hooks of bench.py): after the reservation a step allocates nothing
    (the scratch pointers are stable), and the per-kernel timers return positive device times with the algorithmic bytes of
    the launch shapes."""
    from beetroots_b200 import MLP_BF16_TC, B200Sampler, DevicePosterior, MySamplerParams
    from beetroots_b200.synthetic import make_problem
    from oracle import posterior as OP
    K = 6
    prob = make_problem(24, 20, lambda fm, th: OP.forward_map_compute_all(fm, th, False)["log_f_Theta"], L=10, k_mtm=K)
    post = prob.posterior
    dev = DevicePosterior.from_posterior(post, mlp_mode=MLP_BF16_TC)
    n_max = max(p.size for p in dev.site_pixels)
    dev.reserve_scratch(n_max * (K + 1), post.N)
    dev.compute_all(prob.theta_init)
    dev.init_v_from_grad()
    smp = B200Sampler(MySamplerParams(1e-4, 1e-5, 0.99, [0.5, 0.5], K, True, False), post.D, post.L, post.N, mlp_mode=MLP_BF16_TC)
    smp.attach(post, dev)
    smp.philox_seed = 5
    dev.kernel_profile(enable=True, reset=True)
    smp.generate_new_sample_mtm(1, post)
    smp.generate_new_sample_pmala_rmsprop(2, post)
    prof = dev.kernel_profile(enable=False, reset=True)
    n_pix = post.N // 2
    for name in ("pmala_propose", "pmala_accept", "mtm_gen", "mtm_nl", "mtm_select"):
        ms, nbytes, launches = prof[name]
        assert launches == 2 and ms > 0 and nbytes > 0, (name, prof[name])
    # candidate rows written by mtm_gen: n_pix x K rows of D floats per colour class, plus theta + stencil per pixel
    assert prof["mtm_gen"][1] >= 2 * n_pix * K * post.D * 4
    assert prof["mtm_nl"][1] > prof["mtm_select"][1] > 0
    assert dev.kernel_profile(enable=False)["mtm_nl"][2] == 0          # reset
    dev.close()
