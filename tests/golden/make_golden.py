"""Generate the golden fixtures of tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For every case it (1) builds the reference objects (NeuralNetworkApprox on the stand-in nnbma,
MixingModelsLikelihood, L22DiscreteGradSpatialPrior, SmoothIndicatorPrior, Posterior,
MySampler), (2) records a noise tape while running the reference kernels
``generate_new_sample_mtm`` / ``generate_new_sample_pmala_rmsprop`` for a few iterations,
(3) replays the same tape through the oracle port and asserts agreement (this is what pins the
oracle), and (4) stores inputs, tape and reference outputs so that the GPU box -- which has no
reference tree -- can check both the oracle and the CUDA path against them.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402

refshim.install()

import pandas as pd  # noqa: E402
from nnbma import networks as fake_nnbma  # noqa: E402

from beetroots.modelling.forward_maps.neural_network_approx import NeuralNetworkApprox as RefNN  # noqa: E402
from beetroots.modelling.likelihoods.approx_censored_add_mult import MixingModelsLikelihood as RefLik  # noqa: E402
from beetroots.modelling.posterior import Posterior as RefPosterior  # noqa: E402
from beetroots.modelling.priors.l22_discrete_grad_prior import L22DiscreteGradSpatialPrior as RefL22  # noqa: E402
from beetroots.modelling.priors.smooth_indicator_prior import SmoothIndicatorPrior as RefInd  # noqa: E402
from beetroots.modelling.priors.spatial_prior_params import SpatialPriorParams  # noqa: E402
from beetroots.sampler.my_sampler import MySampler  # noqa: E402
from beetroots.sampler.utils.my_sampler_params import MySamplerParams  # noqa: E402

from beetroots_b200 import modelling as M  # noqa: E402
from beetroots_b200 import synthetic as S  # noqa: E402
from beetroots_b200.network import synthetic_weights  # noqa: E402
from oracle import posterior as OP  # noqa: E402
from oracle.sampler import OracleSampler  # noqa: E402


class TapeRNG:
    """Proxy around numpy Generator recording the draws the kernels consume."""

    def __init__(self, rng, f32=False):
        self.rng, self.log, self.f32 = rng, [], f32

    def standard_normal(self, size=None):
        z = self.rng.standard_normal(size=size)
        if self.f32:            # compact fixtures: the reference consumes float32-representable draws, stored as float32
            z = z.astype(np.float32).astype(np.float64)
        self.log.append(("z", z.copy()))
        return z

    def uniform(self, low=0.0, high=1.0, size=None):
        u = self.rng.uniform(low, high, size=size)
        self.log.append(("u", np.array(u).copy()))
        return u

    def integers(self, *a, **k):
        return self.rng.integers(*a, **k)

    def choice(self, K, p=None):
        # bit-identical to Generator.choice(K, p=p) and consumes one double
        u = self.rng.random()
        self.log.append(("u_sel", u))
        cdf = np.cumsum(p)
        cdf /= cdf[-1]
        return int(np.searchsorted(cdf, u, side="right"))

    def multinomial(self, *a, **k):
        return self.rng.multinomial(*a, **k)

    def normal(self, loc=0.0, scale=1.0, size=None):
        x = self.rng.normal(loc=loc, scale=scale, size=size)
        self.log.append(("normal", (np.asarray(x) - loc) / scale))          # the underlying standard normals
        self.log.append(("normal_raw", np.asarray(x).copy()))
        return x

    def lognormal(self, mean=0.0, sigma=1.0, size=None):
        x = self.rng.lognormal(mean=mean, sigma=sigma, size=size)
        self.log.append(("lognormal", (np.log(np.asarray(x)) - mean) / sigma))
        self.log.append(("lognormal_raw", np.asarray(x).copy()))
        return x

    def __getattr__(self, name):
        return getattr(self.rng, name)


def build_case(H, W, L, K, nnn, seed, censor_boost, drop=()):
    N_full = H * W
    keep = np.array([i for i in range(N_full) if i not in set(drop)])
    N = keep.size
    D = 4
    net = synthetic_weights(seed=seed, n_out=L)
    names = [f"line{i}" for i in range(L)]
    model_name = f"synthetic_{seed}_{L}"
    fake_nnbma.REGISTRY[model_name] = (net, names)
    fixed = {"kappa": None, "P": None, "radm": None, "Avmax": None,
             "angle": float(S.lin_to_scaled(np.array([1.0, 1e5, 1.0, 1.0, 0.0]))[4])}
    ref_fm = RefNN("/nonexistent", model_name, fixed)
    ref_fm.restrict_to_output_subset(names)

    X = np.repeat(np.arange(H), W)[keep]
    Y = np.tile(np.arange(W), H)[keep]
    theta_true = S.smooth_theta_map(H, W, seed + 1)[keep]
    f = ref_fm.evaluate(theta_true)
    sa = 1.38715e-10 * censor_boost
    sm = float(np.log(1.1))
    rng = np.random.default_rng(2023 + seed)
    y = np.maximum(3 * sa, rng.lognormal(-(sm ** 2) / 2, sm, size=(N, L)) * f + rng.normal(0, sa, size=(N, L)))
    omega = 3 * sa
    rl = np.random.default_rng(77 + seed)
    a0 = rl.uniform(-8.9, -8.4, size=L) + np.log10(censor_boost)
    a1 = rl.uniform(0.15, 0.85, size=L)
    with tempfile.TemporaryDirectory() as tmp:
        csv = os.path.join(tmp, "best_params.csv")
        pd.DataFrame({"n": 0, "ell": np.arange(L), "line": names, "a0_best": a0, "a1_best": a1,
                      "target_best": 0.0}).to_csv(csv, index=False)
        ref_lik = RefLik(ref_fm, D, L, N, y, sa, sm, omega, csv, names)
    df = pd.DataFrame({"idx": np.arange(N), "X": X, "Y": Y}).set_index(["X", "Y"])
    tau = np.array([10.0, 1.0, 2.0, 4.0])
    ref_sp = RefL22(SpatialPriorParams("L2-gradient", nnn, tau), "syn", D, N, df, [0, 1, 2, 3])
    lo = S.lin_to_scaled(S.LOWER_LIN)
    hi = S.lin_to_scaled(S.UPPER_LIN)
    ref_ind = RefInd(D, N, 0.1, lo, hi, [0, 1, 2, 3])
    ref_post = RefPosterior(D, L, N, ref_lik, ref_sp, ref_ind)

    # mirrors with the same inputs
    fm = M.NeuralNetworkApprox(net, fixed)
    lik = M.MixingModelsLikelihood(fm, D, L, N, y, sa, sm, omega, a0=a0[None, :], a1=a1[None, :])
    sp = M.L22DiscreteGradSpatialPrior(D, N, X, Y, None, nnn, tau)
    ind = M.SmoothIndicatorPrior(D, N, 0.1, lo[:4], hi[:4])
    post = M.Posterior(D, L, N, lik, sp, ind)
    assert np.array_equal(sp.list_edges, ref_sp.list_edges)
    for k in ref_sp.dict_sites:
        assert np.array_equal(sp.dict_sites[k], ref_sp.dict_sites[k])
    assert np.allclose(lik.log_fm1, ref_lik.log_fm1, rtol=0, atol=0)
    # start a little outside the box for a few pixels so the indicator prior is active
    theta0 = theta_true + 0.05 * np.random.default_rng(5 + seed).standard_normal((N, D))
    theta0[0, 1] = hi[1] + 0.03
    theta0[N - 1, 3] = lo[3] - 0.02
    inputs = dict(H=H, W=W, L=L, K=K, nnn=int(nnn), seed=seed, X=X, Y=Y, y=y, sigma_a=sa, sigma_m=sm,
                  omega=omega, a0=a0, a1=a1, tau=tau, lo=lo[:4], hi=hi[:4], delta=0.1,
                  angle_fixed=fixed["angle"], theta0=theta0)
    return ref_post, post, inputs


def run_case(name, H, W, L, K, nnn, seed, censor_boost, drop=(), kernels=(0, 1, 1, 0), eps=1e-4, compact=False):
    """``compact``: fixtures of the BASELINE config sizes (c2 30x30 K=50, c3 64x64): proposals and Gaussian draws are rounded to
    float32 BEFORE the reference consumes them (so the float32 tape is exactly what it used), v / grad are stored as float32,
    the large forward-map arrays and the model-check block are left out (the small cases cover them)."""
    ref_post, post, inputs = build_case(H, W, L, K, nnn, seed, censor_boost, drop)
    N, D = ref_post.N, ref_post.D
    params = MySamplerParams(eps, 1e-5, 0.99, [0.5, 0.5], K, True, False)
    smp = MySampler(params, D, L, N, rng=np.random.default_rng(42))
    tape = TapeRNG(smp.rng, f32=compact)
    smp.rng = tape
    props = []
    orig = smp.generate_random_start_Theta_1pix

    def rec(Theta, posterior, idx_pix):
        c = orig(Theta, posterior, idx_pix)
        if compact:
            c = c.astype(np.float32).astype(np.float64)
        props.append(c.copy())
        return c

    smp.generate_random_start_Theta_1pix = rec
    theta0 = inputs["theta0"]
    smp.current = ref_post.compute_all(theta0 * 1, compute_derivatives_2nd_order=False)
    smp.v = smp.current["grad"].flatten() ** 2
    smp.j_t = np.zeros((N * D,))

    nbr = M.build_neighbor_table(post.prior_spatial.list_edges, N)
    orc = OracleSampler(post, nbr, eps, 1e-5, 0.99, K)
    orc.init(theta0)
    out = dict(inputs)
    out["eps"] = eps
    cen = (ref_post.likelihood.y <= ref_post.likelihood.omega).mean()
    print(f"[{name}] N={N} censored fraction {cen:.2f}")

    def cmp(a, b, what, tol=1e-9):
        err = np.max(np.abs(a - b) / (1e-300 + np.maximum(np.abs(a), np.abs(b)) + 1e-12))
        assert err < tol, f"{name}: oracle != reference for {what}: {err}"

    # --- compute_all parity (pins oracle/posterior.py)
    cur = smp.current
    cmp(orc.current["forward_map_evals"]["log_f_Theta"], cur["forward_map_evals"]["log_f_Theta"], "log_f")
    # (elementwise relative error: with 10^5 Jacobian entries some are ~1e-6 of the scale, hence the looser bound)
    cmp(orc.current["forward_map_evals"]["grad_log_f_Theta"], cur["forward_map_evals"]["grad_log_f_Theta"], "grad_log_f",
        1e-7 if compact else 1e-9)
    cmp(orc.current["objective_pix_chromatic"], cur["objective_pix_chromatic"], "obj chromatic")
    cmp(orc.current["objective_pix_global"], cur["objective_pix_global"], "obj global")
    cmp(orc.current["grad"], cur["grad"], "grad")
    out["init_log_f"] = cur["forward_map_evals"]["log_f_Theta"]
    out["init_grad_log_f"] = cur["forward_map_evals"]["grad_log_f_Theta"]
    out["init_nll_pix"] = ref_post.likelihood.neglog_pdf(cur["forward_map_evals"], cur["nll_utils"], pixelwise=True)
    out["init_grad_lik"] = ref_post.likelihood.gradient_neglog_pdf(cur["forward_map_evals"], cur["nll_utils"])
    out["init_obj_chrom"] = cur["objective_pix_chromatic"]
    out["init_obj_glob"] = cur["objective_pix_global"]
    out["init_grad"] = cur["grad"]

    sites = list(ref_post.dict_sites.keys())
    out["kernels"] = np.array(kernels)
    for t, kind in enumerate(kernels, start=1):
        tape.log.clear()
        props.clear()
        if kind == 1:
            acc, lpa = _run_pmala(smp, ref_post)
            noise = {}
            for i, s in enumerate(sites):
                noise[s] = {"z": tape.log[2 * i][1], "u": tape.log[2 * i + 1][1]}
                out[f"t{t}_s{s}_z"], out[f"t{t}_s{s}_u"] = noise[s]["z"], noise[s]["u"]
            o_acc, o_lpa = orc.pmala_step(noise)
        else:
            acc, lpa = _run_mtm(smp, ref_post)
            noise, pos = {}, 0
            for i, s in enumerate(sites):
                n_pix = ref_post.dict_sites[s].size
                usel = np.array([e[1] for e in tape.log[pos:pos + n_pix]])
                assert all(e[0] == "u_sel" for e in tape.log[pos:pos + n_pix])
                uacc = tape.log[pos + n_pix][1]
                pos += n_pix + 1
                noise[s] = {"cand": props[i], "u_sel": usel, "u_acc": uacc}
                out[f"t{t}_s{s}_cand"], out[f"t{t}_s{s}_usel"], out[f"t{t}_s{s}_uacc"] = props[i], usel, uacc
            o_acc, o_lpa = orc.mtm_step(noise)
        assert np.array_equal(acc > 0, o_acc > 0), f"{name}: accept decisions differ at t={t}"
        cmp(lpa, o_lpa, f"log accept t={t}", 1e-7)
        cmp(smp.current["Theta"], orc.current["Theta"], f"Theta t={t}")
        cmp(smp.v.reshape(N, D), orc.v, f"v t={t}", 1e-8)
        assert np.array_equal(smp.j_t.reshape(N, D), orc.j_t)
        out[f"t{t}_accept"] = acc
        out[f"t{t}_logpa"] = lpa
        out[f"t{t}_Theta"] = smp.current["Theta"].copy()
        out[f"t{t}_v"] = smp.v.reshape(N, D).copy()
        out[f"t{t}_j"] = smp.j_t.reshape(N, D).copy()
        out[f"t{t}_grad"] = smp.current["grad"].copy()
        out[f"t{t}_obj_chrom"] = smp.current["objective_pix_chromatic"].copy()
        print(f"[{name}] t={t} kind={'PMALA' if kind else 'MTM'} accept={acc.mean():.3f} OK (oracle == reference)")
    if compact:
        for k in list(out):
            if k.endswith("_cand") or k.endswith("_z") or k.endswith("_v") or k.endswith("_grad"):
                out[k] = out[k].astype(np.float32)
            elif k.endswith("_j"):
                out[k] = out[k].astype(np.int16)
        for k in ("init_log_f", "init_grad_log_f", "init_nll_pix", "init_grad_lik", "init_obj_glob"):
            out.pop(k, None)
        out["censored_fraction"] = cen
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(f"[{name}] written ({os.path.getsize(os.path.join(HERE, name + '.npz')) / 1e6:.2f} MB)")
        return
    # --- online model checking (my_sampler.py:158-214): two updates at the final state, replication noise taped
    dmc = {"clppd_online": np.zeros((N, L)), "best_objective": np.inf, "count_pval": 0,
           "p_values_y": np.zeros((N, L)), "p_values_llh": np.zeros((N,))}
    orc.model_check_reset()
    for r in range(2):
        tape.log.clear()
        dobj, nll_full = ref_post.compute_all_for_saver(smp.current["Theta"], smp.current["forward_map_evals"],
                                                        smp.current["nll_utils"])
        dmc = smp._update_model_check_values(dmc, ref_post.likelihood, nll_full, dobj["objective"] * 1)
        za = [e[1] for e in tape.log if e[0] == "normal"][0]
        zm = [e[1] for e in tape.log if e[0] == "lognormal"][0]
        out[f"mc{r}_za"], out[f"mc{r}_zm"] = za, zm
        ea = [e[1] for e in tape.log if e[0] == "normal_raw"][0]
        em = [e[1] for e in tape.log if e[0] == "lognormal_raw"][0]
        # exact replicated observation of the reference (za / zm reproduce it only up to one ulp)
        y_rep = np.maximum(ref_post.likelihood.omega, em * smp.current["forward_map_evals"]["f_Theta"] + ea)
        orc.model_check_update(za, zm, y_rep=y_rep)
    fin = smp._finalize_model_check_values(dict(dmc), ref_post.likelihood, smp.current["forward_map_evals"], nll_full)
    o_c, o_py, o_pl = orc.model_check_get()
    cmp(fin["clppd_online"], o_c, "clppd", 1e-9)
    print("   pval_y mismatches", np.sum(fin["p_values_y"] != o_py), "pval_llh mismatches", np.sum(fin["p_values_llh"] != o_pl),
          "count", dmc["count_pval"], orc.mc["count"])
    assert np.array_equal(fin["p_values_y"], o_py) and np.array_equal(fin["p_values_llh"], o_pl), "model-check p-values"
    out["mc_clppd"], out["mc_pval_y"], out["mc_pval_llh"] = fin["clppd_online"], fin["p_values_y"], fin["p_values_llh"]
    out["mc_dict_objective"] = np.array([dobj["nll"], dobj["objective"]])
    out["mc_nl_prior_spatial"], out["mc_nl_prior_indicator"] = dobj["nl_prior_spatial"], dobj["nl_prior_indicator"]
    print(f"[{name}] model check (2 updates) OK (oracle == reference)")
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)


def _run_pmala(smp, post):
    """Per-pixel outputs of generate_new_sample_pmala_rmsprop: the reference only returns means
    (:906), so re-derive the per-pixel arrays by wrapping np.mean is not possible; instead we
    call the kernel and read the per-pixel arrays through a tiny subclass hook."""
    return _call_with_capture(smp, smp.generate_new_sample_pmala_rmsprop, post)


def _run_mtm(smp, post):
    return _call_with_capture(smp, smp.generate_new_sample_mtm, post)


def _call_with_capture(smp, fn, post):
    """The kernels return (accept_total.mean(), log_proba.mean()).  Capture the arrays by
    temporarily patching ``numpy.ndarray.mean`` is not possible (C type), so patch ``np.mean``
    (used by MTM, :1202) and intercept through frame locals for PMALA via sys.settrace."""
    captured = {}

    def tracer(frame, event, arg):
        if event == "return" and frame.f_code.co_name in (
                "generate_new_sample_pmala_rmsprop", "generate_new_sample_mtm"):
            loc = frame.f_locals
            captured["acc"] = np.array(loc["accept_total"], dtype=float).copy()
            key = "log_proba_accept_total" if "log_proba_accept_total" in loc else "log_rg_total"
            captured["lpa"] = np.array(loc[key], dtype=float).copy()
        return tracer

    sys.settrace(tracer)
    try:
        fn(1, post)
    finally:
        sys.settrace(None)
    return captured["acc"], captured["lpa"]


if __name__ == "__main__" and "--big-only" not in sys.argv:
    # c2-like: 4-neighbour rectangular map, little censoring
    run_case("ref_n4_6x5", H=6, W=5, L=10, K=7, nnn=False, seed=0, censor_boost=1.0,
             kernels=(0, 1, 1, 0, 1, 0))
    # c3-like: heavy censoring, 8 neighbours, non-rectangular (two pixels removed)
    run_case("ref_n8_5x6_censored", H=5, W=6, L=6, K=5, nnn=True, seed=3, censor_boost=6.0,
             drop=(7, 29), kernels=(1, 0, 0, 1, 1))
if __name__ == "__main__":
    if "--big" in sys.argv or "--big-only" in sys.argv:
        # c2 of BASELINE.json: 30x30, D=4, L=10, K=50, 4-neighbour (data/toycases/nn_N30_fixed_angle.yaml), 6 iterations
        run_case("ref_c2_30x30", H=30, W=30, L=10, K=50, nnn=False, seed=11, censor_boost=1.0, kernels=(0, 1, 0, 1, 1, 0),
                 eps=float(os.environ.get("BT_GOLDEN_EPS_C2", "3e-3")), compact=True)
        # c3-like: 64x64, 20-40 % censored, K=10 (the reference's value at N >= 64^2), a step size with real rejections
        run_case("ref_c3_64x64_censored", H=64, W=64, L=10, K=10, nnn=False, seed=12,
                 censor_boost=float(os.environ.get("BT_GOLDEN_BOOST_C3", "1.5")), kernels=(1, 0, 1, 0, 1),
                 eps=float(os.environ.get("BT_GOLDEN_EPS_C3", "1e-2")), compact=True)
    print("golden fixtures written to", HERE)
