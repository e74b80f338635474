"""CPU: the oracle port against (1) the golden vectors recorded from the UNMODIFIED reference
(tests/golden/make_golden.py) and (2) the reference's own known-answer tests for this path."""
import os

import numpy as np
import pytest

from _cases import GOLDEN, GOLDEN_BIG, GOLDEN_CASES, golden_noise, load_golden, problem_from_golden
from oracle import posterior as OP
from oracle.sampler import OracleSampler


def rel(a, b):
    return np.max(np.abs(a - b) / (1e-12 + np.maximum(np.abs(a), np.abs(b))))


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_compute_all_matches_reference(name):
    g = load_golden(name)
    post, nbr = problem_from_golden(g)
    cur = OP.compute_all(post, g["theta0"], nbr)
    assert rel(cur["forward_map_evals"]["log_f_Theta"], g["init_log_f"]) < 1e-12
    assert rel(cur["forward_map_evals"]["grad_log_f_Theta"], g["init_grad_log_f"]) < 1e-10
    assert rel(cur["nll_pix"], g["init_nll_pix"]) < 1e-10
    assert rel(cur["grad_lik"], g["init_grad_lik"]) < 1e-9
    assert rel(cur["objective_pix_chromatic"], g["init_obj_chrom"]) < 1e-10
    assert rel(cur["objective_pix_global"], g["init_obj_glob"]) < 1e-10
    assert rel(cur["grad"], g["init_grad"]) < 1e-9


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_steps_match_reference(name):
    g = load_golden(name)
    post, nbr = problem_from_golden(g)
    orc = OracleSampler(post, nbr, 1e-4, 1e-5, 0.99, int(g["K"]))
    orc.init(g["theta0"])
    for t, kind in enumerate(g["kernels"], start=1):
        noise = golden_noise(g, t, int(kind), post)
        acc, lpa = orc.pmala_step(noise) if kind == 1 else orc.mtm_step(noise)
        assert np.array_equal(acc > 0, g[f"t{t}_accept"] > 0), f"accept decisions differ at t={t}"
        assert rel(lpa, g[f"t{t}_logpa"]) < 1e-7
        assert rel(orc.current["Theta"], g[f"t{t}_Theta"]) < 1e-10
        assert rel(orc.v, g[f"t{t}_v"]) < 1e-8
        assert np.array_equal(orc.j_t, g[f"t{t}_j"])
        assert rel(orc.current["grad"], g[f"t{t}_grad"]) < 1e-8


# ---- the reference's own known-answer tests for this path, restated on the oracle ------------
class _Prior:
    pass


def _grid3():
    from beetroots_b200 import modelling as M
    X, Y = np.repeat(np.arange(3), 3), np.tile(np.arange(3), 3)
    sp = M.L22DiscreteGradSpatialPrior(1, 9, X, Y, None, False, np.ones(1))
    return sp, M.build_neighbor_table(sp.list_edges, 9, 4)


def test_l22_known_answers():
    """/root/reference/tests/modelling/priors/test_l22_discrete_grad_prior.py:61-183."""
    sp, nbr = _grid3()
    Theta = np.zeros((9, 1))
    Theta[4] = 1.0
    had, lap = OP.spatial_sums(Theta, nbr)
    assert np.array_equal(lap[:, 0], [0, -1, 0, -1, 4, -1, 0, -1, 0])      # :67-71
    assert np.array_equal(had[:, 0], [0, 1, 0, 1, 4, 1, 0, 1, 0])          # :92-96
    chrom = (sp.weights * had).sum()
    assert chrom == 8 and 0.5 * chrom == 4                                   # :132-135
    assert np.array_equal(2 * sp.weights * lap, 2 * lap)                     # :168-183


def test_indicator_known_answers():
    """/root/reference/tests/modelling/priors/test_smooth_indicator_prior.py:31-124."""
    p = _Prior()
    p.lower_bounds, p.upper_bounds, p.indicator_margin_scale = np.array([-1.0, -1.0]), np.array([1.0, 1.0]), 0.1
    inside = np.array([[0.0, 0.5]])
    assert OP.indicator_penalty(inside, p)[0] == 0
    out = np.array([[1.0 + 0.1 * 10, 0.0]])           # ten margins outside on one side -> (10)^4
    assert np.isclose(OP.indicator_penalty(out, p)[0], 1e4)                  # :47-48
    one = np.array([[1.1, -1.1]])                      # one margin outside: grad = +-4/Delta
    gr = OP.indicator_gradient(one, p)
    assert np.allclose(gr[0], [4 / 0.1, -4 / 0.1])                           # :87,97


def test_mtm_proposal_density_closed_form():
    """/root/reference/tests/sampler/utils/test_utils.py:51-79: 4 neighbours all at 0, tau = 1:
    log q(x) = log sum_m C(4,m) m exp(-m^2 x^2)."""
    from math import comb
    from beetroots_b200 import modelling as M
    X, Y = np.repeat(np.arange(3), 3), np.tile(np.arange(3), 3)
    sp = M.L22DiscreteGradSpatialPrior(1, 9, X, Y, None, False, np.ones(1))
    nbr = M.build_neighbor_table(sp.list_edges, 9, 4)
    post = M.Posterior(1, 1, 9, None, sp, None)
    orc = OracleSampler(post, nbr, 1e-4, 1e-5, 0.99, 3)
    Theta = np.zeros((9, 1))
    xs = np.array([0.0, 0.3, -1.2, 2.0])
    val = orc.proposal_log_density(Theta, np.array([4]), xs.reshape(1, 4, 1))[0]
    expected = np.log(sum(comb(4, m) * m * np.exp(-(m ** 2) * xs ** 2) for m in range(1, 5)))
    assert np.allclose(val, expected, rtol=1e-12)


def test_mills_ratio():
    """/root/reference/tests/modelling/likelihoods/test_utils.py:6-15."""
    assert np.isclose(OP._norm_pdf_cdf_ratio(0.0), 2 / np.sqrt(2 * np.pi))
    x = np.linspace(-30, 10, 200)
    r = OP._norm_pdf_cdf_ratio(x)
    assert np.all(np.diff(r) < 0)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_model_check_matches_reference(name):
    """clppd / Bayesian p-values after two updates at the final golden state (my_sampler.py:158-267)."""
    g = load_golden(name)
    post, nbr = problem_from_golden(g)
    orc = OracleSampler(post, nbr, 1e-4, 1e-5, 0.99, int(g["K"]))
    T = len(g["kernels"])
    orc.init(g[f"t{T}_Theta"])
    orc.model_check_reset()
    for r in range(2):
        orc.model_check_update(g[f"mc{r}_za"], g[f"mc{r}_zm"])
    c, py, pl = orc.model_check_get()
    assert rel(c, g["mc_clppd"]) < 1e-8
    assert np.mean(py != g["mc_pval_y"]) < 0.01 and np.mean(pl != g["mc_pval_llh"]) < 0.05   # za/zm reproduce y_rep to 1 ulp


# ---- regularisation-weight update: optimiser mirror and spatial potential vs the reference ----
def _mml_golden():
    import json
    with open(os.path.join(GOLDEN, "mml_golden.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("key", sorted(_mml_golden().keys()))
def test_regu_weight_optimiser_and_potential_vs_reference_golden(key):
    """tests/golden/make_golden_mml.py ran the reference's EBayesMMLELogRate / L22 prior; the host mirror
    (beetroots_b200/mml.py) and the oracle's neighbour sums must reproduce both."""
    from beetroots_b200.mml import EBayesMMLELogRate
    ref = _mml_golden()[key]
    name, scale, N0 = key.split("|")
    scale, N0 = float(scale.split("=")[1]), int(N0.split("=")[1])
    g = load_golden(name)
    post, nbr = problem_from_golden(g)
    states = [g["theta0"]] + [g[f"t{t}_Theta"] for t in range(1, len(g["kernels"]) + 1)]
    opt = EBayesMMLELogRate(scale=scale, N0=N0, N1=+np.inf, dim=post.D * post.N, vmin=1e-8, vmax=1e8, homogeneity=2.0,
                            exponent=0.8)
    tau = float(np.mean(g["tau"]))
    for t, th in enumerate(states, start=1):
        had, _ = OP.spatial_sums(th, nbr)
        gX = 0.5 * had.sum(axis=0)                                  # l22_discrete_grad_prior.py:106-128, factor 1/2
        np.testing.assert_allclose(gX, ref["gX_per_dim"][t - 1], rtol=1e-12)
        if t >= N0:
            tau = opt.update(tau, t, float(gX.sum()))
        np.testing.assert_allclose(tau, ref["tau"][t - 1], rtol=1e-12)
        np.testing.assert_allclose(opt.mean_theta, ref["mean_theta"][t - 1], rtol=1e-12)


# ---- optimisation mode (is_stochastic=False): oracle vs the tapes recorded from the reference ----
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_optimisation_mode_matches_reference(name):
    """tests/golden/make_golden_optim.py ran MySampler(is_stochastic=False); replaying its noise through
    OracleSampler.pmala_optim_step / mtm_optim_step reproduces every state (my_sampler.py:908-964, 1050-1086)."""
    g = load_golden(name)
    go = load_golden(name + "_optim")
    post, nbr = problem_from_golden(g)
    orc = OracleSampler(post, nbr, 1e-4, 1e-5, 0.99, int(g["K"]))
    orc.init(g["theta0"])
    for t, kind in enumerate(go["kernels"], start=1):
        if kind == 1:
            acc, _ = orc.pmala_optim_step(go[f"t{t}_z"])
            assert float(acc) == float(go[f"t{t}_accept"])
        else:
            acc = orc.mtm_optim_step({s: {"cand": go[f"t{t}_s{s}_cand"]} for s in post.dict_sites})
            assert np.array_equal(acc, go[f"t{t}_accept"])
        assert rel(orc.current["Theta"], go[f"t{t}_Theta"]) < 1e-9
        assert rel(orc.v, go[f"t{t}_v"]) < 1e-8
        assert abs(orc.current["objective_global"] - float(go[f"t{t}_objective"])) < 1e-8 * abs(float(go[f"t{t}_objective"]))


@pytest.mark.parametrize("name", GOLDEN_BIG)
def test_oracle_follows_the_baseline_sized_reference_tapes(name):
    """c2 (30x30, K=50) and the c3-like 64x64 censored map: the oracle replays the tape of the unmodified reference with identical
    accept / reject decisions at every pixel and iteration (thousands of decisions, 16-18 % rejections in the c3 case)."""
    g = load_golden(name)
    post, nbr = problem_from_golden(g)
    orc = OracleSampler(post, nbr, float(g["eps"]), 1e-5, 0.99, int(g["K"]))
    orc.init(g["theta0"])
    assert rel(orc.current["objective_pix_chromatic"], g["init_obj_chrom"]) < 1e-9
    n_rej = 0
    for t, kind in enumerate(g["kernels"], start=1):
        noise = golden_noise(g, t, int(kind), post)
        noise = {s: {k: np.asarray(v, dtype=np.float64) for k, v in d.items()} for s, d in noise.items()}
        acc, lpa = orc.pmala_step(noise) if kind == 1 else orc.mtm_step(noise)
        assert np.array_equal(acc > 0, g[f"t{t}_accept"] > 0), f"{name}: decisions differ at t={t}"
        n_rej += int((acc == 0).sum())
        assert rel(orc.current["Theta"], g[f"t{t}_Theta"]) < 1e-9
        assert np.allclose(orc.v, g[f"t{t}_v"], rtol=2e-6)                 # stored as float32
        assert np.array_equal(orc.j_t, g[f"t{t}_j"])
    assert n_rej > 0.02 * len(g["kernels"]) * post.N
