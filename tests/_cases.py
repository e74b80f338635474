"""Shared helpers: rebuild problems from the golden fixtures / synthetic recipe (no reference needed)."""
import os

import numpy as np

from beetroots_b200 import modelling as M
from beetroots_b200.network import synthetic_weights

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["ref_n4_6x5", "ref_n8_5x6_censored"]
# BASELINE-sized tapes of the unmodified reference (make_golden.py --big): c2 = 30x30, K=50, 6 iterations; c3-like = 64x64,
# 29 % censored, K=10, step size 1e-2 (16-18 % of the moves rejected), 5 iterations.  Compact storage: float32 noise (exactly
# what the reference consumed), float32 v / grad.
GOLDEN_BIG = ["ref_c2_30x30", "ref_c3_64x64_censored"]


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def problem_from_golden(g):
    L, D = int(g["L"]), 4
    X, Y = g["X"], g["Y"]
    N = X.size
    net = synthetic_weights(seed=int(g["seed"]), n_out=L)
    fixed = {"kappa": None, "P": None, "radm": None, "Avmax": None, "angle": float(g["angle_fixed"])}
    fm = M.NeuralNetworkApprox(net, fixed)
    lik = M.MixingModelsLikelihood(fm, D, L, N, g["y"], float(g["sigma_a"]), float(g["sigma_m"]), float(g["omega"]),
                                   a0=g["a0"][None, :], a1=g["a1"][None, :])
    sp = M.L22DiscreteGradSpatialPrior(D, N, X, Y, None, bool(g["nnn"]), g["tau"])
    ind = M.SmoothIndicatorPrior(D, N, float(g["delta"]), g["lo"], g["hi"])
    post = M.Posterior(D, L, N, lik, sp, ind)
    nbr = M.build_neighbor_table(sp.list_edges, N, 8 if bool(g["nnn"]) else 4)
    return post, nbr


def golden_noise(g, t, kind, post, full=False):
    """Noise of iteration t keyed by site; ``full`` scatters it to (N, ...) arrays for the C ABI."""
    N, D, K = post.N, post.D, int(g["K"])
    out = {}
    for s, idx in post.dict_sites.items():
        if kind == 1:
            z, u = g[f"t{t}_s{s}_z"], g[f"t{t}_s{s}_u"]
            if full:
                zf, uf = np.zeros((N, D)), np.full(N, 0.5)
                zf[idx], uf[idx] = z, u
                z, u = zf, uf
            out[s] = {"z": z, "u": u}
        else:
            c, us, ua = g[f"t{t}_s{s}_cand"], g[f"t{t}_s{s}_usel"], g[f"t{t}_s{s}_uacc"]
            if full:
                cf, usf, uaf = np.zeros((N, K, D)), np.full(N, 0.5), np.full(N, 0.5)
                cf[idx], usf[idx], uaf[idx] = c, us, ua
                c, us, ua = cf, usf, uaf
            out[s] = {"cand": c, "u_sel": us, "u_acc": ua}
    return out


def random_noise(post, K, kind, rng, full=True):
    N, D = post.N, post.D
    out = {}
    for s, idx in post.dict_sites.items():
        if kind == 1:
            z, u = rng.standard_normal((N, D)), rng.uniform(size=N)
            out[s] = {"z": z if full else z[idx], "u": u if full else u[idx]}
        else:
            raise NotImplementedError
    return out


def slice_noise(noise, post):
    """(N, ...) arrays -> per-site arrays for the oracle."""
    out = {}
    for s, idx in post.dict_sites.items():
        out[s] = {k: np.asarray(v, dtype=np.float64)[idx] for k, v in noise[s].items()}
    return out


# ---- deterministic inputs for the saver golden (tests/golden/make_golden_saver.py and tests/test_host.py) ----
class Pow10Scaler:
    def from_scaled_to_lin(self, x):
        return 10.0 ** x


def saver_inputs(t, N=5, D_s=2, L=2):
    rng = np.random.default_rng(100 + t)
    return dict(Theta=rng.standard_normal((N, D_s)),
                forward_map_evals={"log_f_Theta": rng.standard_normal((N, L)), "grad_log_f_Theta": np.zeros((N, D_s, L))},
                nll_utils={"nll": np.float64(rng.uniform()), "nll_au": np.zeros((N, L)), "m_a": rng.standard_normal((N, L))},
                dict_objective={"nll": np.float64(t), "nl_prior_spatial": rng.uniform(size=D_s), "objective": np.float64(2.0 * t)},
                additional_sampling_log={"v": rng.uniform(size=(N, D_s)), "type_t": t % 2, "accepted_t": 0.25 * t,
                                         "log_proba_accept_t": -1.0 * t},
                rng=(np.arange(32, dtype=np.uint8) + t, np.arange(32, dtype=np.uint8)[::-1] + t))
