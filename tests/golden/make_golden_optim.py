"""Golden fixtures for the optimisation mode (is_stochastic=False), from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_optim.py

Same two problems as make_golden.py (inputs are read from tests/golden/<case>.npz by the tests); the reference kernels
`generate_new_sample_pmala_rmsprop` / `generate_new_sample_mtm` run with `MySamplerParams(..., is_stochastic=False)`
(sampler/my_sampler.py:908-964, 1050-1086) with a recording RNG; the tape is replayed through the oracle
(`OracleSampler.pmala_optim_step / mtm_optim_step`), asserted equal, and stored as <case>_optim.npz.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import make_golden as G  # noqa: E402  (installs the shims, imports the reference)


def run_case(name, kernels, **case):
    ref_post, post, inputs = G.build_case(**case)
    N, D, L, K = ref_post.N, ref_post.D, ref_post.L, case["K"]
    params = G.MySamplerParams(1e-4, 1e-5, 0.99, [0.5, 0.5], K, False, False)
    smp = G.MySampler(params, D, L, N, rng=np.random.default_rng(43))
    tape = G.TapeRNG(smp.rng)
    smp.rng = tape
    props = []
    orig = smp.generate_random_start_Theta_1pix

    def rec(Theta, posterior, idx_pix):
        c = orig(Theta, posterior, idx_pix)
        props.append(c.copy())
        return c

    smp.generate_random_start_Theta_1pix = rec
    theta0 = inputs["theta0"]
    smp.current = ref_post.compute_all(theta0 * 1, compute_derivatives_2nd_order=False)
    smp.v = smp.current["grad"].flatten() ** 2
    smp.j_t = np.zeros((N * D,))
    nbr = G.M.build_neighbor_table(post.prior_spatial.list_edges, N)
    orc = G.OracleSampler(post, nbr, 1e-4, 1e-5, 0.99, K)
    orc.init(theta0)

    def cmp(a, b, what, tol=1e-9):
        err = np.max(np.abs(a - b) / (1e-300 + np.maximum(np.abs(a), np.abs(b)) + 1e-12))
        assert err < tol, f"{name}: oracle != reference for {what}: {err}"

    out = {"kernels": np.array(kernels)}
    sites = list(ref_post.dict_sites.keys())
    for t, kind in enumerate(kernels, start=1):
        tape.log.clear()
        props.clear()
        if kind == 1:
            acc, proba = smp.generate_new_sample_pmala_rmsprop(t, ref_post)
            z = [e[1] for e in tape.log if e[0] == "z"][0].reshape(N, D)
            # the reference multiplies the draw in place (z_t *= ..., :922-923): the tape holds a copy taken before
            o_acc, _ = orc.pmala_optim_step(z)
            assert bool(acc) == bool(o_acc), f"{name}: PMALA-optim decision differs at t={t}"
            out[f"t{t}_z"] = z
            out[f"t{t}_accept"] = np.array(float(acc))
        else:
            acc_mean, _ = smp.generate_new_sample_mtm(t, ref_post)
            noise = {s: {"cand": props[i]} for i, s in enumerate(sites)}
            o_acc = orc.mtm_optim_step(noise)
            assert abs(o_acc.mean() - acc_mean) < 1e-12, f"{name}: MTM-optim accept rate differs at t={t}"
            for i, s in enumerate(sites):
                out[f"t{t}_s{s}_cand"] = props[i]
            out[f"t{t}_accept"] = o_acc
        cmp(smp.current["Theta"], orc.current["Theta"], f"Theta t={t}")
        cmp(smp.v.reshape(N, D), orc.v, f"v t={t}", 1e-8)
        out[f"t{t}_Theta"] = smp.current["Theta"].copy()
        out[f"t{t}_v"] = smp.v.reshape(N, D).copy()
        out[f"t{t}_objective"] = np.array(float(smp.current["objective_global"]))
        print(f"[{name}] t={t} kind={'PMALA' if kind else 'MTM'}-optim accept={np.mean(out[f't{t}_accept']):.3f} "
              f"objective={float(smp.current['objective_global']):.6f} OK (oracle == reference)")
    np.savez_compressed(os.path.join(HERE, f"{name}_optim.npz"), **out)


if __name__ == "__main__":
    run_case("ref_n4_6x5", (1, 0, 1, 1, 0, 1), H=6, W=5, L=10, K=7, nnn=False, seed=0, censor_boost=1.0)
    run_case("ref_n8_5x6_censored", (0, 1, 0, 1, 1), H=5, W=6, L=6, K=5, nnn=True, seed=3, censor_boost=6.0, drop=(7, 29))
