"""CPU: host-side logic and the C-ABI library surface (no compute calls without a GPU)."""
import copy
import ctypes
import os
import re

import numpy as np
import pytest

from beetroots_b200 import modelling as M
from beetroots_b200.network import dense_widths, poly_exponents, synthetic_weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol(built_lib):
    from beetroots_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "beetroots_b200.h")).read()
    declared = set(re.findall(r"\b(bt_[a-z0-9_]+)\s*\(", header))
    declared -= {"bt_ctx"}
    assert declared, "no declarations found"
    raw = ctypes.CDLL(built_lib)
    for name in sorted(declared):
        assert hasattr(raw, name), f"{name} declared in include/beetroots_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.bt_version() >= 100


def test_missing_device_fails_loudly(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from beetroots_b200 import _lib
    lib = _lib.load()
    ctx = ctypes.c_void_p()
    rc = lib.bt_ctx_create(ctypes.byref(ctx), 0, 4, 4, 10, 4)
    assert rc != 0 and len(lib.bt_last_error()) > 0


def test_poly_exponents_and_widths():
    e = poly_exponents(4, 3)
    assert e.shape == (34, 3)                      # subnetwork/input_features.json = 34
    assert list(e[0]) == [0, -1, -1] and list(e[4]) == [0, 0, -1] and list(e[33]) == [3, 3, 3]
    w = dense_widths(34, 10, 0.5)
    assert len(w) == 9 and 34 + sum(w) == 1296
    net = synthetic_weights(0, 10)
    assert net.macs_per_row() == 684366            # F1 = 1.369 MFLOP
    assert net.restrict_to_output_subset([0, 3]).n_out == 2


def test_edges_sites_neighbours_match_reference_conventions():
    H, W = 3, 4
    X, Y = np.repeat(np.arange(H), W), np.tile(np.arange(W), H)
    sp = M.L22DiscreteGradSpatialPrior(2, H * W, X, Y, None, False, np.array([1.0, 2.0]))
    assert sp.list_edges.shape == ((H - 1) * W + H * (W - 1), 2)
    assert set(sp.dict_sites) == {0, 1}
    assert np.array_equal(np.sort(np.concatenate(list(sp.dict_sites.values()))), np.arange(H * W))
    nbr = M.build_neighbor_table(sp.list_edges, H * W, 4)
    col = (X + Y) % 2
    for n in range(H * W):
        for m in nbr[n][nbr[n] >= 0]:
            assert col[m] != col[n]                # checkerboard: neighbours have the other colour
    sp8 = M.L22DiscreteGradSpatialPrior(2, H * W, X, Y, None, True, np.array([1.0, 2.0]))
    assert set(sp8.dict_sites) == {0, 1, 2, 3}
    nbr8 = M.build_neighbor_table(sp8.list_edges, H * W, 8)
    assert (nbr8[1 * W + 1] >= 0).sum() == 8
    # neighbour order = order of appearance in the edge list (get_neighboring_pixels, utils.py:100-125)
    n = 1 * W + 1
    e = sp8.list_edges
    sel = e[(e[:, 0] == n) | (e[:, 1] == n)].flatten()
    assert np.array_equal(sel[sel != n], nbr8[n])


def test_sampler_params_asserts_and_deepcopy():
    from beetroots_b200 import B200Sampler, MySamplerParams
    with pytest.raises(AssertionError):
        MySamplerParams(-1.0, 1e-5, 0.99, [0.5, 0.5], 10)
    with pytest.raises(AssertionError):
        MySamplerParams(1e-4, 1e-5, 0.99, [0.5, 0.6], 10)
    p = MySamplerParams(1e-4, 1e-5, 0.99, [0.5, 0.5], 10, True, False)
    s = B200Sampler(p, 4, 10, 12)
    s2 = copy.deepcopy(s)                          # RunMCMC deep-copies the sampler per chain (run_mcmc.py:164-167)
    assert s2.k_mtm == 10 and s2._dev is None
    st, inc = s.get_rng_state()
    s.rng.standard_normal(3)
    s.set_rng_state(bytes(st), bytes(inc))
    a = s.rng.standard_normal(3)
    s.set_rng_state(bytes(st), bytes(inc))
    assert np.array_equal(a, s.rng.standard_normal(3))
    with pytest.raises(NotImplementedError):          # the Hessian-diagonal correction term (posterior.py:259) is not built
        B200Sampler(MySamplerParams(1e-4, 1e-5, 0.99, [0.5, 0.5], 10, True, True), 4, 10, 12)
    opt = B200Sampler(MySamplerParams(1e-4, 1e-5, 0.99, [0.5, 0.5], 10, False, False), 4, 10, 12)
    assert opt.stochastic is False and copy.deepcopy(opt).stochastic is False


def test_likelihood_mirror_shape_error():
    net = synthetic_weights(0, 3)
    fm = M.NeuralNetworkApprox(net, {"kappa": None, "P": None, "radm": None, "Avmax": None, "angle": 0.1})
    assert fm.D_sampling == 4 and fm.list_indices_to_sample_for_nn == [0, 1, 2]
    with pytest.raises(ValueError):
        M.MixingModelsLikelihood(fm, 4, 3, 5, np.ones((4, 3)), 1.0, 0.1, 1.0, a0=np.zeros((1, 3)), a1=np.ones((1, 3)))


# ---- control flow of B200Sampler.sample on a recording stand-in for the device (no CUDA) -------------------------------
class _RecordingDevice:
    """Speaks the DevicePosterior surface B200Sampler uses and records the calls; numbers are arbitrary but deterministic."""

    def __init__(self, N, D, L, objectives):
        self.N, self.D, self.L = N, D, L
        self.n_sites, self.site_keys = 2, [0, 1]
        self.calls, self.objectives, self.n_obj = [], list(objectives), 0
        self.theta = np.zeros((N, D))
        self.chain = []

    def _rec(self, name, *a):
        self.calls.append((name,) + a)

    def compute_all(self, Theta=None):
        if Theta is not None:
            self.theta = np.array(Theta, dtype=float)
        return self.get_current()

    def get_current(self, with_forward_map=True):
        N, D, L = self.N, self.D, self.L
        return {"Theta": self.theta.copy(), "forward_map_evals": {"log_f_Theta": np.zeros((N, L)), "f_Theta": np.ones((N, L))},
                "nll_utils": {"nll": np.float64(1.0)}, "nll_pix": np.ones(N), "grad_lik": np.ones((N, D)),
                "objective_pix_chromatic": np.ones(N), "objective_pix_global": np.ones(N), "objective_chromatic": float(N),
                "objective_global": float(N), "grad": np.ones((N, D))}

    def get_state(self, want_v=True, want_j=True):
        return self.theta.copy(), np.ones((self.N, self.D)), np.zeros((self.N, self.D), dtype=np.int32)

    def get_log_f(self): return np.zeros((self.N, self.L))
    def init_v_from_grad(self): self._rec("init_v")
    def model_check_reset(self): self._rec("mc_reset")
    def begin_iteration(self): self._rec("begin")
    def pmala_site(self, s, *a): self._rec("pmala_site", s); self.theta += 0.25
    def mtm_site(self, s, *a): self._rec("mtm_site", s); self.theta += 0.5
    def mtm_optim_site(self, s, *a): self._rec("mtm_optim_site", s); self.theta += 0.5
    def optim_pmala_step(self, *a): self._rec("optim_pmala"); self.theta += 0.25; return True, 1
    def mtm_finish(self, alpha): self._rec("mtm_finish")
    def reduce_step_stats(self): return np.array([self.N / 2.0, -3.0 * self.N])
    def model_check_update(self, seed, t, *a): self._rec("mc_update", t)
    def model_check_snapshot_clppd(self): self._rec("mc_snapshot", self.n_obj)
    def model_check_update_pvalues(self, seed, t, *a): self._rec("mc_pvalues", t)
    def model_check_get(self, finalize=True): return np.zeros((self.N, self.L)), np.zeros((self.N, self.L)), np.zeros(self.N)
    def spatial_prior_unweighted(self): self._rec("potential"); return np.full(self.D, 2.0 + self.theta[0, 0])
    def update_spatial_weights(self, ps, pi): self._rec("weights", tuple(np.asarray(ps.weights)))
    def chain_reserve(self, cap): self._rec("chain_reserve", cap)
    def chain_length(self): return len(self.chain)
    def chain_push(self): self.chain.append(self.theta.copy())

    def objective_terms(self):
        v = self.objectives[min(self.n_obj, len(self.objectives) - 1)]
        self.n_obj += 1
        return {"nll": np.float64(v), "nl_prior_spatial": np.zeros(self.D), "nl_prior_indicator": np.zeros(self.D),
                "objective": np.float64(v)}


class _ListSaver:
    def __init__(self):
        self.memory, self.logs, self.additional = {}, [], None

    def initialize_memory(self, T_MC, t, **kw): self.memory = {"on": True}
    def update_memory(self, t, additional_sampling_log=None, **kw): self.logs.append((t, dict(additional_sampling_log)))
    def check_need_to_update_memory(self, t): return True
    def check_need_to_save(self, t): return False
    def save_to_file(self): pass
    def save_additional(self, list_arrays, list_names): self.additional = list_names


def _host_problem(N=6, D=4, L=3):
    X, Y = np.repeat(np.arange(N // 2), 2), np.tile(np.arange(2), N // 2)
    sp = M.L22DiscreteGradSpatialPrior(D, N, X, Y, None, False, np.array([10.0, 1.0, 2.0, 4.0]))
    ind = M.SmoothIndicatorPrior(D, N, 0.1, -np.ones(D), np.ones(D))
    lik = M.MixingModelsLikelihood(None, D, L, N, np.ones((N, L)), 1.0, 0.1, 0.5, a0=np.zeros((1, L)), a1=np.ones((1, L)))
    return M.Posterior(D, L, N, lik, sp, ind)


def test_sample_control_flow_regu_weights_and_device_chain():
    """MySampler.sample, sampling mode (my_sampler.py:403-541): weight update before the kernel from t = N0 on, with the
    reference recursion; tau logged every iteration; the device chain takes every thin-th iterate after burn-in until full."""
    from beetroots_b200 import B200Sampler, MySamplerParams
    from beetroots_b200.mml import EBayesMMLELogRate
    post = _host_problem()
    dev = _RecordingDevice(post.N, post.D, post.L, [5.0])
    tau0 = post.prior_spatial.weights.copy()
    smp = B200Sampler(MySamplerParams(1e-4, 1e-5, 0.99, [0.5, 0.5], 3, True, False), post.D, post.L, post.N,
                      rng=np.random.default_rng(0), device_chain=3, device_chain_thin=2)
    smp.attach(post, dev)
    sv = _ListSaver()
    smp.sample(post, sv, 10, Theta_0=np.zeros((post.N, post.D)), disable_progress_bar=True, regu_spatial_N0=4,
               regu_spatial_scale=1.0, T_BI=3)
    opt = EBayesMMLELogRate(scale=1.0, N0=4, N1=+np.inf, dim=post.D * post.N, vmin=1e-8, vmax=1e8, homogeneity=2.0, exponent=0.8)
    tau, theta00 = float(tau0.mean()), 0.0
    kinds = [c[0] for c in dev.calls]
    for t, log in sv.logs:
        if t >= 4:
            tau = opt.update(tau, t, post.D * (2.0 + theta00))
            np.testing.assert_allclose(log["tau"], tau * np.ones(post.D), rtol=1e-14)
        else:
            np.testing.assert_array_equal(log["tau"], tau0)
        theta00 += 1.0 if log["type_t"] == 0 else 0.5                     # two sites per iteration in the stand-in
        assert log["accepted_t"] == 0.5 and log["log_proba_accept_t"] == -3.0
    assert kinds.count("potential") == 7 and kinds.count("weights") == 7 and [t for t, _ in sv.logs] == list(range(1, 11))
    np.testing.assert_allclose(post.prior_spatial.weights, tau * np.ones(post.D), rtol=1e-14)
    # every "potential" comes right before its "weights" and before the iteration's kernel
    for i, k in enumerate(kinds):
        if k == "potential":
            assert kinds[i + 1] == "weights" and kinds[i + 2] == "begin"
    assert ("chain_reserve", 3) in dev.calls and len(dev.chain) == 3         # t = 4, 6, 8 (t > T_BI, even), then full
    assert [c[1] for c in dev.calls if c[0] == "mc_update"] == list(range(4, 11))   # model check only after burn-in
    assert sv.additional == ["clppd", "p-values-y", "p-values-llh", "philox_seed"]


def test_sample_control_flow_optimisation_mode():
    """is_stochastic=False (my_sampler.py:208-212, 222-256, 908-964, 1050-1086, 1199-1200): optimisation kernels, clppd
    snapshot only when the saved objective improves (after burn-in), ESS_OPTIM p-value replications at the end."""
    from beetroots_b200 import B200Sampler, MySamplerParams
    post = _host_problem()
    objs = [9.0, 8.0, 8.5, 7.0, 7.0, 6.0]
    dev = _RecordingDevice(post.N, post.D, post.L, objs)
    smp = B200Sampler(MySamplerParams(1e-4, 1e-5, 0.99, [0.5, 0.5], 3, False, False), post.D, post.L, post.N,
                      rng=np.random.default_rng(1))
    smp.ESS_OPTIM = 5
    smp.attach(post, dev)
    sv = _ListSaver()
    smp.sample(post, sv, 6, Theta_0=np.zeros((post.N, post.D)), disable_progress_bar=True, T_BI=1)
    kinds = [c[0] for c in dev.calls]
    assert "mtm_site" not in kinds and "pmala_site" not in kinds and "mc_update" not in kinds
    n_mtm = sum(1 for _, log in sv.logs if log["type_t"] == 0)
    assert kinds.count("mtm_optim_site") == 2 * n_mtm and kinds.count("mtm_finish") == n_mtm
    assert kinds.count("optim_pmala") == 6 - n_mtm
    for _, log in sv.logs:
        if log["type_t"] == 0:
            assert log["accepted_t"] == log["log_proba_accept_t"] == 0.5   # (mean accept, mean accept), :1199-1200
        else:
            assert log["accepted_t"] is True and log["log_proba_accept_t"] == 1
    # saved objectives 9, 8, 8.5, 7, 7, 6 at t = 1..6; burn-in t = 1; improvements at t = 2 (8 < inf), 4 (7), 6 (6)
    assert [c[1] for c in dev.calls if c[0] == "mc_snapshot"] == [2, 4, 6]
    assert [c[1] for c in dev.calls if c[0] == "mc_pvalues"] == [7, 8, 9, 10, 11]


def test_chain_saver_matches_reference_saver_golden(tmp_path):
    """beetroots_b200.saver.ChainSaver against the memory batches of the UNMODIFIED reference MySaver
    (tests/golden/make_golden_saver.py): same dataset names, shapes, slots and values; npz back end round-trips."""
    from _cases import Pow10Scaler as Pow10, saver_inputs as inputs
    from beetroots_b200.saver import ChainSaver
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "saver_golden.npz"))
    sv = ChainSaver(5, 3, 2, 2, Pow10(), results_path=str(tmp_path / "out"), batch_size=3, freq_save=2,
                    save_forward_map_evals=True, list_idx_sampling=[0, 2])
    sv._h5py = lambda: None                                               # force the npz back end whatever is installed
    batches, T = [], 12
    for t in range(1, T + 1):
        x = inputs(t)
        kw = {k: x[k] for k in ("Theta", "forward_map_evals", "nll_utils", "dict_objective", "additional_sampling_log")}
        if sv.memory == {}:
            sv.initialize_memory(T, t, **kw)
            sv.update_memory(t, rng_state_array=x["rng"][0], rng_inc_array=x["rng"][1], **kw)
        elif sv.check_need_to_update_memory(t):
            sv.update_memory(t, rng_state_array=x["rng"][0], rng_inc_array=x["rng"][1], **kw)
        if sv.check_need_to_save(t):
            batches.append({k: v.copy() for k, v in sv.memory.items()})
            sv.save_to_file()
    assert len(batches) == int(g["n_batches"]) == 2
    for b, mem in enumerate(batches):
        ref_keys = {k.split("__", 1)[1] for k in g.files if k.startswith(f"b{b}__")}
        assert set(mem) == ref_keys
        for k, v in mem.items():
            r = g[f"b{b}__{k}"]
            assert v.shape == r.shape and v.dtype == r.dtype, k
            np.testing.assert_array_equal(v, r, err_msg=k)
    sv.save_additional([np.ones((5, 2)), np.zeros(5)], ["clppd", "p-values-llh"])
    z = np.load(tmp_path / "out" / "mc_chains.npz")
    assert z["list_Theta"].shape == (6, 5, 2) and z["clppd"].shape == (5, 2)
    np.testing.assert_array_equal(z["list_Theta"], np.concatenate([g["b0__list_Theta"], g["b1__list_Theta"]], 0))
