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
