"""Run the UNMODIFIED reference (``/root/reference`` or its travelling copy ``baseline/_ref``) on the synthetic cubes of
``beetroots_b200.synthetic`` -- TEST / BENCHMARK INFRASTRUCTURE, never on the product path.

* ``reference_objects(prob)``: the reference's own ``NeuralNetworkApprox`` (on the stand-in nnbma of ``oracle/refshim``),
  ``MixingModelsLikelihood``, ``L22DiscreteGradSpatialPrior``, ``SmoothIndicatorPrior`` and ``Posterior`` built from the
  inputs of a ``SyntheticProblem`` (the recipe of tests/golden/make_golden.py::build_case);
* ``memory_saver(...)``: the reference's ``MySaver`` with its two h5py-writing methods replaced by recorders (h5py is not
  installed) -- everything else (``initialize_memory``, ``update_memory``, batching) is reference code;
* ``time_reference_kernels(...)``: wall time of ``MySampler.generate_new_sample_mtm`` / ``_pmala_rmsprop``
  (sampler/my_sampler.py:966-1202, 718-906) on the host cores: the CPU arm of ``bench.py``.
"""
import os
import tempfile
import time

import numpy as np

from . import refshim


def set_host_threads(n=None) -> int:
    """Pin every thread pool the reference uses (torch intra-op, numba, BLAS via threadpoolctl) to ``n`` threads
    (default: all host cores) -- torchrun exports OMP_NUM_THREADS=1, which would otherwise cripple the CPU arm."""
    n = int(n or os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ["NUMBA_NUM_THREADS"] = str(n)
    try:
        import torch
        torch.set_num_threads(n)
    except Exception:
        pass
    try:
        import numba
        numba.set_num_threads(min(n, numba.config.NUMBA_NUM_THREADS))
    except Exception:
        pass
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(n)
    except Exception:
        pass
    return n


def reference_objects(prob, names=None):
    """(ref_posterior, ref_forward_map) for a ``beetroots_b200.synthetic.SyntheticProblem``."""
    refshim.install()
    import pandas as pd
    from nnbma import networks as fake_nnbma
    from beetroots.modelling.forward_maps.neural_network_approx import NeuralNetworkApprox as RefNN
    from beetroots.modelling.likelihoods.approx_censored_add_mult import MixingModelsLikelihood as RefLik
    from beetroots.modelling.posterior import Posterior as RefPosterior
    from beetroots.modelling.priors.l22_discrete_grad_prior import L22DiscreteGradSpatialPrior as RefL22
    from beetroots.modelling.priors.smooth_indicator_prior import SmoothIndicatorPrior as RefInd
    from beetroots.modelling.priors.spatial_prior_params import SpatialPriorParams

    post = prob.posterior
    lik, fm = post.likelihood, post.likelihood.forward_map
    N, D, L = post.N, post.D, post.L
    names = [f"line{i}" for i in range(L)] if names is None else names
    model_name = f"synthetic_{id(fm.network)}_{L}"
    fake_nnbma.REGISTRY[model_name] = (fm.network, names)
    ref_fm = RefNN("/nonexistent", model_name, dict(fm.dict_fixed_values_scaled))
    ref_fm.restrict_to_output_subset(names)
    assert prob.a0 is not None and prob.a1 is not None
    const = lambda a: float(a.flat[0]) if np.all(a == a.flat[0]) else a       # noqa: E731
    with tempfile.TemporaryDirectory() as tmp:
        csv = os.path.join(tmp, "best_params.csv")
        pd.DataFrame({"n": 0, "ell": np.arange(L), "line": names, "a0_best": prob.a0, "a1_best": prob.a1,
                      "target_best": 0.0}).to_csv(csv, index=False)
        ref_lik = RefLik(ref_fm, D, L, N, lik.y, const(lik.sigma_a), const(lik.sigma_m), const(lik.omega), csv, names)
    ref_sp = None
    if post.prior_spatial is not None:
        ps = post.prior_spatial
        df = pd.DataFrame({"idx": np.arange(N), "X": ps.X, "Y": ps.Y}).set_index(["X", "Y"])
        ref_sp = RefL22(SpatialPriorParams("L2-gradient", ps.use_next_nearest_neighbors, ps.initial_weights * 1), "syn", D, N,
                        df, list(range(D)))
    pi = post.prior_indicator
    ref_ind = RefInd(D, N, pi.indicator_margin_scale, pi.lower_bounds, pi.upper_bounds, list(range(D)))
    return RefPosterior(D, L, N, ref_lik, ref_sp, ref_ind), ref_fm


class _IdentityScaler:
    def from_scaled_to_lin(self, x):
        return x

    def from_lin_to_scaled(self, x):
        return x


def memory_saver(N, D, D_sampling, L, batch_size=10, freq_save=1, scaler=None, save_forward_map_evals=True):
    """Reference ``MySaver`` (sampler/saver/my_saver.py:9-144) recording its batches in ``.batches`` / ``.additional``
    instead of writing HDF5."""
    refshim.install()
    from beetroots.sampler.saver.my_saver import MySaver

    class MemorySaver(MySaver):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.batches, self.additional = [], {}

        def save_to_file(self):
            self.batches.append({k: np.array(v).copy() for k, v in self.memory.items()})
            self.memory = dict()

        def save_additional(self, list_arrays, list_names):
            for n, a in zip(list_names, list_arrays):
                self.additional[n] = np.array(a).copy()

    return MemorySaver(N, D, D_sampling, L, scaler or _IdentityScaler(), results_path="", batch_size=batch_size,
                       freq_save=freq_save, save_forward_map_evals=save_forward_map_evals, list_idx_sampling=list(range(D)))


def reference_sampler(prob, rng=None):
    refshim.install()
    from beetroots.sampler.my_sampler import MySampler
    from beetroots.sampler.utils.my_sampler_params import MySamplerParams
    kw = prob.sampler_kwargs
    params = MySamplerParams(kw["initial_step_size"], kw["extreme_grad"], kw["history_weight"], list(kw["selection_probas"]),
                             int(kw["k_mtm"]), True, False)
    post = prob.posterior
    return MySampler(params, post.D, post.L, post.N, rng=np.random.default_rng(42) if rng is None else rng)


def time_reference_kernels(prob, pairs=1, warmup=1):
    """Seconds per (one MTM + one PMALA) iteration pair of the unmodified reference kernels on ``prob``'s cube."""
    ref_post, _ = reference_objects(prob)
    smp = reference_sampler(prob)
    N, D = ref_post.N, ref_post.D
    smp.current = ref_post.compute_all(prob.theta_init * 1, compute_derivatives_2nd_order=False)   # my_sampler.py:319-354
    smp.v = smp.current["grad"].flatten() ** 2
    smp.j_t = np.zeros((N * D,))
    t = 1
    for _ in range(warmup):                                    # numba JIT, torch warm-up
        smp.generate_new_sample_mtm(t, ref_post)
        smp.generate_new_sample_pmala_rmsprop(t + 1, ref_post)
        t += 2
    t0 = time.perf_counter()
    for _ in range(pairs):
        smp.generate_new_sample_mtm(t, ref_post)
        smp.generate_new_sample_pmala_rmsprop(t + 1, ref_post)
        t += 2
    return (time.perf_counter() - t0) / pairs
