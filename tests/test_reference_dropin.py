"""CPU, build container only (skipped where /root/reference is absent): the host side of the drop-in boundary run on
UNMODIFIED reference objects -- ``host_inputs_from_posterior`` must extract from a reference ``Posterior`` exactly what
it extracts from the mirrors (same arrays that the C ABI uploads)."""
import os
import sys

import numpy as np
import pytest

from oracle import refshim

pytestmark = pytest.mark.skipif(not refshim.reference_available(), reason="reference tree not present on this box")


def test_reference_posterior_yields_the_same_device_inputs(built_lib):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden as MG                     # builds reference + mirror objects from the same inputs
    from beetroots_b200.device import host_inputs_from_posterior
    ref_post, post, _ = MG.build_case(H=5, W=4, L=6, K=3, nnn=True, seed=2, censor_boost=2.0, drop=(3,))
    a = host_inputs_from_posterior(ref_post)     # unmodified reference Posterior / likelihood / priors / NeuralNetworkApprox
    b = host_inputs_from_posterior(post)         # mirrors
    for k in ("N", "D", "L", "max_degree", "delta", "site_keys"):
        assert a[k] == b[k], k
    for k in ("nbr", "fixed", "sampled_idx", "tau", "tau_init", "lower", "upper"):
        assert np.array_equal(a[k], b[k]), k
    for k in a["obs"]:
        assert np.array_equal(a["obs"][k], b["obs"][k]), k
    for x, y in zip(a["site_pixels"], b["site_pixels"]):
        assert np.array_equal(x, y)
    na, nb = a["net"], b["net"]
    assert np.array_equal(na.exponents, nb.exponents)
    assert np.allclose(na.poly_means, nb.poly_means) and np.allclose(na.poly_stds, nb.poly_stds)
    for wa, wb in zip(na.weights + [na.w_out], nb.weights + [nb.w_out]):
        assert np.array_equal(wa, wb)
    for ba, bb in zip(na.biases + [na.b_out], nb.biases + [nb.b_out]):
        assert np.array_equal(ba, bb)


def test_b200_sampler_has_the_reference_signature():
    import inspect
    refshim.install()
    from beetroots.sampler.my_sampler import MySampler
    from beetroots_b200 import B200Sampler
    ref = inspect.signature(MySampler.sample)
    ours = inspect.signature(B200Sampler.sample)
    assert list(ref.parameters) == list(ours.parameters)
    for name, p in ref.parameters.items():
        if p.default is not inspect._empty and name != "regu_spatial_N0":
            assert ours.parameters[name].default == p.default, name
    ref_init = list(inspect.signature(MySampler.__init__).parameters)
    assert list(inspect.signature(B200Sampler.__init__).parameters)[:len(ref_init)] == ref_init
    for m in ("generate_new_sample_mtm", "generate_new_sample_pmala_rmsprop", "get_rng_state", "set_rng_state",
              "generate_random_start_Theta", "generate_random_start_Theta_1pix"):
        assert hasattr(B200Sampler, m)


def test_saver_auxiliaries_equal_the_reference_nll_utils(built_lib):
    """beetroots_b200.modelling.saver_nll_utils (host glue for the Saver hand-over) against the unmodified reference's
    ``MixingModelsLikelihood.evaluate_all_nll_utils`` at the same Theta: every (N, L) entry the reference saver stores."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden as MG
    from beetroots_b200.modelling import saver_nll_utils
    ref_post, post, inputs = MG.build_case(H=6, W=5, L=7, K=3, nnn=False, seed=4, censor_boost=2.5)
    cur = ref_post.compute_all(inputs["theta0"] * 1, compute_derivatives_2nd_order=False)
    log_f = cur["forward_map_evals"]["log_f_Theta"]
    stored = {k for k in cur["nll_utils"] if "nll_" not in k and "grad" not in k and "hess_diag" not in k
              and np.ndim(cur["nll_utils"][k]) == 2}
    for lik in (ref_post.likelihood, post.likelihood):              # the reference object and its mirror
        aux = saver_nll_utils(lik, log_f)
        assert set(aux) == stored, (set(aux) ^ stored)
        for k in stored:
            np.testing.assert_allclose(aux[k], cur["nll_utils"][k], rtol=1e-12, atol=1e-300, err_msg=k)
    assert 0.05 < aux["censored_mask"].mean() < 0.95 and 0 < aux["lambda_"].min() < aux["lambda_"].max() <= 1
