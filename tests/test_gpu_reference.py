"""GPU drop-in test against the UNMODIFIED reference (baseline/_ref, shipped by oracle/install_reference.py; skipped when
absent): the reference's own ``Posterior`` (NeuralNetworkApprox on the stand-in nnbma, MixingModelsLikelihood,
L22DiscreteGradSpatialPrior, SmoothIndicatorPrior) and its own ``MySaver`` drive ``B200Sampler.sample`` -- and the chain is
compared with the one ``MySampler.sample`` (sampler/my_sampler.py:269-559) produces on the host for the same problem."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _batch_means_se(x, n_batches=8):
    """Monte-Carlo standard error of the mean of a chain (T, ...) by batch means (the chains are autocorrelated)."""
    T = x.shape[0] // n_batches * n_batches
    b = x[:T].reshape((n_batches, T // n_batches) + x.shape[1:]).mean(axis=1)
    return b.std(axis=0, ddof=1) / np.sqrt(n_batches)


H = W = 8
K, T, T_BI = 10, 160, 40          # the reference needs ~1 s per iteration of sample() at 8 x 8 (saver + model check included)
_REF = {}


def _reference_run():
    """One ``MySampler.sample`` run of the unmodified reference on the host, shared by the two parametrisations."""
    if not _REF:
        from oracle import refrun
        from oracle import posterior as OP
        from beetroots_b200.synthetic import make_problem
        prob = make_problem(H, W, lambda fm, th: OP.forward_map_compute_all(fm, th, False)["log_f_Theta"], L=10, k_mtm=K)
        ref_post, _ = refrun.reference_objects(prob)
        refrun.set_host_threads()
        ref_smp = refrun.reference_sampler(prob, rng=np.random.default_rng(7))
        ref_saver = refrun.memory_saver(ref_post.N, ref_post.D, ref_post.D, ref_post.L, batch_size=T, freq_save=1)
        ref_smp.sample(ref_post, ref_saver, T, Theta_0=prob.theta_init * 1, disable_progress_bar=True, T_BI=T_BI)
        _REF.update(prob=prob, ref_post=ref_post, ref_saver=ref_saver)
    return _REF["prob"], _REF["ref_post"], _REF["ref_saver"]


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_reference_posterior_and_saver_drive_b200_sampler(built_lib, mode):
    from oracle import refshim
    if not refshim.reference_available():
        pytest.skip("no reference tree (baseline/_ref): run oracle/install_reference.py in the build container")
    import torch
    assert torch.cuda.is_available()
    from oracle import refrun
    from beetroots_b200 import MLP_BF16_TC, MLP_FP32, B200Sampler, MySamplerParams
    prob, ref_post, ref_saver = _reference_run()
    N, D, L = ref_post.N, ref_post.D, ref_post.L

    # ---- the B200 sampler with the SAME reference objects
    smp = B200Sampler(MySamplerParams(1e-4, 1e-5, 0.99, [0.5, 0.5], K, True, False), D, L, N, rng=np.random.default_rng(8),
                      mlp_mode=MLP_BF16_TC if mode == "bf16" else MLP_FP32)
    saver = refrun.memory_saver(N, D, D, L, batch_size=T, freq_save=1)
    smp.sample(ref_post, saver, T, Theta_0=prob.theta_init * 1, disable_progress_bar=True, T_BI=T_BI)

    # the Saver protocol was served the same way: same datasets, same shapes, same additional arrays
    assert len(saver.batches) == len(ref_saver.batches) >= 1
    a, b = saver.batches[0], ref_saver.batches[0]
    assert set(b) <= set(a), set(b) - set(a)      # every dataset of the reference run, incl. the nll_utils arrays (s_a, m_m, lambda_, ...)
    for k in set(a) & set(b):
        assert a[k].shape == b[k].shape, (k, a[k].shape, b[k].shape)
    assert set(ref_saver.additional) <= set(saver.additional)
    for k in ref_saver.additional:
        assert saver.additional[k].shape == ref_saver.additional[k].shape

    # chains: posterior means agree within Monte-Carlo error, accept rates and objectives are alike
    ca, cb = a["list_Theta"][T_BI:], b["list_Theta"][T_BI:]
    ma, mb = ca.mean(axis=0), cb.mean(axis=0)
    se = np.sqrt(_batch_means_se(ca) ** 2 + _batch_means_se(cb) ** 2) + 1e-12
    z = np.abs(ma - mb) / se
    spread = 0.5 * (ca.std(axis=0) + cb.std(axis=0))
    ratio = np.abs(ma - mb) / spread
    print(f"[reference drop-in {mode}] |dmean| / chain std: median {np.median(ratio):.3f} 95% {np.percentile(ratio, 95):.2f} max "
          f"{ratio.max():.2f}; batch-means z: median {np.median(z):.2f} 95% {np.percentile(z, 95):.2f}")
    # 120 strongly autocorrelated iterates per chain (the reference needs ~1 s per iteration): at this step size a chain is a slow
    # random walk around its start, and batch means under-estimate its Monte-Carlo error (z is printed, not asserted).  The
    # yardstick is the chains' own spread: two independent chains of the same sampler differ by about one spread per pixel.
    assert np.median(ratio) < 1.0 and np.percentile(ratio, 95) < 4.0, "posterior means differ beyond the chains' own spread"
    # 68 % interval widths
    wa = np.percentile(ca, 84, axis=0) - np.percentile(ca, 16, axis=0)
    wb = np.percentile(cb, 84, axis=0) - np.percentile(cb, 16, axis=0)
    assert np.median(np.abs(wa - wb) / (0.5 * (wa + wb))) < 0.6        # (68 % widths from 120 autocorrelated iterates)
    for kernel_type in (0, 1):
        ra = a["list_accepted_t"][a["list_type_t"] == kernel_type].mean()
        rb = b["list_accepted_t"][b["list_type_t"] == kernel_type].mean()
        assert abs(ra - rb) < 0.08, (kernel_type, ra, rb)
    oa, ob = a["list_objective"][T_BI:].mean(), b["list_objective"][T_BI:].mean()
    assert abs(oa - ob) < 0.05 * abs(ob) + 3.0 * (a["list_objective"][T_BI:].std() + b["list_objective"][T_BI:].std())
