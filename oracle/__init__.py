"""CPU oracle for the beetroots sampler step -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A float64 numpy restatement of the reference algorithm on the hot path, each
function citing the reference file:line it follows.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it, and there only as the checker or as the timed CPU baseline -- never as a fallback
for the CUDA path (``beetroots_b200`` raises if its extension is missing).

Pinning (see tests/golden/make_golden.py):
  * everything except the MLP body is pinned against the UNMODIFIED reference imported from
    /root/reference in the build container (MySampler.generate_new_sample_pmala_rmsprop /
    generate_new_sample_mtm, Posterior.compute_all, MixingModelsLikelihood, the priors, the
    MTM proposal density) with a recorded noise tape, and against the reference's own
    known-answer tests (tests/modelling/priors/*, tests/sampler/utils/test_utils.py,
    tests/modelling/likelihoods/test_utils.py);
  * the MLP body (nnbma 1.0.2, absent from /root/reference) is a restatement of its published
    architecture: PARITY UNPINNED at the nnbma boundary.  The reference's own
    NeuralNetworkApprox runs unmodified on top of it (oracle/refshim/nnbma), so Jacobian
    layout, kappa handling, LOGE_10 scaling and column selection are exercised by reference code.
"""
