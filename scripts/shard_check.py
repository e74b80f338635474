"""Run a short Philox chain on a row-band sharded map (torchrun, one rank per GPU) and dump the final state.
Used by tests/test_gpu_multi.py to check that the chain is bit-identical for any number of GPUs."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist

from beetroots_b200 import MLP_BF16_TC, MLP_FP32, B200Sampler, MySamplerParams
from beetroots_b200.device import gpu_log_f
from beetroots_b200.parallel import ShardedDevicePosterior
from beetroots_b200.synthetic import make_problem

out, mode = sys.argv[1], sys.argv[2]
H, W, K = 37, 24, 6
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
# BT_SHARD_ONE_GPU=1: every rank drives its band on GPU 0 and the collectives go through gloo on the host -- the same
# make_band_plan / exchange_halo / reduction code as over NCCL, testable on a single-GPU machine
ONE_GPU = os.environ.get("BT_SHARD_ONE_GPU") == "1"
if ONE_GPU:
    local = 0
torch.cuda.set_device(local)
if world > 1:
    if ONE_GPU:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mlp = MLP_BF16_TC if mode in ("bf16", "samplebf16") else MLP_FP32
NEXT = mode == "next"          # fp32 + the "next" rows: chain store, optimisation mode, regularisation-weight update
prob = make_problem(H, W, lambda fm, th: gpu_log_f(fm, th, device=local), L=10, k_mtm=K, use_next_nearest_neighbors=(mode == "fp32n8"))
post = prob.posterior
XY = {}
if mode == "mask":
    # a non-rectangular, shuffled map (abstract_spatial_prior.py:13-55): pixels and a corner removed, one empty row
    from beetroots_b200 import modelling as M
    X0, Y0 = np.repeat(np.arange(H), W), np.tile(np.arange(W), H)
    keep = np.ones(H * W, bool)
    keep[[7, 20, 333]] = False
    keep[(X0 >= 30) & (Y0 >= 15)] = False
    keep[X0 == 18] = False
    sel = np.where(keep)[0][np.random.default_rng(3).permutation(int(keep.sum()))]
    lik, n = post.likelihood, sel.size
    lik2 = M.MixingModelsLikelihood(lik.forward_map, post.D, post.L, n, lik.y[sel], lik.sigma_a[sel], lik.sigma_m[sel],
                                    lik.omega[sel], lik.log_fm1[sel], lik.log_fp1[sel])
    sp2 = M.L22DiscreteGradSpatialPrior(post.D, n, X0[sel], Y0[sel], None, False, post.prior_spatial.initial_weights)
    post = M.Posterior(post.D, post.L, n, lik2, sp2, post.prior_indicator)
    prob.theta_init = prob.theta_init[sel]
    XY = dict(X=X0[sel], Y=Y0[sel])
sdev = ShardedDevicePosterior(post, H, W, device=local, mlp_mode=mlp, **XY)
smp = B200Sampler(MySamplerParams(1e-4, 1e-5, 0.99, [0.5, 0.5], K, True, False), post.D, post.L, post.N, device=local, mlp_mode=mlp)
smp.attach(post, sdev)
smp.philox_seed = 777
sdev.compute_all(prob.theta_init)
sdev.init_v_from_grad()
stats = []
for t in range(1, 7):
    a = smp.generate_new_sample_mtm(t, post) if t % 2 == 0 else smp.generate_new_sample_pmala_rmsprop(t, post)
    stats.append(a)
th, v, j = sdev.get_state()
terms = sdev.objective_terms()
extra = {}
if NEXT:
    from beetroots_b200.mml import EBayesMMLELogRate
    # (1) on-device chain store: three more iterations, each pushed; summaries gathered over the bands
    sdev.chain_reserve(4)
    sdev.chain_push()
    for t in range(7, 10):
        smp.generate_new_sample_pmala_rmsprop(t, post)
        sdev.chain_push()
    st = sdev.chain_stats(0, None, [10.0, 50.0, 90.0], want_var=True)
    extra.update(chain=sdev.chain_get(), chain_mean=st["mean"], chain_var=st["var"], chain_lo=st["lo"], chain_hi=st["hi"])
    # (2) optimisation mode: one MTM and one PMALA step (global argmin / global objective through all-reduce)
    smp.stochastic = False
    a_mtm = smp.generate_new_sample_mtm(10, post)
    a_pm = smp.generate_new_sample_pmala_rmsprop(11, post)
    extra.update(theta_opt=sdev.get_state()[0], opt_stats=np.array([a_mtm[0], float(a_pm[0])]),
                 opt_objective=sdev.objective_terms()["objective"])
    smp.stochastic = True
    # (3) regularisation-weight update, then two sampling iterations with the new weights
    opt = EBayesMMLELogRate(scale=1.0, N0=1, N1=+np.inf, dim=post.D * post.N, vmin=1e-8, vmax=1e8, homogeneity=2.0, exponent=0.8)
    taus = []
    for t in range(12, 14):
        tau_t = smp.sample_regu_hyperparams(post, opt, t - 11)
        post.prior_spatial.weights = tau_t * 1
        sdev.update_spatial_weights(post.prior_spatial, post.prior_indicator)
        taus.append(tau_t)
        smp.generate_new_sample_mtm(t, post) if t % 2 == 0 else smp.generate_new_sample_pmala_rmsprop(t, post)
    extra.update(taus=np.array(taus), theta_tau=sdev.get_state()[0])
if mode in ("sample", "samplebf16"):
    # the public entry point on the sharded device: B200Sampler.sample with a Saver (only the root writes), regularisation
    # weights updated from iteration 3 on, on-device chain store; every rank seeds its host generator differently
    from beetroots_b200.saver import ChainSaver
    rank = int(os.environ.get("RANK", "0"))
    smp2 = B200Sampler(MySamplerParams(1e-4, 1e-5, 0.99, [0.5, 0.5], K, True, False), post.D, post.L, post.N, device=local,
                       mlp_mode=mlp, rng=np.random.default_rng(5 if rank == 0 else 1000 + rank), device_chain=3, device_chain_thin=2)
    smp2.attach(post, sdev)
    sv = ChainSaver(post.N, post.D, post.D, post.L, results_path="", batch_size=4, freq_save=2, save_forward_map_evals=True)
    smp2.sample(post, sv, 8, Theta_0=prob.theta_init, disable_progress_bar=True, regu_spatial_N0=3, T_BI=2)
    summ = smp2.posterior_summary(list_CI=(68,))
    if rank == 0:
        extra.update(s_theta=sv.saved["list_Theta"], s_v=sv.saved["list_v"], s_logf=sv.saved["list_log_f_Theta"],
                     s_obj=sv.saved["list_objective"], s_type=sv.saved["list_type_t"], s_acc=sv.saved["list_accepted_t"],
                     s_tau=sv.saved["list_tau"], s_clppd=sv.saved["clppd"], s_seed=sv.saved["philox_seed"],
                     s_mmse=summ["Theta_MMSE"], s_final=smp2.current["Theta"])
    else:
        assert sv.saved == {} and sv.memory == {}, "only the root may touch the saver"
if int(os.environ.get("RANK", "0")) == 0:
    np.savez(out, theta=th, v=v, j=j, stats=np.array(stats), objective=terms["objective"], **extra)
    print("world", world, "accept/logpa per iteration", np.array(stats).round(4).tolist(), "objective", terms["objective"])
if world > 1:
    dist.destroy_process_group()
