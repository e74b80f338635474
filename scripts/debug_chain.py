import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from _cases import slice_noise
from beetroots_b200 import B200Sampler, DevicePosterior, MySamplerParams
from beetroots_b200 import modelling as M
from beetroots_b200.synthetic import make_problem
from oracle import posterior as OP
from oracle.sampler import OracleSampler
K, seed = 8, 987654321
prob = make_problem(30, 30, lambda fm, th: OP.forward_map_compute_all(fm, th, False)["log_f_Theta"], L=10, k_mtm=K)
post = prob.posterior
dev = DevicePosterior.from_posterior(post)
cur = dev.compute_all(prob.theta_init); dev.init_v_from_grad()
nbr = M.build_neighbor_table(post.prior_spatial.list_edges, post.N, 4)
orc = OracleSampler(post, nbr, 1e-4, 1e-5, 0.99, K); orc.init(prob.theta_init)
print("obj range", orc.current["objective_pix_chromatic"].min(), orc.current["objective_pix_chromatic"].max(), "grad absmax", np.abs(orc.current["grad"]).max())
print("init obj relerr", np.max(np.abs(cur["objective_pix_chromatic"]-orc.current["objective_pix_chromatic"])/(1+np.abs(orc.current["objective_pix_chromatic"]))))
smp = B200Sampler(MySamplerParams(1e-4, 1e-5, 0.99, [0.5, 0.5], K, True, False), post.D, post.L, post.N)
smp.attach(post, dev); smp.philox_seed = seed
for t, kind in enumerate([1, 0, 1, 0], start=1):
    if kind == 1:
        noise = {}
        for s, key in enumerate(dev.site_keys):
            z, u = dev.debug_philox_pmala(s, seed, t); noise[key] = {"z": z, "u": u}
        smp.generate_new_sample_pmala_rmsprop(t, post)
        o_acc, o_lpa = orc.pmala_step(slice_noise(noise, post))
        lu = np.zeros(post.N)
        for s, key in enumerate(dev.site_keys): lu[post.dict_sites[key]] = np.log(noise[key]["u"][post.dict_sites[key]].astype(np.float64))
    else:
        dev.begin_iteration(); noise = {}
        for s, key in enumerate(dev.site_keys):
            c, us, ua = dev.debug_philox_mtm(s, K, seed, t); noise[key] = {"cand": c, "u_sel": us, "u_acc": ua}
            dev.mtm_site(s, K, seed, t)
        dev.mtm_finish(0.99)
        o_acc, o_lpa = orc.mtm_step(slice_noise(noise, post))
        lu = np.zeros(post.N)
        for s, key in enumerate(dev.site_keys): lu[post.dict_sites[key]] = np.log(noise[key]["u_acc"][post.dict_sites[key]].astype(np.float64))
    acc, lpa = dev.step_stats()
    th, v, _ = dev.get_state()
    diff = np.where((acc > 0) != (o_acc > 0))[0]
    print(f"t={t} kind={kind} acc dev {acc.mean():.3f} orc {o_acc.mean():.3f} flips {diff.size}")
    for n in diff: print("   flip pixel", n, "lpa dev", lpa[n], "orc", o_lpa[n], "log u", lu[n])
    dth = np.abs(th - orc.current["Theta"]).max(axis=1)
    worst = np.argsort(dth)[-3:]
    print("   max dTheta", dth.max(), "worst pixels", worst, dth[worst], "acc", acc[worst], o_acc[worst])
    print("   max rel dv", np.max(np.abs(v - orc.v) / (np.abs(orc.v) + 1e-30)), " max |dlpa|", np.max(np.abs(lpa - o_lpa)))
