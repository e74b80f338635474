import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from beetroots_b200 import MLP_BF16_TC, MLP_FP32, B200Sampler, DevicePosterior, MySamplerParams
from beetroots_b200.synthetic import make_problem
from oracle import posterior as OP
K, T = 10, 200
prob = make_problem(12, 12, lambda fm, th: OP.forward_map_compute_all(fm, th, False)["log_f_Theta"], L=10, k_mtm=K)
post = prob.posterior
kinds = np.random.default_rng(5).integers(0, 2, size=T)
for mode, mlp in (("fp32", MLP_FP32), ("bf16", MLP_BF16_TC)):
    dev = DevicePosterior.from_posterior(post, mlp_mode=mlp)
    cur = dev.compute_all(prob.theta_init); dev.init_v_from_grad()
    smp = B200Sampler(MySamplerParams(1e-4, 1e-5, 0.99, [0.5, 0.5], K, True, False), post.D, post.L, post.N, mlp_mode=mlp)
    smp.attach(post, dev); smp.philox_seed = 2024
    acc = {0: [], 1: []}; lp = {0: [], 1: []}
    th0 = prob.theta_init.copy()
    for t in range(T):
        a, l = (smp.generate_new_sample_mtm(t + 1, post) if kinds[t] == 0 else smp.generate_new_sample_pmala_rmsprop(t + 1, post))
        acc[kinds[t]].append(a); lp[kinds[t]].append(l)
    th, v, j = dev.get_state()
    print(mode, "MTM acc %.3f  PMALA acc %.3f  mean logpa pmala %.4f  |theta-theta0| rms %.4f  v median %.3e obj %.2f" % (
        np.mean(acc[0]), np.mean(acc[1]), np.mean(lp[1]), np.sqrt(np.mean((th - th0) ** 2)), np.median(v), dev.objective_terms()["objective"]))
    print("   first PMALA accs", np.round(acc[1][:6], 3), "first MTM accs", np.round(acc[0][:6], 3))
    dev.close()
