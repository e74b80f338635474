import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from _cases import load_golden, problem_from_golden, golden_noise
from beetroots_b200 import B200Sampler, DevicePosterior, MySamplerParams
np.set_printoptions(precision=5, linewidth=200, suppress=False)
g = load_golden("ref_n4_6x5")
post, _ = problem_from_golden(g)
K = int(g["K"])
dev = DevicePosterior.from_posterior(post)
dev.compute_all(g["theta0"]); dev.init_v_from_grad()
smp = B200Sampler(MySamplerParams(1e-4, 1e-5, 0.99, [0.5, 0.5], K, True, False), post.D, post.L, post.N)
smp.attach(post, dev)
for t, kind in enumerate(g["kernels"][:3], start=1):
    noise = golden_noise(g, t, int(kind), post, full=True)
    if kind == 1: smp.generate_new_sample_pmala_rmsprop(t, post, noise)
    else: smp.generate_new_sample_mtm(t, post, noise)
    acc, lpa = dev.step_stats()
    print("t", t, "kind", kind)
    print(" acc dev", acc.astype(int)); print(" acc ref", g[f"t{t}_accept"].astype(int))
    print(" lpa dev", lpa); print(" lpa ref", g[f"t{t}_logpa"])
    th, v, j = dev.get_state()
    print(" max dTheta", np.abs(th - g[f"t{t}_Theta"]).max(), "max rel dv", np.max(np.abs(v - g[f"t{t}_v"]) / np.abs(g[f"t{t}_v"])))
