import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
from beetroots_b200 import modelling as M, DevicePosterior, MLP_BF16_TC
from beetroots_b200.network import synthetic_weights
L = 10
net = synthetic_weights(0, L)
fm = M.NeuralNetworkApprox(net, {"kappa": None, "P": None, "radm": None, "Avmax": None, "angle": -1.5})
Mr = int(sys.argv[1])
theta = np.random.default_rng(0).uniform(-1.2, 1.2, size=(Mr, 4))
lik = M.MixingModelsLikelihood(fm, 4, L, 8, np.ones((8, L)), 1.0, 0.1, 0.5, a0=np.zeros((1, L)), a1=np.ones((1, L)))
dev = DevicePosterior.from_posterior(M.Posterior(4, L, 8, lik, None, None), mlp_mode=MLP_BF16_TC)
out = dev.forward_map(theta, compute_derivatives=False)
print("ok", np.isfinite(out["log_f_Theta"]).all())
