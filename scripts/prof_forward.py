"""One forward-only evaluation of the bf16 forward map on 2 M random rows (for ncu: the launch is the pre-pass kernel
followed by the pair kernel).  Usage: ncu ... python scripts/prof_forward.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from beetroots_b200 import MLP_BF16_TC, DevicePosterior
from beetroots_b200 import modelling as M
from beetroots_b200.network import synthetic_weights

L, Mr = 10, 2_000_000
net = synthetic_weights(0, L)
fm = M.NeuralNetworkApprox(net, {"kappa": None, "P": None, "radm": None, "Avmax": None, "angle": -1.5})
lik = M.MixingModelsLikelihood(fm, 4, L, 8, np.ones((8, L)), 1.0, 0.1, 0.5, a0=np.zeros((1, L)), a1=np.ones((1, L)))
dev = DevicePosterior.from_posterior(M.Posterior(4, L, 8, lik, None, None), mlp_mode=MLP_BF16_TC)
th = np.random.default_rng(1).uniform(-1.5, 1.5, size=(Mr, 4))
out = dev.forward_map(th)["log_f_Theta"]
print("variant", dev.mlp_kernel_variant(), "rows", Mr, "mean ln f", float(out.mean()))
