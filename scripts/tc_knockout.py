"""Kernel time of the tensor-core forward map with parts knocked out (BT_TC3_FLAGS), rows fixed."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from beetroots_b200 import modelling as M, DevicePosterior, MLP_BF16_TC
from beetroots_b200.network import synthetic_weights
Mrows = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
jac = len(sys.argv) > 2 and sys.argv[2] == "jac"
L = 10
net = synthetic_weights(0, L)
fm = M.NeuralNetworkApprox(net, {"kappa": None, "P": None, "radm": None, "Avmax": None, "angle": -1.5})
theta = np.random.default_rng(0).uniform(-1.2, 1.2, size=(Mrows, 4))
lik = M.MixingModelsLikelihood(fm, 4, L, 8, np.ones((8, L)), 1.0, 0.1, 0.5, a0=np.zeros((1, L)), a1=np.ones((1, L)))
dev = DevicePosterior.from_posterior(M.Posterior(4, L, 8, lik, None, None), mlp_mode=MLP_BF16_TC)
dev.forward_map(theta, compute_derivatives=jac)
dev.mlp_profile(True, True)
for _ in range(3):
    dev.forward_map(theta, compute_derivatives=jac)
ms, rows, launches = dev.mlp_profile(True, False)
cols = Mrows * (4 if jac else 1)
print(f"variant {os.environ.get('BT_TC_VARIANT','1')} ko {os.environ.get('BT_TC3_KO','0')} flags {os.environ.get('BT_TC3_FLAGS','0')} norep {os.environ.get('BT_TC3_NOREP','-')} nofold {os.environ.get('BT_TC3_NOFOLD','-')}: "
      f"{ms / launches:.3f} ms per launch, {cols / (ms / launches) / 1e3:.1f} M columns/s, cycles per 128 columns per SM pair @1.9GHz: {ms / launches * 1e-3 * 1.9e9 / (cols / 128 / 74):.0f}")
