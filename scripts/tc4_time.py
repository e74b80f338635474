"""Kernel time of the tensor-core forward map (variant from BT_TC_VARIANT), rows fixed; BT_TC_DEBUG=1 prints the timeline."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from beetroots_b200 import modelling as M, DevicePosterior, MLP_BF16_TC
from beetroots_b200.network import synthetic_weights
Mrows = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
jac = len(sys.argv) > 2 and sys.argv[2] == "jac"
L = int(sys.argv[3]) if len(sys.argv) > 3 else 10
net = synthetic_weights(0, L)
fm = M.NeuralNetworkApprox(net, {"kappa": None, "P": None, "radm": None, "Avmax": None, "angle": -1.5})
theta = np.random.default_rng(0).uniform(-1.2, 1.2, size=(Mrows // (4 if jac else 1), 4))
lik = M.MixingModelsLikelihood(fm, 4, L, 8, np.ones((8, L)), 1.0, 0.1, 0.5, a0=np.zeros((1, L)), a1=np.ones((1, L)))
dev = DevicePosterior.from_posterior(M.Posterior(4, L, 8, lik, None, None), mlp_mode=MLP_BF16_TC)
dev.forward_map(theta, compute_derivatives=jac)
dev.mlp_profile(True, True)
n = 1 if os.environ.get("BT_TC_DEBUG") else 3
for _ in range(n):
    dev.forward_map(theta, compute_derivatives=jac)
ms, rows, launches = dev.mlp_profile(True, False)
cols = theta.shape[0] * (4 if jac else 1)
per = ms / launches
print(f"variant {dev.mlp_kernel_variant()} ncmax {os.environ.get('BT_TC4_NCMAX','-')} mode {os.environ.get('BT_TC4_MODE','-')} slots {os.environ.get('BT_TC4_SLOTS','-')} jac={jac}: {per:.3f} ms per launch, {cols / per / 1e3:.1f} M columns/s, "
      f"{cols * 1.368732e6 / per / 1e9:.0f} TFLOP/s, cycles per 128 columns per SM pair @1.9GHz: {per * 1e-3 * 1.9e9 / (cols / 128 / 74):.0f}", flush=True)
