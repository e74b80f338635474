import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from beetroots_b200 import modelling as M, DevicePosterior, MLP_BF16_TC, MLP_FP32
from beetroots_b200.network import synthetic_weights
Mrows = int(sys.argv[1]) if len(sys.argv) > 1 else 64
jac = len(sys.argv) > 2 and sys.argv[2] == "jac"
L = 10
net = synthetic_weights(0, L)
fm = M.NeuralNetworkApprox(net, {"kappa": None, "P": None, "radm": None, "Avmax": None, "angle": -1.5})
rng = np.random.default_rng(0)
theta = rng.uniform(-1.2, 1.2, size=(Mrows, 4))
lik = M.MixingModelsLikelihood(fm, 4, L, Mrows, np.ones((Mrows, L)), 1.0, 0.1, 0.5, a0=np.zeros((1, L)), a1=np.ones((1, L)))
dev = DevicePosterior.from_posterior(M.Posterior(4, L, Mrows, lik, None, None))
ref = dev.forward_map(theta, compute_derivatives=jac)
dev.set_mlp_mode(MLP_BF16_TC)
t0 = time.time()
out = dev.forward_map(theta, compute_derivatives=jac)
print("tc call took", time.time() - t0)
d = out["log_f_Theta"] - ref["log_f_Theta"]
print("M", Mrows, "max |d ln f|", np.abs(d).max(), "rms", np.sqrt((d**2).mean()), "ref std", ref["log_f_Theta"].std(axis=0).mean())
print("first rows tc ", out["log_f_Theta"][0, :4], "\n           ref", ref["log_f_Theta"][0, :4])
if jac:
    dj = out["grad_log_f_Theta"] - ref["grad_log_f_Theta"]
    print("max |d jac|", np.abs(dj).max(), "jac scale", np.abs(ref["grad_log_f_Theta"]).max())
    print(out["grad_log_f_Theta"][0, :, 0], ref["grad_log_f_Theta"][0, :, 0])
if os.environ.get("BT_DBG_ROWS"):
    e = np.abs(d).max(axis=1)
    n = min(Mrows, 256)
    print("per-row max err, blocks of 16 rows:", [float(f"{e[i:i+16].max():.3f}") for i in range(0, n, 16)])
    print("per-line max err:", [float(f"{x:.3f}") for x in np.abs(d).max(axis=0)])
