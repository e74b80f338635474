"""Column-major pair kernel (BT_TC_VARIANT=5, mlp_tc4.cuh): agreement with the fp32 CUDA-core kernel and kernel time.
usage: tc4_check.py [n_layers growing n_out] [rows for the timing run]"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from beetroots_b200 import modelling as M, DevicePosterior, MLP_BF16_TC, MLP_FP32
from beetroots_b200.network import synthetic_weights
n_layers = int(sys.argv[1]) if len(sys.argv) > 1 else 10
growing = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
L = int(sys.argv[3]) if len(sys.argv) > 3 else 10
Mrows = int(sys.argv[4]) if len(sys.argv) > 4 else 0
net = synthetic_weights(3, L, n_layers=n_layers, growing_factor=growing)
fm = M.NeuralNetworkApprox(net, {"kappa": None, "P": None, "radm": None, "Avmax": None, "angle": -1.5})
lik = M.MixingModelsLikelihood(fm, 4, L, 8, np.ones((8, L)), 1.0, 0.1, 0.5, a0=np.zeros((1, L)), a1=np.ones((1, L)))
dev = DevicePosterior.from_posterior(M.Posterior(4, L, 8, lik, None, None))
print("variant", dev.mlp_kernel_variant(), flush=True)
worst = 0.0
for Mr, jac in ((1, False), (257, True), (20011, False), (5003, True), (64, False), (300000, False)):
    th = np.random.default_rng(Mr).uniform(-1.5, 1.5, size=(Mr, 4))
    dev.set_mlp_mode(MLP_FP32); ref = dev.forward_map(th, compute_derivatives=jac)
    dev.set_mlp_mode(MLP_BF16_TC); out = dev.forward_map(th, compute_derivatives=jac)
    e = np.abs(out["log_f_Theta"] - ref["log_f_Theta"]).max()
    rms = np.sqrt(np.mean((out["log_f_Theta"] - ref["log_f_Theta"]) ** 2))
    ej = 0.0
    if jac:
        ej = np.abs(out["grad_log_f_Theta"] - ref["grad_log_f_Theta"]).max() / max(1.0, np.abs(ref["grad_log_f_Theta"]).max())
    print(f"M={Mr} jac={jac}: max |d ln f| {e:.3e} rms {rms:.3e} jac rel {ej:.3e}", flush=True)
    worst = max(worst, float(e), float(ej))
print("WORST", worst, flush=True)
if Mrows:
    dev.set_mlp_mode(MLP_BF16_TC)
    for jac in (False, True):
        theta = np.random.default_rng(0).uniform(-1.2, 1.2, size=(Mrows // (4 if jac else 1), 4))
        dev.forward_map(theta, compute_derivatives=jac)
        dev.mlp_profile(True, True)
        for _ in range(3):
            dev.forward_map(theta, compute_derivatives=jac)
        ms, rows, launches = dev.mlp_profile(True, False)
        cols = theta.shape[0] * (4 if jac else 1)
        per = ms / launches
        print(f"TIMING jac={jac}: {per:.3f} ms per launch (pre-pass + pair kernel), {cols / per / 1e3:.1f} M columns/s, "
              f"{cols * 1.368732e6 / per / 1e9:.0f} TFLOP/s, cycles per 128 columns per SM pair @1.9GHz: {per * 1e-3 * 1.9e9 / (cols / 128 / 74):.0f}", flush=True)
