"""Compare the activation buffers dumped by the tc3 kernel (BT_TC3_DUMP) with a numpy evaluation of the network."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from scipy.special import erf
from beetroots_b200 import modelling as M, DevicePosterior, MLP_BF16_TC
from beetroots_b200.network import synthetic_weights
L = 10
Mrows = 128
net = synthetic_weights(0, L)
fm = M.NeuralNetworkApprox(net, {"kappa": None, "P": None, "radm": None, "Avmax": None, "angle": -1.5})
rng = np.random.default_rng(0)
theta = rng.uniform(-1.2, 1.2, size=(Mrows, 4))
lik = M.MixingModelsLikelihood(fm, 4, L, Mrows, np.ones((Mrows, L)), 1.0, 0.1, 0.5, a0=np.zeros((1, L)), a1=np.ones((1, L)))
dev = DevicePosterior.from_posterior(M.Posterior(4, L, Mrows, lik, None, None))
dev.set_mlp_mode(MLP_BF16_TC)
path = "/tmp/tc3_dump.bin"
os.environ["BT_TC3_DUMP"] = path
out = dev.forward_map(theta, compute_derivatives=False)
raw = np.fromfile(path, dtype=np.uint16)
n_kb = raw.size // 2 // (64 * 64)
raw = raw.reshape(2, n_kb, 64 * 64)
# un-swizzle: element (col n, k) at byte (n>>3)*1024 + (n&7)*128 + (((k>>3) ^ (n&7)) << 4) + (k&7)*2
n = np.arange(64)[:, None]; k = np.arange(64)[None, :]
off = ((n >> 3) * 1024 + (n & 7) * 128 + ((((k >> 3) ^ (n & 7)) & 7) << 4) + (k & 7) * 2) // 2
H = np.zeros((128, n_kb * 64), dtype=np.float32)
for c in range(2):
    for kb in range(n_kb):
        bits = raw[c, kb][off].astype(np.uint32) << 16
        H[c * 64:(c + 1) * 64, kb * 64:(kb + 1) * 64] = bits.view(np.float32)
# numpy reference
x = np.concatenate([theta[:, 1:], np.full((Mrows, 1), -1.5)], axis=1)     # NN inputs: P, radm, Avmax, angle
feats = []
for e in net.exponents:
    t = np.ones(Mrows)
    for i in e:
        if i >= 0: t = t * x[:, i]
    feats.append(t)
h = (np.stack(feats, 1) - net.poly_means) / net.poly_stds
for W, b in zip(net.weights, net.biases):
    z = h @ W.T + b
    h = np.concatenate([h, 0.5 * z * (1 + erf(z / np.sqrt(2)))], axis=1)
err = np.abs(H[:, :h.shape[1]] - h)
tol = 0.02 + 0.02 * np.abs(h)
bad = err > tol
print("features:", h.shape[1], "bad fraction per 16-feature group:")
edges = np.cumsum([net.n_feat0] + list(net.block_new))
print("layer edges", edges)
for g in range(0, h.shape[1], 16):
    fr = bad[:, g:g + 16].mean()
    if fr > 0: print(f"  k {g:4d}..{g+15:4d}: bad {fr:.2f}  max err {err[:, g:g+16].max():.3f}   per-col-block(32) {[round(float(bad[cb:cb+32, g:g+16].mean()),2) for cb in range(0,128,32)]}")
print("first bad feature:", int(np.argmax(bad.any(axis=0))) if bad.any() else None)
# fingerprint the last hidden block: which K-blocks does the device result correspond to?
hin = h[:, :864]
W8, b8 = net.weights[8], net.biases[8]
dev8 = H[:, 864:1296]
def g(z): return 0.5 * z * (1 + erf(z / np.sqrt(2)))
def score(mask, bias=True):
    z = (hin * mask) @ W8.T + (b8 if bias else 0)
    return float(np.abs(g(z) - dev8).mean())
kbs = np.arange(864) // 64
print("full", score(np.ones(864)), "nobias", score(np.ones(864), False))
for kb in range(14):
    print("  without kb", kb, score((kbs != kb).astype(float)), " only kb", kb, score((kbs == kb).astype(float)))
print("even", score((kbs % 2 == 0).astype(float)), "odd", score((kbs % 2 == 1).astype(float)))
print("old only", score((kbs < 9).astype(float)), "new only", score((kbs >= 9).astype(float)))
ref8 = g(hin @ W8.T + b8)
# best matching reference feature for device features (by mean abs diff)
for j in list(range(0, 8)) + [100, 127, 128, 129, 200, 255, 256, 257, 300, 343, 344, 345, 400, 431]:
    dd = np.abs(ref8 - dev8[:, j:j + 1]).mean(axis=0)
    print(f"dev feature {j:3d}: best ref {int(dd.argmin()):3d} (err {dd.min():.4f}), own err {dd[j]:.4f}")
zpre = hin @ W8.T
print("pre-activation rms", zpre.std(), " dev8 rms", dev8.std(), "ref8 rms", ref8.std())
