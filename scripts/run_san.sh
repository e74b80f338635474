# compute-sanitizer passes over the tensor-core forward-map kernels (small launches; own timeouts)
mkdir -p gpurun_out
cat > /tmp/san_small.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from beetroots_b200 import modelling as M, DevicePosterior, MLP_BF16_TC
from beetroots_b200.network import synthetic_weights
L = int(sys.argv[1]) if len(sys.argv) > 1 else 10
net = synthetic_weights(3, L)
fm = M.NeuralNetworkApprox(net, {"kappa": None, "P": None, "radm": None, "Avmax": None, "angle": -1.5})
lik = M.MixingModelsLikelihood(fm, 4, L, 8, np.ones((8, L)), 1.0, 0.1, 0.5, a0=np.zeros((1, L)), a1=np.ones((1, L)))
dev = DevicePosterior.from_posterior(M.Posterior(4, L, 8, lik, None, None), mlp_mode=MLP_BF16_TC)
print("variant", dev.mlp_kernel_variant(), flush=True)
for Mr, jac in ((700, False), (300, True), (40000, False)):
    th = np.random.default_rng(Mr).uniform(-1.5, 1.5, size=(Mr, 4))
    out = dev.forward_map(th, compute_derivatives=jac)
    print(Mr, jac, float(out["log_f_Theta"].mean()), flush=True)
print("done", flush=True)
PY
for tool in memcheck synccheck racecheck; do
  echo "=== compute-sanitizer --tool $tool, default kernels (variant 5), 10 lines" > gpurun_out/san_$tool.log
  timeout 300 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_small.py 10 >> gpurun_out/san_$tool.log 2>&1
  echo "exit code $?" >> gpurun_out/san_$tool.log
done
echo "=== variant 4 (round-1 pair kernel fed by the pre-pass) on the 32-line network, gate lifted by BT_TC3P_ANY=1: plain run with a 90 s limit" > gpurun_out/san_v4_32.log
BT_TC_VARIANT=4 timeout 90 python /tmp/san_small.py 32 >> gpurun_out/san_v4_32.log 2>&1
echo "exit code $?" >> gpurun_out/san_v4_32.log
for f in gpurun_out/san_*.log; do echo "--- $f"; tail -6 $f; done
