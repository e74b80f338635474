"""Golden vectors for the regularisation-weight update, from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_mml.py

For every golden case it evaluates the reference's unweighted spatial potential
(`L22DiscreteGradSpatialPrior.neglog_pdf(Theta, arange(N), with_weights=False)`,
modelling/priors/l22_discrete_grad_prior.py:95-128) at theta0 and at every recorded state, and runs the reference
optimiser (`EBayesMMLELogRate`, sampler/utils/mml.py) over that sequence with the constructor arguments of
`MySampler.sample` (sampler/my_sampler.py:369-379).  Output: tests/golden/mml_golden.json.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402

refshim.install()

import pandas as pd  # noqa: E402

from beetroots.modelling.priors.l22_discrete_grad_prior import L22DiscreteGradSpatialPrior as RefL22  # noqa: E402
from beetroots.modelling.priors.spatial_prior_params import SpatialPriorParams  # noqa: E402
from beetroots.sampler.utils.mml import EBayesMMLELogRate as RefOpt  # noqa: E402

out = {}
for name in ("ref_n4_6x5", "ref_n8_5x6_censored"):
    g = np.load(os.path.join(HERE, name + ".npz"))
    X, Y = g["X"], g["Y"]
    N, D = X.size, 4
    df = pd.DataFrame({"idx": np.arange(N), "X": X, "Y": Y}).set_index(["X", "Y"])
    sp = RefL22(SpatialPriorParams("L2-gradient", bool(g["nnn"]), g["tau"]), "syn", D, N, df, [0, 1, 2, 3])
    states = [g["theta0"]] + [g[f"t{t}_Theta"] for t in range(1, len(g["kernels"]) + 1)]
    gX = [sp.neglog_pdf(th, idx_pix=np.arange(N), with_weights=False) for th in states]
    for scale, N0 in ((1.0, 1), (25.0, 3)):
        opt = RefOpt(scale=scale, N0=N0, N1=+np.inf, dim=D * N, vmin=1e-8, vmax=1e8, homogeneity=2.0, exponent=0.8)
        tau, taus, means = float(np.mean(g["tau"])), [], []
        for t, gx in enumerate(gX, start=1):
            if t >= N0:
                tau = float(opt.update(tau, t, float(np.sum(gx))))
            taus.append(tau)
            means.append(float(opt.mean_theta))
        out[f"{name}|scale={scale}|N0={N0}"] = {"gX_per_dim": [list(map(float, v)) for v in gX], "tau": taus,
                                               "mean_theta": means, "tau0": float(np.mean(g["tau"])), "dim": D * N}
with open(os.path.join(HERE, "mml_golden.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote", os.path.join(HERE, "mml_golden.json"), {k: v["tau"][-1] for k, v in out.items()})
