"""Oracle (test infrastructure): float64 numpy restatement of the nnbma emulator network.

nnbma 1.0.2 (poetry.lock:2392) is NOT under /root/reference -> parity unpinned at this
boundary; the architecture follows /root/reference/data/models/meudon_pdr_model_dense/**
(input_features 4, order 3, subnetwork 34 -> n_out, n_layers 10, growing_factor 0.5,
GELU(approximate='none'), no batch-norm, restrictable last layer) and the printed repr in
/root/reference/data/models/test.ipynb cell 1.
"""
from __future__ import annotations

import numpy as np
from scipy.special import erf

_SQRT1_2 = 0.7071067811865476
_INV_SQRT_2PI = 0.3989422804014327


def gelu(x):
    return 0.5 * x * (1.0 + erf(x * _SQRT1_2))


def gelu_prime(x):
    return 0.5 * (1.0 + erf(x * _SQRT1_2)) + x * _INV_SQRT_2PI * np.exp(-0.5 * x * x)


def poly_features(x: np.ndarray, exponents: np.ndarray, with_jac: bool):
    """Monomials of degree 1..order of the inputs (PolynomialNetwork expansion).

    x (M, n_in) -> feats (M, F0) and, if with_jac, d feats / d x  (M, n_in, F0).
    """
    M, n_in = x.shape
    F0, order = exponents.shape
    feats = np.ones((M, F0))
    for j in range(order):
        idx = exponents[:, j]
        sel = idx >= 0
        feats[:, sel] *= x[:, idx[sel]]
    if not with_jac:
        return feats, None
    jac = np.zeros((M, n_in, F0))
    for f in range(F0):
        idx = [i for i in exponents[f] if i >= 0]
        for pos, i in enumerate(idx):
            term = np.ones(M)
            for pos2, i2 in enumerate(idx):
                if pos2 != pos:
                    term = term * x[:, i2]
            jac[:, i, f] += term
    return feats, jac


def mlp_forward(net, x: np.ndarray, with_jac: bool = False):
    """PolynomialNetwork -> standardise -> dense-concat GELU blocks -> linear.

    x (M, n_in) float64.  Returns y (M, n_out) and, if with_jac, the full forward-mode
    Jacobian dy/dx of shape (M, n_out, n_in) (what ``vmap(jacfwd(network.forward))`` returns,
    neural_network_approx.py:83,208-214).
    """
    x = np.asarray(x, dtype=np.float64)
    feats, jac = poly_features(x, net.exponents, with_jac)
    h = (feats - net.poly_means[None, :]) / net.poly_stds[None, :]
    dh = jac / net.poly_stds[None, None, :] if with_jac else None      # (M, n_in, F)
    for W, b in zip(net.weights, net.biases):
        pre = h @ W.T + b[None, :]
        new = gelu(pre)
        if with_jac:
            dpre = dh @ W.T                                           # (M, n_in, new)
            dnew = gelu_prime(pre)[:, None, :] * dpre
            dh = np.concatenate([dh, dnew], axis=2)
        h = np.concatenate([h, new], axis=1)
    y = h @ net.w_out.T + net.b_out[None, :]
    if not with_jac:
        return y, None
    dy = dh @ net.w_out.T                                             # (M, n_in, n_out)
    return y, np.transpose(dy, (0, 2, 1))                             # (M, n_out, n_in)
