"""Host-side description of the nnbma emulator network (weights only, no arithmetic).

The reference forward map ``NeuralNetworkApprox``
(/root/reference/beetroots/modelling/forward_maps/neural_network_approx.py:12-105) wraps a
third-party ``nnbma`` network: ``PolynomialNetwork(input_features=4, order=3)`` feeding a
``DenselyConnected(34 -> n_out, n_layers=10, growing_factor=0.5, GELU(erf))``
(/root/reference/data/models/meudon_pdr_model_dense/**/*.json,
/root/reference/data/models/test.ipynb cell 1).  The real weights are absent from the
reference tree (.MISSING_LARGE_BLOBS:1), so every width below is a runtime value taken from
the weight shapes; nothing in the CUDA path hard-codes 34/1296.

This module only *holds* the parameters and builds deterministic synthetic ones; all
evaluation happens on the GPU (beetroots_b200/csrc) or, for tests, in oracle/.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from itertools import combinations_with_replacement
from typing import List, Sequence

import numpy as np


def poly_exponents(n_in: int, order: int) -> np.ndarray:
    """Monomial table of the polynomial expansion, shape (F, order), entries in [-1, n_in).

    Row f lists the input indices multiplied together for feature f (padded with -1), in
    degree-then-lexicographic order (itertools.combinations_with_replacement), the order
    implied by the surviving standardisation buffers ``poly.means / poly.stds`` in
    /root/reference/data/models/meudon_pdr_model_dense/state_dict.pth (squares have mean 1,
    cross terms mean 0: pattern [1,0,0,0,1,0,0,1,0,1] for degree 2).
    """
    rows = []
    for deg in range(1, order + 1):
        for combo in combinations_with_replacement(range(n_in), deg):
            rows.append(list(combo) + [-1] * (order - deg))
    return np.asarray(rows, dtype=np.int32)


def dense_widths(input_features: int, n_layers: int, growing_factor: float) -> List[int]:
    """Number of new features produced by each of the ``n_layers - 1`` dense-concat blocks.

    Block l maps the current concat (width w_l) to ``int(growing_factor * w_l)`` new
    features that are appended: 34 -> +17 -> +25 -> +38 ... (the
    legacy weights /root/reference/data/models/meudon_pdr_154PAH_model_dense/parameters.pth
    show the pattern Linear(w, w) for factor 1 with n_layers=7 -> 6 blocks + output layer).
    The rounding rule of nnbma is unverified (source absent) which is why the CUDA side
    reads widths from the arrays and never from this function.
    """
    widths, w = [], input_features
    for _ in range(n_layers - 1):
        new = max(1, int(growing_factor * w))
        widths.append(new)
        w += new
    return widths


@dataclass
class DenseMLPWeights:
    """All parameters of PolynomialNetwork -> DenselyConnected, float64 on host."""

    n_in: int                      # NN inputs (4: P, radm, Avmax, angle)
    order: int                     # polynomial order (3)
    exponents: np.ndarray          # (F0, order) int32, see poly_exponents
    poly_means: np.ndarray         # (F0,)
    poly_stds: np.ndarray          # (F0,)
    weights: List[np.ndarray] = field(default_factory=list)   # block l: (new_l, in_l)
    biases: List[np.ndarray] = field(default_factory=list)    # block l: (new_l,)
    w_out: np.ndarray = None       # (n_out, F_total) -- already restricted to the fitted lines
    b_out: np.ndarray = None       # (n_out,)

    @property
    def n_feat0(self) -> int:
        return int(self.exponents.shape[0])

    @property
    def n_out(self) -> int:
        return int(self.w_out.shape[0])

    @property
    def n_feat_total(self) -> int:
        return int(self.w_out.shape[1])

    @property
    def block_in(self) -> List[int]:
        return [int(w.shape[1]) for w in self.weights]

    @property
    def block_new(self) -> List[int]:
        return [int(w.shape[0]) for w in self.weights]

    def macs_per_row(self) -> int:
        """Algorithmic multiply-accumulates of one forward row."""
        return int(sum(w.size for w in self.weights) + self.w_out.size)

    def restrict_to_output_subset(self, rows: Sequence[int]) -> "DenseMLPWeights":
        """Row-slice the last layer (nnbma ``restrict_to_output_subset``,
        neural_network_approx.py:87-105)."""
        rows = np.asarray(rows, dtype=np.int64)
        return DenseMLPWeights(
            self.n_in, self.order, self.exponents, self.poly_means, self.poly_stds,
            list(self.weights), list(self.biases), self.w_out[rows].copy(), self.b_out[rows].copy(),
        )

    def validate(self) -> None:
        w = self.n_feat0
        assert self.poly_means.shape == (w,) and self.poly_stds.shape == (w,)
        for W, b in zip(self.weights, self.biases):
            assert W.ndim == 2 and W.shape[1] == w, (W.shape, w)
            assert b.shape == (W.shape[0],)
            w += W.shape[0]
        assert self.w_out.shape[1] == w and self.b_out.shape == (self.w_out.shape[0],)


def synthetic_weights(
    seed: int = 0,
    n_out: int = 10,
    n_in: int = 4,
    order: int = 3,
    n_layers: int = 10,
    growing_factor: float = 0.5,
    out_center: float = -9.0,
    out_spread: float = 0.3,
    out_gain: float = 0.35,
) -> DenseMLPWeights:
    """Deterministic stand-in weights with the reference architecture (SURVEY.md §8d recipe:
    'synthetic-weight MLP scaled so that log10 f ~ -9 +- 0.3', MLP weight seed 0).

    No forward pass is needed to build them: hidden blocks use a variance-preserving
    N(0, 1/in) init, the last layer is scaled analytically so that the output varies by a few
    tenths of a dex over the unit box, and the output bias puts each line near 10^-9.
    """
    rng = np.random.default_rng(seed)
    exps = poly_exponents(n_in, order)
    f0 = exps.shape[0]
    # standardisation buffers of a N(0,1) input population: exact moments of monomials are not
    # needed, any fixed (mean, std) pair defines a valid network; use values close to the shipped ones.
    deg = (exps >= 0).sum(axis=1)
    means = np.zeros(f0)
    stds = np.ones(f0)
    for f in range(f0):
        idx = exps[f][exps[f] >= 0]
        counts = np.bincount(idx, minlength=n_in)
        # E[prod x_i^c_i] for iid N(0,1): product of (c-1)!! for even c, 0 otherwise
        m1 = np.prod([_dfact(c - 1) if c % 2 == 0 else 0.0 for c in counts])
        m2 = np.prod([_dfact(2 * c - 1) for c in counts])
        means[f] = m1
        stds[f] = np.sqrt(max(m2 - m1 * m1, 1e-12))
    del deg
    net = DenseMLPWeights(n_in, order, exps, means, stds)
    w = f0
    for new in dense_widths(f0, n_layers, growing_factor):
        net.weights.append(rng.standard_normal((new, w)) * np.sqrt(1.6 / w))
        net.biases.append(rng.standard_normal(new) * 0.1)
        w += new
    net.w_out = rng.standard_normal((n_out, w)) * (out_gain / np.sqrt(w))
    net.b_out = out_center + rng.uniform(-out_spread, out_spread, size=n_out)
    net.validate()
    return net


def _dfact(n: int) -> float:
    """Double factorial n!! with (-1)!! = 1."""
    r = 1.0
    while n > 1:
        r *= n
        n -= 2
    return r
