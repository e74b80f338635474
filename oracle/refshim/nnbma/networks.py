"""Stand-in for the absent third-party package nnbma 1.0.2 (test infrastructure only).

Implements just the surface the reference touches (neural_network_approx.py:38-122, 198-239):
``NeuralNetwork.load / set_device / double / eval / forward / evaluate /
restrict_to_output_subset / input_features / output_features / current_output_subset / device``
for the published architecture PolynomialNetwork(order 3) -> DenselyConnected(GELU erf,
dense concat), in torch float64 so the reference's ``vmap(jacfwd(network.forward))`` runs
unmodified.  Weights come from a registry filled by the test (synthetic, deterministic).
"""
import numpy as np
import torch

REGISTRY = {}


class NeuralNetwork(torch.nn.Module):
    def __init__(self, w, names):
        super().__init__()
        self.exponents = torch.as_tensor(np.asarray(w.exponents), dtype=torch.long)
        self.register_buffer("means", torch.as_tensor(w.poly_means, dtype=torch.float64))
        self.register_buffer("stds", torch.as_tensor(w.poly_stds, dtype=torch.float64))
        self.Ws = [torch.as_tensor(W, dtype=torch.float64) for W in w.weights]
        self.bs = [torch.as_tensor(b, dtype=torch.float64) for b in w.biases]
        self.W_out_full = torch.as_tensor(w.w_out, dtype=torch.float64)
        self.b_out_full = torch.as_tensor(w.b_out, dtype=torch.float64)
        self.input_features = w.n_in
        self.outputs_names = list(names)
        self.current_output_subset = list(names)
        self._rows = list(range(len(names)))
        self.device = "cpu"

    @property
    def output_features(self):
        return len(self._rows)

    @staticmethod
    def load(model_name, path_model=None):
        w, names = REGISTRY[model_name]
        return NeuralNetwork(w, names)

    def set_device(self, device):
        self.device = device

    def restrict_to_output_subset(self, output_subset):
        self._rows = [self.outputs_names.index(n) for n in output_subset]
        self.current_output_subset = list(output_subset)

    def forward(self, x):
        feats = []
        for f in range(self.exponents.shape[0]):
            t = None
            for i in self.exponents[f].tolist():
                if i >= 0:
                    t = x[..., i] if t is None else t * x[..., i]
            feats.append(t)
        h = (torch.stack(feats, dim=-1) - self.means) / self.stds
        for W, b in zip(self.Ws, self.bs):
            y = torch.nn.functional.gelu(torch.nn.functional.linear(h, W, b))
            h = torch.cat((h, y), dim=-1)
        rows = torch.as_tensor(self._rows, dtype=torch.long)
        return torch.nn.functional.linear(h, self.W_out_full[rows], self.b_out_full[rows])

    def state_dict(self, *a, **k):
        """Key names of a real nnbma PolynomialNetwork -> DenselyConnected checkpoint (data/models/*/parameters.pth,
        state_dict.pth), so beetroots_b200.device._network_weights can be exercised on it."""
        sd = {"poly.means": self.means, "poly.stds": self.stds}
        for i, (W, b) in enumerate(zip(self.Ws, self.bs)):
            sd[f"subnetwork.layers.{i}.weight"], sd[f"subnetwork.layers.{i}.bias"] = W, b
        sd["subnetwork.output_layer.weight"], sd["subnetwork.output_layer.bias"] = self.W_out_full, self.b_out_full
        return sd

    @property
    def subnetwork(self):
        class _Sub:
            pass
        sub = _Sub()
        sub.current_output_subset_indices = list(self._rows)
        return sub

    order = 3

    def evaluate(self, x: np.ndarray) -> np.ndarray:
        with torch.no_grad():
            return self.forward(torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))).numpy()
