"""Golden memory layout of the reference saver (sampler/saver/my_saver.py) for tests/test_host.py::test_chain_saver_*.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_saver.py
Drives the UNMODIFIED reference ``MySaver`` through 12 iterations (freq_save 2, batch_size 3, 2 of 3 parameters sampled)
with deterministic inputs and stores every batch it would have written (``save_to_file`` is replaced by a recorder: h5py
is not installed here) as tests/golden/saver_golden.npz.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refshim  # noqa: E402

refshim.install()
from beetroots.sampler.saver.my_saver import MySaver  # noqa: E402


sys.path.insert(0, os.path.dirname(HERE))
from _cases import Pow10Scaler as Pow10, saver_inputs as inputs  # noqa: E402


if __name__ == "__main__":
    out, batches = {}, []

    class Rec(MySaver):
        def save_to_file(self):
            batches.append({k: v.copy() for k, v in self.memory.items()})
            self.memory = dict()

    sv = Rec(5, 3, 2, 2, Pow10(), results_path="", batch_size=3, freq_save=2, save_forward_map_evals=True, list_idx_sampling=[0, 2])
    T = 12
    for t in range(1, T + 1):
        x = inputs(t)
        kw = {k: x[k] for k in ("Theta", "forward_map_evals", "nll_utils", "dict_objective", "additional_sampling_log")}
        if sv.memory == {}:
            sv.initialize_memory(T, t, **kw)
            sv.update_memory(t, rng_state_array=x["rng"][0], rng_inc_array=x["rng"][1], **kw)
        elif sv.check_need_to_update_memory(t):
            sv.update_memory(t, rng_state_array=x["rng"][0], rng_inc_array=x["rng"][1], **kw)
        if sv.check_need_to_save(t):
            sv.save_to_file()
    for b, mem in enumerate(batches):
        for k, v in mem.items():
            out[f"b{b}__{k}"] = v
    out["n_batches"] = np.array(len(batches))
    np.savez_compressed(os.path.join(HERE, "saver_golden.npz"), **out)
    print("batches", len(batches), "keys", sorted(batches[0].keys()), "shapes", {k: v.shape for k, v in batches[0].items()})
