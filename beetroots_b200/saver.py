"""ChainSaver: the Saver protocol of the reference without its hard h5py dependency.

``B200Sampler.sample`` (like ``MySampler.sample``) only talks to a saver through
``memory / initialize_memory / update_memory / check_need_to_update_memory / check_need_to_save / save_to_file /
save_additional`` (sampler/saver/abstract_saver.py:130-238, sampler/saver/my_saver.py:9-144).  The reference's own
``MySaver`` works unchanged with ``B200Sampler``; this class is for environments where it cannot be imported (its module
imports h5py at the top).  Same constructor arguments, same batching rules, same dataset names and shapes
(``list_Theta`` in linear space, ``list_<objective term>``, ``list_v``, ``list_type_t``, ``list_accepted_t``,
``list_log_proba_accept_t``, ``list_rng_state``, ``list_rng_inc`` ...).  ``save_to_file`` appends to
``<results_path>/mc_chains.hdf5`` when h5py is importable and otherwise to ``<results_path>/mc_chains.npz`` (whole
datasets, rewritten on every save) -- or keeps everything in ``self.saved`` when ``results_path`` is empty.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import numpy as np

_SKIP_IN_NAME = ("grad", "hess_diag")


def _wanted(key: str, extra: tuple = ()) -> bool:
    return not any(s in key for s in _SKIP_IN_NAME + extra)


class ChainSaver:
    def __init__(self, N: int, D: int, D_sampling: int, L: int, scaler=None, results_path: str = "",
                 batch_size: Optional[int] = None, freq_save: int = 1, save_forward_map_evals: bool = False,
                 list_idx_sampling: Optional[List[int]] = None):
        self.N, self.D, self.D_sampling, self.L = N, D, D_sampling, L
        self.list_idx_sampling = np.arange(D) if list_idx_sampling is None else np.array(list_idx_sampling)
        self.scaler = scaler
        self.batch_size, self.freq_save = batch_size, freq_save
        self.save_forward_map_evals = save_forward_map_evals
        self.t_last_save = self.t_last_init = 0
        self.next_batch_size = self.final_next_batch_size = 0
        self.memory: Dict[str, np.ndarray] = dict()
        self.saved: Dict[str, np.ndarray] = dict()           # everything written so far (npz / in-memory back end)
        self.set_results_path(results_path)

    def set_results_path(self, results_path: str) -> None:
        self.results_path = results_path
        if len(results_path) > 0:
            os.makedirs(results_path, exist_ok=True)             # (several ranks may construct their saver at the same time)

    # ---- when (abstract_saver.py:130-163)
    def check_need_to_save(self, t: int) -> bool:
        return t - self.t_last_init + 1 == self.next_batch_size * self.freq_save

    def check_need_to_update_memory(self, t: int) -> bool:
        return t % self.freq_save == 0

    # ---- what (my_saver.py:9-144)
    def _entries(self, forward_map_evals, nll_utils, dict_objective, additional_sampling_log):
        if self.save_forward_map_evals:
            for k, v in forward_map_evals.items():
                if _wanted(k):
                    yield k, v
        for k, v in nll_utils.items():
            if _wanted(k, ("nll_",)):
                yield k, v
        for k, v in dict_objective.items():
            yield k, v
        for k, v in additional_sampling_log.items():
            yield k, v

    def initialize_memory(self, T_MC: int, t: int, Theta: np.ndarray, forward_map_evals: dict = {}, nll_utils: dict = {},
                          dict_objective: dict = {}, additional_sampling_log: dict = {}) -> None:
        if self.batch_size is None:
            self.batch_size = T_MC
        self.t_last_init = t * 1
        self.next_batch_size = min(self.batch_size, (T_MC - t + 1) // self.freq_save)
        self.final_next_batch_size = B = self.next_batch_size
        pool = getattr(self, "_pool", {})                           # batch arrays of the previous save (already written out)

        def zeros(name, shape, dtype=np.float64):
            # a fresh 0.3 GB batch array at c4 costs ~80 ms of page faults; the previous batch's array is re-zeroed instead
            a = pool.get(name)
            if a is not None and a.shape == tuple(shape) and a.dtype == dtype:
                a.fill(0)
                return a
            return np.zeros(shape, dtype=dtype)

        mem = {"list_Theta": zeros("list_Theta", (B, self.N, self.D_sampling))}
        for k, v in self._entries(forward_map_evals, nll_utils, dict_objective, additional_sampling_log):
            mem[f"list_{k}"] = zeros(f"list_{k}", (B,) + tuple(np.shape(v) if isinstance(v, np.ndarray) or hasattr(v, "shape") else ()))
        mem["list_rng_state"] = zeros("list_rng_state", (B, 32), np.uint8)
        mem["list_rng_inc"] = zeros("list_rng_inc", (B, 32), np.uint8)
        self.memory = mem
        self._pool = {}

    def update_memory(self, t: int, Theta: np.ndarray, forward_map_evals: dict = {}, nll_utils: dict = {},
                      dict_objective: dict = {}, additional_sampling_log: dict = {},
                      rng_state_array: Optional[np.ndarray] = None, rng_inc_array: Optional[np.ndarray] = None) -> None:
        slot = (t - self.t_last_init) // self.freq_save
        if self.scaler is None and self.D_sampling == self.D and np.array_equal(self.list_idx_sampling, np.arange(self.D)):
            self.memory["list_Theta"][slot] = Theta                  # nothing to place or to transform: one copy (33 MB at c4)
        else:
            full = np.zeros((Theta.shape[0], self.D))
            full[:, self.list_idx_sampling] = Theta                  # sampled parameters in their columns, the rest 0
            lin = self.scaler.from_scaled_to_lin(full) if self.scaler is not None else full
            self.memory["list_Theta"][slot] = lin[:, self.list_idx_sampling]
        for k, v in self._entries(forward_map_evals, nll_utils, dict_objective, additional_sampling_log):
            if k not in ("m_a", "s_a", "m_m", "s_m") or k not in dict_objective:
                self.memory[f"list_{k}"][slot] = v
        if rng_state_array is not None and rng_inc_array is not None:
            self.memory["list_rng_state"][slot] = rng_state_array
            self.memory["list_rng_inc"][slot] = rng_inc_array

    # ---- where (abstract_saver.py:190-238)
    def _h5py(self):
        try:
            import h5py
            return h5py if hasattr(h5py, "File") else None
        except Exception:
            return None

    def save_to_file(self) -> None:
        h5 = self._h5py() if self.results_path else None
        if h5 is not None:
            path = os.path.join(self.results_path, "mc_chains.hdf5")
            if self.t_last_init == 1:
                with h5.File(path, "w") as f:
                    for k, v in self.memory.items():
                        f.create_dataset(k, data=v, maxshape=(None,) + v.shape[1:])
            else:
                with h5.File(path, "a") as f:
                    for k, v in self.memory.items():
                        f[k].resize(f[k].shape[0] + self.final_next_batch_size, axis=0)
                        f[k][-self.final_next_batch_size:] = v
        else:
            for k, v in self.memory.items():
                self.saved[k] = v.copy() if (self.t_last_init == 1 or k not in self.saved) else np.concatenate([self.saved[k], v], 0)
            if self.results_path:
                np.savez(os.path.join(self.results_path, "mc_chains.npz"), **self.saved)
        self._pool = self.memory                                     # (copied / written above: reusable for the next batch)
        self.memory = dict()

    def save_additional(self, list_arrays: List[np.ndarray], list_names: List[str]) -> None:
        assert len(list_names) == len(list_arrays)
        h5 = self._h5py() if self.results_path else None
        if h5 is not None:
            with h5.File(os.path.join(self.results_path, "mc_chains.hdf5"), "a") as f:
                for name, arr in zip(list_names, list_arrays):
                    f.create_dataset(name, data=arr)
            return
        for name, arr in zip(list_names, list_arrays):
            self.saved[name] = np.asarray(arr)
        if self.results_path:
            np.savez(os.path.join(self.results_path, "mc_chains.npz"), **self.saved)
