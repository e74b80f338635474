"""Row-band sharding of the pixel map over the GPUs of one node.

The only cross-pixel dependency of the sampler step is the nearest-neighbour stencil of the
spatial prior (l22_discrete_grad_prior.py:9-80) and of the MTM proposal (sampler/utils/utils.py
:100-125), and pixels of one colour class are conditionally independent
(abstract_spatial_prior.py:125-153).  So rank r owns ``H / G`` consecutive rows (X = slow axis,
the ``idx`` order of the reference maps), stores one ghost row above and below, and after each
colour sweep sends its first / last owned row to its neighbours (16 KiB for W=1024, D=4) over
NCCL.  Likelihood, network and RNG are per pixel: the Philox counters are keyed on the GLOBAL
pixel id, so the chain is bit-identical for 1, 2, 4 or 8 GPUs.  The scalar statistics
(accept rate, mean log-ratio, objective terms) are all-reduced.

``make_band_plan`` and ``exchange_halo`` are pure host logic (tested on CPU with gloo).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, Optional

import numpy as np

from . import modelling as M
from .device import MLP_FP32, DevicePosterior


@dataclass
class BandPlan:
    H: int
    W: int
    world: int
    rank: int
    row0: int                 # first owned global row
    row1: int                 # one past the last owned global row
    g0: int                   # first stored global row (row0 or row0 - 1)
    g1: int                   # one past the last stored global row
    local_global_id: np.ndarray   # (N_loc,) global pixel id of every stored pixel, in local order
    owned: np.ndarray             # (N_loc,) uint8
    send_up: Optional[np.ndarray]    # local ids of my first owned row  (-> rank - 1)
    send_down: Optional[np.ndarray]  # local ids of my last owned row   (-> rank + 1)
    recv_up: Optional[np.ndarray]    # local ids of the ghost row above (<- rank - 1)
    recv_down: Optional[np.ndarray]  # local ids of the ghost row below (<- rank + 1)
    owned_gid_by_rank: Optional[list] = None   # global ids of the pixels every rank owns, in that rank's local order (gathers)

    @property
    def n_local(self) -> int:
        return int(self.local_global_id.size)


def band_rows(H: int, world: int, rank: int):
    """Contiguous, balanced row ranges (first H % world ranks get one extra row)."""
    base, rem = divmod(H, world)
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


def make_band_plan(H: int, W: int, world: int, rank: int) -> BandPlan:
    assert 1 <= world <= H, "need at least one row per rank"
    r0, r1 = band_rows(H, world, rank)
    g0 = r0 - 1 if rank > 0 else r0
    g1 = r1 + 1 if rank < world - 1 else r1
    rows = np.arange(g0, g1)
    gid = (rows[:, None] * W + np.arange(W)[None, :]).reshape(-1).astype(np.int64)
    owned = np.repeat((rows >= r0) & (rows < r1), W).astype(np.uint8)

    def local_row(gr):
        return ((gr - g0) * W + np.arange(W)).astype(np.int32)

    return BandPlan(H, W, world, rank, r0, r1, g0, g1, gid, owned,
                    local_row(r0) if rank > 0 else None,
                    local_row(r1 - 1) if rank < world - 1 else None,
                    local_row(r0 - 1) if rank > 0 else None,
                    local_row(r1) if rank < world - 1 else None)


def make_band_plan_xy(X: np.ndarray, Y: np.ndarray, world: int, rank: int) -> BandPlan:
    """Row bands of an arbitrary (non-rectangular, arbitrarily ordered) pixel set given by its index coordinates, as the
    reference's maps are (abstract_spatial_prior.py:13-55: an edge exists only between pixels that are both in the index).
    Bands = contiguous groups of the distinct X values (balanced in rows); ghosts = the pixels of the rows X_first - 1 and
    X_last + 1 where those rows exist; local order = ascending global pixel id (so that neighbour order and colour classes
    of owned pixels are those of the full map); rows on the two sides of a cut are listed by ascending global id on both
    ranks, which is what pairs the halo buffers."""
    X, Y = np.asarray(X).astype(np.int64), np.asarray(Y).astype(np.int64)
    rows = np.unique(X)
    assert 1 <= world <= rows.size, "need at least one row per rank"
    bounds = [band_rows(rows.size, world, r) for r in range(world)]
    gids = np.arange(X.size, dtype=np.int64)

    def owned_of(r):
        a, b = bounds[r]
        return gids[(X >= rows[a]) & (X <= rows[b - 1]) & np.isin(X, rows[a:b])]

    a, b = bounds[rank]
    x_first, x_last = rows[a], rows[b - 1]
    has_up = rank > 0 and rows[a - 1] == x_first - 1              # the row above the cut exists and touches mine
    has_down = rank < world - 1 and rows[b] == x_last + 1
    own = np.isin(X, rows[a:b])
    ghost_up = (X == x_first - 1) if has_up else np.zeros(X.size, bool)
    ghost_down = (X == x_last + 1) if has_down else np.zeros(X.size, bool)
    stored = gids[own | ghost_up | ghost_down]                       # ascending global id
    loc = {int(g): i for i, g in enumerate(stored)}
    ids = lambda m: np.array([loc[int(g)] for g in gids[m]], dtype=np.int32)      # noqa: E731
    owned = own[stored].astype(np.uint8)
    return BandPlan(int(rows.size), int(Y.max() - Y.min() + 1), world, rank, int(x_first), int(x_last) + 1,
                    int(x_first) - (1 if has_up else 0), int(x_last) + 1 + (1 if has_down else 0), stored, owned,
                    ids(own & (X == x_first)) if has_up else None, ids(own & (X == x_last)) if has_down else None,
                    ids(ghost_up) if has_up else None, ids(ghost_down) if has_down else None,
                    [owned_of(r) for r in range(world)])


def exchange_halo(plan: BandPlan, gather: Callable, scatter: Callable, dist, make_buffer: Callable):
    """One halo exchange.  ``gather(name) -> tensor (W, D)`` reads the rows ``plan.<name>`` of the local
    state (name in send_up / send_down), ``scatter(name, tensor)`` writes the rows ``plan.<name>`` (recv_up /
    recv_down), ``make_buffer() -> tensor (W, D)`` allocates a receive buffer, ``dist`` is torch.distributed
    (nccl on the GPUs, gloo in the CPU tests)."""
    if plan.world == 1:
        return
    ops, recvs = [], []
    if plan.send_up is not None:
        ops.append(dist.P2POp(dist.isend, gather("send_up"), plan.rank - 1))
        buf = _buffer(make_buffer, "recv_up")
        ops.append(dist.P2POp(dist.irecv, buf, plan.rank - 1))
        recvs.append(("recv_up", buf))
    if plan.send_down is not None:
        ops.append(dist.P2POp(dist.isend, gather("send_down"), plan.rank + 1))
        buf = _buffer(make_buffer, "recv_down")
        ops.append(dist.P2POp(dist.irecv, buf, plan.rank + 1))
        recvs.append(("recv_down", buf))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    for name, buf in recvs:
        scatter(name, buf)


def _buffer(make_buffer, name):
    """Receive buffer for the ghost row ``name``; ``make_buffer`` may take the name (rows of different lengths on
    non-rectangular maps) or nothing (full rectangle: every row has W pixels)."""
    try:
        return make_buffer(name)
    except TypeError:
        return make_buffer()


def local_posterior(post, plan: BandPlan, X=None, Y=None):
    """Sub-problem of one band: observations sliced, prior graph rebuilt on the stored rows (global
    coordinates, so colours and neighbour order are those of the full map), colour classes restricted to
    the owned pixels."""
    gid = plan.local_global_id
    lik = post.likelihood
    n_loc = gid.size
    lik_loc = M.MixingModelsLikelihood(lik.forward_map, post.D, post.L, n_loc, lik.y[gid], lik.sigma_a[gid],
                                       lik.sigma_m[gid], lik.omega[gid], lik.log_fm1[gid], lik.log_fp1[gid])
    ps = post.prior_spatial
    if X is None:
        X, Y = gid // plan.W, gid % plan.W
    else:
        X, Y = np.asarray(X)[gid], np.asarray(Y)[gid]
    sp = None
    sites: Dict[int, np.ndarray]
    if ps is not None:
        sp = M.L22DiscreteGradSpatialPrior(post.D, n_loc, X, Y, None, ps.use_next_nearest_neighbors, ps.initial_weights)
        sp.weights = np.asarray(ps.weights, dtype=np.float64).copy()
        sites = {k: v[plan.owned[v] > 0] for k, v in sp.dict_sites.items()}
    else:
        sites = {0: np.where(plan.owned > 0)[0]}
    return M.Posterior(post.D, post.L, n_loc, lik_loc, sp, post.prior_indicator, dict_sites=sites), sites


class ShardedDevicePosterior:
    """Same surface as DevicePosterior for B200Sampler, over ``torch.distributed`` ranks (one per GPU)."""

    def __init__(self, posterior, H: int, W: int, device: int = 0, mlp_mode: int = MLP_FP32, X=None, Y=None):
        """``X``, ``Y``: index coordinates of the pixels (any order, the map need not be a full rectangle); omitted for a
        full H x W map in row-major (X slow) order."""
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.N, self.D, self.L = posterior.N, posterior.D, posterior.L
        if X is None:
            assert posterior.N == H * W, "without X, Y the map must be a full H x W rectangle in row-major (X slow) order"
            self.plan = make_band_plan(H, W, self.world, self.rank)
        else:
            assert len(X) == len(Y) == posterior.N
            self.plan = make_band_plan_xy(X, Y, self.world, self.rank)
        if self.world == 1:
            self.local = DevicePosterior.from_posterior(posterior, device=device, mlp_mode=mlp_mode)
        else:
            post_loc, sites = local_posterior(posterior, self.plan, X, Y)
            self.local = DevicePosterior.from_posterior(post_loc, device=device, global_id=self.plan.local_global_id,
                                                        dict_sites=sites, mlp_mode=mlp_mode)
        self.n_sites, self.site_keys = self.local.n_sites, self.local.site_keys
        self.site_pixels = self.local.site_pixels
        self.dev = torch.device("cuda", device)
        # tensors handed to the collectives live on the GPU under NCCL; under gloo (several ranks sharing ONE GPU: the
        # single-device sharding test, tests/test_gpu_parity.py) they are staged through the host
        self.coll_dev = torch.device("cpu") if (self.world > 1 and dist.get_backend() == "gloo") else self.dev
        self._ids = {}
        if self.world > 1:
            for name in ("send_up", "send_down", "recv_up", "recv_down"):
                a = getattr(self.plan, name)
                if a is not None:
                    self._ids[name] = torch.as_tensor(a, dtype=torch.int32, device=self.dev)
        self._owned = self.plan.owned if self.world > 1 else None

    # ---- halo
    def halo_exchange(self):
        if self.world == 1:
            return
        torch = self.torch

        def gather(name):
            t = self._ids[name]
            out = torch.empty((t.numel(), self.D), dtype=torch.float32, device=self.dev)
            self.local.gather_theta_rows(t.data_ptr(), t.numel(), out.data_ptr())
            return out.to(self.coll_dev)                   # (library and torch share the legacy default stream: ordered)

        def scatter(name, buf):
            t = self._ids[name]
            buf = buf.to(self.dev)
            self.local.scatter_theta_rows(t.data_ptr(), t.numel(), buf.data_ptr())

        def make_buffer(name):
            return torch.empty((self._ids[name].numel(), self.D), dtype=torch.float32, device=self.coll_dev)

        exchange_halo(self.plan, gather, scatter, self.dist, make_buffer)

    # ---- DevicePosterior surface used by B200Sampler
    def compute_all(self, Theta=None):
        if Theta is not None:
            assert np.sum(np.isnan(Theta)) == 0
            self.local.set_state(np.ascontiguousarray(Theta[self.plan.local_global_id]) if self.world > 1 else Theta)
        self.local.refresh()
        return self.get_current()

    def init_v_from_grad(self):
        self.local.init_v_from_grad()

    def begin_iteration(self):
        self.local.begin_iteration()

    def pmala_site(self, site, *a, **k):
        self.local.pmala_site(site, *a, **k)
        self.halo_exchange()

    def mtm_site(self, site, *a, **k):
        self.local.mtm_site(site, *a, **k)
        self.halo_exchange()

    def mtm_finish(self, alpha):
        self.local.mtm_finish(alpha)

    def reduce_step_stats(self):
        out = self.local.reduce_step_stats(self._owned)
        return self._allreduce(out)

    def objective_terms(self):
        d = self.local.objective_terms(self._owned)
        if self.world == 1:
            return d
        D = self.D
        vec = np.concatenate([[d["nll"]], d.get("nl_prior_spatial", np.zeros(D)), d.get("nl_prior_indicator", np.zeros(D)),
                              [d["objective"]]])
        return self.local._dict_objective(self._allreduce(vec))

    # ---- optimisation mode: same steps as DevicePosterior, with the halo exchange and the global reductions in between
    def optim_pmala_step(self, eps, eta, alpha, seed, t, z=None):
        loc, lib = self.local, self.local._lib
        from . import _lib as L
        import ctypes as C
        obj_cur = self.objective_terms()["objective"]
        L.check(lib.bt_optim_begin(loc._ctx), "bt_optim_begin")
        z = L.as_f32(z if z is None or self.world == 1 else np.asarray(z)[self.plan.local_global_id])
        accept = False
        for with_noise in (0, 1):
            L.check(lib.bt_optim_pmala_propose(loc._ctx, eps, eta, with_noise, seed, t, L.dptr(z, C.c_float)),
                    "bt_optim_pmala_propose")
            self.halo_exchange()        # ghost rows see an incomplete stencil: take the owner's values
            loc.refresh()
            if self.objective_terms()["objective"] < obj_cur:
                accept = True
                break
        L.check(lib.bt_optim_update_v(loc._ctx, alpha), "bt_optim_update_v")
        L.check(lib.bt_optim_end(loc._ctx, int(accept)), "bt_optim_end")
        return accept, (1 if accept else 0)

    def mtm_optim_site(self, site, K, seed, t, cand=None):
        colsum = self._allreduce(self.local.mtm_optim_weights(site, K, seed, t, cand, self._owned))
        self.local.mtm_optim_select(site, K, int(np.argmin(colsum)))
        self.halo_exchange()

    def model_check_snapshot_clppd(self):
        self.local.model_check_snapshot_clppd()

    def model_check_update_pvalues(self, seed, t, za=None, zm=None):
        self.local.model_check_update_pvalues(seed, t, za, zm)

    def spatial_prior_unweighted(self):
        return self._allreduce(self.local.spatial_prior_unweighted(self._owned))

    def update_spatial_weights(self, ps, pi):
        self.local.update_spatial_weights(ps, pi)

    # the chain store is per band; summaries are gathered like any other per-pixel array
    def chain_reserve(self, capacity):
        self.local.chain_reserve(capacity)

    def chain_push(self):
        self.local.chain_push()

    def chain_length(self):
        return self.local.chain_length()

    def chain_get(self, first=0, count=None):
        loc = self.local.chain_get(first, count)                         # (T, N_loc, D)
        if self.world == 1:
            return loc
        g = self._gather_owned(np.ascontiguousarray(np.swapaxes(loc, 0, 1)))
        return None if g is None else np.swapaxes(g, 0, 1)

    def chain_stats(self, first=0, count=None, percentiles=(), want_var=False):
        st = self.local.chain_stats(first, count, percentiles, want_var)
        if self.world == 1:
            return st
        for k in ("mean", "var"):
            if st[k] is not None:
                st[k] = self._gather_owned(st[k])
        for k in ("lo", "hi"):
            g = self._gather_owned(np.ascontiguousarray(np.swapaxes(st[k], 0, 1)))
            st[k] = None if g is None else np.swapaxes(g, 0, 1)
        return st

    def _allreduce(self, arr):
        if self.world == 1:
            return arr
        t = self.torch.as_tensor(np.asarray(arr, dtype=np.float64), device=self.coll_dev)
        self.dist.all_reduce(t)
        return t.cpu().numpy()

    def broadcast_ints(self, values):
        """The root's integers on every rank (kernel choice, saver decisions, Philox key)."""
        if self.world == 1:
            return list(values)
        t = self.torch.as_tensor(np.asarray(values, dtype=np.int64), device=self.coll_dev)
        self.dist.broadcast(t, src=0)
        return [int(x) for x in t.cpu().numpy()]

    def reserve_scratch(self, max_rows, max_rows_with_jacobian=0):
        self.local.reserve_scratch(max_rows, max_rows_with_jacobian)

    def _gather_owned(self, local_arr, everywhere=False, f32_exact=False):
        """(N_loc, ...) -> (N, ...) on the root rank (None elsewhere); on every rank if ``everywhere``.  Only the root
        talks to the saver, so by default nothing is shipped to, or copied off the device on, the other ranks."""
        if self.world == 1:
            return local_arr
        torch = self.torch
        own = local_arr[self.plan.owned > 0]
        # ``f32_exact``: a float64 array whose values are float32 numbers (Theta, v: float32 on the device) travels as float32
        as_f32 = f32_exact and own.dtype == np.float64
        if as_f32:
            own = own.astype(np.float32)
        by_rank = self.plan.owned_gid_by_rank
        if by_rank is not None:
            sizes = [int(g.size) for g in by_rank]
        else:
            sizes = [(band_rows(self.plan.H, self.world, r)[1] - band_rows(self.plan.H, self.world, r)[0]) * self.plan.W
                     for r in range(self.world)]
        # bands differ in size: gather equal-sized (zero-padded) pieces and trim, which every backend supports
        smax = max(sizes)
        mine = torch.zeros((smax,) + tuple(own.shape[1:]), dtype=torch.as_tensor(own).dtype, device=self.coll_dev)
        mine[:own.shape[0]] = torch.as_tensor(np.ascontiguousarray(own), device=self.coll_dev)
        if everywhere:
            outs = [torch.empty_like(mine) for _ in sizes]
            self.dist.all_gather(outs, mine)
        else:
            outs = [torch.empty_like(mine) for _ in sizes] if self.rank == 0 else None
            self.dist.gather(mine, outs, dst=0)
            if self.rank != 0:
                return None
        cat = torch.cat([o[:s] for o, s in zip(outs, sizes)], 0).cpu().numpy()
        if as_f32:
            cat = cat.astype(np.float64)
        if by_rank is None:
            return cat                                     # full rectangle: bands in rank order = global pixel order
        out = np.empty_like(cat)
        out[np.concatenate(by_rank)] = cat                 # any pixel order: back to the global ids
        return out

    def get_log_f(self):
        return self._gather_owned(self.local.get_log_f())

    def get_state(self, want_v=True, want_j=True):
        th, v, j = self.local.get_state(want_v, want_j)
        return (self._gather_owned(th, f32_exact=True), self._gather_owned(v, f32_exact=True) if want_v else None,
                self._gather_owned(j) if want_j else None)

    def get_current(self, with_forward_map=True):
        cur = self.local.get_current(with_forward_map)
        if self.world == 1:
            return cur
        out = {}
        for k, val in cur.items():
            if isinstance(val, dict):
                out[k] = {kk: (self._gather_owned(vv) if isinstance(vv, np.ndarray) and vv.ndim >= 1 else vv)
                          for kk, vv in val.items()}
            elif isinstance(val, np.ndarray) and val.ndim >= 1:
                out[k] = self._gather_owned(val)
            else:
                out[k] = val
        if self.rank != 0:
            return {}
        for k in ("objective_chromatic", "objective_global"):
            out[k] = float(out["objective_pix_" + k.split("_")[1]].sum())
        out["nll_utils"] = {"nll": np.float64(out["nll_pix"].sum())}
        return out

    def step_stats(self):
        a, lp = self.local.step_stats()
        return self._gather_owned(a), self._gather_owned(lp)

    def model_check_reset(self):
        self.local.model_check_reset()

    def model_check_update(self, seed, t, za=None, zm=None):
        self.local.model_check_update(seed, t, za, zm)

    def model_check_get(self, finalize=True):
        return tuple(self._gather_owned(a) for a in self.local.model_check_get(finalize))

    def kernel_profile(self, enable=True, reset=False):
        return self.local.kernel_profile(enable, reset)

    def synchronize(self):
        self.local.synchronize()

    def kernel_launches(self):
        return self.local.kernel_launches()

    def close(self):
        self.local.close()
