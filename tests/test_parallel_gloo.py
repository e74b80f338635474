"""CPU: the multi-GPU host logic (band plan, halo exchange, sharding-invariant graph) with
world_size-2/3 gloo process groups.  The device kernels are replaced by numpy arrays here; the
exchange code (beetroots_b200/parallel.py exchange_halo) is the one the NCCL path runs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from beetroots_b200 import modelling as M
from beetroots_b200.parallel import band_rows, exchange_halo, local_posterior, make_band_plan, make_band_plan_xy


def test_band_plan_partitions_the_map():
    H, W = 11, 5
    for world in (1, 2, 3, 4):
        seen = np.zeros(H * W, dtype=int)
        for r in range(world):
            p = make_band_plan(H, W, world, r)
            seen[p.local_global_id[p.owned > 0]] += 1
            assert (p.send_up is None) == (r == 0) and (p.send_down is None) == (r == world - 1)
            if p.recv_up is not None:
                assert np.array_equal(p.local_global_id[p.recv_up] // W, np.full(W, p.row0 - 1))
                assert np.all(p.owned[p.recv_up] == 0) and np.all(p.owned[p.send_up] == 1)
            if p.recv_down is not None:
                assert np.array_equal(p.local_global_id[p.recv_down] // W, np.full(W, p.row1))
        assert np.all(seen == 1)
    assert band_rows(10, 4, 0) == (0, 3) and band_rows(10, 4, 3) == (8, 10)


@pytest.mark.parametrize("nnn", [False, True])
def test_local_graph_is_sharding_invariant(nnn):
    """Neighbour ORDER and colours of every owned pixel equal those of the full map (needed for
    bit-identical Philox chains across GPU counts)."""
    H, W, D, L = 9, 6, 4, 3
    N = H * W
    X, Y = np.repeat(np.arange(H), W), np.tile(np.arange(W), H)
    sp = M.L22DiscreteGradSpatialPrior(D, N, X, Y, None, nnn, np.ones(D))
    lik = M.MixingModelsLikelihood(None, D, L, N, np.ones((N, L)), 1.0, 0.1, 0.5, a0=np.zeros((1, L)), a1=np.ones((1, L)))
    post = M.Posterior(D, L, N, lik, sp, None)
    deg = 8 if nnn else 4
    nbr = M.build_neighbor_table(sp.list_edges, N, deg)
    colour_of = np.zeros(N, dtype=int)
    for k, v in sp.dict_sites.items():
        colour_of[v] = k
    for world in (2, 3):
        for r in range(world):
            plan = make_band_plan(H, W, world, r)
            lp, sites = local_posterior(post, plan)
            nbl = M.build_neighbor_table(lp.prior_spatial.list_edges, plan.n_local, deg)
            gid = plan.local_global_id
            for k, loc in sites.items():
                assert np.all(plan.owned[loc] == 1) and np.all(colour_of[gid[loc]] == k)
                for n in loc:
                    mine = nbl[n][nbl[n] >= 0]
                    assert np.array_equal(gid[mine], nbr[gid[n]][nbr[gid[n]] >= 0])


def test_general_band_plan_equals_the_rectangular_one_and_handles_masks():
    """make_band_plan_xy on a full rectangle reproduces make_band_plan; on a non-rectangular, shuffled pixel set (the maps of
    abstract_spatial_prior.py:13-55) every pixel is owned exactly once, the ghost rows are the neighbouring rows that exist,
    the two sides of a cut list the same pixels in the same order, and the local graph of every band keeps the neighbour
    order and colours of the full map."""
    H, W, D, L = 9, 6, 4, 3
    X, Y = np.repeat(np.arange(H), W), np.tile(np.arange(W), H)
    for world in (1, 2, 4):
        for r in range(world):
            a, b = make_band_plan(H, W, world, r), make_band_plan_xy(X, Y, world, r)
            assert np.array_equal(a.local_global_id, b.local_global_id) and np.array_equal(a.owned, b.owned)
            for k in ("send_up", "send_down", "recv_up", "recv_down"):
                assert (getattr(a, k) is None) == (getattr(b, k) is None)
                assert getattr(a, k) is None or np.array_equal(getattr(a, k), getattr(b, k))
    # a masked, shuffled map: two pixels and a whole corner removed, one empty row in the middle (bands without coupling)
    keep = np.ones(H * W, bool)
    keep[[7, 20]] = False
    keep[(X >= 7) & (Y >= 4)] = False
    keep[X == 4] = False
    perm = np.random.default_rng(3).permutation(int(keep.sum()))
    Xm, Ym = X[keep][perm], Y[keep][perm]
    N = Xm.size
    for nnn in (False, True):
        sp = M.L22DiscreteGradSpatialPrior(D, N, Xm, Ym, None, nnn, np.ones(D))
        lik = M.MixingModelsLikelihood(None, D, L, N, np.ones((N, L)), 1.0, 0.1, 0.5, a0=np.zeros((1, L)), a1=np.ones((1, L)))
        post = M.Posterior(D, L, N, lik, sp, None)
        deg = 8 if nnn else 4
        nbr = M.build_neighbor_table(sp.list_edges, N, deg)
        colour_of = np.zeros(N, dtype=int)
        for k, v in sp.dict_sites.items():
            colour_of[v] = k
        for world in (2, 3):
            plans = [make_band_plan_xy(Xm, Ym, world, r) for r in range(world)]
            seen = np.zeros(N, int)
            for r, plan in enumerate(plans):
                gid = plan.local_global_id
                seen[gid[plan.owned > 0]] += 1
                assert np.array_equal(gid[plan.owned > 0], plan.owned_gid_by_rank[r])
                if plan.send_down is not None:                       # my last row <-> the ghost row above of the next rank
                    nxt = plans[r + 1]
                    assert np.array_equal(gid[plan.send_down], nxt.local_global_id[nxt.recv_up])
                    assert np.array_equal(gid[plan.recv_down], nxt.local_global_id[nxt.send_up])
                else:
                    assert r == world - 1 or plans[r + 1].recv_up is None
                lp, sites = local_posterior(post, plan, Xm, Ym)
                nbl = M.build_neighbor_table(lp.prior_spatial.list_edges, plan.n_local, deg)
                for k, loc in sites.items():
                    assert np.all(plan.owned[loc] == 1) and np.all(colour_of[gid[loc]] == k)
                    for n in loc:
                        mine = nbl[n][nbl[n] >= 0]
                        assert np.array_equal(gid[mine], nbr[gid[n]][nbr[gid[n]] >= 0])
            assert np.all(seen == 1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _spawn(fn, make_args, nprocs, tries=3):
    """mp.spawn with a fresh rendezvous port per attempt: between _free_port() closing its probe socket and rank 0 binding the
    port another process can take it (rare, but the suite runs with -x)."""
    for attempt in range(tries):
        try:
            mp.spawn(fn, args=make_args(_free_port()), nprocs=nprocs, join=True)
            return
        except Exception as e:                                    # noqa: BLE001
            msg = str(e)
            if attempt == tries - 1 or not any(k in msg for k in ("Address already in use", "EADDRINUSE", "Connection refused",
                                                                  "connect() timed out", "Socket Timeout")):
                raise


def _worker(rank, world, port, H, W, D, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plan = make_band_plan(H, W, world, rank)
    full = np.arange(H * W * D, dtype=np.float32).reshape(H * W, D)
    state = full[plan.local_global_id].copy()
    state[plan.owned == 0] = -1.0                       # ghosts start stale
    for sweep in range(2):                              # two colour sweeps: owned rows change, ghosts must follow
        state[plan.owned > 0] += 1000.0

        def gather(name):
            return torch.from_numpy(state[getattr(plan, name)].copy())

        def scatter(name, buf):
            state[getattr(plan, name)] = buf.numpy()

        exchange_halo(plan, gather, scatter, dist, lambda: torch.empty((W, D), dtype=torch.float32))
    expect = full[plan.local_global_id] + 2000.0
    ok = np.array_equal(state, expect)
    # scalar statistics: all-reduce of per-rank sums over owned pixels
    t = torch.tensor([float(plan.owned.sum()), float(state[plan.owned > 0].sum())], dtype=torch.float64)
    dist.all_reduce(t)
    ok = ok and t[0].item() == H * W and np.isclose(t[1].item(), float((full + 2000.0).sum()))
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([ok]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_gloo(tmp_path, world):
    port = _free_port()
    _spawn(_worker, lambda p: (world, p, 7, 4, 3, str(tmp_path)), world)
    for r in range(world):
        assert np.load(tmp_path / f"ok{r}.npy")[0]


# ---- the global reductions of the "next" rows (regularisation-weight update, optimisation-mode MTM, chain gather) --------
def _reduction_worker(rank, world, port, H, W, out_dir):
    """Every rank evaluates ITS band with the float64 oracle on the LOCAL graph (local_posterior) and the sharded host
    code (ShardedDevicePosterior._allreduce / _gather_owned, the code the NCCL path runs) combines them; the result
    must equal the oracle on the full map."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from beetroots_b200.parallel import ShardedDevicePosterior
    from oracle import posterior as OP
    D, L, K = 4, 3, 5
    N = H * W
    X, Y = np.repeat(np.arange(H), W), np.tile(np.arange(W), H)
    sp = M.L22DiscreteGradSpatialPrior(D, N, X, Y, None, False, np.array([10.0, 1.0, 2.0, 4.0]))
    lik = M.MixingModelsLikelihood(None, D, L, N, np.ones((N, L)), 1.0, 0.1, 0.5, a0=np.zeros((1, L)), a1=np.ones((1, L)))
    post = M.Posterior(D, L, N, lik, sp, None)
    rng = np.random.default_rng(7)                                   # same stream on every rank
    theta = rng.standard_normal((N, D))
    cand = rng.standard_normal((N, K, D))
    chain = rng.standard_normal((6, N, D))
    nbr = M.build_neighbor_table(sp.list_edges, N, 4)
    plan = make_band_plan(H, W, world, rank)
    lp, sites = local_posterior(post, plan)
    gid = plan.local_global_id
    nbl = M.build_neighbor_table(lp.prior_spatial.list_edges, plan.n_local, 4)
    sh = object.__new__(ShardedDevicePosterior)                      # the host logic without a CUDA context
    sh.torch, sh.dist, sh.world, sh.rank, sh.plan, sh.dev = torch, dist, world, rank, plan, torch.device("cpu")
    sh.coll_dev = sh.dev
    sh.N, sh.D, sh.L = N, D, L
    th_loc = theta[gid]
    ok = True
    # (1) potential of the regularisation-weight update: 1/2 sum_n sum_{i in V_n} (theta_n - theta_i)^2 over OWNED pixels
    had_loc, _ = OP.spatial_sums(th_loc, nbl)
    g_loc = 0.5 * (had_loc * (plan.owned > 0)[:, None]).sum(axis=0)
    had, _ = OP.spatial_sums(theta, nbr)
    ok &= bool(np.allclose(sh._allreduce(g_loc), 0.5 * had.sum(axis=0), rtol=1e-12))
    # (2) optimisation-mode MTM: per-candidate column sums over the owned pixels of a colour class, argmin after all-reduce
    for s, loc in sites.items():
        had_c, _ = OP.spatial_sums(th_loc, nbl, loc, cand[gid[loc]])           # (n_loc, K, D)
        col_loc = (sp.weights * had_c).sum(-1).sum(axis=0)                      # (K,)
        glob = sp.dict_sites[s]
        had_g, _ = OP.spatial_sums(theta, nbr, glob, cand[glob])
        col = (sp.weights * had_g).sum(-1).sum(axis=0)
        red = sh._allreduce(col_loc)
        ok &= bool(np.allclose(red, col, rtol=1e-12)) and int(np.argmin(red)) == int(np.argmin(col))
    # (3) per-pixel arrays and the stored chain come back in global pixel order: on the root by default (the only rank that
    # talks to the saver), on every rank on request
    def G(a):
        g0, g = sh._gather_owned(a), sh._gather_owned(a, everywhere=True)
        assert (g0 is None) == (rank != 0)
        assert g0 is None or (g0.dtype == g.dtype and np.array_equal(g0, g))
        return g

    ok &= bool(np.array_equal(G(th_loc), theta))
    jj = np.arange(N * D, dtype=np.int32).reshape(N, D)                      # the RMSProp counters travel as int32
    got = G(jj[gid])
    ok &= got.dtype == np.int32 and bool(np.array_equal(got, jj))
    ok &= bool(np.array_equal(G(theta[gid, 0]), theta[:, 0]))                # 1-D per-pixel arrays (objectives, accept)
    ch = np.swapaxes(G(np.ascontiguousarray(np.swapaxes(chain[:, gid], 0, 1))), 0, 1)
    ok &= bool(np.array_equal(ch, chain))
    # (4) the root's kernel choice / saver decisions / Philox key reach every rank
    ok &= sh.broadcast_ints([rank * 7 + 3, 2 ** 62 + rank]) == [3, 2 ** 62]
    np.save(os.path.join(out_dir, f"red{rank}.npy"), np.array([ok]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_global_reductions_of_the_next_rows_gloo(tmp_path, world):
    port = _free_port()
    _spawn(_reduction_worker, lambda p: (world, p, 9, 5, str(tmp_path)), world)
    for r in range(world):
        assert np.load(tmp_path / f"red{r}.npy")[0]


# ---------------------------------------------------------------------------------------------------------------------
# sample() on several ranks: one writer, the root's decisions everywhere
class _ShardedStandIn:
    """The surface of parallel.ShardedDevicePosterior that B200Sampler.sample uses, without a GPU: per-pixel arrays exist
    on the root only (the other ranks get None, as from _gather_owned), scalars are the same everywhere, broadcast_ints goes
    through the process group."""

    def __init__(self, N, D, L, rank, world):
        self.N, self.D, self.L, self.rank, self.world = N, D, L, rank, world
        self.n_sites, self.site_keys = 2, [0, 1]
        self.kinds, self.theta = [], np.zeros((N, D))

    def broadcast_ints(self, values):
        t = torch.as_tensor(np.asarray(values, dtype=np.int64))
        dist.broadcast(t, src=0)
        return [int(x) for x in t.numpy()]

    def _root(self, a):
        return a if self.rank == 0 else None

    def compute_all(self, Theta=None):
        N, D = self.N, self.D
        return {"Theta": self.theta, "objective_global": 1.0, "grad": np.ones((N, D))} if self.rank == 0 else {}

    def get_current(self, with_forward_map=True):
        return self.compute_all()

    def get_state(self, want_v=True, want_j=True):
        return (self._root(self.theta.copy()), self._root(np.ones((self.N, self.D))),
                self._root(np.zeros((self.N, self.D), dtype=np.int32)))

    def get_log_f(self): return self._root(np.zeros((self.N, self.L)))
    def init_v_from_grad(self): pass
    def model_check_reset(self): pass
    def begin_iteration(self): pass
    def pmala_site(self, s, *a): self.kinds.append(1); self.theta += 0.25
    def mtm_site(self, s, *a): self.kinds.append(0); self.theta += 0.5
    def mtm_finish(self, alpha): pass
    def reduce_step_stats(self): return np.array([self.N / 2.0, -3.0 * self.N])
    def model_check_update(self, seed, t, *a): self.kinds.append(("mc", seed))
    def model_check_get(self, finalize=True):
        return self._root(np.zeros((self.N, self.L))), self._root(np.zeros((self.N, self.L))), self._root(np.zeros(self.N))
    def objective_terms(self):
        return {"nll": np.float64(2.0), "nl_prior_spatial": np.zeros(self.D), "nl_prior_indicator": np.zeros(self.D),
                "objective": np.float64(2.0)}


def _sample_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from beetroots_b200 import B200Sampler, MySamplerParams
    from beetroots_b200.saver import ChainSaver
    N, D, L, T = 6, 4, 3, 8
    X, Y = np.repeat(np.arange(N // 2), 2), np.tile(np.arange(2), N // 2)
    post = M.Posterior(D, L, N, M.MixingModelsLikelihood(None, D, L, N, np.ones((N, L)), 1.0, 0.1, 0.5, a0=np.zeros((1, L)),
                                                         a1=np.ones((1, L))),
                       M.L22DiscreteGradSpatialPrior(D, N, X, Y, None, False, np.ones(D)),
                       M.SmoothIndicatorPrior(D, N, 0.1, -np.ones(D), np.ones(D)))
    dev = _ShardedStandIn(N, D, L, rank, world)
    # every rank seeds its host generator DIFFERENTLY: kernel choice and Philox key must still be the root's
    smp = B200Sampler(MySamplerParams(1e-4, 1e-5, 0.99, [0.5, 0.5], 3, True, False), D, L, N, rng=np.random.default_rng(100 + rank))
    smp.attach(post, dev)
    saver = ChainSaver(N, D, D, L, results_path=os.path.join(out_dir, "results"), batch_size=2, freq_save=2)
    smp.sample(post, saver, T, Theta_0=np.zeros((N, D)), disable_progress_bar=True, T_BI=2)
    np.save(os.path.join(out_dir, f"kinds{rank}.npy"), np.array([k if isinstance(k, int) else -1 for k in dev.kinds]))
    np.save(os.path.join(out_dir, f"seed{rank}.npy"), np.array([smp.philox_seed] + [k[1] for k in dev.kinds if not isinstance(k, int)]))
    np.save(os.path.join(out_dir, f"wrote{rank}.npy"), np.array([len(saver.saved), len(saver.memory)]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_sample_on_several_ranks_has_one_writer(tmp_path, world):
    """B200Sampler.sample over a sharded device: only the root initialises / updates / saves (one results file, no
    concurrent writers), every rank runs the kernels the ROOT drew with the ROOT's Philox key and takes the same saver
    branches (identical call sequences: no deadlock in the collectives behind them)."""
    port = _free_port()
    _spawn(_sample_worker, lambda p: (world, p, str(tmp_path)), world)
    k0 = np.load(tmp_path / "kinds0.npy")
    for r in range(1, world):
        assert np.array_equal(np.load(tmp_path / f"kinds{r}.npy"), k0)
        assert np.array_equal(np.load(tmp_path / f"seed{r}.npy"), np.load(tmp_path / "seed0.npy"))
        assert np.array_equal(np.load(tmp_path / f"wrote{r}.npy"), [0, 0])       # never touched its saver
    assert np.load(tmp_path / "wrote0.npy")[0] > 0
    assert len([f for f in os.listdir(tmp_path / "results") if f.startswith("mc_chains")]) == 1
