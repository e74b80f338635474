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
from beetroots_b200.parallel import band_rows, exchange_halo, local_posterior, make_band_plan


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


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


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
    mp.spawn(_worker, args=(world, port, 7, 4, 3, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert np.load(tmp_path / f"ok{r}.npy")[0]
