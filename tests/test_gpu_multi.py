"""GPU (>= 2 devices): the row-band sharded chain is bit-identical to the single-GPU chain."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world, out, mode, one_gpu=False):
    script = os.path.join(ROOT, "scripts", "shard_check.py")
    env = dict(os.environ)
    if one_gpu:
        env["BT_SHARD_ONE_GPU"] = "1"
    if world == 1:
        cmd = [sys.executable, script, out, mode]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), script, out, mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]


def _compare(a, b, mode, what):
    for k in ("theta", "v", "j"):
        assert np.array_equal(a[k], b[k]), f"{k} differs between 1 and {what} ({mode})"
    assert np.allclose(a["stats"], b["stats"], rtol=1e-12, atol=1e-12)
    assert np.isclose(a["objective"], b["objective"], rtol=1e-9)
    if mode == "next":
        # chain store and its summaries: per-pixel work, bit-identical; optimisation mode: the global argmin / objective
        # comparison go through an all-reduce (same decisions, same states); after the regularisation-weight update the
        # weights agree to double rounding of the reduced potential, the states to float32 rounding of those weights
        for k in ("chain", "chain_mean", "chain_var", "chain_lo", "chain_hi", "theta_opt"):
            assert np.array_equal(a[k], b[k]), f"{k} differs between 1 and {what}"
        assert np.array_equal(a["opt_stats"], b["opt_stats"]) and np.isclose(a["opt_objective"], b["opt_objective"], rtol=1e-9)
        assert np.allclose(a["taus"], b["taus"], rtol=1e-10) and np.abs(a["taus"][-1] - a["taus"][0]).max() > 0
        assert np.allclose(a["theta_tau"], b["theta_tau"], rtol=1e-5, atol=1e-6)
    if mode.startswith("sample"):
        # B200Sampler.sample + Saver on the sharded device: the root's results are those of the unsharded run (the weight update
        # goes through an all-reduce: weights to double rounding, states to float32 rounding of those weights)
        for k in ("s_type", "s_seed"):
            assert np.array_equal(a[k], b[k]), k
        assert np.allclose(a["s_tau"], b["s_tau"], rtol=1e-10)
        for k in ("s_theta", "s_v", "s_logf", "s_mmse", "s_final", "s_clppd"):
            assert np.allclose(a[k], b[k], rtol=2e-5, atol=2e-6), f"{k} differs between 1 and {what}"
        assert np.allclose(a["s_obj"], b["s_obj"], rtol=1e-6) and np.allclose(a["s_acc"], b["s_acc"], atol=2.0 / 800)


@pytest.mark.parametrize("mode", ["fp32", "bf16", "next", "sample", "samplebf16", "mask"])
def test_sharded_chain_on_one_gpu_is_bit_identical(tmp_path, built_lib, mode):
    """Runs on a single-GPU machine: 3 ranks share GPU 0, each owns a row band (own context, ghost rows), the halo
    exchange and the global reductions go through gloo on the host -- the band plan, exchange and reduction code is the one
    used over NCCL.  The sharded chain must be the unsharded chain bit for bit."""
    ref = str(tmp_path / "w1.npz")
    _run(1, ref, mode)
    out = str(tmp_path / "w3.npz")
    _run(3, out, mode, one_gpu=True)
    _compare(np.load(ref), np.load(out), mode, "3 bands on one GPU")


@pytest.mark.parametrize("mode", ["fp32", "fp32n8", "bf16", "next", "sample", "mask"])
def test_sharded_chain_is_bit_identical(tmp_path, built_lib, mode):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    ref = str(tmp_path / "w1.npz")
    _run(1, ref, mode)
    a = np.load(ref)
    for world in [w for w in (2, 3, 4, 8) if w <= n][:2]:
        out = str(tmp_path / f"w{world}.npz")
        _run(world, out, mode)
        _compare(a, np.load(out), mode, f"{world} GPUs")
