#!/usr/bin/env python
"""Benchmark of the beetroots sampler step (PMALA + MTM) on B200 -- see DESIGN.md §measurement.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one sampler iteration over the whole map.  Timed steps alternate MTM / PMALA
deterministically (the expectation of the reference's selection_probas = (0.5, 0.5) draw,
sampler/my_sampler.py:430-446), so K should be even.  Workload: config c4 of BASELINE.json,
a 1024x1024 synthetic map, D=4, L=10, K_mtm=50, row-band sharded over N GPUs (strong scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pixel MCMC iterations/sec (PMALA+MTM) on 1M-pixel map"
UNIT = "iterations/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=1024, help="map is size x size pixels")
    ap.add_argument("--k-mtm", type=int, default=50)
    ap.add_argument("--mlp", default="auto", choices=["auto", "fp32", "bf16"])
    ap.add_argument("--cpu-size", type=int, default=64, help="side of the bounded CPU-baseline sample map")
    ap.add_argument("--no-extras", action="store_true", help="skip the K_mtm = 10 / 20 runs and the sample() end-to-end")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.th = index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=6)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _port_baseline(size, k_mtm, steps=1, warmup=0):
    """Fallback when no reference tree is on the box: the float64 numpy oracle port of the step (oracle/sampler.py)."""
    import torch
    from beetroots_b200 import modelling as M
    from beetroots_b200.synthetic import make_problem
    from oracle import posterior as OP
    from oracle.sampler import OracleSampler
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    prob = make_problem(size, size, lambda fm, th: OP.forward_map_compute_all(fm, th, False)["log_f_Theta"], L=10,
                        k_mtm=k_mtm)
    post = prob.posterior
    nbr = M.build_neighbor_table(post.prior_spatial.list_edges, post.N, 4)
    orc = OracleSampler(post, nbr, 1e-4, 1e-5, 0.99, k_mtm)
    orc.init(prob.theta_init)
    rng = np.random.default_rng(0)

    def one_pair():
        noise = {}
        for s, idx in post.dict_sites.items():
            nb = nbr[idx]
            pick = nb[np.arange(idx.size), rng.integers(0, 2, idx.size)]
            pick = np.where(pick >= 0, pick, nb[:, 0])
            c = orc.current["Theta"][pick][:, None, :] + rng.standard_normal((idx.size, k_mtm, post.D)) / np.sqrt(
                post.prior_spatial.weights)
            noise[s] = {"cand": c, "u_sel": rng.uniform(size=idx.size), "u_acc": rng.uniform(size=idx.size)}
        orc.mtm_step(noise)
        noise = {s: {"z": rng.standard_normal((idx.size, post.D)), "u": rng.uniform(size=idx.size)}
                 for s, idx in post.dict_sites.items()}
        orc.pmala_step(noise)

    for _ in range(warmup):
        one_pair()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_pair()
    return (time.perf_counter() - t0) / steps, cores


def cpu_baseline(size, k_mtm, pairs=1, warmup=1):
    """The reference's own CPU implementation of the step on the host cores, on a BOUNDED sample of the workload: a
    size x size map of the same recipe (D=4, L=10, K_mtm), ``pairs`` x (one MTM + one PMALA iteration) =
    ``MySampler.generate_new_sample_mtm`` / ``_pmala_rmsprop`` (sampler/my_sampler.py:966-1202, 718-906) of the UNMODIFIED
    reference package (baseline/_ref, oracle/install_reference.py) run through oracle/refrun.py with every thread pool
    pinned to all host cores; the numpy oracle port (kind "port") only if no reference tree is on this box.  The c4 figure
    is the sample's iterations/s scaled LINEARLY by the pixel count -- an upper bound for the reference at 1M pixels, whose
    prior loops are O(N * edges) (never runnable at c4)."""
    from oracle import refshim
    sec = cores = None
    kind = "port"
    if refshim.reference_available():
        try:
            from oracle import posterior as OP
            from oracle import refrun
            from beetroots_b200.synthetic import make_problem
            cores = refrun.set_host_threads()
            prob = make_problem(size, size, lambda fm, th: OP.forward_map_compute_all(fm, th, False)["log_f_Theta"], L=10,
                                k_mtm=k_mtm)
            if warmup:       # numba JIT + torch warm-up on a tiny map of the same recipe
                small = make_problem(8, 8, lambda fm, th: OP.forward_map_compute_all(fm, th, False)["log_f_Theta"], L=10,
                                     k_mtm=k_mtm)
                refrun.time_reference_kernels(small, pairs=1, warmup=1)
            sec = refrun.time_reference_kernels(prob, pairs=pairs, warmup=0)
            kind = "reference"
        except Exception as e:                    # pragma: no cover - stated in the line
            sys.stderr.write(f"[bench] reference arm failed ({type(e).__name__}: {e}); timing the oracle port instead\n")
    if sec is None:
        sec, cores = _port_baseline(size, k_mtm, steps=pairs, warmup=1 if warmup else 0)
    its_small = 2.0 / sec
    scaled = its_small * (size * size) / (1024.0 * 1024.0)
    what = ("unmodified reference MySampler kernels (baseline/_ref)" if kind == "reference"
            else "float64 numpy oracle port (no reference tree on this box)")
    return {"value": scaled, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{size}x{size} map (K_mtm={k_mtm}, L=10, D=4), {pairs} x (1 MTM + 1 PMALA) iterations of the {what} in "
                      f"{sec * pairs:.1f} s = {its_small:.4f} it/s at {size * size} pixels; value = that scaled linearly to "
                      f"1,048,576 pixels (extrapolated: the reference's prior loops are O(N*edges), c4 itself is not runnable)",
            "its_at_sample_size": its_small, "extrapolated": True, "seconds": sec * pairs}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pairs = max(1, min(args.steps // 2, 3))            # bounded: ~9 s per pair at 64 x 64 on 16 cores
    cb = cpu_baseline(args.cpu_size, args.k_mtm, pairs=pairs, warmup=1 if args.warmup else 0)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 / cb["value"] if cb["value"] > 0 else None, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"c4: 1024x1024 map, D=4, L=10, K_mtm={args.k_mtm}, p=(0.5,0.5); timed on a bounded "
                                   f"{args.cpu_size}x{args.cpu_size} sample of the same recipe ({2 * pairs} iterations), "
                                   f"scaled linearly to 1M pixels", "same_config": False},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "its_at_sample_size", "extrapolated")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from beetroots_b200 import MLP_BF16_TC, MLP_FP32, B200Sampler, MySamplerParams
    from beetroots_b200.device import gpu_log_f
    from beetroots_b200.parallel import ShardedDevicePosterior
    from beetroots_b200.synthetic import make_problem

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, (world, args.gpus)
    torch.cuda.set_device(local)
    if world > 1:
        # stdout carries the ONE JSON line and nothing else: NCCL prints its banner (NCCL_DEBUG=VERSION/INFO) on fd 1 when the
        # first communicator is created, so fd 1 points at stderr until that has happened
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            dist.broadcast(warm, src=0)
            dist.gather(warm, [torch.zeros(1, device="cuda") for _ in range(world)] if rank == 0 else None, dst=0)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    H = W = args.size
    K = args.k_mtm
    mode = {"fp32": MLP_FP32, "bf16": MLP_BF16_TC}.get(args.mlp)
    if mode is None:
        mode = MLP_BF16_TC
    prob = make_problem(H, W, lambda fm, th: gpu_log_f(fm, th, device=local), L=10, k_mtm=K)
    post = prob.posterior
    try:
        sdev = ShardedDevicePosterior(post, H, W, device=local, mlp_mode=mode)
    except Exception as e:                       # tensor-core kernel unavailable -> exact fp32 kernel, stated in dtype
        if mode == MLP_FP32:
            raise
        if rank == 0:
            sys.stderr.write(f"[bench] bf16 tensor-core forward map unavailable ({e}); using the fp32 kernel\n")
        mode = MLP_FP32
        sdev = ShardedDevicePosterior(post, H, W, device=local, mlp_mode=mode)
    smp = B200Sampler(MySamplerParams(1e-4, 1e-5, 0.99, [0.5, 0.5], K, True, False), post.D, post.L, post.N,
                      rng=np.random.default_rng(42), device=local, mlp_mode=mode)
    smp.attach(post, sdev)
    smp.philox_seed = 20231019
    sdev.compute_all(prob.theta_init)
    sdev.init_v_from_grad()
    net = post.likelihood.forward_map.network
    flops_row = 2.0 * net.macs_per_row()

    def step(t):
        if t % 2 == 0:
            return smp.generate_new_sample_mtm(t, post)
        return smp.generate_new_sample_pmala_rmsprop(t, post)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_max = max(p.size for p in sdev.local.site_pixels)
    sdev.reserve_scratch(n_max * (K + 1), sdev.local.N)        # every scratch buffer of a step exists before anything is timed

    def timed_steps(n_steps, t0_index):
        """n_steps sampler iterations bracketed by barrier + synchronize: (device ms by CUDA events, max over ranks)."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        w0 = time.perf_counter()
        tt = t0_index
        for _ in range(n_steps):
            step(tt)
            tt += 1
        e1.record()
        barrier()
        wall_ = time.perf_counter() - w0
        ms_ = e0.elapsed_time(e1)
        ms_ = max(ms_, wall_ * 1000.0) if ms_ <= 0 else ms_
        if world > 1:
            x = torch.tensor([ms_], device="cuda", dtype=torch.float64)
            dist.all_reduce(x, op=dist.ReduceOp.MAX)
            ms_ = float(x.item())
        return ms_, wall_, tt

    t = 0
    for _ in range(args.warmup):
        step(t)
        t += 1
    # ---- device-resident timing
    sdev.local.mlp_profile(enable=True, reset=True)
    sdev.local.kernel_profile(enable=True, reset=True)
    launches0 = sdev.local.kernel_launches()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms, wall, t = timed_steps(args.steps, t)
    clk = clocks.stop() if rank == 0 else None
    mlp_ms, mlp_rows, mlp_launches = sdev.local.mlp_profile(enable=False, reset=True)
    kprof = sdev.local.kernel_profile(enable=False, reset=True)
    launches = sdev.local.kernel_launches() - launches0
    value = args.steps / (ms / 1000.0) * (H * W) / (1024.0 * 1024.0)

    # ---- end to end: host Theta in, host Theta out, every step (pinned buffers)
    n_loc = sdev.local.N
    pin_in = torch.empty((n_loc, post.D), dtype=torch.float32).pin_memory()
    pin_out = torch.empty((n_loc, post.D), dtype=torch.float32).pin_memory()
    sdev.local.get_theta_f32_ptr(pin_in.data_ptr())
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        sdev.local.set_theta_f32_ptr(pin_in.data_ptr())          # H2D of the step's input state
        acc, lpa = step(t)
        t += 1
        sdev.local.get_theta_f32_ptr(pin_out.data_ptr())         # D2H of the result (synchronises)
        pin_in.copy_(pin_out)
    barrier()
    e2e_s = time.perf_counter() - w0
    if world > 1:
        tt = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_val = args.steps / e2e_s * (H * W) / (1024.0 * 1024.0)

    # ---- end to end through the public API: B200Sampler.sample(posterior, saver, ...) with a Saver-protocol saver in
    # host memory (the reference's call, sampler/my_sampler.py:269-283): state pulled to the host, converted to float64,
    # handed to the saver every freq_save iterations; model check and objective terms as in the reference loop
    e2e_sample = None
    extras = {}
    if not args.no_extras:
        from beetroots_b200.saver import ChainSaver

        class DiscardingSaver(ChainSaver):
            """Full Saver protocol (batches filled in host memory like MySaver's); a finished batch is dropped instead of
            being appended to a file, so the number measures the sampler's hand-over, not a disk."""

            def save_to_file(self):
                self._pool = self.memory
                self.memory = dict()

        e2e_sample = {}
        n_it = 20            # (a multiple of both freq_save values: MySaver sizes its last batch (T_MC - t + 1) // freq_save)
        for freq in (1, 10):
            saver = DiscardingSaver(post.N, post.D, post.D, post.L, results_path="", batch_size=10 // freq, freq_save=freq,
                                    save_forward_map_evals=False)
            smp2 = B200Sampler(MySamplerParams(1e-4, 1e-5, 0.99, [0.5, 0.5], K, True, False), post.D, post.L, post.N,
                               rng=np.random.default_rng(43), device=local, mlp_mode=mode)
            smp2.attach(post, sdev)
            theta0 = prob.theta_init
            barrier()
            w0 = time.perf_counter()
            smp2.sample(post, saver, n_it, Theta_0=theta0, disable_progress_bar=True, T_BI=0)
            barrier()
            dt = time.perf_counter() - w0
            if world > 1:
                x = torch.tensor([dt], device="cuda", dtype=torch.float64)
                dist.all_reduce(x, op=dist.ReduceOp.MAX)
                dt = float(x.item())
            e2e_sample[f"freq_save_{freq}"] = {"value": n_it / dt * (H * W) / (1024.0 * 1024.0), "unit": UNIT, "iterations": n_it,
                                               "seconds": dt}
        e2e_sample["note"] = ("B200Sampler.sample() incl. the initial compute_all, kernel-type draws, objective terms, model check, "
                              "float64 hand-over of Theta and v to an in-memory Saver (ChainSaver, save_forward_map_evals=False, "
                              "batches of 10 iterations) and the final state pull / save_additional; about 4/5 of the per-iterate "
                              "hand-over time is the saver's own host copies (33 MB float64 arrays at c4), as in the reference's MySaver")
        # ---- the reference's own K_mtm values at N >= 64^2 (nn_N64_fixed_angle.yaml:97): device-resident, same timing rules
        sdev.compute_all(prob.theta_init)
        sdev.init_v_from_grad()
        for k_extra in (10, 20):
            if k_extra == K:
                continue
            smp.k_mtm = k_extra
            for _ in range(2):
                step(t)
                t += 1
            n_e = 6
            ms_e, _, t = timed_steps(n_e, t)
            extras[f"k_mtm_{k_extra}"] = {"value": n_e / (ms_e / 1000.0) * (H * W) / (1024.0 * 1024.0), "unit": UNIT, "steps": n_e,
                                          "warmup": 2, "ms_per_step": ms_e / n_e}
        smp.k_mtm = K

    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_hbm = peaks.get("hbm_gbs", 6500.0)
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PF sustained"
    achieved = (mlp_rows * flops_row / (mlp_ms / 1000.0)) / 1e12 if mlp_ms > 0 else None
    variant = sdev.local.mlp_kernel_variant()
    traffic, traffic_src = None, None
    try:        # DRAM bytes per forward row from the committed ncu --set full capture of the same kernels (not measured here)
        name = {5: "r02_forward_map_traffic.json", 4: "r01e_forward_map_traffic.json"}.get(variant, "r01_mlp_tc_traffic.json")
        tj = json.load(open(os.path.join(ROOT, "profiles", name)))
        if mode == MLP_BF16_TC and mlp_launches:
            traffic = tj["dram_bytes_per_row"] * mlp_rows / mlp_launches
            traffic_src = f"static: profiles/{name} (ncu dram__bytes per forward row) x rows per launch of this run"
    except Exception:
        pass
    hbm = {}
    for name, (k_ms, k_bytes, k_n) in kprof.items():
        if k_n and k_ms > 0:
            gbs = k_bytes / (k_ms / 1000.0) / 1e9
            hbm[name] = {"achieved": gbs, "peak": peak_hbm, "unit": "GB/s", "frac": gbs / peak_hbm, "kernel_ms": k_ms,
                         "launches": k_n, "algorithmic_bytes_per_launch": k_bytes / k_n, "share_of_step": k_ms / ms}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16" if mode == MLP_BF16_TC else "f32", "data": "synthetic",
        "config": {"workload": f"c4: {H}x{W} map, D=4, L=10, K_mtm={K}, timed steps alternate MTM/PMALA (= p (0.5,0.5))",
                   "parallelism": f"row bands x{world}, 1-row halo per colour sweep",
                   "l2": "inputs larger than L2: candidate rows + observation constants ~2 GB per MTM sweep",
                   "mlp": "tcgen05 bf16" if mode == MLP_BF16_TC else "CUDA-core fp32",
                   "wall_s_timed": wall},
        "clocks": clk,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(n_loc * post.D * 4),
                "d2h_bytes_per_step": int(n_loc * post.D * 4 + 16)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": (achieved / peak_tf) if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_flops_per_launch": flops_row * mlp_rows / mlp_launches if mlp_launches else None,
                     "kernel": {5: "forward-map MLP (mlp_pre_mma_kernel + mlp_tc4_kernel)", 4: "forward-map MLP (pre-pass + mlp_tc3 pair kernel)"}.get(variant, "forward-map MLP") if mode == MLP_BF16_TC else "forward-map MLP (mlp_simt_kernel)",
                     "peak_source": peak_src,
                     "rows": int(mlp_rows), "launches": int(mlp_launches), "kernel_ms": mlp_ms,
                     "share_of_step": mlp_ms / ms if ms > 0 else None,
                     "pixel_candidate_evals_per_s_per_gpu": mlp_rows / (mlp_ms / 1000.0) if mlp_ms > 0 else None},
        "roofline_hbm": hbm,
    }
    if e2e_sample is not None:
        line["e2e_sample"] = e2e_sample
    if extras:
        line["extra"] = extras
    if not args.no_cpu_baseline:
        cb = cpu_baseline(args.cpu_size, K, pairs=1, warmup=1)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "its_at_sample_size", "extrapolated")}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
