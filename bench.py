#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native batched interior-point MCP solver.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    (the reference algorithm on the host cores)

Metric (BASELINE.json): converged MCP solves/sec over a batched θ.  Workload: the 2-player lane-change
trajectory game (BASELINE.json configs[2], the config north_star's target is quoted on: "≥1M converged
solves/sec of the 2-player lane-change trajectory game across 8×B200"), benchmark θ distribution and
`tol = 1e-6`, cold start, as in /root/reference/benchmark/path.jl:8,14-17,78-87.  A "step" is one batched
solve of `--batch` θ columns per GPU.  Scaling is weak (fixed per-GPU batch: 2^18 θ per GPU per step).

One JSON line on stdout (rank 0):
  value     converged solves/s, whole job, θ already resident in HBM (device entry point, CUDA events)
  e2e       same metric through the host C-ABI call (pinned host θ → H2D → solve → D2H of x,y,s,…)
  roofline  FP64-pipe roofline of the solve kernel (algorithmic banded-LU flops ÷ kernel time ÷ measured
            DFMA peak); `hbm_io` gives the I/O-floor view against MEASURED_PEAKS.json
  cpu_baseline  the C restatement of the reference (oracle/c) on the box's host cores, bounded sample
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "converged_mcp_solves_per_sec"
UNIT = "solves/s"
TOL = 1e-6           # benchmark/path.jl:8


def build_workload(name: str):
    from mcp_b200 import problems
    if name == "lane_change":
        mcp = problems.lane_change_game().mcp
        gen = lambda B, seed: problems.lane_change_thetas(B, seed=seed)
        desc = "lane_change_2p_H10 (BASELINE configs[2]; benchmark/trajectory_game_benchmark.jl), cold start, tol=1e-6"
    elif name == "readme_qp":
        mcp = problems.readme_qp()
        gen = lambda B, seed: problems.readme_qp_thetas(B, seed=seed)
        desc = "readme_qp (BASELINE configs[0]), cold start, tol=1e-6"
    elif name == "random_qp":
        mcp = problems.random_qp(100, 100)
        gen = lambda B, seed: problems.random_qp_thetas(B, seed=seed)
        desc = "random_convex_qp_100x100 (BASELINE configs[1]), cold start, tol=1e-6"
    else:
        raise SystemExit(f"unknown workload {name}")
    return mcp, gen, desc


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except OSError:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def host_threads() -> int:
    """All host threads this process may use (torchrun pins OMP_NUM_THREADS=1, so do not ask OpenMP)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_baseline(mcp, Θ, sample: int, threads: int = 0):
    """The reference algorithm (C restatement, oracle/c) on the host cores, on the first `sample` θ."""
    from oracle import c_oracle as CO
    CO.build()
    threads = threads or host_threads()
    Θs = np.asfortranarray(Θ[:, :sample])
    CO.solve_batch(mcp.ir, Θs[:, :8], tol=TOL, nthreads=threads)      # warm-up (column ordering, page-in)
    t0 = time.perf_counter()
    r = CO.solve_batch(mcp.ir, Θs, tol=TOL, nthreads=threads)
    dt = time.perf_counter() - t0
    solved = int((r.status == 0).sum())
    return {"value": solved / dt, "unit": UNIT, "cores": int(r.threads), "kind": "port",
            "sample": f"first {Θs.shape[1]} θ of the bench batch, {solved} converged, {dt:.2f} s wall, "
                      f"C restatement of src/solver.jl (Julia cannot run in this image), OpenMP over θ"}, r


def run_reference(args):
    """`--impl reference`: the reference's own algorithm on the host cores (oracle port; no Julia here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    mcp, gen, desc = build_workload(args.workload)
    sample = args.ref_sample
    Θ = gen(sample, 1)
    from oracle import c_oracle as CO
    CO.build()
    threads = host_threads()
    for _ in range(args.warmup):
        CO.solve_batch(mcp.ir, Θ[:, :min(sample, 64)], tol=TOL, nthreads=threads)
    t0 = time.perf_counter()
    solved = 0
    for _ in range(args.steps):
        r = CO.solve_batch(mcp.ir, Θ, tol=TOL, nthreads=threads)
        solved += int((r.status == 0).sum())
    dt = time.perf_counter() - t0
    v = solved / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "batch_per_step": sample, "tol": TOL},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{sample} θ per step (bounded sample of the GPU arm's batch distribution), "
                                       f"C restatement of src/solver.jl, OpenMP over θ on {threads} threads"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)
    return 0


_RESULT_FD = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version line to
    stdout when NCCL_DEBUG=VERSION is set on the box), so fd 1 is pointed at stderr for the duration of the run and
    the result line goes to the saved descriptor."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def _emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="lane_change")
    ap.add_argument("--batch", type=int, default=None, help="θ columns per GPU per step (default per workload)")
    ap.add_argument("--cpu-sample", type=int, default=2048)
    ap.add_argument("--ref-sample", type=int, default=2048)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.batch is None:   # lane-change: 2^18 (tail of never-converging instances amortised); QP: θ is 161 KB/instance
        args.batch = {"lane_change": 1 << 18, "readme_qp": 1 << 20, "random_qp": 1 << 15}.get(args.workload, 1 << 16)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from mcp_b200 import capi, sharding
    from mcp_b200.solver import _handle

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world != args.gpus:
        raise SystemExit("for --gpus N > 1 launch with: python -m torch.distributed.run --nnodes=1 "
                         "--nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N …")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, K, B = max(args.warmup, 3), args.steps, args.batch

    mcp, gen, desc = build_workload(args.workload)
    h = _handle(mcp)
    lib = h._lib
    nx, ny, nt = mcp.unconstrained_dimension, mcp.constrained_dimension, mcp.parameter_dimension
    # weak scaling: every rank draws its own B columns (seed depends on the rank)
    Θ_host = gen(B, 1 + rank)
    dev = torch.device("cuda", local)
    θ_pinned = torch.from_numpy(np.ascontiguousarray(Θ_host.T)).pin_memory()          # [B, nθ] = column-major nθ×B
    θ_dev = θ_pinned.to(dev, non_blocking=True)
    out = {k: torch.empty((B, n), dtype=torch.float64, device=dev) for k, n in (("x", nx), ("y", ny), ("s", ny))}
    kkt = torch.empty(B, dtype=torch.float64, device=dev)
    eps = torch.empty(B, dtype=torch.float64, device=dev)
    outer = torch.empty(B, dtype=torch.int32, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    steps_d = torch.empty(B, dtype=torch.int32, device=dev)
    opts = capi.default_opts(tol=TOL)
    stream = torch.cuda.current_stream()

    def device_step():
        rc = lib.mcpb200_solve_batched_device(h.raw, B, θ_dev.data_ptr(), None, None, None, C.byref(opts),
                                              out["x"].data_ptr(), out["y"].data_ptr(), out["s"].data_ptr(),
                                              kkt.data_ptr(), eps.data_ptr(), outer.data_ptr(), status.data_ptr(),
                                              steps_d.data_ptr(), C.c_void_p(stream.cuda_stream))
        h.check(rc)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fp64_peak = capi.measure_fp64_peak()          # TFLOP/s of the DFMA pipe on this device, measured now
    for _ in range(W):
        device_step()
    barrier()
    kernel_ms, launches, newton = [], 0, 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        ev0.record(stream)
        for _ in range(K):
            device_step()
            # (library-side CUDA events bracket the kernel on this same stream; read after the sync)
        ev1.record(stream)
        barrier()
        total_ms = ev0.elapsed_time(ev1)
    tm = h.timing()                                # last launch: kernel time, Newton steps, solved count
    solved_per_step = tm["solved"]
    total_ms = sharding.reduce_max_ms(total_ms)    # max over ranks
    solved_all = sharding.reduce_sum_int(solved_per_step)
    value = solved_all * K / (total_ms * 1e-3)
    kernel_ms_last = sharding.reduce_max_ms(tm["kernel_ms"])

    # ---- end-to-end through the host C-ABI call: pinned host θ → H2D → solve → D2H --------------------
    host = {k: torch.empty((B, n), dtype=torch.float64).pin_memory() for k, n in (("x", nx), ("y", ny), ("s", ny))}
    hk, he = torch.empty(B, dtype=torch.float64).pin_memory(), torch.empty(B, dtype=torch.float64).pin_memory()
    ho, hs, hn = (torch.empty(B, dtype=torch.int32).pin_memory() for _ in range(3))
    h.set_devices([local])

    def host_step():
        rc = lib.mcpb200_solve_batched(h.raw, B, θ_pinned.data_ptr(), None, None, None, C.byref(opts),
                                       host["x"].data_ptr(), host["y"].data_ptr(), host["s"].data_ptr(),
                                       hk.data_ptr(), he.data_ptr(), ho.data_ptr(), hs.data_ptr(), hn.data_ptr())
        h.check(rc)
        return int((hs == 0).sum())

    host_step()
    barrier()
    t0 = time.perf_counter()
    e2e_solved = 0
    for _ in range(K):
        e2e_solved += host_step()
    barrier()
    e2e_ms = sharding.reduce_max_ms((time.perf_counter() - t0) * 1e3)
    e2e_value = sharding.reduce_sum_int(e2e_solved) / (e2e_ms * 1e-3)
    h2d = B * nt * 8
    d2h = B * ((nx + 2 * ny) * 8 + 8 + 8 + 4 + 4 + 4)

    if rank == 0:
        info = h.info()
        peaks, peak_src = measured_peaks()
        flops_per_launch = info["flops_per_newton_step_band"] * tm["newton_steps"]
        achieved_tf = flops_per_launch / (tm["kernel_ms"] * 1e-3) / 1e12
        io_bytes = B * (8 * (nt + nx + 2 * ny) + 24)
        traffic = None   # DRAM bytes of the solve launch: per-Newton-step figure from the committed ncu capture
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                per_step = json.load(f).get(args.workload, {}).get("dram_bytes_per_newton_step")
            if per_step:
                traffic = float(per_step) * tm["newton_steps"]
        except (OSError, ValueError):
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "batch_per_gpu": B, "global_batch": B * world, "tol": TOL,
                       "parallelism": f"theta-sharded x{world} (independent instances, no collective)",
                       "l2": f"inputs+outputs {B * (nt + nx + 2 * ny) * 8 / 2**20:.0f} MiB per step exceed the 126 MB L2",
                       "solved_fraction": solved_all / (B * world),
                       "newton_steps_per_launch": tm["newton_steps"],
                       "kernel": {k: info[k] for k in ("n_reduced", "kl", "ku", "window_rows", "window_cols",
                                                       "instances_per_cta", "smem_bytes_per_cta", "regs_solve")}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / K},
            "gpu_launches": K * int(tm["launches"]),
            "clocks": clocks.summary(),
            "roofline": {"bound": "fp64", "kernel": "mcp_solve_kernel", "achieved": achieved_tf, "peak": fp64_peak,
                         "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak if fp64_peak else None,
                         "traffic": traffic,
                         "peak_source": "measured now on this device by libmcpb200's DFMA probe (MEASURED_PEAKS.json "
                                        "has no FP64 entry)",
                         "kernel_ms": kernel_ms_last,
                         "flops_per_newton_step": info["flops_per_newton_step_band"],
                         "hbm_io": {"achieved": io_bytes / (tm["kernel_ms"] * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                                    "unit": "GB/s", "frac": io_bytes / (tm["kernel_ms"] * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                    "peak_source": peak_src}},
        }
        if not args.no_cpu_baseline:
            cb, _ = cpu_baseline(mcp, Θ_host, min(args.cpu_sample, B))
            line["cpu_baseline"] = cb
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
