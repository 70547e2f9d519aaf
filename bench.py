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
    CO.solve_batch(mcp.ir, Θs[:, :8], tol=TOL, nthreads=threads, compiled=True)      # warm-up (column ordering, gcc, page-in)
    t0 = time.perf_counter()
    r = CO.solve_batch(mcp.ir, Θs, tol=TOL, nthreads=threads, compiled=True)
    dt = time.perf_counter() - t0
    solved = int((r.status == 0).sum())
    return {"value": solved / dt, "unit": UNIT, "cores": int(r.threads), "kind": "port",
            "sample": f"first {Θs.shape[1]} θ of the bench batch, {solved} converged, {dt:.2f} s wall, "
                      f"C restatement of src/solver.jl with F/∇F as generated C, -O3 -march=native (Julia cannot run in this image), OpenMP over θ"}, r


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
        CO.solve_batch(mcp.ir, Θ[:, :min(sample, 64)], tol=TOL, nthreads=threads, compiled=True)
    t0 = time.perf_counter()
    solved = 0
    for _ in range(args.steps):
        r = CO.solve_batch(mcp.ir, Θ, tol=TOL, nthreads=threads, compiled=True)
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


def _time_device(fn, stream, reps: int):
    """CUDA-event time (ms) of `reps` calls of fn on torch's current stream, after one warm-up call."""
    import torch
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def other_workloads(local: int, fp64_peak: float, hbm_gbs: float, reps: int = 3):
    """The other four BASELINE.json configs, timed in the same run (N = 1 only) at small repetition counts:
    cfg1 README QP 2^20 θ, cfg2 random QP 100×100 cold + the warm-started θ sweep
    (benchmark/quadratic_program_benchmark.jl:51-74), cfg5 lane-change VJP + full Jacobian, cfg4 masked game N = 4.
    Each entry: converged solves/s device-resident (`value`), kernel_ms, the roofline fraction of its solve kernel, and
    the end-to-end figure through the host C-ABI call where the workload has one."""
    import torch
    from mcp_b200 import capi, problems, torch_api
    from mcp_b200.solver import _handle
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream()
    out = {}

    def solve_entry(mcp, Θ_host, tol, x0=None, y0=None, e2e=True, note=None):
        h = _handle(mcp)
        nx, ny, nt = mcp.unconstrained_dimension, mcp.constrained_dimension, mcp.parameter_dimension
        B = Θ_host.shape[1]
        θp = torch.from_numpy(np.ascontiguousarray(Θ_host.T)).pin_memory()
        θd = θp.to(dev)
        sol = dict(x=torch.empty((B, nx), dtype=torch.float64, device=dev), y=torch.empty((B, ny), dtype=torch.float64, device=dev),
                   s=torch.empty((B, ny), dtype=torch.float64, device=dev), kkt=torch.empty(B, dtype=torch.float64, device=dev),
                   eps=torch.empty(B, dtype=torch.float64, device=dev), outer=torch.empty(B, dtype=torch.int32, device=dev),
                   status=torch.empty(B, dtype=torch.int32, device=dev), steps=torch.empty(B, dtype=torch.int32, device=dev))
        dopts = capi.default_opts(tol=tol)

        def step():   # the device entry point on resident tensors, outputs allocated once (as in the headline loop)
            h.check(h._lib.mcpb200_solve_batched_device(
                h.raw, B, θd.data_ptr(), None if x0 is None else x0.data_ptr(), None if y0 is None else y0.data_ptr(), None,
                C.byref(dopts), sol["x"].data_ptr(), sol["y"].data_ptr(), sol["s"].data_ptr(), sol["kkt"].data_ptr(),
                sol["eps"].data_ptr(), sol["outer"].data_ptr(), sol["status"].data_ptr(), sol["steps"].data_ptr(),
                C.c_void_p(stream.cuda_stream)))
        ms = _time_device(step, stream, reps)
        tm, info = h.timing(), h.info()
        solved = int(tm["solved"])
        flops = info["flops_per_newton_step_band"] * tm["newton_steps"]
        io_bytes = B * (8 * (nt + nx + 2 * ny) + 24)
        ent = {"batch": B, "tol": tol, "value": solved / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
               "kernel_ms": tm["kernel_ms"], "kernel_only_value": solved / (tm["kernel_ms"] * 1e-3), "solved_fraction": solved / B, "newton_steps": int(tm["newton_steps"]),
               "roofline": {"bound": "fp64", "achieved": flops / (tm["kernel_ms"] * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                            "frac": flops / (tm["kernel_ms"] * 1e-3) / 1e12 / fp64_peak,
                            "flops_per_newton_step": info["flops_per_newton_step_band"]},
               "hbm_io": {"achieved": io_bytes / (tm["kernel_ms"] * 1e-3) / 1e9, "peak": hbm_gbs, "unit": "GB/s",
                          "frac": io_bytes / (tm["kernel_ms"] * 1e-3) / 1e9 / hbm_gbs}}
        if note:
            ent["note"] = note
        if e2e:
            host = [torch.empty((B, n), dtype=torch.float64).pin_memory() for n in (nx, ny, ny)]
            hk, he = torch.empty(B, dtype=torch.float64).pin_memory(), torch.empty(B, dtype=torch.float64).pin_memory()
            ho, hs, hn = (torch.empty(B, dtype=torch.int32).pin_memory() for _ in range(3))
            opts = capi.default_opts(tol=tol)
            h.set_devices([local])

            def host_step():
                h.check(h._lib.mcpb200_solve_batched(h.raw, B, θp.data_ptr(), None, None, None, C.byref(opts), host[0].data_ptr(),
                                                     host[1].data_ptr(), host[2].data_ptr(), hk.data_ptr(), he.data_ptr(),
                                                     ho.data_ptr(), hs.data_ptr(), hn.data_ptr()))
            host_step()
            t0 = time.perf_counter()
            for _ in range(reps):
                host_step()
            dt = (time.perf_counter() - t0) / reps
            ent["e2e"] = {"value": int((hs == 0).sum()) / dt, "unit": UNIT, "h2d_bytes_per_step": B * nt * 8,
                          "d2h_bytes_per_step": B * ((nx + 2 * ny) * 8 + 28), "ms_per_step": dt * 1e3}
        return ent, sol, θd

    def guard(label, fn):   # one config failing (memory, a box without enough host RAM for 5 GB of pinned θ) must not cost the others
        try:
            fn()
        except Exception as ex:
            out[label + "_error"] = f"{type(ex).__name__}: {ex}"
            torch.cuda.empty_cache()

    def cfg1():
        # cfg1 — README QP, 2^20 θ, defaults (tol 1e-4) and tol 1e-6
        mcp = problems.readme_qp()
        Θ = problems.readme_qp_thetas(1 << 20, seed=1)
        out["cfg1_readme_qp_tol1e-4"], _, _ = solve_entry(mcp, Θ, 1e-4)
        out["cfg1_readme_qp_tol1e-6"], _, _ = solve_entry(mcp, Θ, 1e-6, e2e=False)

    def cfg2():
        # cfg2 — random convex QP 100×100: cold, then the θ sweep (ϕ ← ϕ + 0.01·N(0,1)) warm-started from the cold solution
        mcp = problems.random_qp(100, 100)
        Bq = 1 << 15   # (r1's batch; ≈ 35 ms of the call is the pass-1 tail of the never-converging instances, whatever the batch)
        Θ = problems.random_qp_thetas(Bq, seed=1)
        cold, sol, θd = solve_entry(mcp, Θ, TOL, note="cold start x₀=0, y₀=s₀=1; roofline credits the reference algorithm's flops (dense LU + dense "
                                                       "Schur product, SURVEY.md §8d) — the kernel executes fewer: LDLᵀ for symmetric G_x, Schur terms "
                                                       "that are zero in value skipped (DESIGN.md §8)")
        out["cfg2_random_qp_cold"] = cold
        g = torch.Generator(device=dev)
        g.manual_seed(7)
        θ2 = θd.clone()
        θ2[:, -100:] += 0.01 * torch.randn((Bq, 100), dtype=torch.float64, device=dev, generator=g)
        x0, y0 = sol["x"].clone(), sol["y"].clamp_min(1e-3)
        h = _handle(mcp)
        warm = {k: torch.empty_like(v) for k, v in sol.items()}
        dopts = capi.default_opts(tol=TOL)

        def warm_step():
            h.check(h._lib.mcpb200_solve_batched_device(
                h.raw, Bq, θ2.data_ptr(), x0.data_ptr(), y0.data_ptr(), None, C.byref(dopts), warm["x"].data_ptr(), warm["y"].data_ptr(),
                warm["s"].data_ptr(), warm["kkt"].data_ptr(), warm["eps"].data_ptr(), warm["outer"].data_ptr(), warm["status"].data_ptr(),
                warm["steps"].data_ptr(), C.c_void_p(stream.cuda_stream)))
        ms = _time_device(warm_step, stream, reps)
        tm, info = h.timing(), h.info()
        flops = info["flops_per_newton_step_band"] * tm["newton_steps"]
        out["cfg2_random_qp_warm_sweep"] = {
            "batch": Bq, "tol": TOL, "value": int(tm["solved"]) / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "kernel_ms": tm["kernel_ms"],
            "solved_fraction": int(tm["solved"]) / Bq, "newton_steps": int(tm["newton_steps"]),
            "newton_steps_cold": cold["newton_steps"],
            "roofline": {"bound": "fp64", "achieved": flops / (tm["kernel_ms"] * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": flops / (tm["kernel_ms"] * 1e-3) / 1e12 / fp64_peak},
            "note": "same M, A, b; ϕ perturbed by 0.01·N(0,1); x₀, y₀ = cold solution (y₀ clamped ≥ 1e-3), s₀ = 1, ϵ restarts at 1 "
                    "(src/solver.jl:41,67)"}
        del θd, θ2, sol, warm, x0, y0
        torch.cuda.empty_cache()

    def cfg5():
        # cfg5 — lane-change sensitivities: VJP (adjoint kernel) and the full Jacobian ∂z/∂θ (10 right-hand sides)
        mcp = problems.lane_change_game().mcp
        h = _handle(mcp)
        nx, ny, nt = mcp.unconstrained_dimension, mcp.constrained_dimension, mcp.parameter_dimension
        Bs = 1 << 14
        θd = torch.from_numpy(np.ascontiguousarray(problems.lane_change_thetas(Bs, seed=5, moving=True).T)).to(dev)
        sol = torch_api.solve_device(mcp, θd, tol=TOL)
        zbar = torch.cat([2 * sol["x"], 2 * sol["y"], torch.zeros_like(sol["s"])], dim=1).contiguous()
        ms = _time_device(lambda: torch_api.pullback_device(mcp, θd, sol["x"], sol["y"], sol["s"], sol["eps"], zbar), stream, reps)
        tm = h.timing()
        out["cfg5_lane_change_vjp"] = {"batch": Bs, "value": Bs / (ms * 1e-3), "unit": "VJPs/s", "ms_per_step": ms,
                                       "kernel_ms": tm["kernel_ms"], "note": "adjoint mode: one solve with Cᵀ per instance (src/AutoDiff.jl:59-76)"}
        jac = torch.empty((Bs, nt, nx + 2 * ny), dtype=torch.float64, device=dev)

        def jac_step():
            h.check(h._lib.mcpb200_sensitivities_device(h.raw, Bs, θd.data_ptr(), sol["x"].data_ptr(), sol["y"].data_ptr(), sol["s"].data_ptr(),
                                                        sol["eps"].data_ptr(), jac.data_ptr(), None, None, 0, None, None, None,
                                                        C.c_void_p(stream.cuda_stream)))
        ms = _time_device(jac_step, stream, reps)
        tm = h.timing()
        out["cfg5_lane_change_jacobian"] = {"batch": Bs, "value": Bs / (ms * 1e-3), "unit": "Jacobians/s", "ms_per_step": ms,
                                            "kernel_ms": tm["kernel_ms"], "note": "∂z/∂θ, 700×10 per instance (src/AutoDiff.jl:18-40)"}
        del jac, zbar, sol, θd
        torch.cuda.empty_cache()

    def cfg4():
        # cfg4 — masked game N = 4, H = 30: all 8 ego masks × 256 scenarios, stay-at-rest x₀, tol 1e-4 (the application's settings)
        mcp = problems.masked_game(4, 30).mcp
        Θ = problems.masked_game_thetas(8192, 4, seed=1)
        x0 = torch.from_numpy(np.ascontiguousarray(problems.masked_game_x0(Θ, 4, 30).T)).to(dev)
        out["cfg4_masked_game_n4"], _, _ = solve_entry(mcp, Θ, 1e-4, x0=x0, e2e=False, note="8 ego masks × 1024 scenarios, x₀ = stay-at-rest rollout")

    guard("cfg1_readme_qp", cfg1)
    guard("cfg2_random_qp", cfg2)
    guard("cfg5_lane_change", cfg5)
    guard("cfg4_masked_game_n4", cfg4)
    return out


def measured_traffic(workload: str):
    """DRAM bytes per Newton step of the solve kernel, from the committed one-launch ncu capture that carries its OWN step
    count (`profiles/r2_traffic.json`: fixed-work mode, B × 30 Newton steps in the captured launch)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            return json.load(f).get(workload)
    except (OSError, ValueError):
        return None


def run_single_process(args):
    """`--single-process --gpus N`: ONE process, ONE `mcpb200_solve_batched` call per step after
    `mcpb200_set_devices(0..N-1)` — north_star's "θ batch sharded across the GPUs of one box, gathered on the host" as a
    single plugin call (one host thread per device inside the library, no collective).  Host arrays are plain PAGEABLE
    numpy arrays, as a Julia caller's would be (the library page-locks them for the duration of the call); the same
    call on pinned arrays is timed next to it."""
    import torch
    from mcp_b200 import capi
    from mcp_b200.solver import _handle
    N = args.gpus
    if torch.cuda.device_count() < N:
        raise SystemExit(f"--single-process --gpus {N}: only {torch.cuda.device_count()} GPUs visible")
    mcp, gen, desc = build_workload(args.workload)
    h = _handle(mcp)
    lib = h._lib
    nx, ny, nt = mcp.unconstrained_dimension, mcp.constrained_dimension, mcp.parameter_dimension
    B = args.batch * N
    W, K = max(args.warmup, 1), args.steps
    Θ = np.ascontiguousarray(gen(B, 1).T)                       # [B, nθ] row-major = column-major nθ×B, pageable
    opts = capi.default_opts(tol=TOL)
    h.set_devices(list(range(N)))

    def buffers(pinned):
        def mk(shape, dt):
            a = torch.empty(shape, dtype=dt)
            return a.pin_memory() if pinned else a
        θ = torch.from_numpy(Θ)
        return dict(θ=θ.pin_memory() if pinned else θ, x=mk((B, nx), torch.float64), y=mk((B, ny), torch.float64),
                    s=mk((B, ny), torch.float64), kkt=mk(B, torch.float64), eps=mk(B, torch.float64),
                    outer=mk(B, torch.int32), status=mk(B, torch.int32), steps=mk(B, torch.int32))

    def timed(buf):
        def step():
            h.check(lib.mcpb200_solve_batched(h.raw, B, buf["θ"].data_ptr(), None, None, None, C.byref(opts), buf["x"].data_ptr(),
                                              buf["y"].data_ptr(), buf["s"].data_ptr(), buf["kkt"].data_ptr(), buf["eps"].data_ptr(),
                                              buf["outer"].data_ptr(), buf["status"].data_ptr(), buf["steps"].data_ptr()))
            return int((buf["status"] == 0).sum())
        for _ in range(W):
            step()
        t0 = time.perf_counter()
        solved = sum(step() for _ in range(K))
        dt = time.perf_counter() - t0
        return solved / dt, dt / K * 1e3, h.timing()

    with ClockSampler(0) as clocks:
        v_page, ms_page, tm = timed(buffers(False))
    v_pin, ms_pin, _ = timed(buffers(True))
    line = {"metric": METRIC, "value": tm["solved"] / (tm["kernel_ms"] * 1e-3), "unit": UNIT, "n_gpus": N, "steps": K, "warmup": W,
            "ms_per_step": ms_page, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "mode": "single-process: one mcpb200_solve_batched call over all GPUs (mcpb200_set_devices)",
                       "batch_per_gpu": args.batch, "global_batch": B, "tol": TOL, "launches_per_call": int(tm["launches"]),
                       "value_is": "converged solves ÷ the slowest device's kernel time (library CUDA events)"},
            "e2e": {"value": v_page, "unit": UNIT, "ms_per_step": ms_page, "host_memory": "pageable (page-locked per call by the library)",
                    "h2d_bytes_per_step": B * nt * 8, "d2h_bytes_per_step": B * ((nx + 2 * ny) * 8 + 28)},
            "e2e_pinned": {"value": v_pin, "unit": UNIT, "ms_per_step": ms_pin},
            "gpu_launches": K * int(tm["launches"]), "clocks": clocks.summary()}
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
    ap.add_argument("--single-process", action="store_true",
                    help="one process, one mcpb200_solve_batched call spanning --gpus devices (not the driver's torchrun contract)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-workloads", action="store_true", help="skip the cfg1/2/4/5 entries (N = 1 only)")
    args = ap.parse_args()
    if args.batch is None:   # lane-change: 2^18 (tail of never-converging instances amortised); QP: θ is 161 KB/instance
        args.batch = {"lane_change": 1 << 18, "readme_qp": 1 << 20, "random_qp": 1 << 15}.get(args.workload, 1 << 16)
    if args.impl == "reference":
        return run_reference(args)
    if args.single_process:
        return run_single_process(args)

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
        # DRAM bytes of the solve launches of one step: the per-Newton-step figure of the committed ncu capture (which
        # carries its own step count) × the Newton steps this run's launch executed
        tr = measured_traffic(args.workload)
        traffic = float(tr["dram_bytes_per_newton_step"]) * tm["newton_steps"] if tr else None
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
                         "traffic_source": (tr or {}).get("source"),
                         "algorithmic_io_bytes": io_bytes,
                         "peak_source": "measured now on this device by libmcpb200's DFMA probe (MEASURED_PEAKS.json "
                                        "has no FP64 entry)",
                         "kernel_ms": kernel_ms_last,
                         "flops_per_newton_step": info["flops_per_newton_step_band"],
                         "hbm_io": {"achieved": io_bytes / (tm["kernel_ms"] * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                                    "unit": "GB/s", "frac": io_bytes / (tm["kernel_ms"] * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                    "peak_source": peak_src}},
        }
        if world == 1 and not args.no_other_workloads and args.workload == "lane_change":
            try:
                line["config"]["other_workloads"] = other_workloads(local, fp64_peak, peaks["hbm_gbs"])
            except Exception as ex:     # the headline must survive a failure in the extras — but say so
                line["config"]["other_workloads"] = {"error": f"{type(ex).__name__}: {ex}"}
        if not args.no_cpu_baseline:
            cb, _ = cpu_baseline(mcp, Θ_host, min(args.cpu_sample, B))
            line["cpu_baseline"] = cb
            c1, _ = cpu_baseline(mcp, Θ_host, min(max(args.cpu_sample // 8, 64), B), threads=1)
            line["cpu_baseline_1core"] = c1
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
