"""Adjudicate FP64 parity misses with the extended-precision oracle.

For a set of instances: GPU (if a GPU is present), C oracle and Python FP64 oracle each against
`oracle/ip_oracle_ext.py` (longdouble, full KKT system, refined solves).  Prints and stores, per instance, the
relative distance of each FP64 implementation from the extended-precision trajectory.

    python scripts/adjudicate_parity.py masked  [--B 256] [--all]     # cfg4 N = 4, the 8 masks × 32 scenarios of the test
    python scripts/adjudicate_parity.py lane    [--B 64]
    python scripts/adjudicate_parity.py qp      [--B 64] --all         # cfg2 QP 100×100: the GPU runs LDLᵀ, the oracles pivoted LU
Output: gpurun_out/adjudicate_<name>.json
"""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mcp_b200 import problems          # noqa: E402
from oracle import c_oracle as CO      # noqa: E402
from oracle import ip_oracle_ext as E  # noqa: E402
from oracle.ir_eval import OracleMCP   # noqa: E402


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b)))))


_G = {}


def _setup(which):
    if which == "masked":
        mcp = problems.masked_game(4, 30).mcp
    elif which == "qp":
        mcp = problems.random_qp(100, 100)
    else:
        mcp = problems.lane_change_game().mcp
    _G["mcp"] = mcp
    _G["ome"] = OracleMCP(mcp.ir, extended=True)


def _ext_one(args):
    which, th, x0, tol = args
    if "mcp" not in _G:
        _setup(which)
    e = E.solve_interior_point_ext(_G["ome"], th, x0=x0, tol=tol)
    return (e.status, e.newton_steps, e.outer_iters, e.x.astype(np.float64), e.y.astype(np.float64),
            e.s.astype(np.float64),
            # how far the rounded-to-double extended result is from the extended one (representation floor)
            float(np.max(np.abs(e.x - e.x.astype(np.float64).astype(np.longdouble)))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", choices=["masked", "lane", "qp"])
    ap.add_argument("--B", type=int, default=0)
    ap.add_argument("--all", action="store_true", help="run the extended oracle on every instance, not only suspects")
    ap.add_argument("--workers", type=int, default=os.cpu_count())
    a = ap.parse_args()
    if a.which == "masked":
        B, tol = a.B or 256, 1e-4
        mcp = problems.masked_game(4, 30).mcp
        Θ = problems.masked_game_thetas(B, 4, seed=11)          # tests/test_gpu_parity.py::test_masked_game_parity_statistics
        x0 = problems.masked_game_x0(Θ, 4, 30)
    elif a.which == "qp":
        B, tol = a.B or 64, 1e-6
        mcp = problems.random_qp(100, 100)
        Θ = problems.random_qp_thetas(B, seed=11)               # ::test_dense_symmetric_path_and_fallbacks (first B)
        x0 = None
    else:
        B, tol = a.B or 64, 1e-6
        mcp = problems.lane_change_game().mcp
        Θ = problems.lane_change_thetas(B, seed=2024)           # ::test_lane_change_parity_statistics (first B of 1024)
        x0 = None
    t0 = time.time()
    ref = CO.solve_batch(mcp.ir, Θ, x0=x0, tol=tol)
    print(f"C oracle: {time.time() - t0:.1f}s, solved {(ref.status == 0).sum()}/{B}", flush=True)
    gpu = None
    try:
        import torch
        if torch.cuda.is_available():
            from mcp_b200 import InteriorPoint, solve
            gpu = solve(InteriorPoint(), mcp, Θ, x0=x0, tol=tol)
    except Exception as ex:       # no GPU here: the CPU half still runs
        print("no GPU leg:", ex)
    suspects = list(range(B)) if a.all else []
    if gpu is not None and not a.all:
        for b in range(B):
            if ref.status[b] == 0 and gpu.status[b] == 0:
                e = max(rel(gpu.x[:, b], ref.x[:, b]), rel(gpu.y[:, b], ref.y[:, b]), rel(gpu.s[:, b], ref.s[:, b]))
                if e > 1e-7 or abs(int(gpu.newton_steps[b]) - int(ref.newton_steps[b])) > 0:
                    suspects.append(b)
            elif ref.status[b] != gpu.status[b]:
                suspects.append(b)
        # plus a control group of agreeing instances
        suspects += [b for b in range(0, B, max(1, B // 16)) if b not in suspects]
    print(f"extended-precision oracle on {len(suspects)} instances", flush=True)
    t0 = time.time()
    jobs = [(a.which, Θ[:, b].copy(), None if x0 is None else x0[:, b].copy(), tol) for b in suspects]
    with ProcessPoolExecutor(max_workers=a.workers) as pool:
        ext = list(pool.map(_ext_one, jobs))
    print(f"extended oracle: {time.time() - t0:.1f}s", flush=True)
    rows = []
    for b, (st, steps, outer, ex, ey, es, floor) in zip(suspects, ext):
        row = {"instance": int(b), "ext_status": st, "ext_steps": int(steps),
               "c_status": int(ref.status[b]), "c_steps": int(ref.newton_steps[b]),
               "c_vs_ext": max(rel(ref.x[:, b], ex), rel(ref.y[:, b], ey), rel(ref.s[:, b], es))}
        if gpu is not None:
            row.update({"gpu_status": int(gpu.status[b]), "gpu_steps": int(gpu.newton_steps[b]),
                        "gpu_vs_ext": max(rel(gpu.x[:, b], ex), rel(gpu.y[:, b], ey), rel(gpu.s[:, b], es)),
                        "gpu_vs_c": max(rel(gpu.x[:, b], ref.x[:, b]), rel(gpu.y[:, b], ref.y[:, b]),
                                        rel(gpu.s[:, b], ref.s[:, b]))})
        rows.append(row)
    solved = [r for r in rows if r["ext_status"] == "solved" and r["c_status"] == 0]
    summ = {"which": a.which, "B": B, "tol": tol, "n_ext": len(rows),
            "c_vs_ext_max": max((r["c_vs_ext"] for r in solved), default=None),
            "c_vs_ext_over_bar": [r["instance"] for r in solved if r["c_vs_ext"] > 1e-6]}
    if gpu is not None:
        gs = [r for r in solved if r["gpu_status"] == 0]
        summ.update({"gpu_vs_ext_max": max((r["gpu_vs_ext"] for r in gs), default=None),
                     "gpu_vs_ext_over_bar": [r["instance"] for r in gs if r["gpu_vs_ext"] > 1e-6],
                     "gpu_vs_c_over_bar": [r["instance"] for r in gs if r["gpu_vs_c"] > 1e-6]})
    print(json.dumps(summ, indent=1))
    for r in rows:
        if r["c_vs_ext"] > 1e-7 or r.get("gpu_vs_ext", 0) > 1e-7:
            print(r)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"adjudicate_{a.which}.json"), "w") as f:
        json.dump({"summary": summ, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
