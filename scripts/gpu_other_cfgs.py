"""Throughput of the non-headline configs (BASELINE.json configs[0], [1], [4]) — kernel time by CUDA events."""
import sys, time, json, numpy as np
sys.path.insert(0, ".")
from mcp_b200 import InteriorPoint, solve, problems, solve_jacobian_θ, solve_pullback, capi
from mcp_b200.solver import _handle
out = {}
peak = capi.measure_fp64_peak(); out["fp64_peak_tflops"] = peak
# cfg1: README QP, 2^20 θ
rq = problems.readme_qp(); h = _handle(rq); Θ = problems.readme_qp_thetas(1 << 20, seed=1)
for tol in (1e-4, 1e-6):
    for _ in range(2):
        t = time.time(); sol = solve(InteriorPoint(), rq, Θ, tol=tol); wall = time.time() - t
    tm = h.timing()
    out[f"cfg1_readme_qp_tol{tol:g}"] = dict(B=1 << 20, kernel_ms=tm["kernel_ms"], wall_ms=wall * 1e3, solved=int((sol.status == 0).sum()),
                                            solves_per_s_kernel=(sol.status == 0).sum() / (tm["kernel_ms"] * 1e-3), newton_steps=tm["newton_steps"])
# cfg2: random convex QP 100x100
qp = problems.random_qp(100, 100); hq = _handle(qp); info = hq.info()
B = 2048
Θ = problems.random_qp_thetas(B, seed=1)
for _ in range(2):
    t = time.time(); sol = solve(InteriorPoint(), qp, Θ, tol=1e-6); wall = time.time() - t
tm = hq.timing()
flops = info["flops_per_newton_step_band"] * tm["newton_steps"]
out["cfg2_random_qp_100x100"] = dict(B=B, kernel_ms=tm["kernel_ms"], wall_ms=wall * 1e3, solved=int((sol.status == 0).sum()), newton_steps=tm["newton_steps"],
                                     solves_per_s_kernel=(sol.status == 0).sum() / (tm["kernel_ms"] * 1e-3), flops_per_step=info["flops_per_newton_step_band"],
                                     tflops=flops / (tm["kernel_ms"] * 1e-3) / 1e12, frac_fp64=flops / (tm["kernel_ms"] * 1e-3) / 1e12 / peak,
                                     kernel={k: info[k] for k in ("kl", "ku", "window_rows", "window_cols", "instances_per_cta", "threads_per_instance", "smem_bytes_per_cta")})
# cfg5: lane-change + sensitivities (well-posed θ)
lane = problems.lane_change_game().mcp; hl = _handle(lane)
B = 16384
Θ = problems.lane_change_thetas(B, seed=7, moving=True)
sol = solve(InteriorPoint(), lane, Θ, tol=1e-6)
for _ in range(2):
    t = time.time(); g = solve_pullback(lane, sol, Θ, 2 * sol.x, 2 * sol.y, None); wall = time.time() - t
tm = hl.timing()
out["cfg5_lane_change_vjp"] = dict(B=B, kernel_ms=tm["kernel_ms"], wall_ms=wall * 1e3, vjps_per_s_kernel=B / (tm["kernel_ms"] * 1e-3), solved_fraction=float((sol.status == 0).mean()))
t = time.time(); J = solve_jacobian_θ(lane, sol, Θ); wall = time.time() - t
tm = hl.timing()
out["cfg5_lane_change_jacobian"] = dict(B=B, kernel_ms=tm["kernel_ms"], wall_ms=wall * 1e3, jacobians_per_s_kernel=B / (tm["kernel_ms"] * 1e-3), bytes_out=int(J.nbytes))
print(json.dumps(out, indent=1))
