"""Tiny run of every kernel family for compute-sanitizer (racecheck / memcheck)."""
import sys, numpy as np
sys.path.insert(0, ".")
from mcp_b200 import InteriorPoint, solve, problems, solve_jacobian_θ
which = sys.argv[1:] or ["readme", "lane", "qp12", "qp100", "sens"]
if "readme" in which:
    m = problems.readme_qp(); s = solve(InteriorPoint(), m, problems.readme_qp_thetas(70, seed=1)); print("readme", int((s.status == 0).sum()))
if "lane" in which:
    m = problems.lane_change_game().mcp; Θ = problems.lane_change_thetas(20, seed=1)
    s = solve(InteriorPoint(), m, Θ, tol=1e-6, max_outer_iters=12); print("lane", int((s.status == 0).sum()))
if "sens" in which:
    m = problems.lane_change_game().mcp; Θ = problems.lane_change_thetas(4, seed=5, moving=True)
    s = solve(InteriorPoint(), m, Θ, tol=1e-6); J = solve_jacobian_θ(m, s, Θ); print("sens", J.shape)
if "qp12" in which:
    m = problems.random_qp(12, 10); s = solve(InteriorPoint(), m, problems.random_qp_thetas(6, seed=1, num_primals=12, num_inequalities=10, sparsity_rate=0.5), tol=1e-6); print("qp12", int((s.status == 0).sum()))
if "qp100" in which:
    m = problems.random_qp(100, 100); s = solve(InteriorPoint(), m, problems.random_qp_thetas(3, seed=1), tol=1e-6); print("qp100", int((s.status == 0).sum()))
