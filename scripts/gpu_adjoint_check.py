"""Adjoint pullback (one Cᵀ solve) against the forward pullback (nθ solves of C) on the GPU, plus timings."""
import os, sys, time, json, numpy as np
sys.path.insert(0, ".")
from mcp_b200 import InteriorPoint, solve, problems, solve_pullback
from mcp_b200.solver import _handle

def both(make, Θ, x0=None, tol=1e-6):
    out = {}
    for mode in ("1", "0"):
        os.environ["MCPB200_ADJOINT"] = mode
        mcp = make()
        sol = solve(InteriorPoint(), mcp, Θ, x0=x0, tol=tol)
        g = solve_pullback(mcp, sol, Θ, 2 * sol.x, 2 * sol.y, 0.5 * sol.s)
        t = time.time(); g = solve_pullback(mcp, sol, Θ, 2 * sol.x, 2 * sol.y, 0.5 * sol.s); wall = time.time() - t
        out[mode] = (g, sol.status.copy(), _handle(mcp).timing()["kernel_ms"], wall)
    ga, st, ka, _ = out["1"]; gf, _, kf, _ = out["0"]
    ok = st == 0
    rel = np.max(np.abs(ga[:, ok] - gf[:, ok]), axis=0) / np.maximum(1.0, np.max(np.abs(gf[:, ok]), axis=0))
    return dict(solved=int(ok.sum()), B=int(Θ.shape[1]), rel_err_median=float(np.median(rel)), rel_err_max=float(rel.max()),
                frac_below_1e6=float((rel < 1e-6).mean()), kernel_ms_adjoint=ka, kernel_ms_forward=kf)

res = {}
res["readme"] = both(problems.readme_qp, problems.readme_qp_thetas(4096, seed=1), tol=1e-6)
res["lane_change_moving"] = both(lambda: problems.lane_change_game().mcp, problems.lane_change_thetas(4096, seed=5, moving=True))
Θ = problems.masked_game_thetas(1024, 4, seed=1)
res["masked_n4"] = both(lambda: problems.masked_game(4, 30).mcp, Θ, x0=problems.masked_game_x0(Θ, 4, 30), tol=1e-4)
if len(sys.argv) > 1:
    Θ = problems.masked_game_thetas(256, 10, seed=1)
    res["masked_n10"] = both(lambda: problems.masked_game(10, 30).mcp, Θ, x0=problems.masked_game_x0(Θ, 10, 30), tol=1e-4)
print(json.dumps(res, indent=1))
