"""cfg4 throughput: masked N-player game (N = 4, H = 30), all ego masks x scenarios, stay-at-rest start, tol 1e-4."""
import sys, time, json, numpy as np
sys.path.insert(0, ".")
from mcp_b200 import InteriorPoint, solve, problems, solve_pullback
from mcp_b200.solver import _handle
from oracle import c_oracle as CO
N, H = 4, 30
game = problems.masked_game(N, H); mcp = game.mcp; h = _handle(mcp)
out = {"kernel": h.info()}
for B in ([int(a) for a in sys.argv[1:]] or [2048, 8192]):
    Θ = problems.masked_game_thetas(B, N, seed=1); x0 = problems.masked_game_x0(Θ, N, H)
    for _ in range(2):
        t = time.time(); sol = solve(InteriorPoint(), mcp, Θ, x0=x0, tol=1e-4); wall = time.time() - t
    tm = h.timing()
    out[f"solve_B{B}"] = dict(kernel_ms=tm["kernel_ms"], pass0_ms=tm["pass0_ms"], deferred=tm["deferred"], wall_ms=wall * 1e3, solved=int((sol.status == 0).sum()), newton_steps=tm["newton_steps"],
                              solves_per_s_kernel=float((sol.status == 0).sum() / (tm["kernel_ms"] * 1e-3)), solves_per_s_wall=float((sol.status == 0).sum() / wall))
    t = time.time(); g = solve_pullback(mcp, sol, Θ, 2 * sol.x, None, None); wall = time.time() - t
    tm = h.timing()
    out[f"vjp_B{B}"] = dict(kernel_ms=tm["kernel_ms"], wall_ms=wall * 1e3, vjps_per_s_kernel=B / (tm["kernel_ms"] * 1e-3))
Θs = Θ[:, :256]; t = time.time(); r = CO.solve_batch(mcp.ir, Θs, x0=x0[:, :256], tol=1e-4); dt = time.time() - t
out["cpu_port"] = dict(threads=int(r.threads), solves_per_s=float((r.status == 0).sum() / dt))
print(json.dumps(out, indent=1))
