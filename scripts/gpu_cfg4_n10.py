"""cfg4 at N = 10 (nx = 3000, ny = 3630): parity of a few instances against the C oracle, then throughput."""
import sys, time, json, numpy as np
sys.path.insert(0, ".")
from mcp_b200 import InteriorPoint, solve, problems, solve_pullback
from mcp_b200.solver import _handle
from oracle import c_oracle as CO
N, H = 10, 30
Bs = [int(a) for a in sys.argv[1:]] or [64, 1024]
game = problems.masked_game(N, H); mcp = game.mcp
t = time.time(); h = _handle(mcp); out = {"create_s": time.time() - t, "kernel": h.info()}
Θ = problems.masked_game_thetas(max(Bs), N, seed=1); x0 = problems.masked_game_x0(Θ, N, H)
nref = 4
t = time.time(); ref = CO.solve_batch(mcp.ir, Θ[:, :nref], x0=x0[:, :nref], tol=1e-4); dt = time.time() - t
out["cpu_port"] = dict(threads=int(r.threads) if (r := ref) is not None else 0, seconds=dt, solves_per_s=float((ref.status == 0).sum() / dt),
                       status=ref.status.tolist(), steps=ref.newton_steps.tolist())
for B in Bs:
    for _ in range(2):
        t = time.time(); sol = solve(InteriorPoint(), mcp, Θ[:, :B], x0=x0[:, :B], tol=1e-4); wall = time.time() - t
    tm = h.timing()
    rel = [float(np.abs(sol.x[:, b] - ref.x[:, b]).max() / max(1.0, np.abs(ref.x[:, b]).max())) for b in range(min(nref, B))]
    out[f"solve_B{B}"] = dict(kernel_ms=tm["kernel_ms"], pass0_ms=tm["pass0_ms"], wall_ms=wall * 1e3, solved=int((sol.status == 0).sum()),
                              newton_steps=tm["newton_steps"], solves_per_s_kernel=float((sol.status == 0).sum() / (tm["kernel_ms"] * 1e-3)),
                              status_head=sol.status[:nref].tolist(), steps_head=sol.newton_steps[:nref].tolist(), x_rel_err_vs_oracle=rel)
    t = time.time(); g = solve_pullback(mcp, sol, Θ[:, :B], 2 * sol.x, None, None); wall = time.time() - t
    tm = h.timing()
    out[f"vjp_B{B}"] = dict(kernel_ms=tm["kernel_ms"], wall_ms=wall * 1e3, vjps_per_s_kernel=B / (tm["kernel_ms"] * 1e-3), finite=bool(np.isfinite(g).all()))
print(json.dumps(out, indent=1))
