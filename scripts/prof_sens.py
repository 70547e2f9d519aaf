import sys, numpy as np
sys.path.insert(0, ".")
from mcp_b200 import InteriorPoint, solve, problems, solve_pullback, solve_jacobian_θ
from mcp_b200.solver import _handle
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
mode = sys.argv[2] if len(sys.argv) > 2 else "vjp"
mcp = problems.lane_change_game().mcp
Θ = problems.lane_change_thetas(B, seed=1)
sol = solve(InteriorPoint(), mcp, Θ, tol=1e-6)
for _ in range(2):
    if mode == "vjp":
        g = solve_pullback(mcp, sol, Θ, 2 * sol.x, 2 * sol.y, None)
    else:
        J = solve_jacobian_θ(mcp, sol, Θ)
print(mode, "kernel ms", _handle(mcp).timing()["kernel_ms"])
