import sys, numpy as np
sys.path.insert(0, ".")
from mcp_b200 import InteriorPoint, solve, problems
from mcp_b200.solver import _handle
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B = int(sys.argv[2]) if len(sys.argv) > 2 else 444
game = problems.masked_game(N, 30); mcp = game.mcp
Θ = problems.masked_game_thetas(B, N, seed=1); x0 = problems.masked_game_x0(Θ, N, 30)
sol = solve(InteriorPoint(), mcp, Θ, x0=x0, tol=1e-4)
tm = _handle(mcp).timing()
print("kernel ms", tm["kernel_ms"], "pass0", tm["pass0_ms"], "solved", int((sol.status == 0).sum()), "steps", tm["newton_steps"])
