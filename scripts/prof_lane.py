import sys, numpy as np
sys.path.insert(0, ".")
from mcp_b200 import InteriorPoint, solve, problems
from mcp_b200.solver import _handle
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
lane = problems.lane_change_game(); mcp = lane.mcp
Θ = problems.lane_change_thetas(B, seed=1)
for rep in range(2):
    sol = solve(InteriorPoint(), mcp, Θ, tol=1e-6)
t = _handle(mcp).timing()
print("kernel ms", t["kernel_ms"], "pass0 ms", t["pass0_ms"], "solved", int((sol.status == 0).sum()), "newton steps", t["newton_steps"],
      "deferred", t["deferred"], "pass-0 ns per step per SM", 1e6 * t["pass0_ms"] * 148 / max(1, t["newton_steps"]))
