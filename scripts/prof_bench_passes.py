import sys, numpy as np
sys.path.insert(0, ".")
from mcp_b200 import InteriorPoint, solve, problems
from mcp_b200.solver import _handle
B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
mcp = problems.lane_change_game().mcp
Θ = problems.lane_change_thetas(B, seed=2)
for rep in range(3):
    sol = solve(InteriorPoint(), mcp, Θ, tol=1e-6)
t = _handle(mcp).timing()
print("kernel ms", round(t["kernel_ms"], 1), "pass0", round(t["pass0_ms"], 1), "pass1", round(t["kernel_ms"] - t["pass0_ms"], 1), "solved", t["solved"], "steps", t["newton_steps"], "deferred", t["deferred"])
