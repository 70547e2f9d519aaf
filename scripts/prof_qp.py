import sys, numpy as np
sys.path.insert(0, ".")
from mcp_b200 import InteriorPoint, solve, problems
from mcp_b200.solver import _handle
B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
qp = problems.random_qp(100, 100)
Θ = problems.random_qp_thetas(B, seed=1)
for rep in range(2):
    sol = solve(InteriorPoint(), qp, Θ, tol=1e-6)
tm = _handle(qp).timing()
print("kernel ms", tm["kernel_ms"], "pass0 ms", tm["pass0_ms"], "deferred", tm["deferred"], "solved", int((sol.status == 0).sum()), "steps", int(sol.newton_steps.sum()))
