import sys, numpy as np
sys.path.insert(0, ".")
from mcp_b200 import InteriorPoint, solve, problems
from mcp_b200.solver import _handle
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
mcp = problems.readme_qp()
Θ = problems.readme_qp_thetas(B, seed=1)
for _ in range(2):
    sol = solve(InteriorPoint(), mcp, Θ)
print("kernel ms", _handle(mcp).timing()["kernel_ms"], "solved", int((sol.status == 0).sum()))
