import sys, numpy as np
sys.path.insert(0, ".")
from mcp_b200 import InteriorPoint, solve, problems, solve_jacobian_θ
from oracle.ir_eval import OracleMCP
from oracle import ip_oracle as O
mcp = problems.lane_change_game().mcp
om = OracleMCP(mcp.ir)
Θ = problems.lane_change_thetas(6, seed=5)
sol = solve(InteriorPoint(), mcp, Θ, tol=1e-6)
J = solve_jacobian_θ(mcp, sol, Θ)
for b in range(6):
    if sol.status[b] != 0: continue
    Jz = om.JFz(sol.x[:, b], sol.y[:, b], sol.s[:, b], Θ[:, b], 0.0).toarray()
    Jt = om.JFt(sol.x[:, b], sol.y[:, b], sol.s[:, b], Θ[:, b], 0.0).toarray()
    ref = O.Solution("solved", sol.x[:, b], sol.y[:, b], sol.s[:, b], 0.0, 0.0, 0)
    Jr = O.solve_jacobian_theta(om, ref, Θ[:, b])
    for name, JJ in (("gpu", J[:, :, b]), ("qr ", Jr)):
        resid = Jz @ JJ + Jt
        comp = np.max(np.abs(resid) / (np.abs(Jz) @ np.abs(JJ) + np.abs(Jt) + 1e-300))
        norm = np.max(np.abs(resid)) / (np.linalg.norm(Jz, np.inf) * np.max(np.abs(JJ)) + np.max(np.abs(Jt)))
        rowwise = np.max(np.max(np.abs(resid), axis=1) / (np.sum(np.abs(Jz), axis=1) * np.max(np.abs(JJ)) + 1e-300))
        print(b, name, "componentwise %.2e normwise %.2e rowwise %.2e  maxJ %.2e" % (comp, norm, rowwise, np.abs(JJ).max()))
