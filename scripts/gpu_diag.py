import sys, time, numpy as np
sys.path.insert(0, ".")
from mcp_b200 import InteriorPoint, solve, problems, solve_jacobian_θ, capi
from mcp_b200.solver import _handle
from oracle.ir_eval import OracleMCP
from oracle import ip_oracle as O

lane = problems.lane_change_game(); mcp = lane.mcp
om = OracleMCP(mcp.ir)
Θ = problems.lane_change_thetas(12, seed=5)
sol = solve(InteriorPoint(), mcp, Θ, tol=1e-6)
J = solve_jacobian_θ(mcp, sol, Θ)
for b in range(Θ.shape[1]):
    if sol.status[b] != 0: print(b, "failed"); continue
    ref = O.Solution("solved", sol.x[:, b], sol.y[:, b], sol.s[:, b], 0.0, float(sol.ϵ[b]), 0)
    Jref = O.solve_jacobian_theta(om, ref, Θ[:, b])
    Jz = om.JFz(ref.x, ref.y, ref.s, Θ[:, b], ref.eps).toarray()
    print(b, "scale %.3e maxdiff %.3e rel %.3e cond %.3e" % (np.abs(Jref).max(), np.abs(J[:,:,b]-Jref).max(), np.abs(J[:,:,b]-Jref).max()/max(1,np.abs(Jref).max()), np.linalg.cond(Jz)))

print("fp64 peak TFLOP/s", capi.measure_fp64_peak())
h = _handle(mcp)
print(h.info())
for B in (2048, 16384, 65536):
    Θ = problems.lane_change_thetas(B, seed=1)
    for rep in range(2):
        t = time.time(); sol = solve(InteriorPoint(), mcp, Θ, tol=1e-6); wall = time.time() - t
        tm = h.timing()
    ns = int((sol.status == 0).sum())
    print(f"B={B} wall {wall*1e3:.1f} ms kernel {tm['kernel_ms']:.1f} ms h2d {tm['h2d_ms']:.2f} d2h {tm['d2h_ms']:.2f} solved {ns}/{B} steps {tm['newton_steps']}"
          f" -> {ns/(tm['kernel_ms']*1e-3):.0f} converged solves/s (kernel), {tm['newton_steps']/(tm['kernel_ms']*1e-3)/1e6:.2f} M newton steps/s")
    st = sol.newton_steps
    print("  steps solved mean %.1f, failed mean %.1f, frac failed %.3f" % (st[sol.status==0].mean(), st[sol.status!=0].mean() if (sol.status!=0).any() else 0, (sol.status!=0).mean()))
rq = problems.readme_qp()
hq = _handle(rq)
Θ = problems.readme_qp_thetas(1<<20, seed=1)
for rep in range(2):
    t = time.time(); sol = solve(InteriorPoint(), rq, Θ); wall = time.time()-t
print("readme 1M: wall %.1f ms, kernel %.1f ms, solved %d" % (wall*1e3, hq.timing()['kernel_ms'], (sol.status==0).sum()))
