"""cfg2 (QP 100×100) dense kernel: parity against the C oracle and kernel time of the build selected by MCPB200_DEFS
(D3_SYM=0: pivoted LU only; D3_SYM_LOOKAHEAD=0: LDLᵀ without look-ahead; default: LDLᵀ with look-ahead).
usage: python scripts/gpu_qp_sym.py [parity instances] [timing batches …]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, ".")
from mcp_b200 import InteriorPoint, problems, solve
from mcp_b200.solver import _handle

NPAR = int(sys.argv[1]) if len(sys.argv) > 1 else 128
BATCHES = [int(a) for a in sys.argv[2:]] or [8192]
RTOL = 1e-6


def rel_err(a, b):
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b)))))


def parity(mcp, Θ, tag):
    from oracle import c_oracle as CO
    sol = solve(InteriorPoint(), mcp, Θ, tol=1e-6)
    ref = CO.solve_batch(mcp.ir, Θ, tol=1e-6)
    bad, worst = [], 0.0
    for b in range(Θ.shape[1]):
        same = int(sol.status[b]) == int(ref.status[b])
        if same and ref.status[b] == 0:
            e = max(rel_err(sol.x[:, b], ref.x[:, b]), rel_err(sol.y[:, b], ref.y[:, b]), rel_err(sol.s[:, b], ref.s[:, b]))
            worst = max(worst, e)
            same = abs(int(sol.newton_steps[b]) - int(ref.newton_steps[b])) <= 1 and e <= RTOL
        if not same:
            bad.append((b, int(ref.status[b]), int(sol.status[b]), int(ref.newton_steps[b]), int(sol.newton_steps[b])))
    return {"case": tag, "instances": int(Θ.shape[1]), "match": int(Θ.shape[1] - len(bad)), "solved": int((sol.status == 0).sum()),
            "worst_rel_err": worst, "steps_gpu": int(sol.newton_steps.sum()), "steps_oracle": int(ref.newton_steps.sum()), "bad": bad[:6]}


out = {"defs": os.environ.get("MCPB200_DEFS", "")}
qp = problems.random_qp(100, 100)
Θ = problems.random_qp_thetas(NPAR, seed=11)
out["parity"] = [parity(qp, Θ, "random convex QP (symmetric M)")]
# G_x not symmetric: the symmetry check must send the instance to the pivoted LU
Θa = Θ[:, :16].copy()
Θa[3 + 100 * 7] += 0.25      # vec(M) is column-major: M[r, c] = θ[r + 100 c]
Θa[50 + 100 * 2] -= 0.125
out["parity"].append(parity(qp, Θa, "asymmetric M (LU fallback by the symmetry check)"))
# symmetric but indefinite M: a non-positive pivot must send the step to the pivoted LU
Θi = Θ[:, :16].copy()
for i in (5, 40):
    Θi[i + 100 * i] -= 30.0
out["parity"].append(parity(qp, Θi, "indefinite symmetric M (LU fallback on a non-positive pivot)"))
# dense A (70 % non-zeros): the dense form of the Schur accumulation
Θd = problems.random_qp_thetas(16, seed=3, sparsity_rate=0.3)
out["parity"].append(parity(qp, Θd, "dense A and M (dense Schur accumulation)"))
h = _handle(qp)
info = h.info()
for B in BATCHES:
    Θb = problems.random_qp_thetas(B, seed=1)
    for _ in range(2):
        sol = solve(InteriorPoint(), qp, Θb, tol=1e-6)
    tm = h.timing()
    out[f"B{B}"] = {"kernel_ms": tm["kernel_ms"], "pass0_ms": tm["pass0_ms"], "deferred": tm["deferred"], "solved": int((sol.status == 0).sum()),
                    "newton_steps": int(sol.newton_steps.sum()), "solves_per_s": float((sol.status == 0).sum() / (tm["kernel_ms"] * 1e-3)),
                    "pass0_steps_per_s": None}
print(json.dumps(out, indent=1))
