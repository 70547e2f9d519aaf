"""TEST INFRASTRUCTURE ONLY — literal CPU restatement of the reference's interior-point hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this; nothing under `mcp_b200/` or `csrc/` does.

PARITY PIN STATUS: the Julia reference cannot run in this image (no `julia`), and its own tests
hold no stored vectors — only analytic assertions (`/root/reference/test/runtests.jl:30-38,80-84,
112-115`).  Those known answers are checked in `tests/test_oracle.py`; beyond them (1e-6 agreement on
lane-change / QP100) the parity is **unpinned by the reference** and rests on this restatement being
line-by-line faithful.  Every block below cites the line it restates.  What round 2 added to shrink that
gap (none of it can replace a Julia run): the PROBLEMS this loop is run on are pinned independently of the
product's tracer (`oracle/independent_problems.py`, `tests/test_independent_definitions.py`: F, ∇F_z, ∇F_θ of
the lane-change and masked games equal a second, autograd-based restatement of the reference's sources to
1e-9), and the LOOP's floating-point trajectory is checked against an extended-precision run of the same loop on
the full KKT system (`oracle/ip_oracle_ext.py`; `profiles/r2_adjudicate_*.json`).

Third-party arithmetic the reference delegates to (not under /root/reference):
* UMFPACK sparse LU via LinearSolve.jl ≥2.38 (`src/solver.jl:50,61,83`)  → here SuperLU (`splu`),
  also an exact sparse LU with partial pivoting; results agree to rounding (~1e-13).
* LAPACK geqp3 column-pivoted QR (`src/AutoDiff.jl:39`) → here `scipy.linalg.qr(pivoting=True)`,
  the same LAPACK routine.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla


@dataclass
class Solution:
    status: str          # "solved" | "failed"   (`src/solver.jl:69,86,98,118`)
    x: np.ndarray
    y: np.ndarray
    s: np.ndarray
    kkt_error: float
    eps: float           # ϵ after the last update (`src/solver.jl:111-113,121`)
    outer_iters: int
    # extras for parity diagnostics (not part of the reference's NamedTuple)
    newton_steps: int = 0
    inner_iters_per_outer: tuple = ()
    alphas: tuple = ()


def fraction_to_the_boundary_linesearch(v, d, tau=0.995, decay=0.5, tol=1e-4):
    """`src/solver.jl:127-138`, including its floating-point predicate and NaN semantics."""
    alpha = 1.0
    c = 1 - tau
    while np.any(v + alpha * d < c * v):      # :129 (empty v ⇒ False ⇒ returns 1.0)
        if alpha < tol:                       # :130 — tested BEFORE halving
            return math.nan
        alpha *= decay                        # :134
    return alpha


def solve_interior_point(mcp, theta, x0=None, y0=None, s0=None, tol=1e-4, max_inner_iters=20,
                         max_outer_iters=50, tightening_rate=0.1, loosening_rate=0.5,
                         min_stepsize=1e-4, record=False):
    """`solve(::InteriorPoint, mcp, θ; …)` — `src/solver.jl:35-122`; defaults from `:39-49`."""
    nx, ny = mcp.unconstrained_dimension, mcp.constrained_dimension
    n = nx + 2 * ny
    theta = np.asarray(theta, dtype=np.float64)
    # :39-41 defaults; :64-66 x = x₀ aliases (we copy: the oracle must not clobber test inputs)
    x = np.zeros(nx) if x0 is None else np.array(x0, dtype=np.float64)
    y = np.ones(ny) if y0 is None else np.array(y0, dtype=np.float64)
    s = np.ones(ny) if s0 is None else np.array(s0, dtype=np.float64)
    eye = sp.identity(n, format="csc")

    eps = 1.0                                  # :67
    kkt_error = math.inf                       # :68
    status = "solved"                          # :69
    outer_iters = 1                            # :70
    steps = 0
    per_outer, alphas = [], []
    while kkt_error > tol and eps > tol and outer_iters < max_outer_iters:   # :71
        inner_iters = 1                        # :72
        status = "solved"                      # :73
        while kkt_error > eps and inner_iters < max_inner_iters:             # :75
            F = mcp.F(x, y, s, theta, eps)                                   # :79
            J = mcp.JFz(x, y, s, theta, eps)                                 # :80
            A = (J + tol * eye).tocsc()                                      # :81  tol on ALL n diagonals
            b = -F                                                           # :82
            try:                                                             # :83
                with np.errstate(all="ignore"):
                    dz = spla.splu(A).solve(b)
                ok = True
            except RuntimeError:                                             # exactly singular factor
                ok = False
            if not ok:                                                       # :84-88
                status = "failed"
                break
            dx, dy, ds = dz[:nx], dz[nx:nx + ny], dz[nx + ny:]               # :56-59
            a_s = fraction_to_the_boundary_linesearch(s, ds, tol=min_stepsize)   # :93
            a_y = fraction_to_the_boundary_linesearch(y, dy, tol=min_stepsize)   # :94
            if math.isnan(a_s) or math.isnan(a_y):                           # :96-100
                status = "failed"
                break
            x = x + a_s * dx                                                 # :103  x uses α_s
            s = s + a_s * ds                                                 # :104
            y = y + a_y * dy                                                 # :105
            kkt_error = float(np.max(np.abs(F))) if n else 0.0               # :107  pre-step residual
            if math.isnan(kkt_error) or np.any(np.isnan(F)):
                kkt_error = math.nan                                         # norm(F, Inf) propagates NaN
            inner_iters += 1                                                 # :108
            steps += 1
            if record:
                alphas.append((a_s, a_y))
        eps *= (1 - math.exp(-tightening_rate * inner_iters)) if status == "solved" \
            else (1 + math.exp(-loosening_rate * inner_iters))               # :111-113
        outer_iters += 1                                                     # :114
        per_outer.append(inner_iters)
    if outer_iters == max_outer_iters:                                       # :117-119
        status = "failed"
    return Solution(status, x, y, s, kkt_error, eps, outer_iters, steps, tuple(per_outer), tuple(alphas))


def solve_jacobian_theta(mcp, sol, theta):
    """`_solve_jacobian_θ` — `src/AutoDiff.jl:18-40`: ∂z/∂θ = qr(-∇F_z, ColumnNorm()) \\ ∇F_θ at the
    returned (x, y, s, ϵ); note NO tol·I here (`:27-31`)."""
    theta = np.asarray(theta, dtype=np.float64)
    Jz = mcp.JFz(sol.x, sol.y, sol.s, theta, sol.eps).toarray()              # :27-31, collect at :39
    Jt = mcp.JFt(sol.x, sol.y, sol.s, theta, sol.eps).toarray()              # :33-37 (raises if missing, :19-23)
    Q, R, P = sla.qr(-Jz, pivoting=True)                                     # :39  geqp3
    rhs = Q.T @ Jt
    # Julia's `\` on a QRPivoted of a full-rank square matrix is the exact solve
    out = np.zeros_like(rhs)
    out[P, :] = sla.solve_triangular(R, rhs)
    return out


def vjp_theta(mcp, sol, theta, dx, dy, ds):
    """Pullback of `rrule(solve, …)` — `src/AutoDiff.jl:59-76`: ∂θ = Σ_b (∂z/∂θ)[b,:]ᵀ ∂l/∂b."""
    dzdt = solve_jacobian_theta(mcp, sol, theta)                             # :60
    nx, ny = mcp.unconstrained_dimension, mcp.constrained_dimension
    return (dzdt[:nx].T @ dx + dzdt[nx:nx + ny].T @ dy + dzdt[nx + ny:].T @ ds)   # :65-75


def jvp_theta(mcp, sol, theta, theta_p):
    """Forward rule — `src/AutoDiff.jl:84-117`: z_p = ∂z∂θ · θ_p (`:98`).  Returns (x_p, y_p, s_p).

    (The reference then packs `s` Duals around `solution.y` *values* (`:109-114`, a bug); the
    partials — what this returns — are unaffected.)"""
    dzdt = solve_jacobian_theta(mcp, sol, theta)
    zp = dzdt @ np.asarray(theta_p, dtype=np.float64)
    nx, ny = mcp.unconstrained_dimension, mcp.constrained_dimension
    return zp[:nx], zp[nx:nx + ny], zp[nx + ny:]
