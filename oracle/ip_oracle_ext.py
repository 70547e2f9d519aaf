"""TEST INFRASTRUCTURE ONLY — the interior-point loop of `oracle/ip_oracle.py` (`/root/reference/src/solver.jl:35-138`)
carried out in EXTENDED precision (`numpy.longdouble`: the x87 80-bit format here, ε ≈ 1.1e-19), as the arbiter when
two FP64 implementations (the CUDA kernels, the C oracle) disagree beyond the parity bar: whichever is farther from
this trajectory is the one whose rounding is off.

* iterate, residual, Jacobian entries, step, linesearch predicate, ϵ schedule: all `longdouble`;
* the Newton system `(∇F + tol·I) δz = −F` on the FULL n×n KKT matrix (no condensation — nothing is shared with
  the product's reduced system): factorised once in FP64 by SuperLU, then refined with residuals formed in
  `longdouble` until the correction stalls (classical mixed-precision iterative refinement), so δz is accurate to
  ≈ cond·1e-19 instead of cond·1e-16.

Discrete decisions (inner-iteration counts, step halvings) are therefore those of (near-)exact arithmetic.
Only `tests/` and `scripts/adjudicate_parity.py` may import this.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .ip_oracle import Solution

LD = np.longdouble


def _ftb(v, d, min_step):
    """`fraction_to_the_boundary_linesearch` — `src/solver.jl:127-138` (τ = 0.995, decay = 0.5) in longdouble."""
    alpha = LD(1.0)
    c = LD(1.0) - LD(0.995)
    while np.any(v + alpha * d < c * v):          # :129
        if alpha < min_step:                      # :130
            return LD(np.nan)
        alpha = alpha * LD(0.5)                   # :134
    return alpha


def _refined_solve(rows, cols, vals, n, b, max_refine=8):
    """Solve A x = b with A, b in longdouble: FP64 sparse LU + longdouble residual refinement."""
    A64 = sp.csc_matrix((vals.astype(np.float64), (rows, cols)), shape=(n, n))
    lu = spla.splu(A64)
    x = lu.solve(b.astype(np.float64)).astype(LD)
    last = None
    for _ in range(max_refine):
        r = b.copy()
        np.subtract.at(r, rows, vals * x[cols])   # r = b − A x, every product and sum in longdouble
        nr = float(np.max(np.abs(r))) if n else 0.0
        if last is not None and nr >= 0.5 * last:
            break
        last = nr
        x = x + lu.solve(r.astype(np.float64)).astype(LD)
    return x


def solve_interior_point_ext(omcp, theta, x0=None, y0=None, s0=None, tol=1e-4, max_inner_iters=20,
                             max_outer_iters=50, tightening_rate=0.1, loosening_rate=0.5, min_stepsize=1e-4):
    """`omcp` must be an `OracleMCP(ir, extended=True)`.  Same control flow, line for line, as
    `ip_oracle.solve_interior_point` (`src/solver.jl:63-121`)."""
    nx, ny = omcp.nx, omcp.ny
    n = nx + 2 * ny
    th = np.asarray(theta, dtype=np.float64).astype(LD)
    x = np.zeros(nx, LD) if x0 is None else np.asarray(x0, dtype=np.float64).astype(LD)
    y = np.ones(ny, LD) if y0 is None else np.asarray(y0, dtype=np.float64).astype(LD)
    s = np.ones(ny, LD) if s0 is None else np.asarray(s0, dtype=np.float64).astype(LD)
    k = np.arange(ny)
    rows = np.concatenate([omcp.ir.jz_rows, nx + k, nx + ny + k, nx + ny + k, np.arange(n)]).astype(np.int64)
    cols = np.concatenate([omcp.ir.jz_cols, nx + ny + k, nx + k, nx + ny + k, np.arange(n)]).astype(np.int64)
    tolL, eps = LD(tol), LD(1.0)
    kkt = LD(np.inf)
    status, outer, steps, per_outer = "solved", 1, 0, []
    while kkt > tolL and eps > tolL and outer < max_outer_iters:            # :71
        inner, status = 1, "solved"                                         # :72-73
        while kkt > eps and inner < max_inner_iters:                        # :75
            gh = np.array(omcp._gh(x, y, th), dtype=LD)
            F = np.concatenate([gh[:nx], gh[nx:] - s, s * y - eps])         # :79, mcp.jl:76-80
            jz = np.array(omcp._jz(x, y, th), dtype=LD)
            vals = np.concatenate([jz, -np.ones(ny, LD), s, y, np.full(n, tolL)])   # :80-81 (∇F + tol·I)
            try:
                with np.errstate(all="ignore"):
                    dz = _refined_solve(rows, cols, vals, n, -F)            # :82-83
                ok = bool(np.all(np.isfinite(dz)))
            except RuntimeError:
                ok = False
            if not ok:                                                      # :84-88
                status = "failed"
                break
            dx, dy, ds = dz[:nx], dz[nx:nx + ny], dz[nx + ny:]
            a_s = _ftb(s, ds, LD(min_stepsize))                             # :93
            a_y = _ftb(y, dy, LD(min_stepsize))                             # :94
            if np.isnan(a_s) or np.isnan(a_y):                              # :96-100
                status = "failed"
                break
            x = x + a_s * dx                                                # :103
            s = s + a_s * ds                                                # :104
            y = y + a_y * dy                                                # :105
            kkt = np.max(np.abs(F)) if n else LD(0)                         # :107
            inner += 1                                                      # :108
            steps += 1
        eps = eps * ((LD(1) - np.exp(LD(-tightening_rate) * inner)) if status == "solved"
                     else (LD(1) + np.exp(LD(-loosening_rate) * inner)))    # :111-113
        outer += 1                                                          # :114
        per_outer.append(inner)
    if outer == max_outer_iters:                                            # :117-119
        status = "failed"
    return Solution(status, x, y, s, float(kkt), float(eps), outer, steps, tuple(per_outer), ())
