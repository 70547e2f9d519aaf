"""TEST INFRASTRUCTURE ONLY — independent CPU evaluator of the MCP-IR tape.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs
may import this package; the product path (`mcp_b200/`, `csrc/`) never does.

Builds the reference's compiled callables from an IR object (duck-typed: any object with the
`MCPIR` attributes):

* `F(x, y, s, θ, ϵ)`       ↔ `mcp.F!`      (`/root/reference/src/mcp.jl:76-95`):  [G; H - s; s∘y - ϵ]
* `JFz(x, y, s, θ, ϵ)`     ↔ `mcp.∇F_z!`   (`src/mcp.jl:97-120`): scipy CSC n×n
* `JFt(x, y, s, θ, ϵ)`     ↔ `mcp.∇F_θ!`   (`src/mcp.jl:122-148`): scipy CSC n×nθ

The tape is turned into straight-line Python source and `exec`-compiled (≈50 ns per node), which
is what makes a literal per-instance Newton loop affordable in tests.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp

_FMT = {
    4: "v{i} = v{a} + v{b}", 5: "v{i} = v{a} - v{b}", 6: "v{i} = v{a} * v{b}", 7: "v{i} = v{a} / v{b}",
    8: "v{i} = -v{a}", 9: "v{i} = _sqrt(v{a})", 10: "v{i} = _exp(v{a})", 11: "v{i} = _log(v{a})",
    12: "v{i} = _sin(v{a})", 13: "v{i} = _cos(v{a})", 14: "v{i} = v{a} ** {b}",
}


def _compile(ir, roots, name, extended=False):
    """Python function (x, y, th) -> list of the values of `roots`.  `extended`: transcendental nodes go through
    numpy (dtype-preserving) instead of `math`, so `numpy.longdouble` inputs stay in extended precision."""
    roots = [int(r) for r in roots]
    need = np.zeros(len(ir.op), dtype=bool)
    stack = list(set(roots))
    while stack:
        n = stack.pop()
        if need[n]:
            continue
        need[n] = True
        op = int(ir.op[n])
        if op in (4, 5, 6, 7):
            stack += [int(ir.a[n]), int(ir.b[n])]
        elif op >= 8:
            stack.append(int(ir.a[n]))
    lines = [f"def {name}(x, y, th):"]
    for n in np.nonzero(need)[0]:
        op, a, b = int(ir.op[n]), int(ir.a[n]), int(ir.b[n])
        if op == 0:
            lines.append(f"    v{n} = {float(ir.consts[a])!r}")
        elif op == 1:
            lines.append(f"    v{n} = x[{a}]")
        elif op == 2:
            lines.append(f"    v{n} = y[{a}]")
        elif op == 3:
            lines.append(f"    v{n} = th[{a}]")
        else:
            lines.append("    " + _FMT[op].format(i=n, a=a, b=b))
    lines.append("    return [" + ", ".join(f"v{r}" for r in roots) + "]")
    src = "\n".join(lines).replace("inf", "_inf").replace("nan", "_nan")
    env = {"_sqrt": math.sqrt, "_exp": math.exp, "_log": math.log, "_sin": math.sin, "_cos": math.cos,
           "_inf": math.inf, "_nan": math.nan}
    if extended:
        env.update(_sqrt=np.sqrt, _exp=np.exp, _log=np.log, _sin=np.sin, _cos=np.cos)
    exec(compile(src, f"<ir:{name}>", "exec"), env)
    return env[name]


class OracleMCP:
    """CPU counterpart of `struct PrimalDualMCP` (`src/mcp.jl:13-24`)."""

    def __init__(self, ir, extended=False):
        self.ir = ir
        self.extended = extended
        self.nx, self.ny, self.ntheta = ir.nx, ir.ny, ir.ntheta
        self.unconstrained_dimension, self.constrained_dimension = ir.nx, ir.ny
        self.n = ir.nx + 2 * ir.ny
        self._gh = _compile(ir, ir.gh_nodes, "gh", extended)
        self._jz = _compile(ir, ir.jz_nodes, "jz", extended)
        self._jt = _compile(ir, ir.jt_nodes, "jt", extended) if ir.jt_nodes is not None else None
        nx, ny = self.nx, self.ny
        k = np.arange(ny)
        # ∇F_z pattern: IR block, then -I, diag(s), diag(y)
        self._rows = np.concatenate([ir.jz_rows, nx + k, nx + ny + k, nx + ny + k]).astype(np.int64)
        self._cols = np.concatenate([ir.jz_cols, nx + ny + k, nx + k, nx + ny + k]).astype(np.int64)

    def F(self, x, y, s, theta, eps):
        gh = np.asarray(self._gh(x, y, theta), dtype=np.float64)
        nx = self.nx
        return np.concatenate([gh[:nx], gh[nx:] - s, s * y - eps])

    def JFz(self, x, y, s, theta, eps):
        vals = np.concatenate([np.asarray(self._jz(x, y, theta), dtype=np.float64),
                               -np.ones(self.ny), s, y])
        return sp.csc_matrix((vals, (self._rows, self._cols)), shape=(self.n, self.n))

    def JFt(self, x, y, s, theta, eps):
        if self._jt is None:
            # `src/AutoDiff.jl:19-23`
            raise ValueError("Missing sensitivities. Set `compute_sensitivities = true` when "
                             "constructing the PrimalDualMCP.")
        vals = np.asarray(self._jt(x, y, theta), dtype=np.float64)
        return sp.csc_matrix((vals, (self.ir.jt_rows.astype(np.int64), self.ir.jt_cols.astype(np.int64))),
                             shape=(self.n, self.ntheta))
