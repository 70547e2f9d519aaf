"""MCP-IR: the flat problem description that crosses the C-ABI (`include/mcpb200.h`).

One `MCPIR` holds, for a problem  G(x,y;θ)=0, 0 ≤ H(x,y;θ) ⟂ y ≥ 0  (`/root/reference/src/mcp.jl:1-12`):

* an SSA tape (`op, a, b`, constant pool) in topological order,
* the tape nodes of the `nx` rows of G and the `ny` rows of H,
* the sparse Jacobian of [G; H] w.r.t. [x; y] (CSC order: `rows, cols, node`),
* optionally the sparse Jacobian of [G; H] w.r.t. θ.

The slack/barrier rows `H - s` and `s∘y - ϵ` of the reference's F (`src/mcp.jl:76-80`) and their
Jacobian blocks `-I`, `diag(s)`, `diag(y)` are *not* in the IR: they are structural and every
consumer (CUDA template, oracles) adds them itself.  `full_jacobian_pattern()` gives the
reference's `∇F_z!.rows/.cols` view (`src/mcp.jl:97-120`) for parity checks.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import trace as T


@dataclass
class MCPIR:
    nx: int
    ny: int
    ntheta: int
    op: np.ndarray          # int32[n_nodes]
    a: np.ndarray           # int32[n_nodes]
    b: np.ndarray           # int32[n_nodes]
    consts: np.ndarray      # float64[n_consts]
    gh_nodes: np.ndarray    # int32[nx+ny]   nodes of [G; H]
    jz_rows: np.ndarray     # int32[nnz]     row in [G; H]        (CSC order)
    jz_cols: np.ndarray     # int32[nnz]     col in [x; y]
    jz_nodes: np.ndarray    # int32[nnz]
    jt_rows: Optional[np.ndarray] = None   # Jacobian wrt θ (None ⇔ compute_sensitivities=false)
    jt_cols: Optional[np.ndarray] = None
    jt_nodes: Optional[np.ndarray] = None
    meta: dict = field(default_factory=dict)

    @property
    def n(self) -> int:
        return self.nx + 2 * self.ny

    @property
    def has_sensitivities(self) -> bool:
        return self.jt_nodes is not None

    def full_jacobian_pattern(self):
        """rows/cols of the reference's n×n ∇F_z in CSC order (`src/mcp.jl:110`), plus for each
        entry its source: k ≥ 0 → IR Jacobian entry k; -1 → the `-I` block; -2 → diag(s) (∂(s∘y)/∂y);
        -3 → diag(y) (∂(s∘y)/∂s)."""
        nx, ny = self.nx, self.ny
        r = [self.jz_rows.astype(np.int64)]
        c = [self.jz_cols.astype(np.int64)]
        src = [np.arange(len(self.jz_rows), dtype=np.int64)]
        k = np.arange(ny, dtype=np.int64)
        r += [nx + k, nx + ny + k, nx + ny + k]
        c += [nx + ny + k, nx + k, nx + ny + k]
        src += [np.full(ny, -1), np.full(ny, -2), np.full(ny, -3)]
        r, c, src = np.concatenate(r), np.concatenate(c), np.concatenate(src)
        order = np.lexsort((r, c))
        return r[order], c[order], src[order]

    def constant_entries(self) -> np.ndarray:
        """Indices (into the IR Jacobian) of entries that do not depend on z = (x, y):
        the analogue of `∇F_z!.constant_entries` (`src/mcp.jl:111-118`)."""
        dep = _depends(self, self.jz_nodes, (T.OP_X, T.OP_Y))
        return np.nonzero(~dep)[0]


def _depends(ir: MCPIR, roots: Sequence[int], leaf_ops) -> np.ndarray:
    leaf_ops = set(leaf_ops)
    flag = np.zeros(len(ir.op), dtype=bool)
    for n in range(len(ir.op)):
        op = int(ir.op[n])
        if op in (T.OP_CONST, T.OP_X, T.OP_Y, T.OP_THETA):
            flag[n] = op in leaf_ops
        elif op in (T.OP_ADD, T.OP_SUB, T.OP_MUL, T.OP_DIV):
            flag[n] = flag[ir.a[n]] or flag[ir.b[n]]
        else:
            flag[n] = flag[ir.a[n]]
    return flag[np.asarray(roots, dtype=np.int64)]


def build_ir(g: T.Graph, G: Sequence[T.Expr], H: Sequence[T.Expr], nx: int, ny: int, ntheta: int,
             compute_sensitivities: bool = True) -> MCPIR:
    """Differentiate the traced G, H and pack everything reachable into an `MCPIR`.

    Follows `src/mcp.jl:55-150`: z = [x; y; s] (`:74`), Jacobian wrt z (`:97-120`) and wrt θ
    (`:122-148`, only when `compute_sensitivities`).
    """
    assert len(G) == nx and len(H) == ny
    outs = list(G) + list(H)
    wrt_z = [(T.OP_X, i) for i in range(nx)] + [(T.OP_Y, i) for i in range(ny)]
    jz_rows, jz_cols, jz_exprs = T.sparse_jacobian(g, outs, wrt_z)
    if compute_sensitivities:
        wrt_t = [(T.OP_THETA, i) for i in range(ntheta)]
        jt_rows, jt_cols, jt_exprs = T.sparse_jacobian(g, outs, wrt_t)
    else:
        jt_rows = jt_cols = None
        jt_exprs = []

    # prune to reachable nodes, renumber (order preserved ⇒ still topological)
    roots = [e.id for e in outs] + [e.id for e in jz_exprs] + [e.id for e in jt_exprs]
    keep = np.zeros(len(g.op), dtype=bool)
    stack = list(set(roots))
    while stack:
        n = stack.pop()
        if keep[n]:
            continue
        keep[n] = True
        op = g.op[n]
        if op in (T.OP_ADD, T.OP_SUB, T.OP_MUL, T.OP_DIV):
            stack.append(g.a[n])
            stack.append(g.b[n])
        elif op not in (T.OP_CONST, T.OP_X, T.OP_Y, T.OP_THETA):
            stack.append(g.a[n])
    old_ids = np.nonzero(keep)[0]
    new_id = -np.ones(len(g.op), dtype=np.int64)
    new_id[old_ids] = np.arange(len(old_ids))
    op = np.array([g.op[i] for i in old_ids], dtype=np.int32)
    a = np.array([g.a[i] for i in old_ids], dtype=np.int64)
    b = np.array([g.b[i] for i in old_ids], dtype=np.int64)
    # constants: keep only the used ones
    is_const = op == T.OP_CONST
    used_c = np.unique(a[is_const])
    c_map = -np.ones(max(len(g.consts), 1), dtype=np.int64)
    c_map[used_c] = np.arange(len(used_c))
    consts = np.array([g.consts[i] for i in used_c], dtype=np.float64)
    a[is_const] = c_map[a[is_const]]
    binary = np.isin(op, (T.OP_ADD, T.OP_SUB, T.OP_MUL, T.OP_DIV))
    unary = ~binary & ~np.isin(op, (T.OP_CONST, T.OP_X, T.OP_Y, T.OP_THETA))
    a[binary | unary] = new_id[a[binary | unary]]
    b[binary] = new_id[b[binary]]
    b[~binary & (op != T.OP_POWI)] = -1

    def ids(exprs):
        return np.array([new_id[e.id] for e in exprs], dtype=np.int32)

    return MCPIR(
        nx=nx, ny=ny, ntheta=ntheta,
        op=op, a=a.astype(np.int32), b=b.astype(np.int32), consts=consts,
        gh_nodes=ids(outs),
        jz_rows=jz_rows, jz_cols=jz_cols, jz_nodes=ids(jz_exprs),
        jt_rows=jt_rows, jt_cols=jt_cols,
        jt_nodes=ids(jt_exprs) if compute_sensitivities else None,
    )
