"""Symbolic tracing of user callables G(x, y; θ), H(x, y; θ) into the MCP-IR.

This is the B200 framework's counterpart of the reference's symbolic layer
(`/root/reference/src/mcp.jl:27-52` traces the callables on symbolic vectors made by
`SymbolicTracingUtils.make_variables`; `src/mcp.jl:97-148` takes `sparse_jacobian`s of the
traced residual).  The reference lowers the traced expressions to Julia closures; here they are
lowered to a flat SSA tape (the *MCP-IR*) that the C-ABI library (`csrc/`) turns into CUDA
device functions spliced into the sm_100a kernel template.

Design
------
* `Graph` is a hash-consed expression DAG: every distinct (op, a, b) triple exists once, so
  common sub-expressions are shared for free and structural zeros are detected by the algebraic
  simplifications in `Graph.binary` (x*0, x+0, x-x, constant folding, ...).
* `Expr` is a thin handle with Python operator overloading, so user callables written for numpy
  arrays (`M @ x - theta - A.T @ y`) trace unchanged on object arrays of `Expr`.
* `Graph.gradient` is reverse-mode symbolic differentiation of ONE output over its own
  sub-DAG; `sparse_jacobian` runs it per output row and keeps the structurally non-zero
  entries, returned in CSC order like `SparseArrays.findnz` does (`src/mcp.jl:110`).

The IR op-codes are shared with `include/mcpb200.h` (enum mcpb200_op) and `oracle/`.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np

# --- op-codes (keep in sync with include/mcpb200.h) ------------------------------------------
OP_CONST = 0   # a = index into the constant pool
OP_X = 1       # a = index into x   (unconstrained variable)
OP_Y = 2       # a = index into y   (constrained variable)
OP_THETA = 3   # a = index into θ
OP_ADD = 4
OP_SUB = 5
OP_MUL = 6
OP_DIV = 7
OP_NEG = 8
OP_SQRT = 9
OP_EXP = 10
OP_LOG = 11
OP_SIN = 12
OP_COS = 13
OP_POWI = 14   # a ** b with b an integer literal (stored in the b slot)

_BINARY = {OP_ADD, OP_SUB, OP_MUL, OP_DIV}
_UNARY = {OP_NEG, OP_SQRT, OP_EXP, OP_LOG, OP_SIN, OP_COS}
_LEAF = {OP_CONST, OP_X, OP_Y, OP_THETA}

OP_NAMES = {
    OP_CONST: "const", OP_X: "x", OP_Y: "y", OP_THETA: "theta", OP_ADD: "add", OP_SUB: "sub",
    OP_MUL: "mul", OP_DIV: "div", OP_NEG: "neg", OP_SQRT: "sqrt", OP_EXP: "exp", OP_LOG: "log",
    OP_SIN: "sin", OP_COS: "cos", OP_POWI: "powi",
}


class Graph:
    """Hash-consed expression DAG; node ids are topologically ordered by construction."""

    def __init__(self) -> None:
        self.op: List[int] = []
        self.a: List[int] = []
        self.b: List[int] = []
        self.consts: List[float] = []
        self._memo: Dict[Tuple[int, int, int], int] = {}
        self._const_memo: Dict[float, int] = {}

    # -- node construction ---------------------------------------------------------------
    def _node(self, op: int, a: int, b: int = -1) -> int:
        key = (op, a, b)
        nid = self._memo.get(key)
        if nid is None:
            nid = len(self.op)
            self.op.append(op)
            self.a.append(a)
            self.b.append(b)
            self._memo[key] = nid
        return nid

    def const(self, value: float) -> "Expr":
        value = float(value)
        if value == 0.0:
            value = 0.0  # merge -0.0 and +0.0
        cid = self._const_memo.get(value)
        if cid is None:
            cid = len(self.consts)
            self.consts.append(value)
            self._const_memo[value] = cid
        return Expr(self, self._node(OP_CONST, cid))

    def variables(self, kind: str, n: int) -> np.ndarray:
        """`make_variables` analogue (`src/mcp.jl:37-39`): an object array of n leaf symbols."""
        op = {"x": OP_X, "y": OP_Y, "theta": OP_THETA, "θ": OP_THETA}[kind]
        out = np.empty(n, dtype=object)
        for i in range(n):
            out[i] = Expr(self, self._node(op, i))
        return out

    def const_value(self, nid: int):
        return self.consts[self.a[nid]] if self.op[nid] == OP_CONST else None

    def lift(self, v) -> "Expr":
        if isinstance(v, Expr):
            if v.g is not self:
                raise ValueError("expression belongs to a different Graph")
            return v
        return self.const(v)

    # -- algebra with simplification ---------------------------------------------------------
    def binary(self, op: int, ea: "Expr", eb: "Expr") -> "Expr":
        a, b = ea.id, eb.id
        ca, cb = self.const_value(a), self.const_value(b)
        if ca is not None and cb is not None:
            if op == OP_ADD:
                return self.const(ca + cb)
            if op == OP_SUB:
                return self.const(ca - cb)
            if op == OP_MUL:
                return self.const(ca * cb)
            if op == OP_DIV and cb != 0.0:
                return self.const(ca / cb)
        if op == OP_ADD:
            if ca == 0.0:
                return eb
            if cb == 0.0:
                return ea
            if self.op[b] == OP_NEG:
                return self.binary(OP_SUB, ea, Expr(self, self.a[b]))
            if self.op[a] == OP_NEG:
                return self.binary(OP_SUB, eb, Expr(self, self.a[a]))
            if a > b:  # commutative: canonical operand order improves sharing
                a, b = b, a
        elif op == OP_SUB:
            if cb == 0.0:
                return ea
            if ca == 0.0:
                return self.unary(OP_NEG, eb)
            if a == b:
                return self.const(0.0)
            if self.op[b] == OP_NEG:
                return self.binary(OP_ADD, ea, Expr(self, self.a[b]))
        elif op == OP_MUL:
            if ca == 0.0 or cb == 0.0:
                return self.const(0.0)
            if ca == 1.0:
                return eb
            if cb == 1.0:
                return ea
            if ca == -1.0:
                return self.unary(OP_NEG, eb)
            if cb == -1.0:
                return self.unary(OP_NEG, ea)
            if a > b:
                a, b = b, a
        elif op == OP_DIV:
            if ca == 0.0:
                return self.const(0.0)
            if cb == 1.0:
                return ea
            if cb == -1.0:
                return self.unary(OP_NEG, ea)
            if cb is not None and cb != 0.0 and math.isfinite(1.0 / cb) and (1.0 / cb) * cb == 1.0 \
                    and math.frexp(cb)[0] in (0.5, -0.5):
                # division by an exact power of two is an exact multiplication
                return self.binary(OP_MUL, ea, self.const(1.0 / cb))
        return Expr(self, self._node(op, a, b))

    def unary(self, op: int, ea: "Expr") -> "Expr":
        a = ea.id
        ca = self.const_value(a)
        if ca is not None:
            try:
                if op == OP_NEG:
                    return self.const(-ca)
                if op == OP_SQRT and ca >= 0:
                    return self.const(math.sqrt(ca))
                if op == OP_EXP:
                    return self.const(math.exp(ca))
                if op == OP_LOG and ca > 0:
                    return self.const(math.log(ca))
                if op == OP_SIN:
                    return self.const(math.sin(ca))
                if op == OP_COS:
                    return self.const(math.cos(ca))
            except OverflowError:
                pass
        if op == OP_NEG and self.op[a] == OP_NEG:
            return Expr(self, self.a[a])
        if op == OP_NEG and self.op[a] == OP_SUB:
            return self.binary(OP_SUB, Expr(self, self.b[a]), Expr(self, self.a[a]))
        return Expr(self, self._node(op, a))

    def powi(self, ea: "Expr", n: int) -> "Expr":
        n = int(n)
        if n == 0:
            return self.const(1.0)
        if n == 1:
            return ea
        if n == 2:
            return self.binary(OP_MUL, ea, ea)
        ca = self.const_value(ea.id)
        if ca is not None:
            return self.const(ca ** n)
        return Expr(self, self._node(OP_POWI, ea.id, n))

    # -- differentiation ---------------------------------------------------------------------
    def gradient(self, out: "Expr", wrt_ops: Iterable[int]) -> Dict[Tuple[int, int], "Expr"]:
        """Reverse-mode symbolic gradient of one scalar `out`.

        Returns {(leaf_op, leaf_index): d out / d leaf} for leaves whose op is in `wrt_ops`,
        with structurally zero derivatives omitted.
        """
        wrt_ops = set(wrt_ops)
        # collect the sub-DAG of `out`
        seen = set()
        stack = [out.id]
        while stack:
            n = stack.pop()
            if n in seen:
                continue
            seen.add(n)
            op = self.op[n]
            if op in _BINARY:
                stack.append(self.a[n])
                stack.append(self.b[n])
            elif op in _UNARY or op == OP_POWI:
                stack.append(self.a[n])
        adj: Dict[int, Expr] = {out.id: self.const(1.0)}
        result: Dict[Tuple[int, int], Expr] = {}

        def acc(n: int, e: Expr) -> None:
            if self.const_value(e.id) == 0.0:
                return
            cur = adj.get(n)
            adj[n] = e if cur is None else self.binary(OP_ADD, cur, e)

        for n in sorted(seen, reverse=True):  # ids are topological: parents have larger ids
            g = adj.get(n)
            if g is None:
                continue
            op = self.op[n]
            if op in _LEAF:
                if op in wrt_ops and self.const_value(g.id) != 0.0:
                    result[(op, self.a[n])] = g
                continue
            ea = Expr(self, self.a[n])
            if op == OP_ADD:
                acc(self.a[n], g)
                acc(self.b[n], g)
            elif op == OP_SUB:
                acc(self.a[n], g)
                acc(self.b[n], self.unary(OP_NEG, g))
            elif op == OP_MUL:
                eb = Expr(self, self.b[n])
                acc(self.a[n], self.binary(OP_MUL, g, eb))
                acc(self.b[n], self.binary(OP_MUL, g, ea))
            elif op == OP_DIV:
                eb = Expr(self, self.b[n])
                acc(self.a[n], self.binary(OP_DIV, g, eb))
                # d(a/b)/db = -(a/b)/b ; reuse node n itself for a/b
                acc(self.b[n], self.unary(OP_NEG, self.binary(OP_DIV, self.binary(OP_MUL, g, Expr(self, n)), eb)))
            elif op == OP_NEG:
                acc(self.a[n], self.unary(OP_NEG, g))
            elif op == OP_SQRT:
                acc(self.a[n], self.binary(OP_DIV, g, self.binary(OP_MUL, self.const(2.0), Expr(self, n))))
            elif op == OP_EXP:
                acc(self.a[n], self.binary(OP_MUL, g, Expr(self, n)))
            elif op == OP_LOG:
                acc(self.a[n], self.binary(OP_DIV, g, ea))
            elif op == OP_SIN:
                acc(self.a[n], self.binary(OP_MUL, g, self.unary(OP_COS, ea)))
            elif op == OP_COS:
                acc(self.a[n], self.unary(OP_NEG, self.binary(OP_MUL, g, self.unary(OP_SIN, ea))))
            elif op == OP_POWI:
                k = self.b[n]
                acc(self.a[n], self.binary(OP_MUL, g, self.binary(OP_MUL, self.const(float(k)), self.powi(ea, k - 1))))
            else:  # pragma: no cover
                raise AssertionError(f"unknown op {op}")
        return result

    def depends_on(self, roots: Sequence[int], leaf_ops: Iterable[int]) -> np.ndarray:
        """For each root node id: does its sub-DAG contain a leaf with op in `leaf_ops`?"""
        leaf_ops = set(leaf_ops)
        flag = np.zeros(len(self.op), dtype=bool)
        for n, op in enumerate(self.op):  # forward sweep works because ids are topological
            if op in _LEAF:
                flag[n] = op in leaf_ops
            elif op in _BINARY:
                flag[n] = flag[self.a[n]] or flag[self.b[n]]
            else:
                flag[n] = flag[self.a[n]]
        return flag[np.asarray(list(roots), dtype=np.int64)] if len(roots) else np.zeros(0, dtype=bool)


class Expr:
    """Handle to one node of a `Graph`, with numpy-friendly operator overloading."""

    __slots__ = ("g", "id")
    __array_priority__ = 1000  # make ndarray.__op__(Expr) defer to the element-wise path

    def __init__(self, g: Graph, nid: int) -> None:
        self.g = g
        self.id = nid

    def _b(self, op, other, swap=False):
        if isinstance(other, np.ndarray):
            return NotImplemented
        o = self.g.lift(other)
        return self.g.binary(op, o, self) if swap else self.g.binary(op, self, o)

    def __add__(self, o): return self._b(OP_ADD, o)
    def __radd__(self, o): return self._b(OP_ADD, o, True)
    def __sub__(self, o): return self._b(OP_SUB, o)
    def __rsub__(self, o): return self._b(OP_SUB, o, True)
    def __mul__(self, o): return self._b(OP_MUL, o)
    def __rmul__(self, o): return self._b(OP_MUL, o, True)
    def __truediv__(self, o): return self._b(OP_DIV, o)
    def __rtruediv__(self, o): return self._b(OP_DIV, o, True)
    def __neg__(self): return self.g.unary(OP_NEG, self)
    def __pos__(self): return self

    def __pow__(self, n):
        if isinstance(n, (int, np.integer)) or (isinstance(n, float) and n.is_integer()):
            n = int(n)
            if n >= 0:
                return self.g.powi(self, n)
            return self.g.binary(OP_DIV, self.g.const(1.0), self.g.powi(self, -n))
        if n == 0.5:
            return self.sqrt()
        raise TypeError("only integer powers and 0.5 are supported by the MCP-IR")

    # numpy ufuncs on object arrays dispatch to methods of the same name
    def sqrt(self): return self.g.unary(OP_SQRT, self)
    def exp(self): return self.g.unary(OP_EXP, self)
    def log(self): return self.g.unary(OP_LOG, self)
    def sin(self): return self.g.unary(OP_SIN, self)
    def cos(self): return self.g.unary(OP_COS, self)

    def __repr__(self):
        return f"Expr#{self.id}"


def as_expr_array(g: Graph, values) -> np.ndarray:
    """Flatten whatever a user callable returned (list, ndarray, nested) into Expr[]."""
    flat = np.asarray(values, dtype=object).reshape(-1)
    out = np.empty(flat.shape[0], dtype=object)
    for i, v in enumerate(flat):
        out[i] = g.lift(v)
    return out


def sparse_jacobian(g: Graph, outputs: Sequence[Expr], wrt: Sequence[Tuple[int, int]]):
    """Structurally sparse Jacobian d outputs / d wrt in CSC order.

    `wrt` lists the differentiation variables as (leaf_op, leaf_index) in column order.
    Returns (rows, cols, exprs) sorted column-major (the order `findnz` of a SparseMatrixCSC
    yields, `src/mcp.jl:110`).
    """
    col_of = {key: j for j, key in enumerate(wrt)}
    wrt_ops = {k[0] for k in wrt}
    entries = []
    for i, out in enumerate(outputs):
        for key, d in g.gradient(out, wrt_ops).items():
            j = col_of.get(key)
            if j is not None:
                entries.append((j, i, d))
    entries.sort(key=lambda t: (t[0], t[1]))
    rows = np.array([e[1] for e in entries], dtype=np.int32)
    cols = np.array([e[0] for e in entries], dtype=np.int32)
    exprs = [e[2] for e in entries]
    return rows, cols, exprs
