"""TEST INFRASTRUCTURE ONLY — ctypes front-end of oracle/c/mcp_oracle.c (the C restatement of
`/root/reference/src/solver.jl:35-138`).  Used by tests/ for large-batch parity and by bench.py for
the CPU baseline / `--impl reference` arm.  Never imported by mcp_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libmcp_oracle.so")
SRC = os.path.join(HERE, "c", "mcp_oracle.c")

_i32p, _f64p = C.POINTER(C.c_int32), C.POINTER(C.c_double)


class _Problem(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nt", C.c_int32), ("n_nodes", C.c_int32),
                ("op", _i32p), ("a", _i32p), ("b", _i32p), ("consts", _f64p), ("gh_nodes", _i32p),
                ("jz_nnz", C.c_int32), ("jz_rows", _i32p), ("jz_cols", _i32p), ("jz_nodes", _i32p),
                ("colperm", _i32p), ("eval_fn", C.c_void_p)]


class _Opts(C.Structure):
    _fields_ = [("tol", C.c_double), ("max_inner_iters", C.c_int32), ("max_outer_iters", C.c_int32),
                ("tightening_rate", C.c_double), ("loosening_rate", C.c_double), ("min_stepsize", C.c_double)]


def cpu_tag() -> str:
    """The host CPU's ISA flags: `-march=native` objects built on another machine (this container vs the GPU box) are rebuilt."""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    import hashlib
                    return hashlib.sha1(line.encode()).hexdigest()[:12]
    except OSError:
        pass
    return "unknown"


def build(force: bool = False) -> str:
    stamp = os.path.join(HERE, "_build", "cpu.txt")
    same_cpu = os.path.exists(stamp) and open(stamp).read() == cpu_tag()
    if force or not same_cpu or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        force = force or not same_cpu
        res = subprocess.run(["make", "-C", HERE, "-B" if force else "-s"], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("building the C oracle failed:\n" + res.stdout + res.stderr)
        with open(stamp, "w") as f:
            f.write(cpu_tag())
    return LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.mcp_oracle_solve_batch.restype = C.c_int
        _lib.mcp_oracle_max_threads.restype = C.c_int
    return _lib


def max_threads() -> int:
    return int(_load().mcp_oracle_max_threads())


def column_ordering(ir) -> np.ndarray:
    """Fill-reducing column pre-ordering of the n×n KKT pattern — COLAMD as run by SuperLU, computed
    once per sparsity pattern (what UMFPACK's symbolic analysis does for the reference, `src/solver.jl:61`)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    nx, ny = ir.nx, ir.ny
    n = nx + 2 * ny
    k = np.arange(ny)
    rows = np.concatenate([ir.jz_rows, nx + k, nx + ny + k, nx + ny + k, np.arange(n)])
    cols = np.concatenate([ir.jz_cols, nx + ny + k, nx + k, nx + ny + k, np.arange(n)])
    rng = np.random.default_rng(0)
    A = sp.csc_matrix((rng.uniform(1.0, 2.0, len(rows)), (rows, cols)), shape=(n, n))
    A = A + sp.identity(n) * 10.0
    perm_c = spla.splu(A.tocsc(), permc_spec="COLAMD").perm_c   # column j of A goes to position perm_c[j]
    return np.argsort(perm_c).astype(np.int32)                   # order[k] = column eliminated k-th


def solve_batch(ir, Θ, x0=None, y0=None, s0=None, tol=1e-4, max_inner_iters=20, max_outer_iters=50,
                tightening_rate=0.1, loosening_rate=0.5, min_stepsize=1e-4, nthreads=0, colperm=None, compiled=False):
    """Batched `solve(InteriorPoint(), mcp, θ)` over the columns of Θ (nθ×B) on the host cores."""
    lib = _load()
    Θ = np.asfortranarray(np.asarray(Θ, dtype=np.float64).reshape(ir.ntheta, -1))
    B = Θ.shape[1]
    keep = []

    def i32(a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        keep.append(a)
        return a.ctypes.data_as(_i32p)

    def f64(a, rows=None):
        if a is None:
            return None
        a = np.asfortranarray(np.asarray(a, dtype=np.float64))
        if rows is not None:
            assert a.shape == (rows, B), (a.shape, rows, B)
        keep.append(a)
        return a.ctypes.data_as(_f64p)

    if colperm is None:
        colperm = getattr(ir, "_oracle_colperm", None)
        if colperm is None:
            colperm = column_ordering(ir)
            try:
                ir._oracle_colperm = colperm
            except Exception:
                pass
    fn = None
    if compiled:   # F!/∇F_z! as generated, compiled C — the timing legs of bench.py ask for it; the tests keep
        # the interpreted tape (bit-identical results, no gcc run: the big QP evaluators take a minute to compile)
        from . import c_emit
        fn = c_emit.compiled_eval(ir, cpu_tag())
    p = _Problem(ir.nx, ir.ny, ir.ntheta, len(ir.op), i32(ir.op), i32(ir.a), i32(ir.b), f64(ir.consts),
                 i32(ir.gh_nodes), len(ir.jz_rows), i32(ir.jz_rows), i32(ir.jz_cols), i32(ir.jz_nodes), i32(colperm), fn)
    o = _Opts(tol, max_inner_iters, max_outer_iters, tightening_rate, loosening_rate, min_stepsize)
    nx, ny = ir.nx, ir.ny
    x = np.empty((nx, B), order="F")
    y = np.empty((ny, B), order="F")
    s = np.empty((ny, B), order="F")
    kkt, eps = np.empty(B), np.empty(B)
    outer, status, steps = (np.empty(B, dtype=np.int32) for _ in range(3))
    used = lib.mcp_oracle_solve_batch(
        C.byref(p), C.c_int64(B), f64(Θ), f64(x0, nx), f64(y0, ny), f64(s0, ny), C.byref(o),
        x.ctypes.data_as(_f64p), y.ctypes.data_as(_f64p), s.ctypes.data_as(_f64p), kkt.ctypes.data_as(_f64p),
        eps.ctypes.data_as(_f64p), outer.ctypes.data_as(_i32p), status.ctypes.data_as(_i32p),
        steps.ctypes.data_as(_i32p), C.c_int(nthreads))
    return SimpleNamespace(status=status, x=x, y=y, s=s, kkt_error=kkt, ϵ=eps, eps=eps, outer_iters=outer,
                           newton_steps=steps, threads=used)
