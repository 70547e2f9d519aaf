"""`solve(InteriorPoint(), mcp, θ; x₀, y₀, s₀, tol, …)` on B200 — the host mirror of
`/root/reference/src/solver.jl:35-122` and of the AD rules in `/root/reference/src/AutoDiff.jl`.

Everything numerical happens in libmcpb200.so (CUDA, sm_100a).  This module only validates
arguments, lays batches out the way Julia would (column-major, one instance per column) and calls
the C ABI.  There is no CPU fallback: without a B200 the calls raise `MCPB200Error`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import capi
from .game import ParametricGame, unpack_primals
from .mcp import PrimalDualMCP


class SolverType:
    """`abstract type SolverType` — `src/solver.jl:1`."""


class InteriorPoint(SolverType):
    """`struct InteriorPoint <: SolverType` — `src/solver.jl:2`."""


STATUS = ("solved", "failed")     # `src/solver.jl:69,86,98,118` Symbols :solved / :failed


@dataclass
class Solution:
    """The reference's NamedTuple `(; status, x, y, s, kkt_error, ϵ, outer_iters)` (`src/solver.jl:121`).
    For a single θ the fields are scalars/vectors; for a batch, `x` is nx×B etc. and `status` an
    int array (0 = :solved, 1 = :failed) with `status_symbols()` giving the Symbols."""
    status: object
    x: np.ndarray
    y: np.ndarray
    s: np.ndarray
    kkt_error: object
    ϵ: object
    outer_iters: object
    newton_steps: object = None

    @property
    def eps(self):
        return self.ϵ

    def status_symbols(self):
        if isinstance(self.status, str):
            return self.status
        return np.array(STATUS, dtype=object)[np.asarray(self.status)]


def _handle(mcp: PrimalDualMCP) -> capi.Handle:
    if mcp._handle is None:
        mcp._handle = capi.Handle(mcp.ir)
    return mcp._handle


def _col_major(a, rows: int, B: int, name: str) -> np.ndarray:
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(rows, 1) if B == 1 else a
    if a.shape != (rows, B):
        raise ValueError(f"{name} must have shape ({rows}, {B}), got {a.shape}")
    return np.asfortranarray(a)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def solve(solver_or_game, mcp_or_θ=None, θ=None, *, x0=None, y0=None, s0=None, tol: float = 1e-4,
          max_inner_iters: int = 20, max_outer_iters: int = 50, tightening_rate: float = 0.1,
          loosening_rate: float = 0.5, min_stepsize: float = 1e-4, verbose: bool = False,
          linear_solve_algorithm=None, solver_type=None, **unicode_kw):
    """Mirrors both reference methods:

    * ``solve(InteriorPoint(), mcp, θ; x₀, y₀, s₀, tol, …)`` — `src/solver.jl:35-51`.  θ may be a vector
      (one solve) or an nθ×B matrix (one solve per column — the batched form this framework adds).
    * ``solve(game, θ; solver_type = InteriorPoint(), kwargs...)`` — `src/game.jl:186-205`; returns
      `(primals, variables, kkt_error, status)`.

    `x₀/y₀/s₀` may also be passed by their Julia names.  `linear_solve_algorithm` is accepted for
    signature parity and must be None: the KKT solve is the library's fixed banded LU
    (the reference default is UMFPACK, `src/solver.jl:50`).
    """
    x0 = unicode_kw.pop("x₀", x0)
    y0 = unicode_kw.pop("y₀", y0)
    s0 = unicode_kw.pop("s₀", s0)
    if unicode_kw:
        raise TypeError(f"unexpected keyword arguments {sorted(unicode_kw)}")
    if linear_solve_algorithm is not None:
        raise ValueError("linear_solve_algorithm cannot be chosen: the B200 path has one fixed KKT solver")
    kw = dict(x0=x0, y0=y0, s0=s0, tol=tol, max_inner_iters=max_inner_iters, max_outer_iters=max_outer_iters,
              tightening_rate=tightening_rate, loosening_rate=loosening_rate, min_stepsize=min_stepsize)
    if isinstance(solver_or_game, ParametricGame):
        game, θ_ = solver_or_game, mcp_or_θ
        st = solver_type if solver_type is not None else InteriorPoint()
        sol = _solve_mcp(st, game.mcp, _flatten_blocks(θ_), **kw)            # `src/game.jl:196`
        return GameSolution(primals=unpack_primals(game, sol.x), variables=sol, kkt_error=sol.kkt_error,
                            status=sol.status)
    return _solve_mcp(solver_or_game, mcp_or_θ, θ, **kw)


@dataclass
class GameSolution:
    """`(; primals, variables = (; x, y, s), kkt_error, status)` — `src/game.jl:204`."""
    primals: list
    variables: Solution
    kkt_error: object
    status: object


def _flatten_blocks(θ):
    if isinstance(θ, (list, tuple)) and len(θ) and isinstance(θ[0], (list, tuple, np.ndarray)):
        return np.concatenate([np.asarray(b, dtype=np.float64) for b in θ], axis=0)
    return θ


def _solve_mcp(solver, mcp: PrimalDualMCP, θ, *, x0, y0, s0, tol, max_inner_iters, max_outer_iters,
               tightening_rate, loosening_rate, min_stepsize) -> Solution:
    if not isinstance(solver, InteriorPoint):
        raise TypeError("only InteriorPoint() is implemented (the reference has no other SolverType)")
    θ = np.asarray(θ, dtype=np.float64)
    single = θ.ndim == 1
    nθ = mcp.parameter_dimension
    B = 1 if single else θ.shape[1]
    Θ = _col_major(θ, nθ, B, "θ")
    nx, ny = mcp.unconstrained_dimension, mcp.constrained_dimension
    X0 = None if x0 is None else _col_major(x0, nx, B, "x₀")
    Y0 = None if y0 is None else _col_major(y0, ny, B, "y₀")
    S0 = None if s0 is None else _col_major(s0, ny, B, "s₀")
    h = _handle(mcp)
    x = np.empty((nx, B), order="F")
    y = np.empty((ny, B), order="F")
    s = np.empty((ny, B), order="F")
    kkt = np.empty(B)
    eps = np.empty(B)
    outer = np.empty(B, dtype=np.int32)
    status = np.empty(B, dtype=np.int32)
    steps = np.empty(B, dtype=np.int32)
    opts = capi.default_opts(tol=tol, max_inner_iters=max_inner_iters, max_outer_iters=max_outer_iters,
                             tightening_rate=tightening_rate, loosening_rate=loosening_rate,
                             min_stepsize=min_stepsize)
    rc = h._lib.mcpb200_solve_batched(h.raw, B, _ptr(Θ), _ptr(X0), _ptr(Y0), _ptr(S0), C.byref(opts), _ptr(x),
                                      _ptr(y), _ptr(s), _ptr(kkt), _ptr(eps), _ptr(outer), _ptr(status), _ptr(steps))
    h.check(rc)
    if single:
        return Solution(STATUS[int(status[0])], x[:, 0], y[:, 0], s[:, 0], float(kkt[0]), float(eps[0]),
                        int(outer[0]), int(steps[0]))
    return Solution(status, x, y, s, kkt, eps, outer, steps)


# ---- sensitivities (`src/AutoDiff.jl`) -------------------------------------------------------------
def _sens_call(mcp, θ, sol: Solution, want_jac=False, zbar=None, θ_p=None):
    θ = np.asarray(θ, dtype=np.float64)
    single = θ.ndim == 1
    nθ, nx, ny = mcp.parameter_dimension, mcp.unconstrained_dimension, mcp.constrained_dimension
    n = nx + 2 * ny
    B = 1 if single else θ.shape[1]
    Θ = _col_major(θ, nθ, B, "θ")
    X = _col_major(sol.x, nx, B, "x")
    Y = _col_major(sol.y, ny, B, "y")
    S = _col_major(sol.s, ny, B, "s")
    E = np.ascontiguousarray(np.atleast_1d(np.asarray(sol.ϵ, dtype=np.float64)))
    h = _handle(mcp)
    jac = np.empty((n, nθ, B), order="F") if want_jac else None
    ZB = tb = None
    if zbar is not None:
        ZB = _col_major(zbar, n, B, "z̄")
        tb = np.empty((nθ, B), order="F")
    P, TP, zp = 0, None, None
    if θ_p is not None:
        TP = np.asarray(θ_p, dtype=np.float64)
        if single:
            TP = TP.reshape(nθ, -1, 1)
        P = TP.shape[1]
        TP = np.asfortranarray(TP.reshape(nθ, P, B))
        zp = np.empty((n, P, B), order="F")
    st = np.empty(B, dtype=np.int32)
    rc = h._lib.mcpb200_sensitivities(h.raw, B, _ptr(Θ), _ptr(X), _ptr(Y), _ptr(S), _ptr(E), _ptr(jac), _ptr(ZB),
                                      _ptr(tb), P, _ptr(TP), _ptr(zp), _ptr(st))
    if rc == capi.ERR_NO_SENSITIVITIES:
        # the reference throws ArgumentError here (`src/AutoDiff.jl:19-23`)
        raise ValueError("Missing sensitivities. Set `compute_sensitivities = true` when constructing the "
                         "PrimalDualMCP.")
    h.check(rc)
    if np.any(st != 0):
        # a singular / non-finite KKT matrix at the returned point: the kernel NaN-fills those instances' outputs
        # (the reference's dense QR, `src/AutoDiff.jl:39`, would hand back Inf/NaN or throw)
        import warnings
        warnings.warn(f"sensitivities: the KKT matrix is singular for {int(np.count_nonzero(st))} of {B} instance(s); "
                      "their outputs are NaN", RuntimeWarning, stacklevel=3)
    return jac, tb, zp, st, single


def solve_jacobian_θ(mcp: PrimalDualMCP, solution: Solution, θ) -> np.ndarray:
    """`_solve_jacobian_θ(mcp, solution, θ)` — `src/AutoDiff.jl:18-40`: ∂z/∂θ, n×nθ (×B for a batch)."""
    jac, _, _, _, single = _sens_call(mcp, θ, solution, want_jac=True)
    return jac[:, :, 0] if single else jac


def solve_pullback(mcp: PrimalDualMCP, solution: Solution, θ, dx=None, dy=None, ds=None) -> np.ndarray:
    """The pullback of `rrule(solve, …)` — `src/AutoDiff.jl:52-79`: given cotangents ∂l/∂x, ∂l/∂y, ∂l/∂s
    returns ∂l/∂θ = Σ_b (∂z/∂θ)[b,:]ᵀ ∂l/∂b (`:65-75`).  Missing cotangents are zero."""
    nx, ny = mcp.unconstrained_dimension, mcp.constrained_dimension
    single = np.asarray(θ).ndim == 1
    B = 1 if single else np.asarray(θ).shape[1]

    def part(v, rows):
        return np.zeros((rows, B)) if v is None else np.asarray(v, dtype=np.float64).reshape(rows, B)

    zbar = np.concatenate([part(dx, nx), part(dy, ny), part(ds, ny)], axis=0)
    _, tb, _, _, single = _sens_call(mcp, θ, solution, zbar=zbar)
    return tb[:, 0] if single else tb


def solve_pushforward(mcp: PrimalDualMCP, solution: Solution, θ, θ_p):
    """Forward rule — the ForwardDiff.Dual overload `src/AutoDiff.jl:84-117`: z_p = ∂z∂θ·θ_p (`:98`).
    θ_p is nθ×P (×B); returns (x_p, y_p, s_p).  (The reference re-wraps the `s` partials around
    `solution.y` *values*, `:109-114` — the partials returned here are the same either way.)"""
    nx, ny = mcp.unconstrained_dimension, mcp.constrained_dimension
    _, _, zp, _, single = _sens_call(mcp, θ, solution, θ_p=θ_p)
    if single:
        zp = zp[:, :, 0]
    return zp[:nx], zp[nx:nx + ny], zp[nx + ny:]


def value_and_gradient(f_of_solution, grad_of_solution, mcp, θ, **solve_kw):
    """Convenience equivalent of `Zygote.gradient(θ -> f(solve(InteriorPoint(), mcp, θ)), θ)`
    (`test/runtests.jl:75-80`): `grad_of_solution(sol)` returns (∂f/∂x, ∂f/∂y, ∂f/∂s)."""
    sol = solve(InteriorPoint(), mcp, θ, **solve_kw)
    dx, dy, ds = grad_of_solution(sol)
    return f_of_solution(sol), solve_pullback(mcp, sol, θ, dx, dy, ds)
