"""Host-side mirror of the reference's `PrimalDualMCP` (`/root/reference/src/mcp.jl:13-210`).

Same constructors, same keyword names, same fields; construction traces the callables into the
MCP-IR (`trace.py`, `ir.py`).  The numerical work happens in `libmcpb200.so` (see `solver.py`).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np

from . import trace as T
from .ir import MCPIR, build_ir


class PrimalDualMCP:
    """Primal-dual KKT system of  G(x,y;θ)=0, 0 ≤ H(x,y;θ) ⟂ y ≥ 0  (`src/mcp.jl:1-24`).

    Constructors (all mirror the reference):

    * ``PrimalDualMCP(G, H, unconstrained_dimension=…, constrained_dimension=…,
      parameter_dimension=…, compute_sensitivities=True)`` — callables ``G(x, y, θ=…)``,
      ``H(x, y, θ=…)`` accepting symbolic vectors and a θ keyword (`src/mcp.jl:27-52`).
    * ``PrimalDualMCP.from_symbolic(G_sym, H_sym, graph, nx, ny, nθ, …)`` — already traced
      expressions (`src/mcp.jl:55-150`).
    * ``PrimalDualMCP.from_K(K, lower_bounds, upper_bounds, parameter_dimension=…)`` — callable
      ``K(z, θ=…)`` with bounds (`src/mcp.jl:155-177`), and ``from_K_symbolic`` (`:182-210`).
    """

    def __init__(self, G: Callable, H: Callable, *, unconstrained_dimension: int,
                 constrained_dimension: int, parameter_dimension: int,
                 compute_sensitivities: bool = True):
        g = T.Graph()
        x = g.variables("x", unconstrained_dimension)        # `src/mcp.jl:37-39`
        y = g.variables("y", constrained_dimension)
        th = g.variables("theta", parameter_dimension)
        G_sym = T.as_expr_array(g, G(x, y, θ=th) if _wants_unicode(G) else G(x, y, theta=th))   # :40
        H_sym = T.as_expr_array(g, H(x, y, θ=th) if _wants_unicode(H) else H(x, y, theta=th))   # :41
        self._init_symbolic(g, G_sym, H_sym, unconstrained_dimension, constrained_dimension,
                            parameter_dimension, compute_sensitivities)

    # -- alternate constructors -------------------------------------------------------------
    @classmethod
    def from_symbolic(cls, graph: T.Graph, G_sym, H_sym, nx: int, ny: int, ntheta: int,
                      compute_sensitivities: bool = True) -> "PrimalDualMCP":
        self = cls.__new__(cls)
        self._init_symbolic(graph, T.as_expr_array(graph, G_sym), T.as_expr_array(graph, H_sym),
                            nx, ny, ntheta, compute_sensitivities)
        return self

    @classmethod
    def from_K(cls, K: Callable, lower_bounds: Sequence[float], upper_bounds: Sequence[float], *,
               parameter_dimension: int, compute_sensitivities: bool = True) -> "PrimalDualMCP":
        """`PrimalDualMCP(K, lb, ub; parameter_dimension)` — `src/mcp.jl:155-177`."""
        lb = np.asarray(lower_bounds, dtype=np.float64)
        ub = np.asarray(upper_bounds, dtype=np.float64)
        unc, con = split_bounds(lb, ub)
        g = T.Graph()
        # z symbols: unconstrained entries become x-leaves, constrained ones y-leaves, so that the
        # split of `src/mcp.jl:193-199` is already baked into the leaves.
        xs = g.variables("x", len(unc))
        ys = g.variables("y", len(con))
        z = np.empty(len(lb), dtype=object)
        z[unc] = xs
        z[con] = ys
        th = g.variables("theta", parameter_dimension)
        K_sym = T.as_expr_array(g, K(z, θ=th) if _wants_unicode(K) else K(z, theta=th))
        return cls.from_symbolic(g, K_sym[unc], K_sym[con], len(unc), len(con), parameter_dimension,
                                 compute_sensitivities)

    @classmethod
    def from_K_symbolic(cls, graph: T.Graph, K_sym, z_is_constrained: Sequence[bool], ntheta: int,
                        compute_sensitivities: bool = True) -> "PrimalDualMCP":
        """Symbolic K form (`src/mcp.jl:182-210`).  The caller must have created the z symbols so that
        unconstrained entries are x-leaves and constrained entries y-leaves, in order (`:193-199`)."""
        con = np.asarray(z_is_constrained, dtype=bool)
        K_sym = T.as_expr_array(graph, K_sym)
        return cls.from_symbolic(graph, K_sym[~con], K_sym[con], int((~con).sum()), int(con.sum()),
                                 ntheta, compute_sensitivities)

    # -- common ----------------------------------------------------------------------------------
    def _init_symbolic(self, g, G_sym, H_sym, nx, ny, ntheta, compute_sensitivities):
        if len(G_sym) != nx or len(H_sym) != ny:
            raise ValueError(f"G/H returned {len(G_sym)}/{len(H_sym)} rows, expected {nx}/{ny}")
        self.unconstrained_dimension = int(nx)      # `src/mcp.jl:20-23`
        self.constrained_dimension = int(ny)
        self.parameter_dimension = int(ntheta)
        self.compute_sensitivities = bool(compute_sensitivities)
        self.ir: MCPIR = build_ir(g, list(G_sym), list(H_sym), nx, ny, ntheta, compute_sensitivities)
        self._handle = None   # lazily created device-side problem (solver.py)

    def close(self):
        if self._handle is not None:
            self._handle.close()
            self._handle = None


def split_bounds(lb: np.ndarray, ub: np.ndarray):
    """`src/mcp.jl:191-194`: all upper bounds +Inf, lower bounds -Inf or 0."""
    if not (np.all(np.isinf(ub) & (ub > 0)) and np.all((np.isinf(lb) & (lb < 0)) | (lb == 0))):
        raise AssertionError("PrimalDualMCP assumes upper bounds = Inf and lower bounds ∈ {-Inf, 0}")
    unc = np.nonzero(np.isinf(lb))[0]
    con = np.nonzero(~np.isinf(lb))[0]
    return unc, con


def _wants_unicode(fn) -> bool:
    """Julia callables take `θ` as the keyword; accept Python callables spelling it `θ` or `theta`."""
    import inspect
    try:
        params = inspect.signature(fn).parameters
    except (TypeError, ValueError):
        return False
    return "θ" in params
