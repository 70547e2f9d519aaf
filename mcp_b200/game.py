"""Game → MCP front-end, mirroring `/root/reference/src/game.jl`.

`OptimizationProblem` (`src/game.jl:2-6`), `ParametricGame` (`:16-44`), `game_to_mcp` (`:47-157`),
`solve(game, θ)` (`:186-205`) and `num_players` (`:208-210`).

Block vectors: where the reference passes `BlockArrays` (`x[Block(i)]`), callables here receive a
`Blocks` object — `x[i]` is player i's sub-vector (0-based) and `x.flat` the concatenation.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import trace as T
from .mcp import PrimalDualMCP


class Blocks:
    """Minimal stand-in for a BlockVector: a flat object array plus block boundaries."""

    def __init__(self, flat: np.ndarray, dims: Sequence[int]):
        self.flat = flat if isinstance(flat, np.ndarray) else np.asarray(flat)
        self.dims = [int(d) for d in dims]
        assert sum(self.dims) == len(self.flat)
        self._off = np.concatenate([[0], np.cumsum(self.dims)]).astype(int)

    def __getitem__(self, i: int) -> np.ndarray:
        return self.flat[self._off[i]:self._off[i + 1]]

    def __len__(self) -> int:
        return len(self.dims)

    def blocks(self) -> List[np.ndarray]:
        return [self[i] for i in range(len(self.dims))]


@dataclass
class OptimizationProblem:
    """`src/game.jl:2-6`."""
    objective: Callable
    private_equality: Optional[Callable] = None
    private_inequality: Optional[Callable] = None


@dataclass
class GameDims:
    x: List[int]
    θ: List[int]
    λ: List[int]
    μ: List[int]
    λ̃: int
    μ̃: int


def _flat(g: T.Graph, v) -> np.ndarray:
    return T.as_expr_array(g, v)


def game_to_mcp(*, test_point: Sequence[Sequence[float]], test_parameter: Sequence[Sequence[float]],
                problems: Sequence[OptimizationProblem], shared_equality: Optional[Callable] = None,
                shared_inequality: Optional[Callable] = None):
    """`game_to_mcp` — `src/game.jl:47-157`.  Returns K, z, bounds and dims, with z's unconstrained
    entries traced as x-leaves and constrained entries as y-leaves (the split of `src/mcp.jl:193-199`)."""
    N = len(problems)
    assert N == len(test_point)                                            # :55
    dx = [len(b) for b in test_point]                                      # dimensions(), :159-183
    dθ = [len(b) for b in test_parameter]
    tp = Blocks(np.concatenate([np.asarray(b, dtype=np.float64) for b in test_point]), dx)
    tθ = Blocks(np.concatenate([np.asarray(b, dtype=np.float64) for b in test_parameter]), dθ)
    dλ = [0 if p.private_equality is None else len(np.atleast_1d(p.private_equality(tp, tθ[i])))
          for i, p in enumerate(problems)]
    dμ = [0 if p.private_inequality is None else len(np.atleast_1d(p.private_inequality(tp, tθ[i])))
          for i, p in enumerate(problems)]
    dλs = 0 if shared_equality is None else len(np.atleast_1d(shared_equality(tp, tθ)))
    dμs = 0 if shared_inequality is None else len(np.atleast_1d(shared_inequality(tp, tθ)))
    dims = GameDims(dx, dθ, dλ, dμ, dλs, dμs)

    g = T.Graph()
    nx_primal, nλ, nμ = sum(dx), sum(dλ), sum(dμ)
    n_unc = nx_primal + nλ + dλs
    n_con = nμ + dμs
    # z = [x; λ; λ̃; μ; μ̃] (`:120-131`): the first three groups are unconstrained (`:133-147`)
    unc = g.variables("x", n_unc)
    con = g.variables("y", n_con)
    x = Blocks(unc[:nx_primal], dx)                                        # :68-70
    λ = Blocks(unc[nx_primal:nx_primal + nλ], dλ)                          # :71-73
    λs = unc[nx_primal + nλ:]                                              # :77
    μ = Blocks(con[:nμ], dμ)                                               # :74-76
    μs = con[nμ:]                                                          # :78
    θ = Blocks(g.variables("theta", sum(dθ)), dθ)                          # :79-81

    fs = [p.objective(x, θ[i]) for i, p in enumerate(problems)]            # :84-86
    gs = [None if p.private_equality is None else _flat(g, p.private_equality(x, θ[i]))
          for i, p in enumerate(problems)]                                 # :87-89
    hs = [None if p.private_inequality is None else _flat(g, p.private_inequality(x, θ[i]))
          for i, p in enumerate(problems)]                                 # :90-92
    gsh = None if shared_equality is None else _flat(g, shared_equality(x, θ))       # :94
    hsh = None if shared_inequality is None else _flat(g, shared_inequality(x, θ))   # :95

    def dot(a, b):
        acc = g.const(0.0)
        for u, v in zip(a, b):
            acc = acc + u * v
        return acc

    grads = []
    off = 0
    for i in range(N):                                                     # :98-103
        L = g.lift(fs[i])
        if gs[i] is not None:
            L = L - dot(λ[i], gs[i])
        if hs[i] is not None:
            L = L - dot(μ[i], hs[i])
        if gsh is not None:
            L = L - dot(λs, gsh)
        if hsh is not None:
            L = L - dot(μs, hsh)
        dL = g.gradient(L, {T.OP_X})
        zero = g.const(0.0)
        grads += [dL.get((T.OP_X, off + k), zero) for k in range(dx[i])]
        off += dx[i]

    K = list(grads)                                                        # :107-118
    for gi in gs:
        if gi is not None:
            K += list(gi)
    if gsh is not None:
        K += list(gsh)
    for hi in hs:
        if hi is not None:
            K += list(hi)
    if hsh is not None:
        K += list(hsh)
    assert len(K) == n_unc + n_con
    lower = np.concatenate([np.full(n_unc, -np.inf), np.zeros(n_con)])    # :133-139
    upper = np.full(n_unc + n_con, np.inf)                                 # :141-147
    return dict(graph=g, K_symbolic=K, lower_bounds=lower, upper_bounds=upper, dims=dims,
                n_unconstrained=n_unc, n_constrained=n_con, parameter_dimension=sum(dθ))


class ParametricGame:
    """`src/game.jl:16-44`.  Always builds sensitivities, like the reference (`:42` uses the
    constructor default `compute_sensitivities = true`, `src/mcp.jl:188`)."""

    def __init__(self, *, test_point, test_parameter, problems, shared_equality=None,
                 shared_inequality=None, compute_sensitivities: bool = True):
        comp = game_to_mcp(test_point=test_point, test_parameter=test_parameter, problems=problems,
                           shared_equality=shared_equality, shared_inequality=shared_inequality)
        self.problems = list(problems)
        self.shared_equality = shared_equality
        self.shared_inequality = shared_inequality
        self.dims: GameDims = comp["dims"]
        K = comp["K_symbolic"]
        nu = comp["n_unconstrained"]
        self.mcp = PrimalDualMCP.from_symbolic(comp["graph"], K[:nu], K[nu:], nu, comp["n_constrained"],
                                               comp["parameter_dimension"], compute_sensitivities)


def num_players(game: ParametricGame) -> int:
    """`src/game.jl:208-210`."""
    return len(game.problems)


def unpack_primals(game: ParametricGame, x: np.ndarray):
    """Per-player slices of x (`src/game.jl:199-202`); works on a vector or an (nx, B) matrix."""
    ends = np.cumsum(game.dims.x)
    return [x[(0 if i == 0 else ends[i - 1]):ends[i]] for i in range(num_players(game))]
