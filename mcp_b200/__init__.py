"""mcp_b200 — B200-native batched interior-point solver for mixed complementarity problems.

Public surface mirrors `/root/reference/src/MixedComplementarityProblems.jl:16`
(`export PrimalDualMCP, solve, ParametricGame, OptimizationProblem`) plus the batched solve
(θ as an nθ×B matrix) and the sensitivity rules of `src/AutoDiff.jl`.
"""
from .mcp import PrimalDualMCP
from .game import OptimizationProblem, ParametricGame, num_players
from .solver import (InteriorPoint, SolverType, Solution, GameSolution, solve, solve_jacobian_θ,
                     solve_pullback, solve_pushforward)

__all__ = ["PrimalDualMCP", "OptimizationProblem", "ParametricGame", "num_players", "InteriorPoint",
           "SolverType", "Solution", "GameSolution", "solve", "solve_jacobian_θ", "solve_pullback",
           "solve_pushforward"]
