"""mcp_b200 — B200-native batched interior-point solver for mixed complementarity problems.

Public surface mirrors `/root/reference/src/MixedComplementarityProblems.jl:16`
(`export PrimalDualMCP, solve, ParametricGame, OptimizationProblem`) plus the batched solve.
"""
from .mcp import PrimalDualMCP
from .game import OptimizationProblem, ParametricGame, num_players

__all__ = ["PrimalDualMCP", "OptimizationProblem", "ParametricGame", "num_players"]
