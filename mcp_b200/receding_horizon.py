"""Warm-started receding-horizon loop, device-resident — the batched counterpart of the reference's
`WarmStartRecedingHorizonStrategy` (`/root/reference/examples/utils.jl:195-235,274-308`) and of the masked-game rollout
built on it (`/root/reference/examples/parametric_masked_game_solver.jl:19-42`), for any trajectory game produced by the
game front-end whose players are planar double integrators (every game of the reference: `examples/lane_change.jl:48-51`,
`examples/train_and_test_utils.jl:393-397`).

The layout is read from `game.dims`, nothing is hard-wired to one game: player i's primal block is
[states (4 × H, time-major); controls (2 × H)] (`unpack_trajectory`, `examples/utils.jl:2-16`), the blocks of `dims.x`
give the offsets, θ_i = [state_i (4); params_i] (`pack_parameters`, `:27-29`), the horizon is `dims.x[i] / 6`.

Per planning step and per instance (reference line numbers in `examples/utils.jl`):
  * θ = pack_parameters(state, params)                                   (`:27-29,297`)
  * if the last solve was `:solved`: warm start x₀, y₀ from it            (`:209-216`)
    else: x₀ = [zero-input rollout from the current state; 0 multipliers], y₀ = ones      (`:217-227`)
  * solve; remember the solution only if it solved                        (`:231-235`)
  * for the next `turn_length` simulation steps apply the planned controls in order, advancing the dynamics (`:287-307`)
x, y, θ and the states stay in HBM between steps; the solve is libmcpb200's kernel, the glue around it is a fixed
number of batched torch ops per step (one gather-free einsum for the rollout guess, one `where` per warm-started
vector, one matmul pair for the dynamics) — no per-stage or per-player Python loops.
"""
from __future__ import annotations

import torch

from . import problems
from .torch_api import solve_device


def masked_game_parameters(goals: torch.Tensor, masks: torch.Tensor) -> torch.Tensor:
    """Per-player parameters of the masked game: player 1 carries the mask under test, the others all ones
    (`parametric_masked_game_solver.jl:19`: `vcat(goal_i, i == 1 ? mask : ones(N))`).
    goals [B, N, 2], masks [B, N]  →  params [B, N, 2 + N]."""
    B, N, _ = goals.shape
    m = torch.ones((B, N, N), dtype=goals.dtype, device=goals.device)
    m[:, 0, :] = masks
    return torch.cat([goals, m], dim=2)


class BatchedRecedingHorizon:
    def __init__(self, game, horizon: int = None, turn_length: int = 1, **solve_opts):
        self.game, self.mcp = game, game.mcp
        dims = game.dims
        self.N = len(dims.x)
        assert all(d == dims.x[0] and d % 6 == 0 for d in dims.x), "players must be planar double integrators (4 states + 2 controls per stage)"
        self.H = dims.x[0] // 6
        assert horizon is None or horizon == self.H, f"the game was built with horizon {self.H}"
        assert all(d == dims.θ[0] for d in dims.θ)
        self.ppp = dims.θ[0] - 4                       # params_per_player (`examples/utils.jl:169-171`)
        assert 1 <= turn_length <= self.H
        self.turn_length = turn_length
        self.nx, self.ny = self.mcp.unconstrained_dimension, self.mcp.constrained_dimension
        self.opts = solve_opts
        self.A = torch.tensor(problems.DI_A, dtype=torch.float64)
        self.Bm = torch.tensor(problems.DI_B, dtype=torch.float64)
        # A^t for the zero-input rollout, t = 0 … H-1  (`zero_input_trajectory`, `:181-192`)
        pw = [torch.eye(4, dtype=torch.float64)]
        for _ in range(self.H - 1):
            pw.append(pw[-1] @ self.A)
        self.Apow = torch.stack(pw)                    # [H, 4, 4]
        self.last = None                               # (x, y, ever-solved mask)
        self._dev = None

    def _to(self, dev):
        if self._dev != dev:
            self.A, self.Bm, self.Apow = self.A.to(dev), self.Bm.to(dev), self.Apow.to(dev)
            self._dev = dev

    def _pack_theta(self, state, params):
        # state [B, N, 4], params [B, N, ppp]  →  θ [B, N·(4+ppp)]
        return torch.cat([state, params], dim=2).reshape(state.shape[0], -1).contiguous()

    def _rollout_guess(self, state):
        """[pack_trajectory(zero-input rollout); zeros(multipliers)] (`:219-227`): states A^t x₀, zero controls."""
        B, H, N = state.shape[0], self.H, self.N
        xs = torch.einsum("tij,bnj->bnti", self.Apow, state).reshape(B, N, 4 * H)            # time-major states
        prim = torch.cat([xs, torch.zeros((B, N, 2 * H), dtype=torch.float64, device=state.device)], dim=2).reshape(B, 6 * H * N)
        return torch.cat([prim, torch.zeros((B, self.nx - 6 * H * N), dtype=torch.float64, device=state.device)], dim=1)

    def plan(self, state: torch.Tensor, params: torch.Tensor):
        """One planning step: returns the solution dict of the batched solve (x, y, s, status, …)."""
        B = state.shape[0]
        self._to(state.device)
        θ = self._pack_theta(state, params)
        x0 = self._rollout_guess(state)
        y0 = torch.ones((B, self.ny), dtype=torch.float64, device=state.device)
        if self.last is not None:
            lx, ly, ok = self.last
            x0 = torch.where(ok[:, None], lx, x0)
            y0 = torch.where(ok[:, None], ly, y0)
        sol = solve_device(self.mcp, θ, x0=x0, y0=y0, **self.opts)
        ok = sol["status"] == 0
        if self.last is None:
            self.last = (sol["x"].clone(), sol["y"].clone(), ok.clone())
        else:
            lx, ly, lok = self.last
            self.last = (torch.where(ok[:, None], sol["x"], lx), torch.where(ok[:, None], sol["y"], ly), ok | lok)
        return sol

    def controls(self, sol, k: int = 0):
        """The k-th planned control of every player, [B, N, 2]."""
        H, N = self.H, self.N
        prim = sol["x"][:, : 6 * H * N].reshape(-1, N, 6 * H)
        return prim[:, :, 4 * H + 2 * k: 4 * H + 2 * k + 2]

    def advance(self, state, u):
        return state @ self.A.T + u @ self.Bm.T

    def step(self, state: torch.Tensor, params: torch.Tensor):
        """Plan once, then simulate `turn_length` steps along the plan.  state [B, N, 4], params [B, N, ppp]
        (CUDA float64).  Returns (next_state, solution dict)."""
        sol = self.plan(state, params)
        for k in range(self.turn_length):
            state = self.advance(state, self.controls(sol, k))
        return state, sol
