"""Warm-started receding-horizon loop, device-resident — the batched counterpart of the reference's
`WarmStartRecedingHorizonStrategy` (`/root/reference/examples/utils.jl:195-235,274-308`) for trajectory games built
with `problems.lane_change_game`-style layouts (per player [states 4×H; controls 2×H], double-integrator dynamics).

Per simulation step and per instance (reference line numbers in `examples/utils.jl`):
  * θ = pack_parameters(state, params)                                   (`:27-29,297`)
  * if the last solve was `:solved`: warm start x₀, y₀ from it            (`:209-216`)
    else: x₀ = zero-input rollout from the current state, y₀ = ones      (`:217-227`)
  * solve; remember the solution only if it solved                        (`:231-235`)
  * apply the first planned control of every player, advance the dynamics (`:307`, turn_length = 1)
x, y, θ and the states stay in HBM between steps; only the solve is a libmcpb200 kernel, the packing / dynamics
step are a handful of elementwise torch ops (plumbing).
"""
from __future__ import annotations

import torch

from . import problems
from .torch_api import solve_device


class BatchedRecedingHorizon:
    def __init__(self, game, horizon: int, n_players: int = 2, params_per_player: int = 1, n_eq=None, **solve_opts):
        self.game, self.mcp = game, game.mcp
        self.H, self.N, self.ppp = horizon, n_players, params_per_player
        self.nx, self.ny = self.mcp.unconstrained_dimension, self.mcp.constrained_dimension
        self.opts = solve_opts
        self.A = torch.tensor(problems.DI_A, dtype=torch.float64)
        self.Bm = torch.tensor(problems.DI_B, dtype=torch.float64)
        self.last = None          # (x, y, solved mask)

    def _pack_theta(self, state, params):
        # state [B, N, 4], params [B, N, ppp]  →  θ [B, N·(4+ppp)]
        return torch.cat([state, params], dim=2).reshape(state.shape[0], -1).contiguous()

    def _rollout_guess(self, state):
        B, H, N = state.shape[0], self.H, self.N
        A = self.A.to(state.device)
        x0 = torch.zeros((B, self.nx), dtype=torch.float64, device=state.device)
        st = state.clone()
        for t in range(H):
            for i in range(N):
                x0[:, i * 6 * H + 4 * t: i * 6 * H + 4 * t + 4] = st[:, i]
            st = st @ A.T
        return x0

    def step(self, state: torch.Tensor, params: torch.Tensor):
        """state [B, N, 4], params [B, N, ppp] (CUDA float64).  Returns (next_state, solution dict)."""
        B = state.shape[0]
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
        # first control of every player, then one dynamics step
        H, N = self.H, self.N
        u = torch.stack([sol["x"][:, i * 6 * H + 4 * H: i * 6 * H + 4 * H + 2] for i in range(N)], dim=1)   # [B, N, 2]
        nxt = state @ self.A.to(state.device).T + u @ self.Bm.to(state.device).T
        return nxt, sol
