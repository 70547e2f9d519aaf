"""TEST INFRASTRUCTURE ONLY — problem definitions of the trajectory-game configs written a SECOND time,
straight from the reference's sources, with no use of the product's tracer / IR / game front-end.

Why: `oracle/ir_eval.py` evaluates the tape the product's own `trace.py` / `game.py` / `problems.py`
produced, so a wrong Lagrangian gradient, constraint row or θ packing in that front-end would be invisible to
every GPU-vs-oracle parity test (both sides would solve the same wrong problem).  Here the same K(z; θ) is
built from vectorised torch-float64 code and `torch.autograd` (exact derivatives to rounding, nothing shared
with `mcp_b200.trace`), following

* `game_to_mcp`              `/root/reference/src/game.jl:47-157`   (Lagrangians :98-103, K :107-118, z :120-131)
* `build_mcp_components`     `/root/reference/examples/utils.jl:87-178` (costs :96-106, g̃ :109-123, h̃ :126-155)
* `unpack_trajectory` / `unpack_parameters`  `examples/utils.jl:2-16,32-41`
* lane-change game           `/root/reference/examples/lane_change.jl:2-55`
* masked N-player game       `/root/reference/examples/train_and_test_utils.jl:341-401`
* F = [G; H − s; s∘y − ϵ], z = [x; y; s]   `/root/reference/src/mcp.jl:72-80`, split by bounds `:193-199`

The un-vendored TrajectoryGamesBase / TrajectoryGamesExamples pieces (double integrator, polygon half-spaces,
box-bound constraints) follow SURVEY.md Appendix A, as everything else in this repo must.

Only `tests/` may import this module.
"""
from __future__ import annotations

import math

import numpy as np
import torch

DT = 0.1
F64 = torch.float64


def _double_integrator():
    # TrajectoryGamesExamples.planar_double_integrator(dt = 0.1, m = 1): (px, py, vx, vy), (ax, ay)
    A = torch.eye(4, dtype=F64)
    A[0, 2] = DT
    A[1, 3] = DT
    B = torch.zeros(4, 2, dtype=F64)
    B[0, 0] = B[1, 1] = 0.5 * DT * DT
    B[2, 0] = B[3, 1] = DT
    return A, B


class IndependentTrajectoryGame:
    """K(z; θ) of an N-player trajectory game exactly as `game_to_mcp` stacks it, for one of the two game
    families of BASELINE.json (`kind` = "lane_change" | "masked")."""

    def __init__(self, kind: str, N: int, horizon: int, **kw):
        self.kind, self.N, self.H = kind, N, horizon
        self.A, self.B = _double_integrator()
        if kind == "lane_change":
            assert N == 2
            lane_width, num_lanes, height = kw.get("lane_width", 2.0), kw.get("num_lanes", 2), kw.get("height", 50.0)
            centers = [(i - 0.5) * lane_width for i in range(1, num_lanes + 1)]       # lane_change.jl:3
            xmin, xmax = centers[0] - 0.5 * lane_width, centers[-1] + 0.5 * lane_width  # :4-9
            ymin, ymax = 0.0, height
            self.state_lb = [-math.inf, -math.inf, -10.0, 0.0]                        # :49
            self.state_ub = [math.inf, math.inf, 10.0, 10.0]
            self.ctrl_lb, self.ctrl_ub = [-5.0, -5.0], [3.0, 3.0]                     # :50
            self.extra = 1                                                            # params_per_player = 1
        elif kind == "masked":
            length = kw.get("length", 10.0)
            xmin, xmax, ymin, ymax = -0.5 * length, 0.5 * length, -0.5 * length, 0.5 * length   # :341-349
            self.state_lb = [-math.inf, -math.inf, -2.0, -2.0]                        # :394
            self.state_ub = [math.inf, math.inf, 2.0, 2.0]
            self.ctrl_lb, self.ctrl_ub = [-1.0, -1.0], [1.0, 1.0]                     # :395
            self.extra = N + 2                                                        # time_test.jl:23-24
        else:
            raise ValueError(kind)
        # rectangle given by its vertices counter-clockwise from (xmin, ymin): half-spaces a·p ≤ b per edge
        self.halfspaces = torch.tensor([[0.0, -1.0, -ymin], [1.0, 0.0, xmax], [0.0, 1.0, ymax], [-1.0, 0.0, -xmin]],
                                       dtype=F64)
        H = horizon
        self.primal = 6 * H * N                      # utils.jl:157-160
        self.n_eq = 4 * N * H                        # initial state + (H-1) dynamics rows, :109-123
        self.nx = self.primal + self.n_eq            # x = [τ; λ̃]
        n_box_s = sum(math.isfinite(v) for v in self.state_lb) + sum(math.isfinite(v) for v in self.state_ub)
        n_box_c = sum(math.isfinite(v) for v in self.ctrl_lb) + sum(math.isfinite(v) for v in self.ctrl_ub)
        self.ny = H * (1 + 4 * N + n_box_c * N + n_box_s * N)
        self.ntheta = N * (4 + self.extra)

    # ---- unpacking ------------------------------------------------------------------------------------------
    def _traj(self, tau):
        """per player i: states [H,4], controls [H,2]  (utils.jl:2-16: [vec(X 4×H); vec(U 2×H)], column-major)."""
        H = self.H
        xs, us = [], []
        for i in range(self.N):
            blk = tau[6 * H * i: 6 * H * (i + 1)]
            xs.append(blk[:4 * H].reshape(H, 4))
            us.append(blk[4 * H:].reshape(H, 2))
        return xs, us

    def _theta_blocks(self, th):
        w = 4 + self.extra
        return [th[w * i: w * (i + 1)] for i in range(self.N)]

    # ---- the three ingredients ------------------------------------------------------------------------------
    def costs(self, tau, th):
        xs, us = self._traj(tau)
        out = []
        for i, thi in enumerate(self._theta_blocks(th)):
            if self.kind == "lane_change":                                            # lane_change.jl:18-24
                lane = thi[-1]
                stage = ((xs[i][:, 0] - lane) ** 2 + 0.5 * (xs[i][:, 2] ** 2 + (xs[i][:, 3] - 2.0) ** 2)
                         + 0.1 * (us[i] ** 2).sum(dim=1))
            else:                                                                     # train_and_test_utils.jl:364-370
                N = self.N
                goal, mask = thi[len(thi) - (N + 2): len(thi) - N], thi[len(thi) - N:]
                stage = (((xs[i][:, :2] - goal) ** 2).sum(dim=1) + (xs[i][:, 2:] ** 2).sum(dim=1)
                         + 0.1 * (us[i] ** 2).sum(dim=1))
                for j in range(N):
                    if j != i:
                        stage = stage + 2.0 * (mask[i] * mask[j]) / ((xs[i][:, :2] - xs[j][:, :2]) ** 2).sum(dim=1)
            out.append(stage.sum() / self.H)                                          # reducer: mean over stages
        return out

    def shared_equality(self, tau, th):
        xs, us = self._traj(tau)
        X = torch.cat(xs, dim=1)                          # [H, 4N] joint state per stage
        init = torch.cat([b[:4] for b in self._theta_blocks(th)])
        rows = [X[0] - init]                              # utils.jl:114
        nxt = torch.cat([xs[i][:-1] @ self.A.T + us[i][:-1] @ self.B.T for i in range(self.N)], dim=1)
        rows.append((X[1:] - nxt).reshape(-1))            # :117-120, stage-major
        return torch.cat(rows)

    def shared_inequality(self, tau, th):
        xs, us = self._traj(tau)
        H, N = self.H, self.N
        if self.kind == "lane_change":                    # lane_change.jl:39-46
            h1 = ((xs[0][:, :2] - xs[1][:, :2]) ** 2).sum(dim=1) - 4.0
        else:                                             # train_and_test_utils.jl:380-388: the constant [1]
            h1 = torch.ones(H, dtype=F64) + 0.0 * xs[0][:, 0]
        a, b = self.halfspaces[:, :2], self.halfspaces[:, 2]
        # per stage, per player, per half-space: b − a·p  
        h2 = torch.stack([b - xs[i][:, :2] @ a.T for i in range(N)], dim=1).reshape(-1)   # [H, N, 4]
        U = torch.cat(us, dim=1)                          # [H, 2N]
        X = torch.cat(xs, dim=1)                          # [H, 4N]

        def box(V, lb, ub):
            lb, ub = lb * N, ub * N
            lo = [k for k, v in enumerate(lb) if math.isfinite(v)]
            hi = [k for k, v in enumerate(ub) if math.isfinite(v)]
            return torch.cat([V[:, lo] - torch.tensor([lb[k] for k in lo], dtype=F64),
                              torch.tensor([ub[k] for k in hi], dtype=F64) - V[:, hi]], dim=1).reshape(-1)

        h3 = box(U, self.ctrl_lb, self.ctrl_ub)           # utils.jl:140-145
        h4 = box(X, self.state_lb, self.state_ub)         # utils.jl:147-152
        return torch.cat([h1, h2, h3, h4])                # :154

    # ---- K and F --------------------------------------------------------------------------------------------
    def K(self, x, y, th):
        """[∇_{τ_i} L_i (i = 1..N); g̃; h̃]  with  L_i = f_i − λ̃·g̃ − μ̃·h̃   (game.jl:98-118)."""
        tau, lam = x[:self.primal], x[self.primal:]
        if not tau.requires_grad:
            tau = tau.clone().requires_grad_(True)
        g = self.shared_equality(tau, th)
        h = self.shared_inequality(tau, th)
        shared = (lam * g).sum() + (y * h).sum()
        grads = []
        for i, f in enumerate(self.costs(tau, th)):
            dL = torch.autograd.grad(f - shared, tau, create_graph=True, retain_graph=True)[0]
            grads.append(dL[6 * self.H * i: 6 * self.H * (i + 1)])
        return torch.cat(grads + [g, h])

    def F(self, x, y, s, th, eps):
        """mcp.jl:76-80."""
        t = lambda v: torch.as_tensor(np.asarray(v, dtype=np.float64))
        x, y, s, th = t(x), t(y), t(s), t(th)
        K = self.K(x, y, th).detach()
        return torch.cat([K[:self.nx], K[self.nx:] - s, s * y - eps]).numpy()

    def jacobians(self, x, y, s, th, eps):
        """Dense ∇F_z (n×n, z = [x; y; s]) and ∇F_θ (n×nθ)."""
        t = lambda v: torch.as_tensor(np.asarray(v, dtype=np.float64))
        x, y, s, th = t(x), t(y), t(s), t(th)
        nx, ny = self.nx, self.ny

        def Ffun(x_, y_, s_, th_):
            K = self.K(x_, y_, th_)
            return torch.cat([K[:nx], K[nx:] - s_, s_ * y_ - eps])

        Jx, Jy, Js, Jt = torch.autograd.functional.jacobian(Ffun, (x, y, s, th), vectorize=True)
        return torch.cat([Jx, Jy, Js], dim=1).numpy(), Jt.numpy()
