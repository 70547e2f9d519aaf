"""Builders for the workloads named in BASELINE.json `configs` — the *definitions* of the problems,
restated from the reference's tests, benchmark and examples.  They only produce traced
`PrimalDualMCP`s / `ParametricGame`s and synthetic θ batches; no solver logic lives here.

cfg1  README / test QP             `/root/reference/test/runtests.jl:16-22`, `README.md:51-65`
test  2-player clamp game          `test/runtests.jl:88-116`
cfg2  random convex QP             `benchmark/quadratic_program_benchmark.jl:7-90`
cfg3  lane-change trajectory game  `examples/lane_change.jl:2-55`, `examples/utils.jl:2-178`,
                                   `benchmark/trajectory_game_benchmark.jl:36-87`
"""
from __future__ import annotations

import numpy as np

from .game import Blocks, OptimizationProblem, ParametricGame
from .mcp import PrimalDualMCP


# ------------------------------------------------------------------------------------------------
# cfg1: README QP   min ½xᵀMx − θᵀx  s.t. Ax − b ≥ 0     (`test/runtests.jl:9-22`)
# ------------------------------------------------------------------------------------------------
README_M = np.array([[2.0, 1.0], [1.0, 2.0]])
README_A = np.eye(2)
README_b = np.array([1.0, 1.0])


def readme_qp(compute_sensitivities: bool = True) -> PrimalDualMCP:
    def G(x, y, θ):
        return README_M @ x - θ - README_A.T @ y          # `test/runtests.jl:21`

    def H(x, y, θ):
        return README_A @ x - README_b                    # `test/runtests.jl:22`

    return PrimalDualMCP(G, H, unconstrained_dimension=2, constrained_dimension=2,
                         parameter_dimension=2, compute_sensitivities=compute_sensitivities)


def readme_qp_from_K(compute_sensitivities: bool = True) -> PrimalDualMCP:
    """`AlternativeCallableConstructor` — `test/runtests.jl:23-28,53-63`."""
    def K(z, θ):
        x, y = z[:2], z[2:]
        return np.concatenate([README_M @ x - θ - README_A.T @ y, README_A @ x - README_b])

    lb = np.concatenate([np.full(2, -np.inf), np.zeros(2)])
    ub = np.full(4, np.inf)
    return PrimalDualMCP.from_K(K, lb, ub, parameter_dimension=2,
                                compute_sensitivities=compute_sensitivities)


def readme_qp_thetas(B: int, seed: int = 1) -> np.ndarray:
    """θ ~ U[0,1)² (`README.md:54` uses `rand(rng, 2)`); returned column-major nθ×B like Julia."""
    rng = np.random.default_rng(seed)
    return np.asfortranarray(rng.random((2, B)))


# ------------------------------------------------------------------------------------------------
# test game: each player  min ‖x_i − θ_i‖²  s.t. |x_i| ≤ lim     (`test/runtests.jl:88-106`)
# ------------------------------------------------------------------------------------------------
def clamp_game(lim: float = 0.5) -> ParametricGame:
    def make(i):
        return OptimizationProblem(
            objective=lambda x, θi: sum((x[i] - θi) ** 2),
            private_inequality=lambda x, θi: np.concatenate([-x[i] + lim, x[i] + lim]),
        )

    return ParametricGame(test_point=[[1, 1], [1, 1]], test_parameter=[[1, 1], [1, 1]],
                          problems=[make(0), make(1)])


# ------------------------------------------------------------------------------------------------
# cfg2: random convex QP, parameters carried in θ = [vec(M); vec(A); b; ϕ]
#       (`benchmark/quadratic_program_benchmark.jl:51-90`)
# ------------------------------------------------------------------------------------------------
def qp_unpack(θ, num_primals, num_inequalities):
    """`unpack_parameters` — `quadratic_program_benchmark.jl:77-90` (column-major reshapes)."""
    n, m = num_primals, num_inequalities
    M = θ[: n * n].reshape((n, n), order="F")
    A = θ[n * n: n * n + m * n].reshape((m, n), order="F")
    b = θ[n * n + m * n: n * n + m * (n + 1)]
    ϕ = θ[n * n + m * (n + 1):]
    return M, A, b, ϕ


def random_qp(num_primals: int = 100, num_inequalities: int = 100,
              compute_sensitivities: bool = False) -> PrimalDualMCP:
    """`generate_test_problem(::QuadraticProgramBenchmark)` — `quadratic_program_benchmark.jl:7-48`.
    Built through the (G, H) form: the reference's own `K(z, θ)` is positional (`:34`) and cannot be
    called by the keyword-calling constructor (`src/mcp.jl:166`)."""
    n, m = num_primals, num_inequalities

    def G(x, y, θ):
        M, A, b, ϕ = qp_unpack(θ, n, m)
        return M @ x - ϕ - A.T @ y                         # :20

    def H(x, y, θ):
        M, A, b, ϕ = qp_unpack(θ, n, m)
        return A @ x - b                                   # :31

    return PrimalDualMCP(G, H, unconstrained_dimension=n, constrained_dimension=m,
                         parameter_dimension=n * n + m * n + m + n,
                         compute_sensitivities=compute_sensitivities)


def random_qp_theta(rng: np.random.Generator, num_primals=100, num_inequalities=100,
                    sparsity_rate=0.9) -> np.ndarray:
    """`generate_random_parameter(::QuadraticProgramBenchmark)` — `:51-74`.  (Julia's MersenneTwister
    stream cannot be reproduced; the same arrays are fed to oracle, CPU baseline and GPU.)"""
    n, m = num_primals, num_inequalities
    P = rng.standard_normal((n, n)) * (rng.random((n, n)) < (1 - sparsity_rate))
    M = P.T @ P
    A = rng.standard_normal((m, n)) * (rng.random((m, n)) < (1 - sparsity_rate))
    b = rng.standard_normal(m)
    ϕ = rng.standard_normal(n)
    return np.concatenate([M.reshape(-1, order="F"), A.reshape(-1, order="F"), b, ϕ])


def random_qp_thetas(B: int, seed: int = 1, **kw) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return np.asfortranarray(np.stack([random_qp_theta(rng, **kw) for _ in range(B)], axis=1))


# ------------------------------------------------------------------------------------------------
# cfg3: 2-player lane-change trajectory game
# ------------------------------------------------------------------------------------------------
DT = 0.1
# planar double integrator (TrajectoryGamesExamples.planar_double_integrator, dt=0.1, m=1;
# SURVEY.md Appendix A): state (px, py, vx, vy), control (ax, ay)
DI_A = np.array([[1, 0, DT, 0], [0, 1, 0, DT], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float64)
DI_B = np.array([[0.5 * DT * DT, 0], [0, 0.5 * DT * DT], [DT, 0], [0, DT]], dtype=np.float64)
STATE_LB = np.array([-np.inf, -np.inf, -10.0, 0.0])       # `examples/lane_change.jl:49`
STATE_UB = np.array([np.inf, np.inf, 10.0, 10.0])
CONTROL_LB = np.array([-5.0, -5.0])                       # `examples/lane_change.jl:50`
CONTROL_UB = np.array([3.0, 3.0])


def road_environment(lane_width=2.0, num_lanes=2, height=50.0):
    """`setup_road_environment` — `examples/lane_change.jl:2-12`: lane centres and the rectangle's
    half-spaces a·p ≤ b in counter-clockwise edge order (LazySets `constraints_list` of the VPolygon)."""
    centers = [(i - 0.5) * lane_width for i in range(1, num_lanes + 1)]
    x0, x1 = centers[0] - 0.5 * lane_width, centers[-1] + 0.5 * lane_width
    halfspaces = [((0.0, -1.0), 0.0), ((1.0, 0.0), x1), ((0.0, 1.0), height), ((-1.0, 0.0), -x0)]
    return centers, halfspaces, (x0, x1, 0.0, height)


def _unpack_trajectory(x: Blocks, horizon: int, n_players: int = 2):
    """`unpack_trajectory` — `examples/utils.jl:2-16`: per player [states(4×H, time-major); controls(2×H)];
    returns joint per-time lists xs[t] (8,) and us[t] (4,) stacked over players."""
    xs, us = [], []
    for t in range(horizon):
        xs.append(np.concatenate([x[i][4 * t: 4 * t + 4] for i in range(n_players)]))
        us.append(np.concatenate([x[i][4 * horizon + 2 * t: 4 * horizon + 2 * t + 2] for i in range(n_players)]))
    return xs, us


def box_constraints(v, lb, ub):
    """`TrajectoryGamesBase.get_constraints_from_box_bounds`:
    [v[finite lb] − lb; ub − v[finite ub]]."""
    lo = [v[i] - lb[i] for i in range(len(lb)) if np.isfinite(lb[i])]
    hi = [ub[i] - v[i] for i in range(len(ub)) if np.isfinite(ub[i])]
    return lo + hi


def lane_change_game(horizon: int = 10, height: float = 50.0, num_lanes: int = 2, lane_width: float = 2.0,
                     compute_sensitivities: bool = True) -> ParametricGame:
    """`generate_test_problem(::TrajectoryGameBenchmark)` — `benchmark/trajectory_game_benchmark.jl:36-57`
    → `build_mcp_components` (`examples/utils.jl:87-178`) on `setup_trajectory_game`
    (`examples/lane_change.jl:15-55`), `params_per_player = 1`.  nx=200, ny=250, nθ=10 at H=10."""
    _, halfspaces, _ = road_environment(lane_width, num_lanes, height)
    N, H = 2, horizon

    def stage_cost(ii):
        def cost(xj, uj, t, θi):                           # `examples/lane_change.jl:18-24`
            xi, ui = xj[4 * ii: 4 * ii + 4], uj[2 * ii: 2 * ii + 2]
            lane_preference = θi[-1]
            return ((xi[0] - lane_preference) ** 2
                    + 0.5 * ((xi[2] - 0.0) ** 2 + (xi[3] - 2.0) ** 2)
                    + 0.1 * (ui[0] ** 2 + ui[1] ** 2))
        return cost

    def objective(ii):
        cost = stage_cost(ii)

        def f(x, θi):                                      # `examples/utils.jl:96-106`
            xs, us = _unpack_trajectory(x, H, N)
            total = 0.0
            for t in range(H):                             # discount_factor = 1.0 (`lane_change.jl:35`)
                total = total + cost(xs[t], us[t], t, θi)
            return total / H                               # reducer (`lane_change.jl:27-29`)
        return f

    def shared_equality(x, θ):                             # `examples/utils.jl:109-123`
        xs, us = _unpack_trajectory(x, H, N)
        init = np.concatenate([θ[i][:4] for i in range(N)])   # unpack_parameters, `utils.jl:32-41`
        rows = list(xs[0] - init)
        for t in range(1, H):
            nxt = np.concatenate([DI_A @ xs[t - 1][4 * i: 4 * i + 4] + DI_B @ us[t - 1][2 * i: 2 * i + 2]
                                  for i in range(N)])
            rows += list(xs[t] - nxt)
        return np.array(rows, dtype=object)

    def shared_inequality(x, θ):                           # `examples/utils.jl:126-155`
        xs, us = _unpack_trajectory(x, H, N)
        h1 = []
        for xj in xs:                                      # coupling, `examples/lane_change.jl:39-46`
            d0, d1 = xj[0] - xj[4], xj[1] - xj[5]
            h1.append(d0 ** 2 + d1 ** 2 - 4.0)
        h2 = []
        for xj in xs:                                      # environment: product(constraints, positions)
            for i in range(N):
                p = xj[4 * i: 4 * i + 2]
                for (a, b) in halfspaces:
                    h2.append(-(a[0] * p[0] + a[1] * p[1]) + b)
        h3 = []
        clb, cub = np.tile(CONTROL_LB, N), np.tile(CONTROL_UB, N)
        for uj in us:                                      # actuator limits, `utils.jl:140-145`
            h3 += box_constraints(uj, clb, cub)
        h4 = []
        slb, sub = np.tile(STATE_LB, N), np.tile(STATE_UB, N)
        for xj in xs:                                      # state limits, `utils.jl:147-152`
            h4 += box_constraints(xj, slb, sub)
        return np.array(h1 + h2 + h3 + h4, dtype=object)

    primal_dim = H * (4 + 2)                               # `utils.jl:157-160`
    return ParametricGame(
        test_point=[np.zeros(primal_dim)] * N,
        test_parameter=[np.zeros(4 + 1)] * N,             # state + params_per_player=1 (`:169-171`)
        problems=[OptimizationProblem(objective=objective(i)) for i in range(N)],
        shared_equality=shared_equality, shared_inequality=shared_inequality,
        compute_sensitivities=compute_sensitivities)


def lane_change_thetas(B: int, seed: int = 1, num_lanes=2, lane_width=2.0, height=50.0,
                       moving: bool = False) -> np.ndarray:
    """`generate_random_parameter(::TrajectoryGameBenchmark)` — `trajectory_game_benchmark.jl:62-87`:
    θ = [p₁ ~ U(road), 0, 0, lane₁; p₂ ~ U(road), 0, 0, lane₂], lanes drawn from the lane centres.

    NB the benchmark's zero initial velocity sits exactly on the state bound v_y ≥ 0
    (`examples/lane_change.jl:49`), so the equality `x₁ = initial_state` and that inequality are linearly
    dependent: ∇F_z is numerically singular at the solution (cond ≈ 1e16) and sensitivities are not well
    defined there.  `moving=True` draws v_x ~ U(-1,1), v_y ~ U(0.5,3) instead (the example's own start has
    v_y = 1, `examples/lane_change.jl:58`) for well-posed sensitivity tests."""
    centers, _, (x0, x1, y0, y1) = road_environment(lane_width, num_lanes, height)
    rng = np.random.default_rng(seed)
    θ = np.zeros((10, B), order="F")
    for i in range(2):
        θ[5 * i + 0] = rng.uniform(x0, x1, B)
        θ[5 * i + 1] = rng.uniform(y0, y1, B)
        θ[5 * i + 4] = rng.choice(np.asarray(centers), B)
        if moving:
            θ[5 * i + 2] = rng.uniform(-1.0, 1.0, B)
            θ[5 * i + 3] = rng.uniform(0.5, 3.0, B)
    return θ


def lane_change_zero_input_x0(θ: np.ndarray, horizon: int = 10, n_eq: int = 80) -> np.ndarray:
    """Zero-input rollout initial guess — `examples/utils.jl:181-192,219-227`: constant-velocity
    states, zero controls, zero equality multipliers.  θ is nθ×B; returns nx×B."""
    B = θ.shape[1]
    H = horizon
    out = np.zeros((2 * 6 * H + n_eq, B), order="F")
    for i in range(2):
        st = θ[5 * i: 5 * i + 4].copy()
        base = i * 6 * H
        for t in range(H):
            out[base + 4 * t: base + 4 * t + 4] = st
            st = DI_A @ st
    return out


# ------------------------------------------------------------------------------------------------
# cfg4: N-player masked trajectory game (player selection), horizon 30
#       `examples/train_and_test_utils.jl:362-401` (`setup_trajectory_game(; environment, N)`),
#       environment = square of side 10 (`setup_road_environment(; length)`, `:341-349`),
#       built with `params_per_player = N + 2` (`examples/time_test.jl:23-24`): θ_i = [state(4); goal(2); mask(N)]
# ------------------------------------------------------------------------------------------------
def masked_game(N: int = 4, horizon: int = 30, length: float = 10.0,
                compute_sensitivities: bool = True) -> ParametricGame:
    """nx = 10·N·H, ny = H·(12N+1), nθ = N·(N+6)  (SURVEY.md §8 table)."""
    H = horizon
    h = 0.5 * length
    halfspaces = [((0.0, -1.0), h), ((1.0, 0.0), h), ((0.0, 1.0), h), ((-1.0, 0.0), h)]
    state_lb, state_ub = np.array([-np.inf, -np.inf, -2.0, -2.0]), np.array([np.inf, np.inf, 2.0, 2.0])   # :394
    ctrl_lb, ctrl_ub = np.array([-1.0, -1.0]), np.array([1.0, 1.0])                                        # :395

    def objective(ii):
        def f(x, θi):                                        # stage cost `:364-370`, mean over time `:372-374`
            xs, us = _unpack_trajectory(x, H, N)
            goal, mask = θi[4:6], θi[6:6 + N]                # θi[end-(N+1):end-N], θi[end-(N-1):end]
            total = 0.0
            for t in range(H):
                xi, ui = xs[t][4 * ii: 4 * ii + 4], us[t][2 * ii: 2 * ii + 2]
                c = ((xi[0] - goal[0]) ** 2 + (xi[1] - goal[1]) ** 2 + xi[2] ** 2 + xi[3] ** 2
                     + 0.1 * (ui[0] ** 2 + ui[1] ** 2))
                for jj in range(N):
                    if jj != ii:
                        xj = xs[t][4 * jj: 4 * jj + 4]
                        c = c + 2.0 * (mask[ii] * mask[jj]) / ((xi[0] - xj[0]) ** 2 + (xi[1] - xj[1]) ** 2)
                total = total + c
            return total / H
        return f

    def shared_equality(x, θ):                               # `examples/utils.jl:109-123`
        xs, us = _unpack_trajectory(x, H, N)
        init = np.concatenate([θ[i][:4] for i in range(N)])
        rows = list(xs[0] - init)
        for t in range(1, H):
            nxt = np.concatenate([DI_A @ xs[t - 1][4 * i: 4 * i + 4] + DI_B @ us[t - 1][2 * i: 2 * i + 2]
                                  for i in range(N)])
            rows += list(xs[t] - nxt)
        return np.array(rows, dtype=object)

    def shared_inequality(x, θ):                             # `examples/utils.jl:126-155`
        xs, us = _unpack_trajectory(x, H, N)
        h1 = [1.0 + 0.0 * xs[t][0] for t in range(H)]        # coupling_constraints returns [1] per stage (`:380-388`)
        h2 = []
        for xj in xs:
            for i in range(N):
                p = xj[4 * i: 4 * i + 2]
                for (a, b) in halfspaces:
                    h2.append(-(a[0] * p[0] + a[1] * p[1]) + b)
        h3, h4 = [], []
        clb, cub = np.tile(ctrl_lb, N), np.tile(ctrl_ub, N)
        slb, sub = np.tile(state_lb, N), np.tile(state_ub, N)
        for uj in us:
            h3 += box_constraints(uj, clb, cub)
        for xj in xs:
            h4 += box_constraints(xj, slb, sub)
        return np.array(h1 + h2 + h3 + h4, dtype=object)

    return ParametricGame(
        test_point=[np.zeros(6 * H)] * N, test_parameter=[np.zeros(4 + 2 + N)] * N,
        problems=[OptimizationProblem(objective=objective(i)) for i in range(N)],
        shared_equality=shared_equality, shared_inequality=shared_inequality,
        compute_sensitivities=compute_sensitivities)


def masked_game_thetas(B: int, N: int = 4, seed: int = 1, side: float = 5.0, min_dist: float = 1.0,
                       all_masks: bool = True) -> np.ndarray:
    """Scenario distribution of `scripts/data_generation.py:5-38,52-55`: positions and goals ~ U[-side/2, side/2]²
    with pairwise distance ≥ min_dist, zero velocity; θ_i = [state(4); goal(2); mask(N)], player 1 carries the
    mask (ego = 1, the others swept over {0,1}), the other players all ones
    (`examples/parametric_masked_game_solver.jl:19`, `game_with_masks.jl:13,24`)."""
    rng = np.random.default_rng(seed)

    def spread():
        pts = []
        while len(pts) < N:
            p = rng.uniform(-0.5 * side, 0.5 * side, 2)
            if all(np.hypot(*(p - q)) >= min_dist for q in pts):
                pts.append(p)
        return np.array(pts)

    nθ = N * (N + 6)
    θ = np.zeros((nθ, B), order="F")
    n_masks = 2 ** (N - 1)
    scenario = None
    for b in range(B):
        if scenario is None or not all_masks or b % n_masks == 0:
            scenario = (spread(), spread())
        pos, goal = scenario
        m = b % n_masks if all_masks else n_masks - 1
        ego_mask = np.array([1.0] + [float((m >> k) & 1) for k in range(N - 1)])
        for i in range(N):
            base = i * (N + 6)
            θ[base: base + 2, b] = pos[i]
            θ[base + 4: base + 6, b] = goal[i]
            θ[base + 6: base + 6 + N, b] = ego_mask if i == 0 else 1.0
    return θ


def masked_game_x0(θ: np.ndarray, N: int = 4, horizon: int = 30) -> np.ndarray:
    """Stay-at-rest rollout initial guess (zero-input trajectory from zero velocity, `examples/utils.jl:181-192`)."""
    B = θ.shape[1]
    H = horizon
    out = np.zeros((10 * N * H, B), order="F")
    for i in range(N):
        st = θ[i * (N + 6): i * (N + 6) + 4]
        for t in range(H):
            out[i * 6 * H + 4 * t: i * 6 * H + 4 * t + 4] = st
    return out
