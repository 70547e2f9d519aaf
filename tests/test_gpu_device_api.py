"""GPU tests of the device-resident front-end (mcp_b200/torch_api.py, receding_horizon.py): the §8(f) rows
"fused solve + loss + VJP" and "warm-started receding-horizon loop on device"."""
import numpy as np
import pytest

from mcp_b200 import InteriorPoint, problems, solve, solve_pullback

pytestmark = pytest.mark.gpu


def test_autograd_matches_reference_gradient(readme_mcp):
    """`Zygote.gradient(θ -> Σx²+Σy², θ)` of the README QP — test/runtests.jl:75-84 — through torch autograd,
    all tensors resident on the GPU."""
    import torch
    from mcp_b200.torch_api import MCPSolve
    Θ = torch.tensor([[-0.5, 0.5], [0.3, 0.7]], dtype=torch.float64, device="cuda", requires_grad=True)
    x, y, s, status = MCPSolve.apply(readme_mcp, Θ)
    loss = (x ** 2).sum() + (y ** 2).sum()
    loss.backward()
    assert torch.all(status == 0)
    np.testing.assert_allclose(Θ.grad[0].cpu().numpy(), [-7.0000583, -4.9997786], atol=1e-5)
    # second instance against the host path
    sol = solve(InteriorPoint(), readme_mcp, np.array([0.3, 0.7]))
    g = solve_pullback(readme_mcp, sol, np.array([0.3, 0.7]), 2 * sol.x, 2 * sol.y, None)
    np.testing.assert_allclose(Θ.grad[1].cpu().numpy(), g, rtol=1e-9, atol=1e-12)


def test_autograd_lane_change_batch(lane_game):
    import torch
    from mcp_b200.torch_api import MCPSolve
    mcp = lane_game.mcp
    Θh = problems.lane_change_thetas(64, seed=9, moving=True)
    Θ = torch.tensor(np.ascontiguousarray(Θh.T), device="cuda", requires_grad=True)
    x, y, s, status = MCPSolve.apply(mcp, Θ, None, None, dict(tol=1e-6))
    (x[:, :4] ** 2).sum().backward()                      # a loss on the first planned state of player 1
    sol = solve(InteriorPoint(), mcp, Θh, tol=1e-6)
    dx = np.zeros_like(sol.x)
    dx[:4] = 2 * sol.x[:4]
    g = solve_pullback(mcp, sol, Θh, dx, None, None)
    ok = sol.status == 0
    np.testing.assert_array_equal(status.cpu().numpy(), sol.status)
    np.testing.assert_allclose(Θ.grad.cpu().numpy().T[:, ok], g[:, ok], rtol=1e-9, atol=1e-9)
    # finite-difference check of one well-posed instance, lane preference of player 1 (θ[4])
    b = int(np.nonzero(ok)[0][0])
    h = 1e-5
    vals = []
    for sgn in (+1, -1):
        θb = Θh[:, b].copy()
        θb[4] += sgn * h
        vals.append(np.sum(solve(InteriorPoint(), mcp, θb, tol=1e-6).x[:4] ** 2))
    fd = (vals[0] - vals[1]) / (2 * h)
    assert abs(fd - g[4, b]) <= 1e-3 * max(1.0, abs(fd))


def test_receding_horizon_matches_cpu_loop(lane_game):
    """Three closed-loop steps of 32 games: the device-resident loop against the same logic with the C oracle
    (warm start from the last solved solution, zero-input rollout otherwise; `examples/utils.jl:195-235`)."""
    import torch
    from mcp_b200.receding_horizon import BatchedRecedingHorizon
    from oracle import c_oracle as CO
    mcp = lane_game.mcp
    B, H = 32, 10
    Θ0 = problems.lane_change_thetas(B, seed=21, moving=True)
    state = np.stack([Θ0[0:4].T, Θ0[5:9].T], axis=1)                 # [B, 2, 4]
    params = np.stack([Θ0[4:5].T, Θ0[9:10].T], axis=1)               # [B, 2, 1]
    rh = BatchedRecedingHorizon(lane_game, horizon=H, tol=1e-4)
    st_d, pr_d = torch.tensor(state, device="cuda"), torch.tensor(params, device="cuda")
    st_c = state.copy()
    last = None
    for step in range(3):
        st_d, sol_d = rh.step(st_d, pr_d)
        # CPU loop
        θ = np.concatenate([st_c, params], axis=2).reshape(B, -1).T
        x0 = problems.lane_change_zero_input_x0(np.asfortranarray(θ), H)
        y0 = np.ones((250, B))
        if last is not None:
            x0 = np.where(last[2][None, :], last[0], x0)
            y0 = np.where(last[2][None, :], last[1], y0)
        ref = CO.solve_batch(mcp.ir, θ, x0=np.asfortranarray(x0), y0=np.asfortranarray(y0), tol=1e-4)
        ok = ref.status == 0
        last = (np.where(ok[None, :], ref.x, last[0]) if last else ref.x.copy(),
                np.where(ok[None, :], ref.y, last[1]) if last else ref.y.copy(), ok | (last[2] if last else False))
        u = np.stack([ref.x[40:42].T, ref.x[100:102].T], axis=1)     # first control of each player
        st_c = st_c @ problems.DI_A.T + u @ problems.DI_B.T
        np.testing.assert_array_equal(sol_d["status"].cpu().numpy(), ref.status)
        both = ok
        assert both.sum() >= B // 2
        np.testing.assert_allclose(st_d.cpu().numpy()[both], st_c[both], rtol=1e-6, atol=1e-7)
        st_c[~both] = st_d.cpu().numpy()[~both]    # failed instances wander: resynchronise them


def test_receding_horizon_masked_game_matches_cpu_loop():
    """f2 for the masked N-player game (`examples/parametric_masked_game_solver.jl:19-42`): 3 closed-loop steps of all
    8 ego masks of 4 scenarios (N = 4, horizon 10), device-resident loop against the same logic driven by the C oracle —
    goal/mask packing (player 1 carries the mask), warm start from the last solved plan, stay-at-rest rollout otherwise."""
    import torch
    from mcp_b200.receding_horizon import BatchedRecedingHorizon, masked_game_parameters
    from oracle import c_oracle as CO
    N, H, B = 4, 10, 32
    game = problems.masked_game(N, H)
    mcp = game.mcp
    Θ0 = problems.masked_game_thetas(B, N, seed=5)                       # θ_i = [state(4); goal(2); mask(N)]
    blocks = Θ0.T.reshape(B, N, 6 + N)
    state, goals, masks = blocks[:, :, :4].copy(), blocks[:, :, 4:6].copy(), blocks[:, 0, 6:].copy()
    rh = BatchedRecedingHorizon(game, tol=1e-4)
    assert (rh.N, rh.H, rh.ppp) == (N, H, N + 2)
    st_d = torch.tensor(state, device="cuda")
    pr_d = masked_game_parameters(torch.tensor(goals, device="cuda"), torch.tensor(masks, device="cuda"))
    np.testing.assert_array_equal(pr_d.cpu().numpy(), blocks[:, :, 4:])     # the packing reproduces the data generator's θ
    st_c, last = state.copy(), None
    nx, ny = mcp.unconstrained_dimension, mcp.constrained_dimension
    for step in range(3):
        st_d, sol_d = rh.step(st_d, pr_d)
        θ = np.concatenate([st_c, blocks[:, :, 4:]], axis=2).reshape(B, -1).T
        x0 = problems.masked_game_x0(np.asfortranarray(θ), N, H)
        # (a moving start: the rollout guess is A^t x, not "stay at rest")
        for i in range(N):
            stt = st_c[:, i].T.copy()
            for t in range(H):
                x0[i * 6 * H + 4 * t: i * 6 * H + 4 * t + 4] = stt
                stt = problems.DI_A @ stt
        y0 = np.ones((ny, B))
        if last is not None:
            x0 = np.where(last[2][None, :], last[0], x0)
            y0 = np.where(last[2][None, :], last[1], y0)
        ref = CO.solve_batch(mcp.ir, θ, x0=np.asfortranarray(x0), y0=np.asfortranarray(y0), tol=1e-4)
        ok = ref.status == 0
        last = (np.where(ok[None, :], ref.x, last[0]) if last else ref.x.copy(),
                np.where(ok[None, :], ref.y, last[1]) if last else ref.y.copy(), ok | (last[2] if last else False))
        u = np.stack([ref.x[i * 6 * H + 4 * H: i * 6 * H + 4 * H + 2].T for i in range(N)], axis=1)
        st_c = st_c @ problems.DI_A.T + u @ problems.DI_B.T
        np.testing.assert_array_equal(sol_d["status"].cpu().numpy(), ref.status)
        assert ok.sum() >= (3 * B) // 4
        np.testing.assert_allclose(st_d.cpu().numpy()[ok], st_c[ok], rtol=1e-6, atol=1e-7)
        st_c[~ok] = st_d.cpu().numpy()[~ok]
    # masking players out changes the ego's closed-loop path: the 8 masks of one scenario must not all coincide
    ego = st_d.cpu().numpy()[:8, 0, :2]
    assert np.max(np.abs(ego - ego[7])) > 1e-4


def test_receding_horizon_turn_length(lane_game):
    """turn_length = 2: one plan, two simulated steps along it (`examples/utils.jl:287-307`)."""
    import torch
    from mcp_b200.receding_horizon import BatchedRecedingHorizon
    B = 8
    Θ0 = problems.lane_change_thetas(B, seed=3, moving=True)
    state = torch.tensor(np.stack([Θ0[0:4].T, Θ0[5:9].T], axis=1), device="cuda")
    params = torch.tensor(np.stack([Θ0[4:5].T, Θ0[9:10].T], axis=1), device="cuda")
    rh = BatchedRecedingHorizon(lane_game, turn_length=2, tol=1e-4)
    nxt, sol = rh.step(state, params)
    u0, u1 = rh.controls(sol, 0), rh.controls(sol, 1)
    want = rh.advance(rh.advance(state, u0), u1)
    np.testing.assert_allclose(nxt.cpu().numpy(), want.cpu().numpy(), rtol=0, atol=0)
    # the planned second state of a solved instance is where the dynamics put it after the first control
    ok = (sol["status"] == 0).cpu().numpy()
    H = rh.H
    planned = sol["x"][:, 4:8].cpu().numpy()                 # player 1, stage 2 state
    np.testing.assert_allclose(planned[ok], rh.advance(state, u0)[:, 0].cpu().numpy()[ok], atol=1e-3)
