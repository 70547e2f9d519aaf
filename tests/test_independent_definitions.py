"""The traced problem definitions (trace.py → ir.py → game.py → problems.py) against definitions written a second
time, directly from the reference's sources, with torch autograd (`oracle/independent_problems.py`): F, ∇F_z and
∇F_θ at random (x, y, s, θ).  This is what pins the *problems* the parity tests solve — independently of the tracer.

Reference: `/root/reference/src/game.jl:98-131`, `examples/utils.jl:87-178`, `examples/lane_change.jl:2-55`,
`examples/train_and_test_utils.jl:362-401`, `src/mcp.jl:72-80`.
"""
import numpy as np
import pytest

from mcp_b200 import problems
from oracle.independent_problems import IndependentTrajectoryGame
from oracle.ir_eval import OracleMCP


def _random_point(ind, rng, kind):
    x = rng.normal(size=ind.nx)
    y = rng.uniform(0.1, 2.0, ind.ny)
    s = rng.uniform(0.1, 2.0, ind.ny)
    th = rng.normal(size=ind.ntheta)
    if kind == "masked":
        # masks in [0, 1] like the application; positions spread so the 1/d² coupling is well scaled
        N, w = ind.N, 6 + ind.N
        for i in range(N):
            th[w * i + 6: w * (i + 1)] = rng.uniform(0.0, 1.0, N)
        H = ind.H
        for i in range(N):
            x[6 * H * i: 6 * H * i + 4 * H: 4] += 3.0 * i
    return x, y, s, th


def _check(game, ind, kind, seed, n_points=3, jac=True):
    om = OracleMCP(game.mcp.ir)
    assert (om.nx, om.ny, om.ntheta) == (ind.nx, ind.ny, ind.ntheta)
    rng = np.random.default_rng(seed)
    for _ in range(n_points):
        x, y, s, th = _random_point(ind, rng, kind)
        eps = float(rng.uniform(1e-3, 1.0))
        F_ir = om.F(x, y, s, th, eps)
        F_in = ind.F(x, y, s, th, eps)
        scale = max(1.0, np.max(np.abs(F_in)))
        assert np.max(np.abs(F_ir - F_in)) <= 1e-9 * scale
        if jac:
            Jz_in, Jt_in = ind.jacobians(x, y, s, th, eps)
            Jz_ir = om.JFz(x, y, s, th, eps).toarray()
            Jt_ir = om.JFt(x, y, s, th, eps).toarray()
            assert np.max(np.abs(Jz_ir - Jz_in)) <= 1e-9 * max(1.0, np.max(np.abs(Jz_in)))
            assert np.max(np.abs(Jt_ir - Jt_in)) <= 1e-9 * max(1.0, np.max(np.abs(Jt_in)))
            # the IR's sparsity pattern must cover every structurally non-zero entry the autograd Jacobian has
            assert np.count_nonzero(Jz_in) <= om.JFz(x, y, s, th, eps).nnz


def test_lane_change_definition_matches_independent_restatement():
    game = problems.lane_change_game()                    # BASELINE configs[2]: H = 10, nx = 200, ny = 250, nθ = 10
    ind = IndependentTrajectoryGame("lane_change", 2, 10)
    assert (ind.nx, ind.ny, ind.ntheta) == (200, 250, 10)
    _check(game, ind, "lane_change", seed=11)


@pytest.mark.parametrize("N,H", [(3, 4), (4, 6)])
def test_masked_game_definition_matches_independent_restatement(N, H):
    game = problems.masked_game(N, H)
    ind = IndependentTrajectoryGame("masked", N, H)
    _check(game, ind, "masked", seed=7 + N)


def test_masked_game_full_size_residual_matches():
    """cfg4 at the application's size (N = 4, H = 30): F only (the dense autograd Jacobian is 4140²)."""
    game = problems.masked_game(4, 30)
    ind = IndependentTrajectoryGame("masked", 4, 30)
    assert (ind.nx, ind.ny, ind.ntheta) == (1200, 1470, 40)
    _check(game, ind, "masked", seed=3, n_points=2, jac=False)
