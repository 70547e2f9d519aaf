"""CPU tests of the oracle: pinned against every known answer the reference's own tests hold
(`/root/reference/test/runtests.jl`), against the committed golden vectors, and the C restatement
against the Python one.  Also covers the tracer/IR (the analogue of the reference's symbolic layer).
"""
import json
import math
import os

import numpy as np
import pytest

from mcp_b200 import problems, trace as T
from mcp_b200.mcp import PrimalDualMCP
from oracle import c_oracle as CO
from oracle import ip_oracle as O
from oracle.ir_eval import OracleMCP

GOLD = os.path.join(os.path.dirname(__file__), "golden")
θ_TEST = np.array([-0.5, 0.5])   # test/runtests.jl:19


def check_solution(sol, θ):
    """`check_solution` — test/runtests.jl:30-38, verbatim tolerances."""
    G = problems.README_M @ sol.x - θ - problems.README_A.T @ sol.y
    H = problems.README_A @ sol.x - problems.README_b
    assert np.all(np.abs(G) <= 5e-3)
    assert np.all(H >= 0)
    assert np.all(sol.y >= 0)
    assert sol.y @ H <= 5e-3
    assert np.all(sol.s <= 5e-3)
    assert sol.kkt_error <= 5e-3
    assert sol.status == "solved"


def test_basic_callable_constructor():
    """test/runtests.jl:40-51."""
    om = OracleMCP(problems.readme_qp().ir)
    sol = O.solve_interior_point(om, θ_TEST)
    check_solution(sol, θ_TEST)
    # analytic KKT point the assertions imply: x* = [1,1], y* = Mx* − θ = [3.5, 2.5]
    np.testing.assert_allclose(sol.x, [1, 1], atol=5e-3)
    np.testing.assert_allclose(sol.y, [3.5, 2.5], atol=5e-3)


def test_alternative_callable_constructor():
    """test/runtests.jl:53-63: K(z; θ) + bounds gives the same MCP."""
    a = O.solve_interior_point(OracleMCP(problems.readme_qp().ir), θ_TEST)
    b = O.solve_interior_point(OracleMCP(problems.readme_qp_from_K().ir), θ_TEST)
    check_solution(b, θ_TEST)
    np.testing.assert_allclose(a.x, b.x, rtol=1e-12)
    np.testing.assert_allclose(a.y, b.y, rtol=1e-12)
    assert a.outer_iters == b.outer_iters


def test_autodifferentiation():
    """test/runtests.jl:65-85: reverse ≈ forward ≈ finite differences at atol 1e-3; ∇f = [−7, −5]."""
    om = OracleMCP(problems.readme_qp().ir)

    def f(θ):
        sol = O.solve_interior_point(om, θ)
        return np.sum(sol.x ** 2) + np.sum(sol.y ** 2)

    sol = O.solve_interior_point(om, θ_TEST)
    rev = O.vjp_theta(om, sol, θ_TEST, 2 * sol.x, 2 * sol.y, np.zeros(2))
    fwd = np.zeros(2)
    for q in range(2):
        e = np.zeros(2)
        e[q] = 1.0
        xp, yp, sp = O.jvp_theta(om, sol, θ_TEST, e)
        fwd[q] = 2 * sol.x @ xp + 2 * sol.y @ yp
    fd = np.array([(f(θ_TEST + h) - f(θ_TEST - h)) / 2e-6 for h in (np.array([1e-6, 0]), np.array([0, 1e-6]))])
    np.testing.assert_allclose(rev, fd, atol=1e-3)
    np.testing.assert_allclose(rev, fwd, atol=1e-3)
    np.testing.assert_allclose(rev, [-7, -5], atol=1e-3)


def test_missing_sensitivities_raises():
    """src/AutoDiff.jl:19-23."""
    om = OracleMCP(problems.readme_qp(compute_sensitivities=False).ir)
    sol = O.solve_interior_point(om, θ_TEST)
    with pytest.raises(ValueError, match="Missing sensitivities"):
        O.solve_jacobian_theta(om, sol, θ_TEST)


def test_parametric_game():
    """test/runtests.jl:88-116: primals ≈ clamp(θ_i, −0.5, 0.5) at atol 10·tol."""
    game = problems.clamp_game()
    assert game.dims.x == [2, 2] and game.dims.μ == [4, 4] and game.dims.λ == [0, 0]
    om = OracleMCP(game.mcp.ir)
    θ = np.array([-1.0, 0.0, 1.0, 1.0])
    tol = 1e-4
    sol = O.solve_interior_point(om, θ, tol=tol)
    assert sol.status == "solved"
    np.testing.assert_allclose(sol.x[:2], np.clip(θ[:2], -0.5, 0.5), atol=10 * tol)
    np.testing.assert_allclose(sol.x[2:4], np.clip(θ[2:], -0.5, 0.5), atol=10 * tol)


# ---- linesearch semantics (src/solver.jl:127-138) -----------------------------------------------
def test_linesearch_semantics():
    ls = O.fraction_to_the_boundary_linesearch
    assert ls(np.array([]), np.array([])) == 1.0                       # empty ⇒ 1.0
    assert ls(np.array([1.0]), np.array([5.0])) == 1.0                 # moving away from the boundary
    assert ls(np.array([1.0]), np.array([-0.995])) == 1.0              # lands exactly on (1−τ)v
    assert ls(np.array([1.0]), np.array([-1.0])) == 0.5
    assert ls(np.array([1.0, 1.0]), np.array([-1.0, -7.0])) == 0.125
    assert ls(np.array([1.0]), np.array([-0.9 * 2.0 ** 14])) == 2.0 ** -14  # smallest accepted step for tol = 1e-4
    assert math.isnan(ls(np.array([1.0]), np.array([-0.9 * 2.0 ** 15])))    # α < tol is tested BEFORE halving
    assert ls(np.array([1.0]), np.array([np.nan])) == 1.0              # NaN compares false ⇒ α = 1
    assert math.isnan(ls(np.array([0.0]), np.array([-1.0])))           # v = 0 can never satisfy the predicate


def test_outer_iteration_cap_marks_failed():
    """src/solver.jl:117-119: hitting max_outer_iters is a failure even if the iterate is fine."""
    om = OracleMCP(problems.readme_qp().ir)
    sol = O.solve_interior_point(om, θ_TEST, max_outer_iters=3)
    assert sol.status == "failed" and sol.outer_iters == 3


def test_infeasible_instance_fails():
    """Players spawned < 2 m apart make the lane-change game infeasible (collision constraint):
    the solver runs into the outer cap."""
    om = OracleMCP(problems.lane_change_game().mcp.ir)
    θ = np.array([1.0, 10.0, 0, 0, 1.0, 1.2, 10.1, 0, 0, 3.0])
    sol = O.solve_interior_point(om, θ, tol=1e-6)
    assert sol.status == "failed"


# ---- golden vectors ------------------------------------------------------------------------------------
def _small():
    with open(os.path.join(GOLD, "small.json")) as f:
        return json.load(f)


def test_golden_readme_qp():
    g = _small()["readme_qp_default"]
    # survey probe: inner iterations per outer pass, 10 Newton steps, outer_iters = 7
    assert g["inner_iters_per_outer"] == [4, 3, 1, 4, 1, 3] and g["newton_steps"] == 10 and g["outer_iters"] == 7
    om = OracleMCP(problems.readme_qp().ir)
    sol = O.solve_interior_point(om, g["theta"])
    np.testing.assert_allclose(sol.x, g["x"], rtol=1e-12)
    np.testing.assert_allclose(sol.y, g["y"], rtol=1e-12)
    np.testing.assert_allclose(sol.s, g["s"], rtol=1e-9)
    assert sol.eps == pytest.approx(g["eps"], rel=1e-12) and sol.kkt_error == pytest.approx(g["kkt_error"], rel=1e-9)
    np.testing.assert_allclose(O.solve_jacobian_theta(om, sol, g["theta"]), g["dzdtheta"], rtol=1e-8, atol=1e-12)


@pytest.mark.parametrize("name,builder,kw", [
    ("lane_change_seed1.npz", lambda: problems.lane_change_game().mcp, {}),
    ("random_qp_12x10_seed1.npz", lambda: problems.random_qp(12, 10), {}),
])
def test_golden_batches_python_and_c(name, builder, kw):
    d = np.load(os.path.join(GOLD, name))
    mcp = builder()
    om = OracleMCP(mcp.ir)
    Θ, tol = d["theta"], float(d["tol"])
    c = CO.solve_batch(mcp.ir, Θ, tol=tol)
    for b in range(Θ.shape[1]):
        sol = O.solve_interior_point(om, Θ[:, b], tol=tol)
        assert (sol.status == "solved") == (d["status"][b] == 0)
        assert sol.newton_steps == d["newton_steps"][b] and sol.outer_iters == d["outer_iters"][b]
        np.testing.assert_allclose(sol.x, d["x"][:, b], rtol=1e-9, atol=1e-12)
        # the C restatement (different sparse LU, same algorithm) must follow the same trajectory
        assert c.status[b] == d["status"][b]
        assert c.newton_steps[b] == d["newton_steps"][b] and c.outer_iters[b] == d["outer_iters"][b]
        scale = lambda v: max(1.0, np.max(np.abs(v)))
        assert np.max(np.abs(c.x[:, b] - d["x"][:, b])) / scale(d["x"][:, b]) < 1e-6
        assert np.max(np.abs(c.y[:, b] - d["y"][:, b])) / scale(d["y"][:, b]) < 1e-6
        assert np.max(np.abs(c.s[:, b] - d["s"][:, b])) / scale(d["s"][:, b]) < 1e-6


def test_golden_qp100_c_oracle():
    """cfg2 at the benchmark's size: θ regenerated from the seed (guarded by its SHA-256), the C restatement against the
    Python oracle's frozen trajectory."""
    import hashlib
    d = np.load(os.path.join(GOLD, "random_qp_100x100_seed5.npz"))
    Θ = problems.random_qp_thetas(int(d["B"]), seed=int(d["seed"]))
    assert hashlib.sha256(np.ascontiguousarray(Θ).tobytes()).hexdigest() == str(d["theta_sha256"])
    mcp = problems.random_qp(100, 100)
    c = CO.solve_batch(mcp.ir, Θ, tol=float(d["tol"]))
    np.testing.assert_array_equal(c.status, d["status"])
    np.testing.assert_array_equal(c.newton_steps, d["newton_steps"])
    np.testing.assert_array_equal(c.outer_iters, d["outer_iters"])
    for k in ("x", "y", "s"):
        got, want = getattr(c, k), d[k]
        assert np.max(np.abs(got - want)) / max(1.0, np.max(np.abs(want))) < 1e-9


def test_c_oracle_readme_batch_and_threads():
    g = _small()["readme_qp_batch"]
    ir = problems.readme_qp().ir
    Θ = np.array(g["theta"]).T
    one = CO.solve_batch(ir, Θ, nthreads=1)
    many = CO.solve_batch(ir, Θ, nthreads=4)
    for b, ref in enumerate(g["sols"]):
        assert one.newton_steps[b] == ref["newton_steps"] and one.outer_iters[b] == ref["outer_iters"]
        np.testing.assert_allclose(one.x[:, b], ref["x"], rtol=1e-10)
        np.testing.assert_allclose(one.y[:, b], ref["y"], rtol=1e-10)
    np.testing.assert_array_equal(one.x, many.x)          # threading must not change results
    np.testing.assert_array_equal(one.newton_steps, many.newton_steps)


def test_c_oracle_warm_start_and_empty():
    ir = problems.readme_qp().ir
    om = OracleMCP(ir)
    Θ = problems.readme_qp_thetas(4, seed=3)
    cold = CO.solve_batch(ir, Θ)
    warm = CO.solve_batch(ir, Θ, x0=cold.x, y0=np.maximum(cold.y, 1e-3))
    for b in range(4):
        ref = O.solve_interior_point(om, Θ[:, b], x0=cold.x[:, b], y0=np.maximum(cold.y[:, b], 1e-3))
        assert warm.newton_steps[b] == ref.newton_steps
        np.testing.assert_allclose(warm.x[:, b], ref.x, rtol=1e-9)
    empty = CO.solve_batch(ir, np.zeros((2, 0)))
    assert empty.x.shape == (2, 0)


# ---- tracer / IR -------------------------------------------------------------------------------------------
def _fd_jacobian(f, v, h=1e-6):
    f0 = f(v)
    J = np.zeros((len(f0), len(v)))
    for j in range(len(v)):
        e = np.zeros(len(v))
        e[j] = h
        J[:, j] = (f(v + e) - f(v - e)) / (2 * h)
    return J


def test_tracer_jacobian_matches_finite_differences():
    rng = np.random.default_rng(0)

    def G(x, y, θ):
        return np.array([x[0] * x[1] + θ[0] * np.sin(x[2]) - y[0], x[2] ** 3 / (1.0 + x[0] ** 2) + np.exp(θ[1] * x[1]),
                         np.sqrt(x[0] ** 2 + 1.0) * y[1] - np.log(2.0 + x[1] ** 2) + np.cos(θ[0])], dtype=object)

    def H(x, y, θ):
        return np.array([x[0] - θ[0] * x[1] ** 2, x[2] / (2.0 + θ[1] ** 2) + x[0] * x[1] * x[2]], dtype=object)

    mcp = PrimalDualMCP(G, H, unconstrained_dimension=3, constrained_dimension=2, parameter_dimension=2)
    om = OracleMCP(mcp.ir)
    x, y, s, θ = rng.normal(size=3), rng.random(2) + 0.5, rng.random(2) + 0.5, rng.normal(size=2)
    Jz = om.JFz(x, y, s, θ, 0.3).toarray()
    Jt = om.JFt(x, y, s, θ, 0.3).toarray()
    fz = lambda z: om.F(z[:3], z[3:5], z[5:], θ, 0.3)
    ft = lambda t: om.F(x, y, s, t, 0.3)
    np.testing.assert_allclose(Jz, _fd_jacobian(fz, np.concatenate([x, y, s])), atol=1e-7)
    np.testing.assert_allclose(Jt, _fd_jacobian(ft, θ), atol=1e-7)
    # H does not depend on y here ⇒ no structural entries in the ∇_y H block
    assert not np.any((mcp.ir.jz_rows >= 3) & (mcp.ir.jz_cols >= 3))


def test_ir_structure_lane_change():
    """Dimensions and sparsity derived in SURVEY.md §8: nx=200, ny=250, nθ=10, nnz(∇zF)=1990 of which
    1330 z-constant, nnz(∇θF)=28."""
    ir = problems.lane_change_game().mcp.ir
    assert (ir.nx, ir.ny, ir.ntheta, ir.n) == (200, 250, 10, 700)
    rows, cols, src = ir.full_jacobian_pattern()
    assert len(rows) == 1990
    assert np.all(np.diff(cols) >= 0)                      # CSC order (src/mcp.jl:110)
    assert len(ir.constant_entries()) + 250 == 1330        # IR constants + the −I block
    assert len(ir.jt_rows) == 28
    # G_y = −H_xᵀ on the primal rows
    nx = ir.nx
    gy = {(int(r), int(c - nx)) for r, c in zip(ir.jz_rows, ir.jz_cols) if r < nx and c >= nx}
    hx = {(int(c), int(r - nx)) for r, c in zip(ir.jz_rows, ir.jz_cols) if r >= nx and c < nx}
    assert gy == hx


def test_tracer_simplifications_and_sharing():
    g = T.Graph()
    x = g.variables("x", 2)
    assert (x[0] * 0.0).id == g.const(0.0).id and (x[0] + 0.0).id == x[0].id and (x[0] - x[0]).id == g.const(0.0).id
    assert (x[0] * x[1]).id == (x[1] * x[0]).id            # commutative canonicalisation ⇒ CSE
    assert (x[0] / 4.0).id == (x[0] * 0.25).id             # exact power-of-two division
    assert (-(-x[0])).id == x[0].id
    d = g.gradient(x[0] * x[0] * x[1], {T.OP_X})
    assert set(d) == {(T.OP_X, 0), (T.OP_X, 1)}


def test_bounds_assertion():
    """src/mcp.jl:191: upper bounds must be Inf and lower bounds −Inf or 0."""
    with pytest.raises(AssertionError):
        PrimalDualMCP.from_K(lambda z, θ: z, [0.0, 1.0], [np.inf, np.inf], parameter_dimension=1)
    with pytest.raises(AssertionError):
        PrimalDualMCP.from_K(lambda z, θ: z, [0.0, -np.inf], [1.0, np.inf], parameter_dimension=1)


def test_ir_structure_masked_game():
    """cfg4 dimensions and sparsity from SURVEY.md §8 (N = 4, H = 30): nx = 10·N·H, ny = H(12N+1), nθ = N(N+6),
    nnz(∇zF) = 12 970 of which 8 110 z-constant, nnz(∇θF) = 1 216; the oracle converges in 12–20 Newton steps."""
    ir = problems.masked_game(4, 30).mcp.ir
    assert (ir.nx, ir.ny, ir.ntheta) == (1200, 1470, 40)
    assert len(ir.jz_rows) + 3 * ir.ny == 12970
    assert len(ir.constant_entries()) + ir.ny == 8110
    assert len(ir.jt_rows) == 1216
    Θ = problems.masked_game_thetas(4, 4, seed=2)
    r = CO.solve_batch(ir, Θ, x0=problems.masked_game_x0(Θ, 4, 30), tol=1e-4)
    assert np.all(r.status == 0) and np.all((r.newton_steps >= 10) & (r.newton_steps <= 24))
    # θ layout: player 1 carries the swept mask, the others all ones (parametric_masked_game_solver.jl:19)
    assert np.all(Θ[6, :] == 1.0) and set(np.unique(Θ[7:10, :])) <= {0.0, 1.0} and np.all(Θ[16:20, :] == 1.0)
