"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): same converged status, iteration counts within ±1, x, y, s within
1e-6 relative (FP64), sensitivities within 1e-5.
"""
import numpy as np
import pytest

from mcp_b200 import InteriorPoint, solve, problems
from oracle.ir_eval import OracleMCP
from oracle import ip_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-6     # north_star: x, y, s within 1e-6 relative
SENS_TOL = 1e-5  # north_star: sensitivities within 1e-5


def rel_err(a, b):
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b))))) if len(b) else 0.0


_EXT = {}


def rounding_sensitive(mcp, θ, x_ref, tol, x0=None, y0=None, **kw):
    """True when the FP64 oracle's own result for this instance is farther than a tenth of the parity bar from the
    extended-precision trajectory (`oracle/ip_oracle_ext.py`): the trajectory then hinges on rounding and no FP64
    implementation can be held to 1e-6 on it (measured records in profiles/r2_adjudicate_*.json)."""
    from oracle import ip_oracle_ext as E
    key = id(mcp.ir)
    if key not in _EXT:
        _EXT[key] = OracleMCP(mcp.ir, extended=True)
    e = E.solve_interior_point_ext(_EXT[key], θ, x0=x0, y0=y0, tol=tol, **kw)
    return e.status != "solved" or rel_err(np.asarray(x_ref), e.x.astype(np.float64)) > 0.1 * RTOL


def golden_ill_conditioned(name):
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "ill_conditioned.json")) as f:
        return {int(k) for k in json.load(f)[name]}


def compare_batch(mcp, Θ, sol, tol, x0=None, y0=None, min_match=1.0, **kw):
    om = OracleMCP(mcp.ir)
    B = Θ.shape[1]
    n_ok = 0
    worst = 0.0
    bad = []
    for b in range(B):
        ref = O.solve_interior_point(om, Θ[:, b], tol=tol, x0=None if x0 is None else x0[:, b],
                                     y0=None if y0 is None else y0[:, b], **kw)
        same_status = (ref.status == "solved") == (sol.status[b] == 0)
        if ref.status == "solved":
            it_ok = (abs(ref.newton_steps - int(sol.newton_steps[b])) <= 1
                     and abs(ref.outer_iters - int(sol.outer_iters[b])) <= 1)
            e = max(rel_err(sol.x[:, b], ref.x), rel_err(sol.y[:, b], ref.y), rel_err(sol.s[:, b], ref.s))
        else:
            # failed (infeasible) instances wander chaotically until the outer cap — even the Python and C
            # oracles disagree on their step counts — so only the status is comparable
            it_ok, e = True, 0.0
        if same_status and it_ok and e <= RTOL:
            n_ok += 1
        else:
            bad.append((b, ref.status, int(sol.status[b]), ref.newton_steps, int(sol.newton_steps[b]), e))
        worst = max(worst, e)
    frac = n_ok / B
    assert frac >= min_match, f"only {n_ok}/{B} instances match the oracle (worst rel err {worst:.3e}); first bad: {bad[:5]}"
    return frac, worst


def test_readme_qp_known_answer(readme_mcp):
    """`test/runtests.jl:40-51` + the frozen oracle vector."""
    sol = solve(InteriorPoint(), readme_mcp, np.array([-0.5, 0.5]))
    assert sol.status == "solved"
    np.testing.assert_allclose(sol.x, [1.00007288, 1.00010203], rtol=0, atol=2e-8)
    np.testing.assert_allclose(sol.y, [3.5002478, 2.50027695], rtol=0, atol=2e-8)
    assert sol.outer_iters == 7 and sol.newton_steps == 10
    # the reference's own assertions
    G = problems.README_M @ sol.x - np.array([-0.5, 0.5]) - problems.README_A.T @ sol.y
    H = problems.README_A @ sol.x - problems.README_b
    assert np.all(np.abs(G) <= 5e-3) and np.all(H >= 0) and np.all(sol.y >= 0)
    assert sol.y @ H <= 5e-3 and np.all(sol.s <= 5e-3) and sol.kkt_error <= 5e-3


def test_readme_qp_batch(readme_mcp):
    Θ = problems.readme_qp_thetas(512, seed=1)
    sol = solve(InteriorPoint(), readme_mcp, Θ)
    compare_batch(readme_mcp, Θ, sol, tol=1e-4)
    sol6 = solve(InteriorPoint(), readme_mcp, Θ, tol=1e-6)
    compare_batch(readme_mcp, Θ, sol6, tol=1e-6)


def test_clamp_game(clamp_game):
    """`test/runtests.jl:108-115`."""
    θ = [[-1.0, 0.0], [1.0, 1.0]]
    res = solve(clamp_game, θ, tol=1e-4)
    assert res.status == "solved"
    for i, th in enumerate(θ):
        np.testing.assert_allclose(res.primals[i], np.clip(th, -0.5, 0.5), atol=1e-3)


def test_lane_change_batch(lane_game):
    mcp = lane_game.mcp
    Θ = problems.lane_change_thetas(48, seed=1)
    sol = solve(InteriorPoint(), mcp, Θ, tol=1e-6)
    frac, worst = compare_batch(mcp, Θ, sol, tol=1e-6, min_match=1.0)     # measured: 48/48
    print(f"lane-change parity: {frac:.3f} matched, worst rel err {worst:.2e}")


def test_lane_change_warm_start(lane_game):
    mcp = lane_game.mcp
    Θ = problems.lane_change_thetas(16, seed=3)
    x0 = problems.lane_change_zero_input_x0(Θ)
    sol = solve(InteriorPoint(), mcp, Θ, x0=x0, tol=1e-4)
    compare_batch(mcp, Θ, sol, tol=1e-4, x0=x0, min_match=1.0)


def test_sensitivities_readme(readme_mcp):
    from mcp_b200 import solve_jacobian_θ, solve_pullback, solve_pushforward
    θ = np.array([-0.5, 0.5])
    sol = solve(InteriorPoint(), readme_mcp, θ)
    om = OracleMCP(readme_mcp.ir)
    ref = O.solve_interior_point(om, θ)
    J = solve_jacobian_θ(readme_mcp, sol, θ)
    Jref = O.solve_jacobian_theta(om, ref, θ)
    np.testing.assert_allclose(J, Jref, rtol=SENS_TOL, atol=1e-7)
    g = solve_pullback(readme_mcp, sol, θ, 2 * sol.x, 2 * sol.y, None)
    np.testing.assert_allclose(g, [-7.0000583, -4.9997786], atol=1e-5)   # ∇(Σx²+Σy²), test/runtests.jl:75-84
    xp, yp, sp = solve_pushforward(readme_mcp, sol, θ, np.eye(2))
    np.testing.assert_allclose(np.vstack([xp, yp, sp]), Jref, rtol=SENS_TOL, atol=1e-7)


def test_sensitivities_lane_change(lane_game):
    """cfg5: batched ∂z/∂θ on well-posed instances (moving start, see `problems.lane_change_thetas`)."""
    from mcp_b200 import solve_jacobian_θ, solve_pullback
    mcp = lane_game.mcp
    Θ = problems.lane_change_thetas(24, seed=5, moving=True)
    sol = solve(InteriorPoint(), mcp, Θ, tol=1e-6)
    J = solve_jacobian_θ(mcp, sol, Θ)
    zbar = np.concatenate([2 * sol.x, 2 * sol.y, 0 * sol.s], axis=0)
    g = solve_pullback(mcp, sol, Θ, 2 * sol.x, 2 * sol.y, None)
    om = OracleMCP(mcp.ir)
    checked = 0
    for b in range(Θ.shape[1]):
        if sol.status[b] != 0:
            continue
        ref = O.Solution("solved", sol.x[:, b], sol.y[:, b], sol.s[:, b], 0.0, float(sol.ϵ[b]), 0)
        Jref = O.solve_jacobian_theta(om, ref, Θ[:, b])
        scale = max(1.0, np.max(np.abs(Jref)))
        assert np.max(np.abs(J[:, :, b] - Jref)) / scale < SENS_TOL, (b, scale)
        gref = Jref.T @ zbar[:, b]
        assert np.max(np.abs(g[:, b] - gref)) / max(1.0, np.max(np.abs(gref))) < SENS_TOL
        checked += 1
    assert checked >= 16


def test_adjoint_pullback_matches_forward(monkeypatch):
    """The pullback runs in adjoint mode (one solve with Cᵀ per instance); the forward mode (one solve of C per column
    of ∇F_θ, the structure of `src/AutoDiff.jl:18-40,59-76`) is kept behind a switch (MCPB200_ADJOINT=0).  Both must give the same
    θ̄ — on the masked game (nθ = 40: 40 factorisations vs 1) and on the lane-change game."""
    from mcp_b200 import solve_pullback

    def pull(make, Θ, x0, tol, mode):
        monkeypatch.setenv("MCPB200_ADJOINT", mode)
        mcp = make()
        sol = solve(InteriorPoint(), mcp, Θ, x0=x0, tol=tol)
        return sol, solve_pullback(mcp, sol, Θ, 2 * sol.x, 2 * sol.y, 0.5 * sol.s)

    Θ = problems.masked_game_thetas(32, 4, seed=3)
    x0 = problems.masked_game_x0(Θ, 4, 30)
    cases = [(lambda: problems.masked_game(4, 30).mcp, Θ, x0, 1e-4),
             (lambda: problems.lane_change_game().mcp, problems.lane_change_thetas(64, seed=5, moving=True), None, 1e-6)]
    for make, Θc, x0c, tol in cases:
        sol, ga = pull(make, Θc, x0c, tol, "1")
        _, gf = pull(make, Θc, x0c, tol, "0")
        ok = sol.status == 0
        assert ok.sum() >= 0.8 * Θc.shape[1]
        rel = np.max(np.abs(ga[:, ok] - gf[:, ok]), axis=0) / np.maximum(1.0, np.max(np.abs(gf[:, ok]), axis=0))
        assert rel.max() < SENS_TOL, rel.max()


def test_pushforward_matches_jacobian_contraction(lane_game):
    """The Dual overload's z_p = ∂z/∂θ · θ_p (`src/AutoDiff.jl:98`) is computed with the tangents as right-hand sides
    (one solve per tangent); it must equal the contraction of the full Jacobian, for several tangents at once."""
    from mcp_b200 import solve_jacobian_θ, solve_pushforward
    mcp = lane_game.mcp
    Θ = problems.lane_change_thetas(16, seed=5, moving=True)
    sol = solve(InteriorPoint(), mcp, Θ, tol=1e-6)
    J = solve_jacobian_θ(mcp, sol, Θ)                                  # n × nθ × B
    rng = np.random.default_rng(0)
    θp = rng.standard_normal((10, 3, Θ.shape[1]))                      # nθ × P × B
    xp, yp, sp = solve_pushforward(mcp, sol, Θ, θp)
    zp = np.concatenate([xp, yp, sp], axis=0)                          # n × P × B
    ok = np.nonzero(sol.status == 0)[0]
    assert len(ok) >= 8
    for b in ok:
        ref = J[:, :, b] @ θp[:, :, b]
        assert np.max(np.abs(zp[:, :, b] - ref)) / max(1.0, np.max(np.abs(ref))) < SENS_TOL


def test_sensitivities_degenerate_backward_error(lane_game):
    """On the benchmark's own θ (zero velocity on the v_y ≥ 0 bound) ∇F_z is numerically singular, so two
    exact solvers need not agree entry-wise; what must hold is a small backward error of
    ∇F_z · (∂z/∂θ) = −∇F_θ  (`src/AutoDiff.jl:39`)."""
    from mcp_b200 import solve_jacobian_θ
    mcp = lane_game.mcp
    Θ = problems.lane_change_thetas(6, seed=5)
    sol = solve(InteriorPoint(), mcp, Θ, tol=1e-6)
    J = solve_jacobian_θ(mcp, sol, Θ)
    om = OracleMCP(mcp.ir)
    for b in range(Θ.shape[1]):
        if sol.status[b] != 0:
            continue
        Jz = om.JFz(sol.x[:, b], sol.y[:, b], sol.s[:, b], Θ[:, b], 0.0).toarray()
        Jt = om.JFt(sol.x[:, b], sol.y[:, b], sol.s[:, b], Θ[:, b], 0.0).toarray()
        resid = Jz @ J[:, :, b] + Jt
        rowwise = np.max(np.max(np.abs(resid), axis=1) / (np.sum(np.abs(Jz), axis=1) * np.max(np.abs(J[:, :, b])) + 1e-300))
        assert rowwise < 1e-12      # (the reference-style dense QR reaches ~1e-15 here, the condensed LU ~1e-28)


# ---- committed golden vectors (tests/golden, generated by the oracle) ------------------------------------
import json
import os

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name,builder", [
    ("lane_change_seed1.npz", lambda: problems.lane_change_game().mcp),
    ("random_qp_12x10_seed1.npz", lambda: problems.random_qp(12, 10)),
])
def test_against_golden_batches(name, builder):
    d = np.load(os.path.join(GOLD, name))
    mcp = builder()
    sol = solve(InteriorPoint(), mcp, d["theta"], tol=float(d["tol"]))
    np.testing.assert_array_equal(sol.status, d["status"])
    assert np.all(np.abs(sol.newton_steps - d["newton_steps"]) <= 1)
    assert np.all(np.abs(sol.outer_iters - d["outer_iters"]) <= 1)
    for got, want in ((sol.x, d["x"]), (sol.y, d["y"]), (sol.s, d["s"])):
        for b in range(want.shape[1]):
            assert rel_err(got[:, b], want[:, b]) <= RTOL


def test_against_golden_qp100():
    """cfg2 at the benchmark's size against the Python oracle's frozen trajectory (θ regenerated from the seed; the
    fixture holds its SHA-256).  The GPU factorises by LDLᵀ, the oracle by pivoted sparse LU."""
    import hashlib
    d = np.load(os.path.join(GOLD, "random_qp_100x100_seed5.npz"))
    Θ = problems.random_qp_thetas(int(d["B"]), seed=int(d["seed"]))
    if hashlib.sha256(np.ascontiguousarray(Θ).tobytes()).hexdigest() != str(d["theta_sha256"]):
        pytest.skip("numpy's generator stream differs from the one the fixture was made with")
    sol = solve(InteriorPoint(), problems.random_qp(100, 100), Θ, tol=float(d["tol"]))
    np.testing.assert_array_equal(sol.status, d["status"])
    assert np.all(np.abs(sol.newton_steps - d["newton_steps"]) <= 1)
    assert np.all(np.abs(sol.outer_iters - d["outer_iters"]) <= 1)
    for got, want in ((sol.x, d["x"]), (sol.y, d["y"]), (sol.s, d["s"])):
        for b in range(want.shape[1]):
            assert rel_err(got[:, b], want[:, b]) <= RTOL


def test_against_golden_small(readme_mcp, clamp_game):
    with open(os.path.join(GOLD, "small.json")) as f:
        g = json.load(f)
    gb = g["readme_qp_batch"]
    Θ = np.array(gb["theta"]).T
    sol = solve(InteriorPoint(), readme_mcp, Θ)
    for b, ref in enumerate(gb["sols"]):
        assert sol.newton_steps[b] == ref["newton_steps"] and sol.outer_iters[b] == ref["outer_iters"]
        np.testing.assert_allclose(sol.x[:, b], ref["x"], rtol=1e-9)
        np.testing.assert_allclose(sol.y[:, b], ref["y"], rtol=1e-9)
        assert sol.ϵ[b] == pytest.approx(ref["eps"], rel=1e-12)
    r6 = g["readme_qp_tol1e-6"]
    s6 = solve(InteriorPoint(), readme_mcp, np.array(r6["theta"]), tol=1e-6)
    assert s6.outer_iters == r6["outer_iters"] and s6.newton_steps == r6["newton_steps"]
    np.testing.assert_allclose(s6.x, r6["x"], rtol=1e-9)
    gc = g["clamp_game"]
    sc = solve(InteriorPoint(), clamp_game.mcp, np.array(gc["theta"]), tol=gc["tol"])
    assert sc.status == gc["status"] and sc.newton_steps == gc["newton_steps"]
    np.testing.assert_allclose(sc.x, gc["x"], rtol=1e-7, atol=1e-12)
    # golden sensitivities of the README QP
    from mcp_b200 import solve_jacobian_θ
    gd = g["readme_qp_default"]
    sd = solve(InteriorPoint(), readme_mcp, np.array(gd["theta"]))
    np.testing.assert_allclose(solve_jacobian_θ(readme_mcp, sd, np.array(gd["theta"])), gd["dzdtheta"],
                               rtol=SENS_TOL, atol=1e-9)


# ---- cfg2: random convex QP (benchmark/quadratic_program_benchmark.jl) -------------------------------------
def test_random_qp_100x100():
    """The benchmark's own size: 100 primals, 100 inequalities, θ = [vec(M); vec(A); b; ϕ] (20 200 entries),
    dense condensed system; cold start and the warm-started θ sweep of SURVEY.md §8d."""
    mcp = problems.random_qp(100, 100)
    Θ = problems.random_qp_thetas(12, seed=1)
    sol = solve(InteriorPoint(), mcp, Θ, tol=1e-6)
    from oracle import c_oracle as CO
    ref = CO.solve_batch(mcp.ir, Θ, tol=1e-6)
    np.testing.assert_array_equal(sol.status, ref.status)
    ok = sol.status == 0
    assert ok.sum() >= 6
    assert np.all(np.abs(sol.newton_steps[ok] - ref.newton_steps[ok]) <= 1)
    for b in np.nonzero(ok)[0]:
        assert rel_err(sol.x[:, b], ref.x[:, b]) <= RTOL and rel_err(sol.y[:, b], ref.y[:, b]) <= RTOL
        assert rel_err(sol.s[:, b], ref.s[:, b]) <= RTOL
    # θ sweep: perturb ϕ, warm start from the previous solution (y₀ clamped away from zero)
    rng = np.random.default_rng(5)
    Θ2 = Θ.copy()
    Θ2[-100:] += 0.01 * rng.standard_normal((100, Θ.shape[1]))
    x0, y0 = sol.x.copy(), np.maximum(sol.y, 1e-3)
    warm = solve(InteriorPoint(), mcp, Θ2, x0=x0, y0=y0, tol=1e-6)
    refw = CO.solve_batch(mcp.ir, Θ2, x0=x0, y0=y0, tol=1e-6)
    np.testing.assert_array_equal(warm.status, refw.status)
    for b in np.nonzero(warm.status == 0)[0]:
        assert abs(int(warm.newton_steps[b]) - int(refw.newton_steps[b])) <= 1
        assert rel_err(warm.x[:, b], refw.x[:, b]) <= RTOL


@pytest.mark.parametrize("n,m,threads", [(37, 53, 512), (64, 40, 512), (96, 100, 512), (104, 96, 512), (111, 128, 256)])
def test_dense_kernel_shapes(n, m, threads):
    """The register-tiled dense kernel pads rows to 32-lane chunks and columns to 16-warp groups; these shapes hit
    the partial chunk, the exact-multiple and the largest cases (the right-hand side is column n); the last one does
    not fit the register-tiled kernel's shared memory and must fall back to the shared-memory dense kernel."""
    mcp = problems.random_qp(n, m)
    from mcp_b200.solver import _handle
    info = _handle(mcp).info()
    assert info["threads_per_instance"] == threads, info
    Θ = problems.random_qp_thetas(6, seed=n + m, num_primals=n, num_inequalities=m, sparsity_rate=0.8)
    sol = solve(InteriorPoint(), mcp, Θ, tol=1e-6)
    from oracle import c_oracle as CO
    ref = CO.solve_batch(mcp.ir, Θ, tol=1e-6)
    np.testing.assert_array_equal(sol.status, ref.status)
    ok = sol.status == 0
    assert ok.sum() >= 2
    assert np.all(np.abs(sol.newton_steps[ok] - ref.newton_steps[ok]) <= 1)
    for b in np.nonzero(ok)[0]:
        assert rel_err(sol.x[:, b], ref.x[:, b]) <= RTOL and rel_err(sol.y[:, b], ref.y[:, b]) <= RTOL
        assert rel_err(sol.s[:, b], ref.s[:, b]) <= RTOL


def _qp_vs_c_oracle(mcp, Θ):
    """Every instance: same status as the C oracle; solved ones: Newton steps within ±1, x, y, s to the bar."""
    from oracle import c_oracle as CO
    sol = solve(InteriorPoint(), mcp, Θ, tol=1e-6)
    ref = CO.solve_batch(mcp.ir, Θ, tol=1e-6)
    np.testing.assert_array_equal(sol.status, ref.status)
    ok = np.nonzero(sol.status == 0)[0]
    assert np.all(np.abs(sol.newton_steps[ok] - ref.newton_steps[ok]) <= 1)
    for b in ok:
        assert rel_err(sol.x[:, b], ref.x[:, b]) <= RTOL and rel_err(sol.y[:, b], ref.y[:, b]) <= RTOL
        assert rel_err(sol.s[:, b], ref.s[:, b]) <= RTOL
    return sol, len(ok)


def test_dense_symmetric_path_and_fallbacks():
    """r2: when G_x (from θ) is symmetric the dense kernel factorises the condensed matrix by LDLᵀ without pivoting
    (kernel_template.cuh, d3p_*), with the Schur product summed over the non-zeros of H_x only when H_x is sparse in
    value.  The paths an instance can take — symmetric + sparse A (the benchmark's own generator), symmetric + dense A,
    G_x not symmetric (symmetry check ⇒ pivoted LU), G_x symmetric but indefinite (non-positive pivot ⇒ pivoted LU
    from that Newton step on) — all against the C oracle, which always runs the reference's pivoted sparse LU."""
    mcp = problems.random_qp(100, 100)
    Θ = problems.random_qp_thetas(48, seed=11)
    _, n = _qp_vs_c_oracle(mcp, Θ)
    assert n >= 44
    _, n = _qp_vs_c_oracle(mcp, problems.random_qp_thetas(12, seed=3, sparsity_rate=0.3))   # dense A, dense M
    assert n >= 10
    Θa = Θ[:, :12].copy()
    Θa[3 + 100 * 7] += 0.25      # vec(M) is column-major: M[r, c] = θ[r + 100 c]
    Θa[50 + 100 * 2] -= 0.125
    _, n = _qp_vs_c_oracle(mcp, Θa)
    assert n >= 10
    Θi = Θ[:, :12].copy()
    for i in (5, 40):
        Θi[i + 100 * i] -= 30.0  # two negative directions: LDLᵀ meets a negative pivot in the first Newton step
    _, n = _qp_vs_c_oracle(mcp, Θi)
    assert n >= 6


def test_dense_problem_beyond_one_cta():
    """A dense condensed system whose factorisation window (200 × 202 doubles = 323 KB) does not fit the shared memory
    of one SM: r1 refused it (`MCPB200_ERR_UNSUPPORTED`); the window now moves to the instance's global block next to
    its vectors (slow, but solved) — same parity bar against the C oracle."""
    from mcp_b200.solver import _handle
    n, m = 200, 120
    mcp = problems.random_qp(n, m)
    info = _handle(mcp).info()
    assert info["n_reduced"] == n and info["window_rows"] == n
    Θ = problems.random_qp_thetas(6, seed=31, num_primals=n, num_inequalities=m)
    _, solved = _qp_vs_c_oracle(mcp, Θ)
    assert solved >= 4


def test_random_qp_batch_properties():
    """cfg2 at a batch the oracle cannot follow (every CTA solves several instances back to back, both passes run):
    size-independent checks on ALL instances with the QP's own data — stationarity M x − ϕ − Aᵀy, primal
    feasibility A x − b = s ≥ 0, complementarity at the ϵ scale (the reference's `check_solution`-style assertions,
    test/runtests.jl:30-38) — and a random sample against the C oracle."""
    from oracle import c_oracle as CO
    n = m = 100
    mcp = problems.random_qp(n, m)
    B = 1200
    Θ = problems.random_qp_thetas(B, seed=21)
    sol = solve(InteriorPoint(), mcp, Θ, tol=1e-6)
    assert np.all((sol.status == 0) | (sol.status == 1))
    ok = np.nonzero(sol.status == 0)[0]
    assert len(ok) >= 0.98 * B
    M = Θ[:n * n].reshape(n, n, B, order="F")
    A = Θ[n * n:n * n + m * n].reshape(m, n, B, order="F")
    b, ϕ = Θ[n * n + m * n:n * n + m * n + m], Θ[-n:]
    G = np.einsum("ijb,jb->ib", M, sol.x) - ϕ - np.einsum("kib,kb->ib", A, sol.y)
    H = np.einsum("kjb,jb->kb", A, sol.x) - b
    scale = 1.0 + np.abs(sol.y[:, ok]).max(axis=0)
    assert np.max(np.abs(G[:, ok]).max(axis=0) / scale) <= 1e-5
    assert np.max(np.abs(H[:, ok] - sol.s[:, ok])) <= 1e-5
    assert np.all(sol.y[:, ok] > 0) and np.all(sol.s[:, ok] > 0)
    assert np.max(sol.s[:, ok] * sol.y[:, ok]) < 1e-2
    assert np.all((sol.kkt_error[ok] <= 1e-6) | (sol.ϵ[ok] <= 1e-6))
    idx = np.random.default_rng(1).choice(B, 24, replace=False)
    ref = CO.solve_batch(mcp.ir, Θ[:, idx], tol=1e-6)
    np.testing.assert_array_equal(sol.status[idx], ref.status)
    for k, bb in enumerate(idx):
        if ref.status[k] == 0:
            assert abs(int(ref.newton_steps[k]) - int(sol.newton_steps[bb])) <= 1
            assert rel_err(sol.x[:, bb], ref.x[:, k]) <= RTOL and rel_err(sol.y[:, bb], ref.y[:, k]) <= RTOL


# ---- edge cases ---------------------------------------------------------------------------------------------------
def test_empty_and_single_batches(readme_mcp):
    empty = solve(InteriorPoint(), readme_mcp, np.zeros((2, 0)))
    assert empty.x.shape == (2, 0) and empty.status.shape == (0,)
    one = solve(InteriorPoint(), readme_mcp, np.array([[0.3], [0.7]]))
    ref = O.solve_interior_point(OracleMCP(readme_mcp.ir), [0.3, 0.7])
    np.testing.assert_allclose(one.x[:, 0], ref.x, rtol=1e-9)


def test_ragged_batch_sizes(readme_mcp):
    """Batch sizes that are not multiples of the warp / CTA / grid size must all be covered exactly once."""
    om = OracleMCP(readme_mcp.ir)
    for B in (1, 31, 33, 2369, 2371):
        Θ = problems.readme_qp_thetas(B, seed=B)
        sol = solve(InteriorPoint(), readme_mcp, Θ)
        assert np.all(sol.status == 0) and np.all(sol.newton_steps > 0)
        for b in (0, B // 2, B - 1):
            np.testing.assert_allclose(sol.x[:, b], O.solve_interior_point(om, Θ[:, b]).x, rtol=1e-9)


def test_initial_point_kwargs_and_in_place(readme_mcp):
    """x₀, y₀, s₀ all given (src/solver.jl:39-41); outputs may alias them like the reference's x = x₀ (:64-66)."""
    om = OracleMCP(readme_mcp.ir)
    rng = np.random.default_rng(0)
    B = 40
    Θ = problems.readme_qp_thetas(B, seed=9)
    x0, y0, s0 = rng.normal(size=(2, B)), rng.random((2, B)) + 0.1, rng.random((2, B)) + 0.1
    sol = solve(InteriorPoint(), readme_mcp, Θ, x0=x0, y0=y0, s0=s0)
    for b in range(B):
        ref = O.solve_interior_point(om, Θ[:, b], x0=x0[:, b], y0=y0[:, b], s0=s0[:, b])
        assert (ref.status == "solved") == (sol.status[b] == 0)
        if ref.status == "solved":
            assert abs(ref.newton_steps - int(sol.newton_steps[b])) <= 1
            assert rel_err(sol.x[:, b], ref.x) <= RTOL and rel_err(sol.y[:, b], ref.y) <= RTOL


def test_option_passthrough(readme_mcp):
    """tol, iteration caps, rates and min_stepsize reach the kernel (src/solver.jl:42-49)."""
    om = OracleMCP(readme_mcp.ir)
    θ = np.array([-0.5, 0.5])
    for kw in (dict(max_outer_iters=3), dict(max_inner_iters=3), dict(tightening_rate=0.3, loosening_rate=0.2),
               dict(tol=1e-8), dict(min_stepsize=0.6)):
        ref = O.solve_interior_point(om, θ, **kw)
        sol = solve(InteriorPoint(), readme_mcp, θ, **kw)
        assert sol.status == ref.status and sol.outer_iters == ref.outer_iters and sol.newton_steps == ref.newton_steps, kw
        np.testing.assert_allclose(sol.x, ref.x, rtol=1e-8)
        assert sol.ϵ == pytest.approx(ref.eps, rel=1e-10)


def test_infeasible_and_nan_instances(lane_game):
    """A colliding start is infeasible ⇒ :failed at the outer cap; a NaN θ must not hang or poison neighbours."""
    mcp = lane_game.mcp
    Θ = problems.lane_change_thetas(4, seed=1)
    Θ[:, 1] = [1.0, 10.0, 0, 0, 1.0, 1.2, 10.1, 0, 0, 3.0]
    Θ[0, 2] = np.nan
    sol = solve(InteriorPoint(), mcp, Θ, tol=1e-6)
    assert sol.status[1] == 1 and sol.outer_iters[1] == 50
    om = OracleMCP(mcp.ir)
    for b in (0, 3):                       # the neighbours of the NaN instance are unaffected
        ref = O.solve_interior_point(om, Θ[:, b], tol=1e-6)
        assert (ref.status == "solved") == (sol.status[b] == 0)
        if ref.status == "solved":
            assert rel_err(sol.x[:, b], ref.x) <= RTOL


def test_large_batch_properties(lane_game):
    """At bench scale the oracle is too slow to run in full: check size-independent properties — every
    instance written exactly once, converged instances satisfy the KKT residual bound they claim, and a
    sample agrees with the C oracle."""
    from oracle import c_oracle as CO
    mcp = lane_game.mcp
    B = 20000
    Θ = problems.lane_change_thetas(B, seed=123)
    sol = solve(InteriorPoint(), mcp, Θ, tol=1e-6)
    assert np.all((sol.status == 0) | (sol.status == 1)) and np.all(sol.outer_iters >= 2)
    ok = sol.status == 0
    assert 0.9 < ok.mean() < 0.99
    assert np.all(np.isfinite(sol.x[:, ok])) and np.all(sol.y[:, ok] > 0) and np.all(sol.s[:, ok] > 0)
    assert np.all((sol.kkt_error[ok] <= 1e-6) | (sol.ϵ[ok] <= 1e-6))     # the loop's own exit test (:71)
    assert np.max(sol.s[:, ok] * sol.y[:, ok]) < 1e-2                     # complementarity at the ϵ scale
    idx = np.random.default_rng(0).choice(B, 64, replace=False)
    ref = CO.solve_batch(mcp.ir, Θ[:, idx], tol=1e-6)
    agree = 0
    for k, b in enumerate(idx):
        if ref.status[k] == 0 and sol.status[b] == 0 and abs(int(ref.newton_steps[k]) - int(sol.newton_steps[b])) <= 1 \
                and rel_err(sol.x[:, b], ref.x[:, k]) <= RTOL:
            agree += 1
        elif ref.status[k] == 1 and sol.status[b] == 1:
            agree += 1
    assert agree == 64


# ---- cfg4: masked N-player game (examples/train_and_test_utils.jl:362-401), N = 4, horizon 30 ------------------
def test_masked_game_n4():
    """nx = 1200, ny = 1470, nθ = 40: all 2^(N-1) ego masks of two scenarios, stay-at-rest initial guess, tol 1e-4
    (the application's settings), against the C oracle; plus the sensitivity backward error."""
    from oracle import c_oracle as CO
    game = problems.masked_game(4, 30)
    mcp = game.mcp
    assert (mcp.unconstrained_dimension, mcp.constrained_dimension, mcp.parameter_dimension) == (1200, 1470, 40)
    Θ = problems.masked_game_thetas(16, 4, seed=1)
    x0 = problems.masked_game_x0(Θ, 4, 30)
    sol = solve(InteriorPoint(), mcp, Θ, x0=x0, tol=1e-4)
    ref = CO.solve_batch(mcp.ir, Θ, x0=x0, tol=1e-4)
    np.testing.assert_array_equal(sol.status, ref.status)
    assert (sol.status == 0).sum() >= 12
    for b in np.nonzero(ref.status == 0)[0]:
        assert abs(int(sol.newton_steps[b]) - int(ref.newton_steps[b])) <= 1
        assert rel_err(sol.x[:, b], ref.x[:, b]) <= RTOL and rel_err(sol.y[:, b], ref.y[:, b]) <= RTOL
        assert rel_err(sol.s[:, b], ref.s[:, b]) <= RTOL
    # masking a player out changes the ego's plan: the sweep must not return identical trajectories
    assert np.max(np.abs(sol.x[:, 0] - sol.x[:, 7])) > 1e-3
    from mcp_b200 import solve_pullback
    g = solve_pullback(mcp, sol, Θ, 2 * sol.x, None, None)
    assert g.shape == (40, 16) and np.all(np.isfinite(g[:, sol.status == 0]))


def test_masked_game_n10():
    """The application's largest configuration (SURVEY.md §8 cfg4, N = 10): nx = 3000, ny = 3630, nθ = 160, 200 k tape
    nodes.  The evaluation is compiled as separately linked units, the state lives in the global block (large-state
    mode) and the window (57 × 113) in shared memory.  Four scenarios against the C oracle."""
    from oracle import c_oracle as CO
    game = problems.masked_game(10, 30)
    mcp = game.mcp
    assert (mcp.unconstrained_dimension, mcp.constrained_dimension, mcp.parameter_dimension) == (3000, 3630, 160)
    Θ = problems.masked_game_thetas(4, 10, seed=1)
    x0 = problems.masked_game_x0(Θ, 10, 30)
    sol = solve(InteriorPoint(), mcp, Θ, x0=x0, tol=1e-4)
    ref = CO.solve_batch(mcp.ir, Θ, x0=x0, tol=1e-4)
    np.testing.assert_array_equal(sol.status, ref.status)
    assert (sol.status == 0).sum() >= 3
    for b in np.nonzero(ref.status == 0)[0]:
        assert abs(int(sol.newton_steps[b]) - int(ref.newton_steps[b])) <= 1
        assert rel_err(sol.x[:, b], ref.x[:, b]) <= RTOL and rel_err(sol.y[:, b], ref.y[:, b]) <= RTOL
        assert rel_err(sol.s[:, b], ref.s[:, b]) <= RTOL
    from mcp_b200 import solve_pullback
    g = solve_pullback(mcp, sol, Θ, 2 * sol.x, None, None)
    assert g.shape == (160, 4) and np.all(np.isfinite(g[:, sol.status == 0]))


def test_masked_game_jacobian_consistency():
    """cfg4 sensitivities in large-state mode (16 right-hand sides per factorisation pass): the full Jacobian must
    satisfy ∇F_z · (∂z/∂θ) = −∇F_θ (`src/AutoDiff.jl:39`) to backward-error level, and contracting it with z̄ must
    reproduce the adjoint pullback."""
    from mcp_b200 import solve_jacobian_θ, solve_pullback
    game = problems.masked_game(4, 30)
    mcp = game.mcp
    Θ = problems.masked_game_thetas(4, 4, seed=7)
    x0 = problems.masked_game_x0(Θ, 4, 30)
    sol = solve(InteriorPoint(), mcp, Θ, x0=x0, tol=1e-4)
    assert (sol.status == 0).sum() >= 3
    J = solve_jacobian_θ(mcp, sol, Θ)                                   # n × nθ × B
    g = solve_pullback(mcp, sol, Θ, 2 * sol.x, 2 * sol.y, 0.5 * sol.s)
    zbar = np.concatenate([2 * sol.x, 2 * sol.y, 0.5 * sol.s], axis=0)
    om = OracleMCP(mcp.ir)
    for b in np.nonzero(sol.status == 0)[0]:
        gref = J[:, :, b].T @ zbar[:, b]
        assert np.max(np.abs(g[:, b] - gref)) / max(1.0, np.max(np.abs(gref))) < SENS_TOL
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    for b in np.nonzero(sol.status == 0)[0]:      # every solved instance (r1 checked one)
        Jz = om.JFz(sol.x[:, b], sol.y[:, b], sol.s[:, b], Θ[:, b], float(sol.ϵ[b]))
        Jt = om.JFt(sol.x[:, b], sol.y[:, b], sol.s[:, b], Θ[:, b], float(sol.ϵ[b]))
        Jt = Jt.toarray() if hasattr(Jt, "toarray") else np.asarray(Jt)
        R = Jz @ J[:, :, b] + Jt
        scale = abs(Jz).max() * max(1.0, np.abs(J[:, :, b]).max())
        assert np.max(np.abs(R)) / scale < 1e-9, (b, np.max(np.abs(R)) / scale)
        # … and entry-wise against the oracle's own solve of ∇F_z X = −∇F_θ on the full 4140-dimensional system (sparse LU
        # in the role of the reference's QR, `src/AutoDiff.jl:27-39`; the matrix is well conditioned at these points)
        Jref = spla.splu(sp.csc_matrix(Jz)).solve(-Jt)
        assert np.max(np.abs(J[:, :, b] - Jref)) / max(1.0, np.max(np.abs(Jref))) < SENS_TOL, b


def test_lane_change_parity_statistics(lane_game):
    """1 024 random lane-change instances against the C oracle: same status, Newton steps within ±1, x/y/s within
    1e-6 relative — on EVERY instance (measured on B200: 1 024 / 1 024, identical step counts on all 986 solved ones,
    GPU within 9.3e-9 and the C oracle within 3.4e-13 of the extended-precision trajectory,
    profiles/r2_adjudicate_lane_change.json).  An instance may miss the bar only if the FP64 oracle itself is off the
    extended-precision trajectory (decided at run time) AND it is named in tests/golden/ill_conditioned.json (none is)."""
    from oracle import c_oracle as CO
    mcp = lane_game.mcp
    B = 1024
    Θ = problems.lane_change_thetas(B, seed=2024)
    sol = solve(InteriorPoint(), mcp, Θ, tol=1e-6)
    ref = CO.solve_batch(mcp.ir, Θ, tol=1e-6)
    same_status = sol.status == ref.status
    solved = ref.status == 0
    ok = same_status.copy()
    worst = 0.0
    for b in np.nonzero(solved & same_status)[0]:
        e = max(rel_err(sol.x[:, b], ref.x[:, b]), rel_err(sol.y[:, b], ref.y[:, b]), rel_err(sol.s[:, b], ref.s[:, b]))
        steps_ok = abs(int(sol.newton_steps[b]) - int(ref.newton_steps[b])) <= 1
        ok[b] = steps_ok and e <= RTOL
        if steps_ok:
            worst = max(worst, e)
    print(f"parity: {ok.sum()}/{B} to the bar, status agreement {same_status.sum()}/{B}, "
          f"identical step counts {int((sol.newton_steps[solved] == ref.newton_steps[solved]).sum())}/{int(solved.sum())}, "
          f"worst rel err among step-matched {worst:.2e}")
    allowed = golden_ill_conditioned("lane_change_seed2024_B1024_tol1e-6")
    for b in np.nonzero(~ok)[0]:
        assert int(b) in allowed, f"instance {b} misses the bar and is not a named rounding-sensitive instance"
        assert rounding_sensitive(mcp, Θ[:, b], ref.x[:, b], 1e-6), f"instance {b}: the FP64 oracle follows the exact trajectory, the GPU does not"


def test_masked_game_parity_statistics():
    """256 masked-game instances (N = 4: all 8 ego masks of 32 scenarios, the data-generation sweep of SURVEY.md §8d)
    against the C oracle — the path with shape-grouped evaluation, large-state mode and cooperative instances."""
    from oracle import c_oracle as CO
    mcp = problems.masked_game(4, 30).mcp
    B = 256
    Θ = problems.masked_game_thetas(B, 4, seed=11)
    x0 = problems.masked_game_x0(Θ, 4, 30)
    sol = solve(InteriorPoint(), mcp, Θ, x0=x0, tol=1e-4)
    ref = CO.solve_batch(mcp.ir, Θ, x0=x0, tol=1e-4)
    same_status = sol.status == ref.status
    solved = ref.status == 0
    ok = same_status.copy()
    worst = 0.0
    for b in np.nonzero(solved & same_status)[0]:
        e = max(rel_err(sol.x[:, b], ref.x[:, b]), rel_err(sol.y[:, b], ref.y[:, b]), rel_err(sol.s[:, b], ref.s[:, b]))
        steps_ok = abs(int(sol.newton_steps[b]) - int(ref.newton_steps[b])) <= 1
        ok[b] = steps_ok and e <= RTOL
        if steps_ok:
            worst = max(worst, e)
    print(f"masked-game parity: {ok.sum()}/{B} to the bar, status agreement {same_status.sum()}/{B}, worst rel err {worst:.2e}")
    assert solved.sum() >= 0.9 * B
    assert same_status.all()
    # Measured (profiles/r2_adjudicate_masked_n4.json): 253 / 256 to the bar; on the other three — instances 130, 184, 185 —
    # the FP64 C oracle itself is 6.4e-2 / 4.3e-3 / 1.9e-6 away from the extended-precision trajectory (different
    # Newton-step counts: 61 vs 63, 70 vs 103 on the first two): their trajectories hinge on rounding.  Exactly those
    # named instances may miss, and the reason is re-verified here.
    allowed = golden_ill_conditioned("masked_n4_seed11_B256_tol1e-4")
    for b in np.nonzero(~ok)[0]:
        assert int(b) in allowed, f"instance {b} misses the bar and is not a named rounding-sensitive instance"
        assert rounding_sensitive(mcp, Θ[:, b], ref.x[:, b], 1e-4, x0=x0[:, b]), f"instance {b}: not rounding-sensitive"


def test_in_library_multi_device_sharding(lane_game):
    """`mcpb200_set_devices`: one host call, θ columns split in contiguous blocks over the GPUs of the box (one
    host thread per device, no collective).  Results must be identical to the single-device run."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs ≥ 2 GPUs (run under `gpurun --gpus 2`)")
    from mcp_b200.solver import _handle
    mcp = lane_game.mcp
    Θ = problems.lane_change_thetas(3001, seed=77)
    h = _handle(mcp)
    h.set_devices([0])
    one = solve(InteriorPoint(), mcp, Θ, tol=1e-6)
    ndev = torch.cuda.device_count()
    h.set_devices(list(range(ndev)))           # every GPU of the box (2 under `gpurun --gpus 2`, 8 under `--gpus 8`)
    two = solve(InteriorPoint(), mcp, Θ, tol=1e-6)
    t = h.timing()
    h.set_devices([0])
    np.testing.assert_array_equal(one.status, two.status)
    np.testing.assert_array_equal(one.newton_steps, two.newton_steps)
    np.testing.assert_array_equal(one.x, two.x)          # same kernel, same inputs ⇒ bit-identical
    np.testing.assert_array_equal(one.y, two.y)
    assert t["launches"] == 2 * ndev and t["solved"] == int((one.status == 0).sum())   # 2 passes on each device


def test_sensitivities_singular_instance_is_nan(readme_mcp):
    """A singular KKT matrix at the supplied point (y = s = 0 ⇒ D = 0/0) must come back as NaN plus a warning, never as
    whatever the caller's output buffer held (the reference's QR, `src/AutoDiff.jl:39`, hands back non-finite values)."""
    from mcp_b200 import solve_jacobian_θ, solve_pullback, solve_pushforward
    from mcp_b200.solver import Solution
    Θ = problems.readme_qp_thetas(4, seed=3)
    sol = solve(InteriorPoint(), readme_mcp, Θ)
    bad = Solution(sol.status.copy(), sol.x.copy(), sol.y.copy(), sol.s.copy(), sol.kkt_error, sol.ϵ, sol.outer_iters,
                   sol.newton_steps)
    bad.y[:, 1] = 0.0
    bad.s[:, 1] = 0.0
    with pytest.warns(RuntimeWarning, match="singular"):
        g = solve_pullback(readme_mcp, bad, Θ, 2 * bad.x, 2 * bad.y, None)
    assert np.all(np.isnan(g[:, 1])) and np.all(np.isfinite(g[:, [0, 2, 3]]))
    with pytest.warns(RuntimeWarning):
        J = solve_jacobian_θ(readme_mcp, bad, Θ)
    assert np.all(np.isnan(J[:, :, 1])) and np.all(np.isfinite(J[:, :, 0]))
    with pytest.warns(RuntimeWarning):
        xp, yp, sp = solve_pushforward(readme_mcp, bad, Θ, np.tile(np.eye(2)[:, :, None], (1, 1, 4)))
    assert np.all(np.isnan(xp[:, :, 1])) and np.all(np.isfinite(xp[:, :, 2]))


def test_gpu_solutions_satisfy_independent_kkt(lane_game):
    """GPU lane-change solutions checked against the problem definition written independently of the tracer
    (`oracle/independent_problems.py`): the reference's own `check_solution` assertions (`test/runtests.jl:30-38`)
    at the scale of tol = 1e-6, with G and H evaluated by torch autograd straight from the reference's formulas."""
    from oracle.independent_problems import IndependentTrajectoryGame
    ind = IndependentTrajectoryGame("lane_change", 2, 10)
    mcp = lane_game.mcp
    Θ = problems.lane_change_thetas(32, seed=9)
    sol = solve(InteriorPoint(), mcp, Θ, tol=1e-6)
    checked = 0
    for b in np.nonzero(sol.status == 0)[0]:
        F = ind.F(sol.x[:, b], sol.y[:, b], sol.s[:, b], Θ[:, b], 0.0)
        G, Hms, sy = F[:200], F[200:450], F[450:]
        H = Hms + sol.s[:, b]
        assert np.max(np.abs(G)) <= 1e-5 and np.min(H) >= -1e-5 and np.min(sol.y[:, b]) >= 0
        assert np.max(np.abs(Hms)) <= 1e-5 and np.max(sy) <= 1e-4 and float(sol.y[:, b] @ H) <= 250 * 1e-4
        checked += 1
    assert checked >= 28


def _lcp_like_mcp(compute_sensitivities=False):
    """A small MCP whose H depends on y (an LCP-style coupling B y): G = M x − θ − Aᵀ y, H = A x − b + B y."""
    from mcp_b200.mcp import PrimalDualMCP
    M = np.array([[2.0, 1, 0], [1, 2, 0.5], [0, 0.5, 3]])
    A = np.array([[1.0, 0, 1], [0, 1, 0], [1, 1, 0]])
    Bm = np.array([[0.5, 0.1, 0], [0.1, 0.4, 0], [0, 0, 0.3]])
    b = np.array([1.0, 1, 0.5])
    return PrimalDualMCP(lambda x, y, θ: M @ x - θ - A.T @ y, lambda x, y, θ: A @ x - b + Bm @ y,
                         unconstrained_dimension=3, constrained_dimension=3, parameter_dimension=3,
                         compute_sensitivities=compute_sensitivities)


def test_h_depending_on_y_matches_oracle():
    """∇_y H ≠ 0 (`src/mcp.jl:76-80` allows any H(x, y; θ)): the (nx+ny)-dimensional mode against the Python oracle, which
    solves the full n×n system as the reference does — same status, identical step counts, x/y/s within 1e-6."""
    mcp = _lcp_like_mcp()
    Θ = np.asfortranarray(np.random.default_rng(4).uniform(-1.0, 2.0, (3, 256)))
    for tol in (1e-4, 1e-6):
        sol = solve(InteriorPoint(), mcp, Θ, tol=tol)
        compare_batch(mcp, Θ, sol, tol=tol, min_match=1.0)


def test_h_depending_on_y_sensitivities():
    """Sensitivities in the (nx+ny)-dimensional mode (r2): ∂z/∂θ from J_B [Z_x; Z_y] = −[∇_θG; ∇_θH], Z_s = −(s/y) Z_y,
    against the oracle's QR on the full n×n system (`src/AutoDiff.jl:18-40`); the pullback and the pushforward must
    be its contractions (`:42-117`), and finite differences of the solve agree."""
    from mcp_b200 import solve_jacobian_θ, solve_pullback, solve_pushforward
    mcp = _lcp_like_mcp(compute_sensitivities=True)
    Θ = np.asfortranarray(np.random.default_rng(9).uniform(-1.0, 2.0, (3, 32)))
    sol = solve(InteriorPoint(), mcp, Θ, tol=1e-6)
    assert np.all(sol.status == 0)
    J = solve_jacobian_θ(mcp, sol, Θ)
    g = solve_pullback(mcp, sol, Θ, 2 * sol.x, 2 * sol.y, 3 * sol.s)
    tp = np.random.default_rng(10).standard_normal((3, 2, 32))
    xp, yp, sp = solve_pushforward(mcp, sol, Θ, tp)
    om = OracleMCP(mcp.ir)
    for b in range(Θ.shape[1]):
        ref = O.Solution("solved", sol.x[:, b], sol.y[:, b], sol.s[:, b], 0.0, float(sol.ϵ[b]), 0)
        Jref = O.solve_jacobian_theta(om, ref, Θ[:, b])
        scale = max(1.0, np.max(np.abs(Jref)))
        assert np.max(np.abs(J[:, :, b] - Jref)) / scale < SENS_TOL, b
        zbar = np.concatenate([2 * sol.x[:, b], 2 * sol.y[:, b], 3 * sol.s[:, b]])
        np.testing.assert_allclose(g[:, b], Jref.T @ zbar, rtol=SENS_TOL, atol=1e-7)
        np.testing.assert_allclose(np.vstack([xp[:, :, b], yp[:, :, b], sp[:, :, b]]), Jref @ tp[:, :, b], rtol=SENS_TOL, atol=1e-7)
    # finite differences of the solve itself (tight tolerance so that the ϵ-path contribution is below the bar)
    h = 1e-5
    b = 0
    tight = dict(tol=1e-9)
    base = solve(InteriorPoint(), mcp, Θ[:, b], **tight)
    Jb = solve_jacobian_θ(mcp, base, Θ[:, b])
    for q in range(3):
        θp, θm = Θ[:, b].copy(), Θ[:, b].copy()
        θp[q] += h
        θm[q] -= h
        sp_, sm_ = solve(InteriorPoint(), mcp, θp, **tight), solve(InteriorPoint(), mcp, θm, **tight)
        fd = (np.concatenate([sp_.x, sp_.y, sp_.s]) - np.concatenate([sm_.x, sm_.y, sm_.s])) / (2 * h)
        np.testing.assert_allclose(Jb[:, q], fd, rtol=2e-3, atol=2e-4)


def test_full_y_mode_on_lane_change_matches_condensed(lane_game, monkeypatch):
    """The same mode forced (MCPB200_FULL_Y=1) on the lane-change game, a 450-dimensional banded system: it must follow the
    condensed kernel's trajectory (identical Newton-step counts, x/y/s within 1e-6)."""
    from mcp_b200 import capi
    from mcp_b200.solver import _handle
    mcp = lane_game.mcp
    Θ = problems.lane_change_thetas(64, seed=12)
    ref = solve(InteriorPoint(), mcp, Θ, tol=1e-6)
    monkeypatch.setenv("MCPB200_FULL_Y", "1")
    game2 = problems.lane_change_game(compute_sensitivities=False)
    h = _handle(game2.mcp)
    assert h.info()["n_reduced"] == 450
    got = solve(InteriorPoint(), game2.mcp, Θ, tol=1e-6)
    np.testing.assert_array_equal(got.status, ref.status)
    ok = ref.status == 0
    assert np.all(np.abs(got.newton_steps[ok] - ref.newton_steps[ok]) <= 1)
    for b in np.nonzero(ok)[0]:
        assert rel_err(got.x[:, b], ref.x[:, b]) <= RTOL and rel_err(got.y[:, b], ref.y[:, b]) <= RTOL
        assert rel_err(got.s[:, b], ref.s[:, b]) <= RTOL
