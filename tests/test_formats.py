"""On-disk formats of the reference's scripts: CPU-only."""
import json

import numpy as np

from mcp_b200 import formats, problems


def test_scenario_csv_round_trip(tmp_path):
    states, goals = formats.generate_agents_and_goals(4, rng=3)
    assert states.shape == (4, 4) and goals.shape == (4, 2) and np.all(states[:, 2:] == 0.0)
    d = np.linalg.norm(states[:, None, :2] - states[None, :, :2], axis=-1) + 10 * np.eye(4)
    assert d.min() >= 1.0 and np.abs(states[:, :2]).max() <= 2.5
    p = tmp_path / "agents_and_goals.csv"
    formats.write_scenario_csv(str(p), states, goals)
    assert p.read_text().splitlines()[0] == "id,x,y,vx,vy,goal_x,goal_y"      # scripts/data_generation.py:44
    s2, g2 = formats.read_scenario_csv(str(p))
    np.testing.assert_array_equal(s2, states.reshape(-1))
    np.testing.assert_array_equal(g2, goals.reshape(-1))


def test_theta_matches_the_benchmark_generator():
    """θ built from a scenario is laid out exactly as `problems.masked_game_thetas` does it."""
    Θ = problems.masked_game_thetas(8, 4, seed=2)
    for b in (0, 5):
        θ = Θ[:, b]
        states = np.array([θ[i * 10: i * 10 + 4] for i in range(4)])
        goals = np.array([θ[i * 10 + 4: i * 10 + 6] for i in range(4)])
        np.testing.assert_array_equal(formats.masked_game_theta(states, goals, θ[6:10]), θ)


def test_result_json_keys_and_shapes(tmp_path):
    N, H = 3, 5
    x = np.arange(6 * H * N + 7, dtype=float)            # private blocks first, then shared equality duals
    states, goals = np.arange(4 * N, dtype=float), -np.arange(2 * N, dtype=float)
    res = formats.result_dict(x, states, goals, N, H)
    assert set(res) == {f"Player {i} {k}" for i in (1, 2, 3) for k in
                        ("Initial State", "Goal", "Trajectory", "Control", "Latest Initial State", "Latest Control")}
    xs2 = np.array(res["Player 2 Trajectory"][0])
    us2 = np.array(res["Player 2 Control"][0])
    assert xs2.shape == (H, 4) and us2.shape == (H, 2)
    np.testing.assert_array_equal(xs2.reshape(-1), x[30:50])       # states of player 2, time-major
    np.testing.assert_array_equal(us2.reshape(-1), x[50:60])
    assert res["Player 2 Latest Initial State"] == xs2[1].tolist() and res["Player 2 Latest Control"] == us2[0].tolist()
    p = tmp_path / "result.json"
    formats.write_result_json(str(p), res)
    assert json.loads(p.read_text()) == res
