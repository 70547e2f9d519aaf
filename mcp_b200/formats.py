"""On-disk formats of the reference's application scripts — host-side glue, no GPU work.

* scenario CSV  — `scripts/data_generation.py:40-48`: header `id,x,y,vx,vy,goal_x,goal_y`, one row per agent;
* result JSON   — `examples/parametric_masked_game_solver.jl:60-66`: keys `"Player i Initial State"`, `"Player i Goal"`,
  `"Player i Trajectory"`, `"Player i Control"`, `"Player i Latest Initial State"`, `"Player i Latest Control"`.

The PATH comparison harness of `benchmark/path.jl` needs the closed PATH binary and is not part of this package.
"""
from __future__ import annotations

import csv
import json
import math
from typing import Sequence, Tuple

import numpy as np

CSV_HEADER = ["id", "x", "y", "vx", "vy", "goal_x", "goal_y"]


def generate_agents_and_goals(N: int, bounds=(-2.5, 2.5), min_distance: float = 1.0, rng=None):
    """`generate_agents_and_goals` — `scripts/data_generation.py:5-38`: rejection-sampled positions and goals with a
    pairwise distance of at least `min_distance`, coordinates rounded to 4 decimals, zero initial velocity (the
    reference multiplies its random velocity by 0, `:22-25`).  Returns (states[N,4], goals[N,2])."""
    rng = np.random.default_rng(rng)

    def spread():
        pts = []
        while len(pts) < N:
            p = [round(float(rng.uniform(bounds[0], bounds[1])), 4) for _ in range(2)]
            if all(math.dist(p, q) >= min_distance for q in pts):
                pts.append(p)
        return np.array(pts)

    pos = spread()
    return np.hstack([pos, np.zeros((N, 2))]), spread()


def write_scenario_csv(path: str, states: np.ndarray, goals: np.ndarray) -> None:
    """`save_to_csv` — `scripts/data_generation.py:40-48`."""
    states, goals = np.asarray(states, dtype=float).reshape(-1, 4), np.asarray(goals, dtype=float).reshape(-1, 2)
    with open(path, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(CSV_HEADER)
        for i, (s, g) in enumerate(zip(states, goals)):
            w.writerow([i + 1, *[repr(float(v)) for v in s], *[repr(float(v)) for v in g]])


def read_scenario_csv(path: str) -> Tuple[np.ndarray, np.ndarray]:
    """Returns (initial_states[4N], goals[2N]) stacked over players in id order, the form `run_example` takes
    (`examples/parametric_masked_game_solver.jl:13-14`)."""
    with open(path, newline="") as fh:
        rows = list(csv.DictReader(fh))
    if not rows or any(k not in rows[0] for k in CSV_HEADER):
        raise ValueError(f"{path}: expected header {','.join(CSV_HEADER)}")
    rows.sort(key=lambda r: int(r["id"]))
    states = np.array([[float(r[k]) for k in ("x", "y", "vx", "vy")] for r in rows]).reshape(-1)
    goals = np.array([[float(r[k]) for k in ("goal_x", "goal_y")] for r in rows]).reshape(-1)
    return states, goals


def masked_game_theta(states: Sequence[float], goals: Sequence[float], mask: Sequence[float]) -> np.ndarray:
    """θ of the masked game from a scenario: θ_i = [state_i(4); goal_i(2); mask_i(N)], player 1 carries `mask`, the
    others all ones (`examples/parametric_masked_game_solver.jl:19`, state prefix `examples/utils.jl:27-29`)."""
    states, goals = np.asarray(states, dtype=float).reshape(-1, 4), np.asarray(goals, dtype=float).reshape(-1, 2)
    N = states.shape[0]
    mask = np.asarray(mask, dtype=float)
    if mask.shape != (N,):
        raise ValueError(f"mask must have {N} entries")
    return np.concatenate([np.concatenate([states[i], goals[i], mask if i == 0 else np.ones(N)]) for i in range(N)])


def player_trajectories(x: np.ndarray, N: int, horizon: int):
    """Per player (xs[H,4], us[H,2]) from the primals of one solve (`unpack_trajectory`, `examples/utils.jl:2-16`:
    per player [states(4×H, time-major); controls(2×H)]); `x` holds the N private blocks first."""
    per = 6 * horizon
    out = []
    for i in range(N):
        blk = np.asarray(x[i * per:(i + 1) * per], dtype=float)
        out.append((blk[:4 * horizon].reshape(horizon, 4), blk[4 * horizon:].reshape(horizon, 2)))
    return out


def result_dict(x: np.ndarray, states: Sequence[float], goals: Sequence[float], N: int, horizon: int) -> dict:
    """The dictionary `run_example(…; save=true)` returns (`examples/parametric_masked_game_solver.jl:58-68`) for the
    plan of the LAST solve of a roll-out: trajectory / control wrapped in one-element lists as there."""
    states, goals = np.asarray(states, dtype=float).reshape(-1, 4), np.asarray(goals, dtype=float).reshape(-1, 2)
    res = {}
    for i, (xs, us) in enumerate(player_trajectories(x, N, horizon), start=1):
        res[f"Player {i} Initial State"] = states[i - 1].tolist()
        res[f"Player {i} Goal"] = goals[i - 1].tolist()
        res[f"Player {i} Trajectory"] = [xs.tolist()]
        res[f"Player {i} Control"] = [us.tolist()]
        res[f"Player {i} Latest Initial State"] = xs[1].tolist() if horizon > 1 else xs[0].tolist()
        res[f"Player {i} Latest Control"] = us[0].tolist()
    return res


def write_result_json(path: str, result: dict) -> None:
    with open(path, "w") as fh:
        json.dump(result, fh)
