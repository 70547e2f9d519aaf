"""Regenerates the golden vectors under tests/golden/ from the Python oracle (oracle/ip_oracle.py).

The reference's tests hold no stored vectors, and Julia cannot run here, so these are
RESTATEMENT-DERIVED regression vectors: they freeze the oracle's trajectory (status, iteration counts,
x/y/s, ϵ, kkt_error) so that the C oracle, the CUDA path and future edits are all checked against the
same numbers.  Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mcp_b200 import problems  # noqa: E402
from oracle import ip_oracle as O  # noqa: E402
from oracle.ir_eval import OracleMCP  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def pack(sol):
    return dict(status=sol.status, x=sol.x.tolist(), y=sol.y.tolist(), s=sol.s.tolist(), kkt_error=sol.kkt_error,
                eps=sol.eps, outer_iters=sol.outer_iters, newton_steps=sol.newton_steps,
                inner_iters_per_outer=list(sol.inner_iters_per_outer))


def main():
    out = {}
    # cfg1: README QP at the test θ (test/runtests.jl:19), default kwargs and tol=1e-6
    om = OracleMCP(problems.readme_qp().ir)
    θ = [-0.5, 0.5]
    s4 = O.solve_interior_point(om, θ)
    out["readme_qp_default"] = dict(theta=θ, tol=1e-4, **pack(s4))
    out["readme_qp_default"]["dzdtheta"] = O.solve_jacobian_theta(om, s4, θ).tolist()
    out["readme_qp_default"]["grad_sum_sq"] = O.vjp_theta(om, s4, θ, 2 * s4.x, 2 * s4.y, 0 * s4.s).tolist()
    out["readme_qp_tol1e-6"] = dict(theta=θ, tol=1e-6, **pack(O.solve_interior_point(om, θ, tol=1e-6)))
    Θ = problems.readme_qp_thetas(16, seed=7)
    out["readme_qp_batch"] = dict(theta=Θ.T.tolist(), tol=1e-4,
                                  sols=[pack(O.solve_interior_point(om, Θ[:, b])) for b in range(16)])
    # clamp game (test/runtests.jl:108-115)
    og = OracleMCP(problems.clamp_game().mcp.ir)
    out["clamp_game"] = dict(theta=[-1.0, 0.0, 1.0, 1.0], tol=1e-4,
                             **pack(O.solve_interior_point(og, [-1.0, 0.0, 1.0, 1.0], tol=1e-4)))
    with open(os.path.join(HERE, "small.json"), "w") as f:
        json.dump(out, f, indent=1)
    # cfg3: lane-change, benchmark θ (seed 1), cold start, tol = 1e-6 (benchmark/path.jl:8)
    ol = OracleMCP(problems.lane_change_game().mcp.ir)
    Θ = problems.lane_change_thetas(12, seed=1)
    sols = [O.solve_interior_point(ol, Θ[:, b], tol=1e-6) for b in range(Θ.shape[1])]
    np.savez_compressed(
        os.path.join(HERE, "lane_change_seed1.npz"), theta=Θ, tol=1e-6,
        status=np.array([0 if s.status == "solved" else 1 for s in sols], dtype=np.int32),
        x=np.stack([s.x for s in sols], axis=1), y=np.stack([s.y for s in sols], axis=1),
        s=np.stack([s.s for s in sols], axis=1), kkt_error=np.array([s.kkt_error for s in sols]),
        eps=np.array([s.eps for s in sols]), outer_iters=np.array([s.outer_iters for s in sols], dtype=np.int32),
        newton_steps=np.array([s.newton_steps for s in sols], dtype=np.int32))
    # cfg2 (small instance of the random convex QP, 12 primals / 10 inequalities)
    qp = problems.random_qp(12, 10)
    oq = OracleMCP(qp.ir)
    Θ = problems.random_qp_thetas(6, seed=1, num_primals=12, num_inequalities=10, sparsity_rate=0.5)
    sols = [O.solve_interior_point(oq, Θ[:, b], tol=1e-6) for b in range(Θ.shape[1])]
    np.savez_compressed(
        os.path.join(HERE, "random_qp_12x10_seed1.npz"), theta=Θ, tol=1e-6,
        status=np.array([0 if s.status == "solved" else 1 for s in sols], dtype=np.int32),
        x=np.stack([s.x for s in sols], axis=1), y=np.stack([s.y for s in sols], axis=1),
        s=np.stack([s.s for s in sols], axis=1), eps=np.array([s.eps for s in sols]),
        outer_iters=np.array([s.outer_iters for s in sols], dtype=np.int32),
        newton_steps=np.array([s.newton_steps for s in sols], dtype=np.int32))
    qp100_golden()
    print("golden vectors written to", HERE)


def qp100_golden():
    """cfg2 at the benchmark's own size (100 primals, 100 inequalities, θ of 20 200 entries).  θ is NOT stored (646 KB for
    four instances): it is regenerated from the seed by `problems.random_qp_thetas` and guarded by its SHA-256."""
    import hashlib
    qp = problems.random_qp(100, 100)
    oq = OracleMCP(qp.ir)
    seed, B = 5, 4
    Θ = problems.random_qp_thetas(B, seed=seed)
    sols = [O.solve_interior_point(oq, Θ[:, b], tol=1e-6) for b in range(B)]
    np.savez_compressed(
        os.path.join(HERE, "random_qp_100x100_seed5.npz"), seed=seed, B=B, tol=1e-6,
        theta_sha256=hashlib.sha256(np.ascontiguousarray(Θ).tobytes()).hexdigest(),
        status=np.array([0 if s.status == "solved" else 1 for s in sols], dtype=np.int32),
        x=np.stack([s.x for s in sols], axis=1), y=np.stack([s.y for s in sols], axis=1),
        s=np.stack([s.s for s in sols], axis=1), eps=np.array([s.eps for s in sols]),
        outer_iters=np.array([s.outer_iters for s in sols], dtype=np.int32),
        newton_steps=np.array([s.newton_steps for s in sols], dtype=np.int32))


if __name__ == "__main__":
    main()
