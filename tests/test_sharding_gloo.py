"""The N>1 path on CPU: two gloo ranks shard a θ batch the way bench.py / libmcpb200 do (contiguous
column blocks, no data-path collective), solve their shard (the C oracle stands in for the GPU here),
and reduce timing/counters exactly like the bench.  The union must equal the single-process result."""
import os
import socket

import numpy as np
import pytest

from mcp_b200 import problems, sharding
from oracle import c_oracle as CO


def test_shard_ranges_cover_batch():
    for B in (0, 1, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ir = problems.readme_qp().ir
    Θ = problems.readme_qp_thetas(B, seed=11)
    b0, b1 = sharding.shard_range(B, rank, world)
    sol = CO.solve_batch(ir, Θ[:, b0:b1], nthreads=1)
    ms = sharding.reduce_max_ms(10.0 * (rank + 1))
    solved = sharding.reduce_sum_int(int((sol.status == 0).sum()))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), x=sol.x, b0=b0, b1=b1, ms=ms, solved=solved)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding(tmp_path):
    import torch.multiprocessing as mp
    B, world = 37, 2
    mp.spawn(_worker, args=(world, _free_port(), B, str(tmp_path)), nprocs=world, join=True)
    full = CO.solve_batch(problems.readme_qp().ir, problems.readme_qp_thetas(B, seed=11), nthreads=1)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    x = np.concatenate([p["x"] for p in parts], axis=1)
    np.testing.assert_array_equal(x, full.x)
    assert all(float(p["ms"]) == 20.0 for p in parts)            # max over ranks
    assert all(int(p["solved"]) == int((full.status == 0).sum()) for p in parts)
