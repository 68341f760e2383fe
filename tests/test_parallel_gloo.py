"""CPU, world_size 2, gloo: the data-parallel host logic (sharding, gradient averaging that feeds
the SGD grad_scale, max-over-ranks timing, result gathering)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    from efficientdet_b200 import parallel
    assert parallel.init_from_env("gloo") == (rank, world)
    lo, hi = parallel.shard_range(8, rank, world)
    data = torch.arange(8, dtype=torch.float32)[lo:hi]
    # each replica: gradient of mean over its shard of sum_i w*x_i  -> x_mean(shard)
    g = torch.stack([data.mean(), torch.tensor(float(rank + 1))])
    scale = parallel.allreduce_gradients_(g)
    mx = parallel.max_over_ranks(10.0 + rank)
    gathered = parallel.gather_concat(data[:, None])
    out[rank] = (lo, hi, (g * scale).tolist(), mx, gathered[:, 0].tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_dp_host_logic_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0][:2] == (0, 4) and out[1][:2] == (4, 8)
    for r in range(world):
        g = out[r][2]
        # averaged gradient == gradient of the global-batch mean when shards are equal-sized
        assert abs(g[0] - np.arange(8).mean()) < 1e-6 and abs(g[1] - 1.5) < 1e-6
        assert out[r][3] == 11.0
        assert out[r][4] == list(range(8))


def test_shard_range_rejects_ragged():
    import pytest
    from efficientdet_b200 import parallel
    with pytest.raises(ValueError):
        parallel.shard_range(7, 0, 2)
    assert parallel.shard_range(6, 2, 3) == (4, 6)
    assert parallel.world() == (0, 1)


def test_plan_buckets():
    from efficientdet_b200 import parallel
    # D4-like: heads 5.3 M, BiFPN 2.9 M, backbone stages 7-5 14 M, rest 2.9 M (+ stem)
    total = 25_000_000
    marks = [(100, 19_700_000), (400, 16_800_000), (700, 2_900_000), (1200, 0)]
    b = parallel.plan_buckets(marks, total)
    assert b == [(100, 19_700_000, total), (400, 16_800_000, 19_700_000), (700, 2_900_000, 16_800_000),
                 (1200, 0, 2_900_000)]
    # buckets tile the range without gaps, in descending order, and end at the last launch
    assert b[-1][0] == 1200 and all(x[1] == y[2] for x, y in zip(b, b[1:]))
    # D0 with a frozen backbone: 0.63 M trainable parameters -> one bucket, one segment
    marks = [(60, 3_980_000), (200, 3_630_000)]
    assert parallel.plan_buckets(marks, 4_260_000) == [(200, 3_630_000, 4_260_000)]
    # small buckets merge forward; a trailing empty range only extends the last segment
    assert parallel.plan_buckets([(10, 90), (20, 50), (30, 50)], 100, min_elems=20) == [(30, 50, 100)]
    assert parallel.plan_buckets([(10, 70), (20, 40), (30, 40)], 100, min_elems=20) == [(10, 70, 100), (30, 40, 70)]
    import pytest
    with pytest.raises(ValueError):
        parallel.plan_buckets([(10, 50), (20, 60)], 100, 1)


def _bucket_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from efficientdet_b200 import parallel
    parallel.init_from_env("gloo")
    g = torch.arange(100, dtype=torch.float32) * (rank + 1)
    one = g.clone()
    parallel.allreduce_gradients_(one)
    for _, lo, hi in parallel.plan_buckets([(1, 60), (2, 25), (3, 0)], 100, min_elems=10):
        parallel.allreduce_gradients_(g[lo:hi])
    out[rank] = bool(torch.equal(g, one))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_allreduce_equals_single_allreduce_world2():
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_bucket_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]
