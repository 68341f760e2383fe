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
