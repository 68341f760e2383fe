"""Data-parallel plumbing (one process per GPU, torch.distributed).

The reference's only parallelism is synchronous data parallelism via
tf.distribute.MirroredStrategy (train_tpu.py:233-247, :315-316): every replica runs the step on
its shard of the global batch, per-replica gradients are summed across replicas with the loss
pre-scaled by 1/replicas (i.e. gradients are averaged), BatchNorm statistics and the loss
normalisers (#positive anchors) stay per replica.  Inference is embarrassingly batch-parallel
("replicas only": no collective).  Backend: NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU
tests.
"""
import os

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend=None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT)."""
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1 or dist.is_initialized():
        return world()
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    dist.init_process_group(backend)
    return world()


def shard_range(global_batch, rank, world_size):
    """[begin, end) rows of the global batch owned by `rank` (global = per_replica * replicas,
    train_tpu.py:315-316)."""
    if global_batch % world_size:
        raise ValueError("global batch %d is not divisible by %d replicas" % (global_batch, world_size))
    per = global_batch // world_size
    return rank * per, (rank + 1) * per


def allreduce_gradients_(flat_grads):
    """In-place SUM over replicas of the flat gradient range; the optimizer kernel applies the
    1/replicas factor (effdet_sgd_momentum_step grad_scale).  Returns the averaging factor."""
    _, n = world()
    if n > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    return 1.0 / n


def plan_buckets(marks, total, min_elems=1 << 20):
    """Gradient buckets for the overlapped all-reduce.  The flat gradient buffer is laid out in creation
    order (backbone, BiFPN, heads) and the backward pass produces it from the END: `marks` lists, in
    backward-production order, (index one past the last launch of a backward segment, lowest flat index that
    segment completes).  Returns [(launch_end, lo, hi)]: after launches [.., launch_end) have run,
    flat[lo:hi) is final and can be all-reduced while the next segment computes.  Buckets smaller than
    `min_elems` are merged into the following one (a collective has a fixed launch cost of tens of
    microseconds); the last bucket always closes the range."""
    out, hi = [], int(total)
    for i, (launch_end, lo) in enumerate(marks):
        lo = int(lo)
        if lo > hi:
            raise ValueError("bucket marks must descend through the flat buffer")
        last = i == len(marks) - 1
        if hi - lo >= min_elems or (last and hi > lo):
            out.append((int(launch_end), lo, hi))
            hi = lo
        elif last:                       # nothing left to reduce: the trailing launches join the last bucket
            if out:
                out[-1] = (int(launch_end), out[-1][1], out[-1][2])
            else:
                out.append((int(launch_end), lo, lo))
    return out


def max_over_ranks(value, device=None):
    """Timing reduction used by bench.py: device time = max over ranks."""
    _, n = world()
    if n == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def gather_concat(t):
    """Concatenate per-rank outputs along dim 0 on every rank (inference result collection)."""
    _, n = world()
    if n == 1:
        return t
    parts = [torch.empty_like(t) for _ in range(n)]
    dist.all_gather(parts, t.contiguous())
    return torch.cat(parts, 0)
