"""Host-side sharding of the hot path over the GPUs of one box.

Elements (benchmark04/05), vector index ranges (benchmark01/02) and matrix rows (benchmark03) are
independent, so the path shards with NO data-path collective: every rank owns one contiguous range, the
basis matrices (and benchmark03's x) are replicated, and the only exchange is the all-reduce of the scalar
partial sums behind the `norm:` columns.  This module is the single place that logic lives; bench.py and the
tests use it, the C++ drivers restate it (utils/bench_common.h: shard_range).
"""
import math


def shard_range(total, rank, world, multiple=32):
    """[begin, end) owned by `rank`: contiguous, in units of `multiple` (32 keeps warp-interleaved groups of
    the _Coa layout and 16-byte vectors from straddling ranks), remainders spread over the first ranks, the
    tail (< multiple) to the last rank.  Ranks may own an empty range when total < world * multiple."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    if total < 0 or multiple < 1:
        raise ValueError("bad total/multiple")
    units, tail = divmod(total, multiple)
    base, extra = divmod(units, world)
    begin_u = rank * base + min(rank, extra)
    end_u = begin_u + base + (1 if rank < extra else 0)
    begin, end = begin_u * multiple, end_u * multiple
    if rank == world - 1:
        end += tail
    return begin, end


def all_ranges(total, world, multiple=32):
    return [shard_range(total, r, world, multiple) for r in range(world)]


def global_sumsq(local_sumsq, device=None):
    """all-reduce (sum) of one double over the default process group (NCCL on the GPUs, gloo in the CPU tests);
    returns the global value on every rank.  With no process group it is the identity."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(local_sumsq)
    t = torch.tensor([float(local_sumsq)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def global_norm(local_sumsq, device=None):
    return math.sqrt(global_sumsq(local_sumsq, device))


def max_over_ranks(values, device=None):
    """element-wise max of a list of floats over ranks (device-timed milliseconds -> job time)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]
