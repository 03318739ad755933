"""Multi-GPU host logic on CPU: element-range partition and the scalar all-reduce, world_size 2 over gloo.
The device operator is stood in for by the oracle (this test is about the sharding, not the kernel)."""
import importlib
import math
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sharding = importlib.import_module("gpu-benchmarking_b200.sharding")


@pytest.mark.parametrize("total", [0, 1, 31, 32, 33, 64, 1000, 131072, 2097152, 67104])
@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("multiple", [1, 2, 32])
def test_ranges_partition_exactly(total, world, multiple):
    r = sharding.all_ranges(total, world, multiple)
    assert r[0][0] == 0 and r[-1][1] == total
    for (b0, e0), (b1, e1) in zip(r, r[1:]):
        assert e0 == b1 and b0 <= e0
    for b, e in r[:-1]:
        assert b % multiple == 0 and e % multiple == 0          # interleaved groups never straddle ranks
    sizes = [e - b for b, e in r]
    assert max(sizes) - min(sizes) <= multiple + total % multiple  # balanced to one unit (+ the tail)


def test_bad_arguments():
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)
    with pytest.raises(ValueError):
        sharding.shard_range(-1, 0, 1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nq, nelmt, coa, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        nm = nq - 1
        b = oracle.gen_basis(nm, nq)
        begin, end = sharding.shard_range(nelmt, rank, world)
        n = end - begin
        # element-dependent input, generated for the GLOBAL element index so the shards differ
        rng = np.random.default_rng(4242)
        full = rng.standard_normal(nelmt * nm ** 3)
        local_in = full[begin * nm ** 3: end * nm ** 3]
        if coa:
            local_in = oracle.to_coa(local_in, n, nm ** 3)
        out = oracle.bwdtrans_hex(nq, nq, nq, n, b, b, b, local_in, coa=coa) if n else np.zeros(0)
        local = oracle.sumsq(out) if n else 0.0
        norm = sharding.global_norm(local)
        tmax = sharding.max_over_ranks([1.0 + rank, 5.0 - rank])
        q.put((rank, begin, end, norm, tmax))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("coa", [False, True])
def test_two_rank_gloo_norm_matches_single_process(coa):
    import torch.multiprocessing as mp
    import oracle
    nq, nelmt, world = 4, 4096 + 64, 2
    nm = nq - 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nq, nelmt, coa, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    b = oracle.gen_basis(nm, nq)
    full = np.random.default_rng(4242).standard_normal(nelmt * nm ** 3)
    want = math.sqrt(oracle.sumsq(oracle.bwdtrans_hex(nq, nq, nq, nelmt, b, b, b, full)))
    assert [(r[1], r[2]) for r in res] == sharding.all_ranges(nelmt, world)
    for _, _, _, norm, tmax in res:
        assert abs(norm - want) / want < 1e-12     # summation order changes with the rank count: not bit-exact
        assert tmax == [2.0, 5.0]
