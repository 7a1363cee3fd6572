"""world_size-2 gloo test of the multi-GPU host logic (radixhashjoin_b200/distributed.py) on CPU:
rank partitioning -> counts exchange -> all_to_all of the tuples -> local join -> digest reduce.
The kernels are replaced by a numpy partitioner that restates the device hash, and the local join
by the oracle, so what is exercised is exactly the exchange logic the GPU path uses."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _oracle as O
from radixhashjoin_b200 import workloads as W
from radixhashjoin_b200.distributed import ShardedJoin, cpu_partition_by_rank, rank_of_values

LOG2_LOCAL = 13


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gbits = LOG2_LOCAL + (world.bit_length() - 1)
    n_local = 1 << LOG2_LOCAL
    w = W.uniform_unique(LOG2_LOCAL, "cpu", row_offset=rank * n_local, log2_global=gbits)

    def join_fn(R, S):
        p = O.oracle_join(W.to_numpy_tuples(R), W.to_numpy_tuples(S))
        return p, len(p)

    sj = ShardedJoin(world, rank, lambda T: cpu_partition_by_rank(T, world), join_fn)
    pairs, count, (nR, nS) = sj.step(w.R, w.S)
    # every value this rank received belongs to it
    cnt, s, x = O.pairs_digest(pairs)
    tot = torch.tensor([cnt, nR, nS], dtype=torch.int64)
    dist.all_reduce(tot)
    parts = [None] * world
    dist.all_gather_object(parts, (s, x))
    q.put((rank, int(tot[0]), int(tot[1]), int(tot[2]), parts))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_sharded_join_matches_closed_form(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    gbits = LOG2_LOCAL + (world.bit_length() - 1)
    exp = W.uniform_unique_global_digest(gbits)
    for rank, cnt, nR, nS, parts in res:
        assert cnt == exp[0] == 1 << gbits          # every probe tuple found its partner on some rank
        assert nR == nS == 1 << gbits               # nothing lost or duplicated in the exchange
        s = sum(p[0] for p in parts) & ((1 << 64) - 1)
        x = 0
        for p in parts:
            x ^= p[1]
        assert (s, x) == exp[1:]


def test_rank_function_is_a_function_of_the_value():
    rng = np.random.default_rng(1)
    v = rng.integers(0, 2**63, 10000, dtype=np.uint64)
    for world in (1, 2, 4, 8):
        r = rank_of_values(v, world)
        assert r.min() >= 0 and r.max() < world
        assert np.array_equal(r, rank_of_values(v.copy(), world))
        if world > 1:
            assert np.bincount(r, minlength=world).min() > 10000 / world * 0.8


def test_broadcast_strategy_rule():
    """the build side is replicated only when shuffling both sides would move more (SURVEY 8e small-build caveat)"""
    from radixhashjoin_b200.distributed import broadcast_is_cheaper
    assert broadcast_is_cheaper(1 << 24, 1 << 30, 8)          # BASELINE config 3
    assert broadcast_is_cheaper(1 << 30, 1 << 24, 2)          # either side may be the small one
    assert not broadcast_is_cheaper(1 << 27, 1 << 27, 8)      # config 5 shape: shuffle
    assert not broadcast_is_cheaper(1 << 24, 1 << 30, 1)      # one rank: nothing to exchange
    assert not broadcast_is_cheaper(1 << 28, 1 << 30, 8)      # build x world > probe


def test_sharded_workload_generators_tile_the_global_relations():
    """rank r of N generates exactly rows [r n/N, (r+1) n/N) of the global relation pair, and the per-rank closed-form
    digests add up to the global one -- what bench.py's multi-GPU verification relies on."""
    import torch
    from radixhashjoin_b200 import workloads as W
    M = (1 << 64) - 1
    for make, glob in ((lambda r, n: W.foreign_key(9, 13, rank=r, world=n), W.foreign_key(9, 13)),
                       (lambda r, n: W.zipf_probe(13, rank=r, world=n), W.zipf_probe(13))):
        for world in (2, 4):
            parts = [make(r, world) for r in range(world)]
            assert torch.equal(torch.cat([p.R for p in parts]), glob.R)
            assert torch.equal(torch.cat([p.S for p in parts]), glob.S)
            cnt = sum(p.expected[0] for p in parts)
            s = sum(p.expected[1] for p in parts) & M
            x = 0
            for p in parts:
                x ^= p.expected[2]
            assert (cnt, s, x) == tuple(glob.expected)
    # uniform: rows by offset, digest by range
    g = W.uniform_unique(12)
    a = W.uniform_unique(11, row_offset=0, log2_global=12)
    b = W.uniform_unique(11, row_offset=2048, log2_global=12)
    assert torch.equal(torch.cat([a.S, b.S]), g.S) and torch.equal(torch.cat([a.R, b.R]), g.R)
    d0 = W.uniform_unique_global_digest(12, lo=0, hi=2048)
    d1 = W.uniform_unique_global_digest(12, lo=2048, hi=4096)
    assert (d0[0] + d1[0], (d0[1] + d1[1]) & M, d0[2] ^ d1[2]) == tuple(g.expected)


def _e2e_worker(rank, world, port, q):
    """HostResidentSteps (bench.py's e2e at N > 1) over gloo: the host buffers are plain CPU tensors, the sharded join is the
    numpy partitioner + oracle join of the test above."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from radixhashjoin_b200.distributed import HostResidentSteps
    gbits = LOG2_LOCAL + (world.bit_length() - 1)
    n_local = 1 << LOG2_LOCAL
    w = W.uniform_unique(LOG2_LOCAL, "cpu", row_offset=rank * n_local, log2_global=gbits)
    R, S = w.R.clone(), w.S.clone()
    out = torch.zeros((2 * n_local, 2), dtype=torch.int64)

    def join_fn(a, b):
        p = O.oracle_join(W.to_numpy_tuples(a), W.to_numpy_tuples(b))
        t = torch.from_numpy(p.view(np.uint64).reshape(-1, 2).view(np.int64).copy())
        out[:len(p)] = t
        return out[:len(p)], len(p)

    sj = ShardedJoin(world, rank, lambda T: cpu_partition_by_rank(T, world), join_fn)

    def step():
        pairs, count, _ = sj.step(R, S)
        return pairs, count

    def alloc(shape):
        return torch.empty(shape, dtype=torch.int64)

    def alloc_fails_on_rank0(shape):
        if rank == 0:
            raise RuntimeError("CUDA error: out of memory (simulated)")
        return alloc(shape)

    res = {}
    # (a) everything fits: host copies are made, the device tensors are refilled from them in every step
    hs = HostResidentSteps(step, R, S, out, world, dist=dist, alloc_host=alloc, sync=lambda: None, avail_fn=lambda: 1 << 40)
    assert hs.ok and hs.h2d_bytes == 16 * 2 * n_local
    R.zero_()
    S.zero_()            # the next step must bring the shards back from the host copies
    dt, count, d2h = hs.run(2, warmup=1)
    cnt, s, x = O.pairs_digest(hs.hout[:count].numpy().view(np.uint64).reshape(-1, 2).copy().view(O.PAIR).reshape(-1))
    parts = [None] * world
    dist.all_gather_object(parts, (cnt, s, x))
    res["a"] = (dt > 0, d2h == 16 * count, parts, torch.equal(R, w.R) and torch.equal(S, w.S))
    # (b) ONE rank sees too little host memory: every rank skips, with a reason, and nobody allocates
    hs = HostResidentSteps(step, R, S, out, world, dist=dist, alloc_host=alloc, sync=lambda: None,
                           avail_fn=(lambda: 0) if rank == 1 else (lambda: 1 << 40))
    res["b"] = (hs.ok, hs.why, hs.hR is None)
    # (c) ONE rank's allocation fails: every rank skips
    hs = HostResidentSteps(step, R, S, out, world, dist=dist, alloc_host=alloc_fails_on_rank0, sync=lambda: None,
                           avail_fn=lambda: 1 << 40)
    res["c"] = (hs.ok, hs.why, hs.hR is None)
    q.put((rank, res))
    dist.destroy_process_group()


def test_host_resident_steps_agree_across_ranks():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_e2e_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    gbits = LOG2_LOCAL + 1
    exp = W.uniform_unique_global_digest(gbits)
    for rank in range(world):
        timed, d2h_ok, parts, refilled = res[rank]["a"]
        assert timed and d2h_ok and refilled
        assert sum(p[0] for p in parts) == exp[0]
        assert sum(p[1] for p in parts) & ((1 << 64) - 1) == exp[1]
        x = 0
        for p in parts:
            x ^= p[2]
        assert x == exp[2]
        for case in ("b", "c"):
            ok, why, freed = res[rank][case]
            assert ok is False and why and freed
    assert "host memory" in res[1]["b"][1] and "host memory" in res[0]["b"][1]
    assert "simulated" in res[0]["c"][1] and "another rank" in res[1]["c"][1]


def test_host_memory_available_is_bounded_by_the_machine():
    import psutil
    from radixhashjoin_b200.distributed import host_memory_available
    a = host_memory_available()
    assert 0 < a <= psutil.virtual_memory().total
