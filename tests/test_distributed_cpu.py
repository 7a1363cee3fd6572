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
