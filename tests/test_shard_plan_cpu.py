"""The host-side planning of the multi-GPU joins, which needs no GPU: rhj_shard_plan_make and rhj_pipe_sym_bytes are pure
functions of the sizes (include/rhj.h).  What the N = 2 / 4 / 8 runs of bench.py rely on: rank bits + pass-1 sub-digit bits fit
one pass of at most 1024 digits, the second pass at most 1024, the final partitions hold ~2048 build tuples (one shared-memory
table of k_join), every rank computes the same symmetric-block size, the 12-byte wire format shrinks it."""
import ctypes

import pytest

from radixhashjoin_b200 import _lib

RHJ_OK, RHJ_ERR_ARG = 0, 2


def plan(nR, nS, world):
    p = _lib.ShardPlan()
    rc = _lib.load().rhj_shard_plan_make(nR, nS, world, ctypes.byref(p))
    return rc, p


@pytest.mark.parametrize("world", [1, 2, 4, 8, 16])
@pytest.mark.parametrize("log2_per_gpu", [10, 16, 20, 24, 27, 28])
def test_shard_plan_properties(world, log2_per_gpu):
    n_global = (1 << log2_per_gpu) * world          # weak scaling: bench.py's shape (config 5: 2^28 per GPU x 8)
    rc, p = plan(n_global, n_global, world)
    assert rc == RHJ_OK
    assert p.world == world and (1 << p.rank_bits) == world
    assert p.bits_total == p.bits_pass1 + p.bits_pass2
    assert p.rank_bits + p.bits_pass1 <= 10          # one pass separates destinations and pass-1 partitions: <= 1024 digits
    assert p.bits_pass2 <= 10
    per_rank = n_global // world
    if per_rank > 2560:                              # more than one table load: partitioned down to ~2048 build tuples
        assert (per_rank >> p.bits_total) <= 2048 or p.bits_total == p.bits_pass1 + 10
        assert p.bits_total == 0 or (per_rank >> (p.bits_total - 1)) > 2048
    else:
        assert p.bits_total == 0
    assert p.build_is_S == 0


def test_shard_plan_build_side_and_bad_arguments():
    rc, p = plan(1 << 30, 1 << 24, 8)                # config 3 with the relations swapped: S is the build side
    assert rc == RHJ_OK and p.build_is_S == 1
    rc, q = plan(1 << 24, 1 << 30, 8)
    assert rc == RHJ_OK and q.build_is_S == 0 and q.bits_total == p.bits_total      # sized by the smaller side per rank
    for bad_world in (0, 3, 6, 32, -1):
        assert plan(1 << 20, 1 << 20, bad_world)[0] == RHJ_ERR_ARG
    assert _lib.load().rhj_shard_plan_make(1 << 20, 1 << 20, 2, None) == RHJ_ERR_ARG


def sym_bytes(p, world, rank, chunks, nR, nS, wire):
    cfg = _lib.PipeCfg()
    cfg.world, cfg.rank, cfg.chunks, cfg.wire_bytes = world, rank, chunks, wire
    cfg.nR_local_max, cfg.nS_local_max = nR, nS
    return int(_lib.load().rhj_pipe_sym_bytes(ctypes.byref(p), ctypes.byref(cfg)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_pipe_symmetric_block_is_the_same_on_every_rank_and_sized_for_180_gb(world):
    n = 1 << 27                                      # bench.py at N > 1: 2^27 + 2^27 tuples per GPU
    rc, p = plan(n * world, n * world, world)
    assert rc == RHJ_OK
    sizes = {sym_bytes(p, world, r, 4, n, n, 16) for r in range(world)}
    assert len(sizes) == 1                           # symmetric memory: the same layout arithmetic everywhere
    b16 = sizes.pop()
    b12 = sym_bytes(p, world, 0, 4, n, n, 12)
    data = 2 * 2 * n * 16                            # both relations, double-buffered by step parity
    assert data < b16 < 1.5 * data                   # fixed-capacity regions: expected size + headroom, not a multiple
    assert 0.70 < b12 / b16 < 0.80                   # 12-byte records on the wire
    assert b16 < 40 << 30                            # next to 8 GiB of inputs, staging and the final partitions in 180 GB
    # more chunks = smaller regions with relatively more headroom
    assert sym_bytes(p, world, 0, 8, n, n, 16) > b16


def test_pipe_symmetric_block_rejects_bad_configurations():
    rc, p = plan(1 << 22, 1 << 22, 4)
    assert rc == RHJ_OK
    ok = sym_bytes(p, 4, 0, 4, 1 << 20, 1 << 20, 16)
    assert ok > 0
    assert sym_bytes(p, 2, 0, 4, 1 << 20, 1 << 20, 16) == 0      # world differs from the plan's
    assert sym_bytes(p, 4, 4, 4, 1 << 20, 1 << 20, 16) == 0      # rank out of range
    assert sym_bytes(p, 4, 0, 0, 1 << 20, 1 << 20, 16) == 0      # no chunks
    assert sym_bytes(p, 4, 0, 99, 1 << 20, 1 << 20, 16) == 0     # too many chunks
    assert sym_bytes(p, 4, 0, 4, 1 << 20, 1 << 20, 8) == 0       # unknown wire format
    assert sym_bytes(p, 4, 0, 4, 1 << 20, 1 << 20, 0) == ok      # 0 = the 16-byte default
