"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI
(include/rhj.h via radixhashjoin_b200.api), against the oracle on the same inputs.

Bar: BIT-EXACT.  Integer work only -- the sorted multiset of (rowidR,rowidS) pairs must be
identical to the oracle's (the reference leaves the order unspecified), histograms must be equal
counter by counter, partitions equal as per-bucket multisets.
"""
import os

import numpy as np
import pytest
import torch

import _oracle as O
import query_oracle as Q
from radixhashjoin_b200 import (EMIT_COUNT_THEN_WRITE, EMIT_FUSED, DIGIT_HASH, DIGIT_RAW, PAIR_DTYPE, RhjError,
                                Relation, Result)
from radixhashjoin_b200 import workloads as W

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def to_dev(t):
    """numpy TUPLE array -> (n,2) int64 CUDA tensor"""
    a = np.ascontiguousarray(t).view(np.uint64).reshape(-1, 2).view(np.int64)
    return torch.from_numpy(a.copy()).to(DEV)


def pairs_np(p):
    return p.cpu().numpy().view(np.uint64).reshape(-1, 2).view(PAIR_DTYPE).reshape(-1)


def tuples_np(t):
    return t.cpu().numpy().view(np.uint64).reshape(-1, 2).view(O.TUPLE).reshape(-1)


def rand_rel(rng, n, dom, id_base=0):
    return O.as_tuples(rng.permutation(n).astype(np.uint64) + np.uint64(id_base),
                       rng.integers(0, dom, n, dtype=np.uint64))


def check_join(engine, R, S, emit, expect=None):
    if expect is None:
        expect = O.oracle_join(R, S)
    cap = max(len(expect), 1)
    out, n = engine.join_device(to_dev(R), to_dev(S), capacity=cap, emit=emit)
    assert n == len(expect)
    got = pairs_np(out)
    assert np.array_equal(O.sort_pairs(got), O.sort_pairs(expect))
    return engine.last_plan()


# ---- step (1): histogram == HistogramJob/global sum, on the reference's own bucket function -------
@pytest.mark.parametrize("n", [0, 1, 511, 4096, 4097, 100000, 1 << 20])
def test_histogram_equals_reference_histogram(engine, n):
    rng = np.random.default_rng(n + 1)
    T = rand_rel(rng, n, 1 << 40)
    _, hist = O.oracle_partition(T, 256)                     # relation_info.histogram, structs.cpp:168-173
    got = engine.histogram(to_dev(T), 8, 0, DIGIT_RAW).cpu().numpy().view(np.uint64)
    assert np.array_equal(got, hist)


def test_histogram_skewed_digit(engine):
    """every tuple in one bucket (warp-aggregated counters must not lose counts)"""
    T = O.as_tuples(np.arange(100000, dtype=np.uint64), np.full(100000, 0x1234_5600, dtype=np.uint64))
    got = engine.histogram(to_dev(T), 8, 0, DIGIT_RAW).cpu().numpy()
    assert got[0] == 100000 and got.sum() == 100000


# ---- steps (2)+(3): partition == hash_relation as per-bucket multisets ---------------------------
@pytest.mark.parametrize("n,bits", [(0, 8), (1, 8), (4096, 8), (50000, 8), (1 << 20, 8), (300000, 9), (70000, 3)])
def test_partition_equals_reference_partition(engine, n, bits):
    rng = np.random.default_rng(n + bits)
    T = rand_rel(rng, n, 1 << 30)
    exp, hist = O.oracle_partition(T, 1 << bits)
    out, off = engine.partition(to_dev(T), bits, 0, DIGIT_RAW)
    off = off.cpu().numpy().view(np.uint64)
    assert np.array_equal(off, np.concatenate([[0], np.cumsum(hist)]).astype(np.uint64))
    got = tuples_np(out)
    mask = np.uint64((1 << bits) - 1)
    assert np.array_equal(got["payload"] & mask, exp["payload"] & mask)      # same bucket sequence
    key = lambda a: np.lexsort((a["key"], a["payload"], a["payload"] & mask))
    assert np.array_equal(got[key(got)], exp[key(exp)])                      # same multiset per bucket


def test_partition_hash_digits_is_a_permutation(engine):
    rng = np.random.default_rng(77)
    T = rand_rel(rng, 200000, 1 << 50)
    out, off = engine.partition(to_dev(T), 9, 23, DIGIT_HASH)
    got = tuples_np(out)
    assert int(off[-1]) == len(T)
    assert np.array_equal(np.sort(got, order=["key", "payload"]), np.sort(T, order=["key", "payload"]))


# ---- steps (4)+(5): the join ------------------------------------------------------------------------
EDGE = [(0, 0, 10), (0, 9, 10), (9, 0, 10), (1, 1, 1), (2, 3, 1), (5, 4, 2), (100, 100, 1 << 40), (4096, 4096, 1 << 20),
        (4097, 4097, 1 << 20), (4097, 100000, 3000), (43000, 43100, 500), (20000, 3, 7), (3, 20000, 7),
        (300000, 200000, 1 << 18), (1 << 20, 1 << 20, 1 << 19)]


@pytest.mark.parametrize("emit", [EMIT_FUSED, EMIT_COUNT_THEN_WRITE])
@pytest.mark.parametrize("nR,nS,dom", EDGE)
def test_join_equals_oracle(engine, nR, nS, dom, emit):
    rng = np.random.default_rng(nR * 7919 + nS * 31 + dom % 1000)
    R, S = rand_rel(rng, nR, dom), rand_rel(rng, nS, dom, 1 << 33)
    check_join(engine, R, S, emit)


@pytest.mark.parametrize("emit", [EMIT_FUSED, EMIT_COUNT_THEN_WRITE])
def test_join_u64_extremes(engine, emit):
    """values and row ids at the u64 limits; 0xFFFF.. must not collide with any sentinel"""
    rng = np.random.default_rng(5)
    vals = np.array([0, 1, 2**32, 2**63, 2**64 - 1, 2**64 - 256, 255, 256], dtype=np.uint64)
    R = O.as_tuples(rng.integers(2**40, 2**64 - 1, 6000, dtype=np.uint64), vals[rng.integers(0, len(vals), 6000)])
    S = O.as_tuples(rng.integers(2**40, 2**64 - 1, 5000, dtype=np.uint64), vals[rng.integers(0, len(vals), 5000)])
    check_join(engine, R, S, emit)


@pytest.mark.parametrize("emit", [EMIT_FUSED, EMIT_COUNT_THEN_WRITE])
def test_join_build_overflow_single_key(engine, emit):
    """one build key repeated 10000 times: no radix bit can split it -> multi-round table loads"""
    R = O.as_tuples(np.arange(10000, dtype=np.uint64), np.full(10000, 7, dtype=np.uint64))
    S = O.as_tuples(np.arange(300, dtype=np.uint64) + np.uint64(50000),
                    np.array([7, 8, 9] * 100, dtype=np.uint64))
    plan = check_join(engine, R, S, emit)
    assert plan["build_is_S"] == 1
    check_join(engine, S, R, emit)


@pytest.mark.parametrize("emit", [EMIT_FUSED, EMIT_COUNT_THEN_WRITE])
def test_join_two_pass_plan(engine, emit):
    """large enough for two radix passes (bits_total > 9), duplicates on both sides"""
    rng = np.random.default_rng(123)
    n = 3 << 20
    R, S = rand_rel(rng, n, n // 2), rand_rel(rng, n, n // 2, 1 << 40)
    plan = check_join(engine, R, S, emit)
    assert plan["bits_pass2"] > 0 and plan["bits_total"] == plan["bits_pass1"] + plan["bits_pass2"]


def test_join_zipf_probe_skew(engine):
    """Zipf probe keys: one partition holds ~1/20 of the probe side -> probe chunking across CTAs"""
    w = W.zipf_probe(20)
    R, S = W.to_numpy_tuples(w.R), W.to_numpy_tuples(w.S)
    exp = O.oracle_join(R, S)
    assert O.pairs_digest(exp) == tuple(w.expected)
    for emit in (EMIT_FUSED, EMIT_COUNT_THEN_WRITE):
        check_join(engine, R, S, emit, exp)


def test_fused_emitter_reports_needed_capacity(engine):
    rng = np.random.default_rng(8)
    R, S = rand_rel(rng, 5000, 50), rand_rel(rng, 5000, 50)
    exp = O.oracle_join(R, S)
    with pytest.raises(RhjError) as ei:
        engine.join_device(to_dev(R), to_dev(S), capacity=10, emit=EMIT_FUSED)
    assert ei.value.code == 4 and ei.value.needed == len(exp)


def test_count_then_write_two_calls(engine):
    rng = np.random.default_rng(9)
    R, S = rand_rel(rng, 60000, 900), rand_rel(rng, 50000, 900)
    exp = O.oracle_join(R, S)
    n = engine.join_count_device(to_dev(R), to_dev(S))
    assert n == len(exp)
    out = torch.empty((n, 2), dtype=torch.int64, device=DEV)
    engine.join_write_device(out)
    assert np.array_equal(O.sort_pairs(pairs_np(out)), O.sort_pairs(exp))


def test_join_host_and_result_mirror(engine):
    """host relations in, reference-shaped Result out (Result.h:19-38)"""
    rng = np.random.default_rng(10)
    R, S = rand_rel(rng, 30000, 2000), rand_rel(rng, 40000, 2000)
    exp = O.oracle_join(R, S)
    res = Result(engine).multiRadixHashJoin(Relation(R), Relation(S))
    assert not res.isEmpty() and len(res.pairs) == len(exp)
    walked = np.concatenate(list(res.pages()))
    assert np.array_equal(O.sort_pairs(walked), O.sort_pairs(exp))
    assert res.size == (len(exp) % 8191 or 8191)
    empty = Result(engine).multiRadixHashJoin(Relation(R[:10]), Relation(O.as_tuples([1], [2**60])))
    assert empty.isEmpty()


# ---- full-size properties (sizes the oracle cannot finish quickly): closed-form digest ---------------
@pytest.mark.parametrize("make", [lambda: W.uniform_unique(24, DEV), lambda: W.foreign_key(16, 25, DEV),
                                  lambda: W.zipf_probe(24, DEV)], ids=["uniform24", "fk16x25", "zipf24"])
@pytest.mark.parametrize("emit", [EMIT_FUSED, EMIT_COUNT_THEN_WRITE])
def test_closed_form_digest_large(engine, make, emit):
    w = make()
    out, n = engine.join_device(w.R, w.S, capacity=w.expected[0], emit=emit)
    assert (n,) + engine.pairs_digest(out)[1:] == tuple(w.expected)


def test_baseline_config2_uniform_2p27(engine):
    """BASELINE.json configs[1]: 2^27 x 2^27 unique uniform u64, 1:1 -- count and digest in closed form"""
    w = W.uniform_unique(27, DEV)
    out, n = engine.join_device(w.R, w.S, capacity=1 << 27, emit=EMIT_FUSED)
    assert (n,) + engine.pairs_digest(out)[1:] == tuple(w.expected)
    plan = engine.last_plan()
    assert plan["bits_total"] == 16


def test_baseline_config3_foreign_key_full_size(engine):
    """BASELINE.json configs[2] at its full single-GPU size: 2^24-tuple build x 2^30-tuple probe (16 GiB), every probe
    tuple matches one build row -- count and digest in closed form (the oracle does not fit host RAM at this size)."""
    free, _ = torch.cuda.mem_get_info()
    if free < 100 * (1 << 30):
        pytest.skip("needs ~90 GiB of free HBM")
    w = W.foreign_key(24, 30, DEV)
    out, n = engine.join_device(w.R, w.S, capacity=1 << 30, emit=EMIT_FUSED)
    assert (n,) + engine.pairs_digest(out)[1:] == tuple(w.expected)
    del out, w
    torch.cuda.empty_cache()


def test_baseline_config4_zipf_2p28(engine):
    """BASELINE.json configs[3] at full size: Zipf(1.0) probe keys over 2^28 unique build keys; the hottest key draws 1/28
    of the probe side, which forces the exact (histogram) path for the probe relation and probe-chunked work items."""
    free, _ = torch.cuda.mem_get_info()
    if free < 60 * (1 << 30):
        pytest.skip("needs ~50 GiB of free HBM")
    w = W.zipf_probe(28, DEV)
    out, n = engine.join_device(w.R, w.S, capacity=1 << 28, emit=EMIT_FUSED)
    assert (n,) + engine.pairs_digest(out)[1:] == tuple(w.expected)
    assert engine.last_plan()["optimistic_pass1"] & 2 == 0      # the skewed probe side kept its histograms
    del out, w
    torch.cuda.empty_cache()


# ---- neighbours: filters, gathers, checksum -----------------------------------------------------------
def test_filter_gather_sum_equal_oracle(engine):
    rng = np.random.default_rng(11)
    n = 100003
    col = rng.integers(0, 1000, n, dtype=np.uint64)
    dcol = torch.from_numpy(col.view(np.int64)).to(DEV)
    out = np.empty(n, dtype=np.uint64)
    for op in "><=":
        k = O.liborc().orc_filter(col.ctypes.data, n, ord(op), 500, out.ctypes.data)
        got = engine.filter(dcol, op, 500).cpu().numpy().view(np.uint64)
        assert np.array_equal(got, out[:k])
    # a second predicate over the survivors of the first (Query.cpp applies filters in sequence)
    first = engine.filter(dcol, ">", 200)
    second = engine.filter(dcol, "<", 400, rowids=first).cpu().numpy().view(np.uint64)
    assert np.array_equal(second, np.nonzero((col > 200) & (col < 400))[0].astype(np.uint64))
    rows = rng.integers(0, n, 70001, dtype=np.uint64)
    drows = torch.from_numpy(rows.view(np.int64)).to(DEV)
    t = tuples_np(engine.gather_tuples(dcol, drows))
    assert np.array_equal(t["key"], rows) and np.array_equal(t["payload"], col[rows.astype(np.int64)])
    assert engine.gather_sum(dcol, drows) == O.liborc().orc_column_sum(col.ctypes.data, rows.ctypes.data, len(rows))


def test_shuffle_partition_keeps_equal_values_together(engine):
    rng = np.random.default_rng(12)
    T = rand_rel(rng, 100000, 5000)
    out, counts = engine.shuffle_partition(to_dev(T), 8)
    got = tuples_np(out)
    assert sum(counts) == len(T)
    assert np.array_equal(np.sort(got, order=["key", "payload"]), np.sort(T, order=["key", "payload"]))
    from radixhashjoin_b200.distributed import rank_of_values
    # the numpy restatement used by the CPU (gloo) tests is the device's rank function
    assert counts == np.bincount(rank_of_values(T["payload"], 8), minlength=8).tolist()
    owner = {}
    at = 0
    for r, c in enumerate(counts):
        for v in np.unique(got["payload"][at:at + c]):
            assert owner.setdefault(int(v), r) == r
        at += c


# ---- the reference's own golden vectors through the CUDA join ------------------------------------------
def test_small_workload_golden_through_gpu_join(engine, small_dir, small_joins_golden):
    """small.work driven exactly like Query::execute with every join done by librhj.so: all 50 lines of
    small/small.result and all 94 per-join digests the unmodified reference logged."""
    rels = Q.load_workload(small_dir)
    queries = Q.parse_work(os.path.join(small_dir, "small.work"))
    expected = open(os.path.join(small_dir, "small.result")).read().split("\n")
    trace = []
    lines = [Q.execute(q, rels, lambda R, S: engine.join_host(R, S), trace) for q in queries]
    assert lines == expected[:50]
    assert sorted(Q.join_trace_record(*t) for t in trace) == small_joins_golden


# ---- update_intermediate on the GPU ----------------------------------------------------
def _rows(cols):
    """multiset of intermediate rows as a sorted 2-D array"""
    a = np.stack(cols, axis=1) if len(cols) else np.empty((0, 0), dtype=np.uint64)
    return a[np.lexsort(a.T[::-1])] if a.size else a


@pytest.mark.parametrize("match_on_S", [0, 1])
def test_intermediate_expand_equals_oracle(engine, match_on_S):
    """case 2 (intermediate.cpp:108-125): N:M expansion of an existing intermediate by a join result"""
    rng = np.random.default_rng(31 + match_on_S)
    n_rows, n_pairs = 50000, 70000
    inter = [rng.integers(0, 3000, n_rows, dtype=np.uint64), None, rng.integers(0, 10**6, n_rows, dtype=np.uint64), None]
    pairs = np.empty(n_pairs, dtype=PAIR_DTYPE)
    pairs["keyR"] = rng.integers(0, 3500, n_pairs, dtype=np.uint64)
    pairs["keyS"] = rng.integers(0, 3500, n_pairs, dtype=np.uint64)
    pairs = np.unique(pairs)
    t1, t2 = (1, 0) if match_on_S else (0, 1)     # the binding already joined is table2 (S side) iff match_on_S
    exp = Q.update_intermediate(inter, pairs, t1, t2)
    carried, new = engine.intermediate_expand_host(inter[0], pairs, match_on_S, [inter[0], inter[2]])
    got = _rows([carried[0], new, carried[1]])
    assert len(new) == len(exp[0]) > n_rows
    assert np.array_equal(got, _rows([exp[0], exp[1], exp[2]]))


@pytest.mark.parametrize("wide", [False, True])
def test_intermediate_filter_equals_oracle(engine, wide):
    """case 3 (intermediate.cpp:130-138): keep rows whose (row id, row id) pair is in the join result;
    wide = row ids >= 2^32 (composite-key fallback: match on the first id, verify the second)"""
    rng = np.random.default_rng(41)
    n_rows, n_pairs = 60000, 40000
    base = np.uint64(2**40) if wide else np.uint64(0)
    a = rng.integers(0, 300, n_rows, dtype=np.uint64) + base
    b = rng.integers(0, 300, n_rows, dtype=np.uint64) + base
    c = rng.integers(0, 10**9, n_rows, dtype=np.uint64)
    pairs = np.empty(n_pairs, dtype=PAIR_DTYPE)
    pairs["keyR"] = rng.integers(0, 300, n_pairs, dtype=np.uint64) + base
    pairs["keyS"] = rng.integers(0, 300, n_pairs, dtype=np.uint64) + base
    pairs = np.unique(pairs)
    keep = np.isin(a.astype(object) * 2**64 + b.astype(object), pairs["keyR"].astype(object) * 2**64 + pairs["keyS"].astype(object)) \
        if wide else np.isin((a << np.uint64(32)) | b, (pairs["keyR"] << np.uint64(32)) | pairs["keyS"])
    got = engine.intermediate_filter_host(a, b, pairs, [a, b, c])
    assert 0 < len(got[0]) == int(keep.sum())
    assert np.array_equal(_rows(got), _rows([a[keep], b[keep], c[keep]]))


def _gpu_update_fn(engine):
    """update_intermediate with cases 2 and 3 on the GPU (the Python twin of host/intermediate.cpp)"""
    def gpu_update(inter, pairs, t1, t2):
        e1, e2 = inter[t1] is None, inter[t2] is None
        if e1 and e2:
            return Q.update_intermediate(inter, pairs, t1, t2)
        live = [i for i, col in enumerate(inter) if col is not None]
        new = [None] * len(inter)
        if e1 or e2:
            full, fresh = (t2, t1) if e1 else (t1, t2)
            carried, col = engine.intermediate_expand_host(inter[full], pairs, 1 if e1 else 0, [inter[i] for i in live])
            new[fresh] = col
        else:
            carried = engine.intermediate_filter_host(inter[t1], inter[t2], pairs, [inter[i] for i in live])
        for i, colv in zip(live, carried):
            new[i] = colv
        return new
    return gpu_update


def test_small_workload_with_gpu_intermediate(engine, small_dir):
    """small.work with BOTH the join and update_intermediate on the GPU: all 50 checksum lines"""
    rels = Q.load_workload(small_dir)
    queries = Q.parse_work(os.path.join(small_dir, "small.work"))
    expected = open(os.path.join(small_dir, "small.result")).read().split("\n")
    upd = _gpu_update_fn(engine)
    lines = [Q.execute(q, rels, lambda R, S: engine.join_host(R, S), update_fn=upd) for q in queries]
    assert lines == expected[:50]


def test_edge_workload_golden_through_gpu(engine, edge_dir, edge_joins_golden):
    """edge.work (tests/golden/make_edge.py: one relation bound twice, a same-binding predicate, empty joins, a
    one-row relation, a one-value join column, values >= 2^32, a 288000-row intermediate filtered by a third join)
    with every join done by librhj.so -- all 13 lines the unmodified reference printed and its 14 per-join records --
    and again with update_intermediate on the GPU as well."""
    rels = Q.load_workload(edge_dir, "edge.init")
    queries = Q.parse_work(os.path.join(edge_dir, "edge.work"))
    expected = open(os.path.join(edge_dir, "edge.result")).read().split("\n")
    trace = []
    lines = [Q.execute(q, rels, lambda R, S: engine.join_host(R, S), trace) for q in queries]
    assert lines == expected[:13]
    assert sorted(Q.join_trace_record(*t) for t in trace) == edge_joins_golden
    upd = _gpu_update_fn(engine)
    lines = [Q.execute(q, rels, lambda R, S: engine.join_host(R, S), update_fn=upd) for q in queries]
    assert lines == expected[:13]


HOST_BIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "radixhashjoin_b200", "host", "_build")


@pytest.mark.skipif(not os.path.exists(os.path.join(HOST_BIN, "join_b200_full")),
                    reason="drop-in binaries are built in the dev container (need the reference sources)")
def test_reference_program_with_dropin_translation_units(small_dir):
    """The reference PROGRAM itself (join.cpp, Query.cpp, ... unmodified) linked with our Result.cpp +
    intermediate.cpp and librhj.so: `cat small.init small.work | ./join` == small.result."""
    import subprocess
    cwd = os.path.dirname(small_dir)
    data = open(os.path.join(small_dir, "small.init"), "rb").read() + open(os.path.join(small_dir, "small.work"), "rb").read()
    out = subprocess.run([os.path.join(HOST_BIN, "join_b200_full")], input=data, cwd=cwd, capture_output=True, timeout=600)
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    assert out.stdout.decode() == open(os.path.join(small_dir, "small.result")).read()


@pytest.mark.skipif(not os.path.exists(os.path.join(HOST_BIN, "join_b200_full")),
                    reason="drop-in binaries are built in the dev container (need the reference sources)")
def test_reference_program_with_dropin_translation_units_edge_workload(edge_dir):
    """the same program on edge.work: same-binding predicates go through host/intermediate.cpp's parse_table, the third
    join of a query through rhj_intermediate_filter_host, a query without joins prints 0."""
    import subprocess
    cwd = os.path.dirname(edge_dir)
    data = open(os.path.join(edge_dir, "edge.init"), "rb").read() + open(os.path.join(edge_dir, "edge.work"), "rb").read()
    for binary in ("join_b200", "join_b200_full"):
        out = subprocess.run([os.path.join(HOST_BIN, binary)], input=data, cwd=cwd, capture_output=True, timeout=600)
        assert out.returncode == 0, out.stderr.decode()[-2000:]
        assert out.stdout.decode() == open(os.path.join(edge_dir, "edge.result")).read()


# ---- multi-GPU, exact exchange (rhj_shardx_*): emulated ranks on one GPU ------------------------------------
@pytest.mark.parametrize("world,n_local,dom", [(1, 50000, 20000), (2, 60000, 50000), (4, 40000, 1 << 40), (8, 300000, 1 << 20),
                                               (4, 500, 200), (2, 3000, 1000)])
def test_dma_shard_join_emulated_ranks(world, n_local, dom):
    """rhj_shardx_*: pass 1 into local staging, chunks copied to the destinations (tensor copies stand
    in for the peer DMA), pass 2 over (source, partition) pieces, join.  Union == oracle."""
    from radixhashjoin_b200 import RadixHashJoin
    from radixhashjoin_b200.distributed import rank_of_values
    rng = np.random.default_rng(world * 77 + n_local)
    Rg, Sg = rand_rel(rng, world * n_local, dom), rand_rel(rng, world * n_local, dom, 1 << 35)
    expect = O.sort_pairs(O.oracle_join(Rg, Sg))
    engines = [RadixHashJoin(0) for _ in range(world)]
    shards = [(to_dev(Rg[r * n_local:(r + 1) * n_local]), to_dev(Sg[r * n_local:(r + 1) * n_local])) for r in range(world)]
    plan = engines[0].shard_plan(len(Rg), len(Sg), world)
    ndig = world << plan.bits_pass1
    stage = [[torch.empty((n_local, 2), dtype=torch.int64, device=DEV) for _ in range(2)] for _ in range(world)]
    hist = [[torch.empty(ndig, dtype=torch.int64, device=DEV) for _ in range(2)] for _ in range(world)]
    for r in range(world):
        engines[r].shardx_begin(plan)
        for rel in (0, 1):
            engines[r].shardx_pass1(plan, rel, shards[r][rel], stage[r][rel], hist[r][rel])
    lay = [[None, None] for _ in range(world)]
    for rel in (0, 1):
        all_hist = torch.stack([hist[r][rel] for r in range(world)])                     # the all-gather
        for r in range(world):
            lay[r][rel] = engines[r].shardx_layout(plan, r, rel, all_hist)
    recv = [[torch.empty((max(lay[r][rel][3], 1), 2), dtype=torch.int64, device=DEV) for rel in (0, 1)] for r in range(world)]
    for rel in (0, 1):
        assert sum(lay[r][rel][3] for r in range(world)) == world * n_local
        for r in range(world):                                                           # the peer copies
            so, sc, do, _, worst = lay[r][rel]
            assert worst == max(lay[q][rel][3] for q in range(world))   # every rank knows the largest share
            for d in range(world):
                recv[d][rel][do[d]:do[d] + sc[d]].copy_(stage[r][rel][so[d]:so[d] + sc[d]])
    torch.cuda.synchronize()
    got = []
    for r in range(world):
        for rel in (0, 1):
            mine = tuples_np(recv[r][rel][:lay[r][rel][3]])
            assert (rank_of_values(mine["payload"], world) == r).all()
            engines[r].shardx_pass2(plan, rel, recv[r][rel][:lay[r][rel][3]])
        out = torch.empty((max(len(expect), 1), 2), dtype=torch.int64, device=DEV)
        pairs, n = engines[r].shardx_join(plan, out)
        got.append(pairs_np(pairs))
    got = np.concatenate(got)
    assert len(got) == len(expect)
    assert np.array_equal(O.sort_pairs(got), expect)
    for e in engines:
        e.close()


@pytest.mark.parametrize("hot", [0, 200000])
def test_dma_shard_join_histogram_free_second_pass(hot, monkeypatch):
    """Received shards whose pass-1 partition sizes look Poisson take the histogram-free second pass (fixed-capacity
    final partitions, plan bits 4 | 8).  With `hot` copies of one probe value forced through it, one region overflows:
    that rank redoes its second passes through the exact path and still emits exactly the oracle's pairs."""
    from radixhashjoin_b200 import RadixHashJoin
    if hot:
        monkeypatch.setenv("RHJ_FORCE_OPT", "1")
    world, n_local = 2, 1 << 20
    rng = np.random.default_rng(2024 + hot)
    Rg, Sg = rand_rel(rng, world * n_local, 1 << 62), rand_rel(rng, world * n_local, 1 << 62, 1 << 34)
    Sg["payload"][::2] = Rg["payload"][:n_local]                         # half of the probe side matches
    if hot:
        Sg["payload"][1:2 * hot:2] = np.uint64(12345)                     # one hot probe value (absent from R)
    exp = O.pairs_digest(O.oracle_join(Rg, Sg))
    engines = [RadixHashJoin(0) for _ in range(world)]
    shards = [(to_dev(Rg[r * n_local:(r + 1) * n_local]), to_dev(Sg[r * n_local:(r + 1) * n_local])) for r in range(world)]
    plan = engines[0].shard_plan(len(Rg), len(Sg), world)
    ndig = world << plan.bits_pass1
    stage = [[torch.empty((n_local, 2), dtype=torch.int64, device=DEV) for _ in range(2)] for _ in range(world)]
    hist = [[torch.empty(ndig, dtype=torch.int64, device=DEV) for _ in range(2)] for _ in range(world)]
    for r in range(world):
        engines[r].shardx_begin(plan)
        for rel in (0, 1):
            engines[r].shardx_pass1(plan, rel, shards[r][rel], stage[r][rel], hist[r][rel])
    lay = [[None, None] for _ in range(world)]
    for rel in (0, 1):
        all_hist = torch.stack([hist[r][rel] for r in range(world)])
        for r in range(world):
            lay[r][rel] = engines[r].shardx_layout(plan, r, rel, all_hist)
    recv = [[torch.empty((max(lay[r][rel][3], 1), 2), dtype=torch.int64, device=DEV) for rel in (0, 1)] for r in range(world)]
    for rel in (0, 1):
        for r in range(world):
            so, sc, do, _, worst = lay[r][rel]
            assert worst == max(lay[q][rel][3] for q in range(world))   # every rank knows the largest share
            for d in range(world):
                recv[d][rel][do[d]:do[d] + sc[d]].copy_(stage[r][rel][so[d]:so[d] + sc[d]])
    torch.cuda.synchronize()
    total, ssum, sxor, bits = 0, 0, 0, []
    for r in range(world):
        for rel in (0, 1):
            engines[r].shardx_pass2(plan, rel, recv[r][rel][:lay[r][rel][3]])
        out = torch.empty((exp[0], 2), dtype=torch.int64, device=DEV)
        pairs, n = engines[r].shardx_join(plan, out)
        d = engines[r].pairs_digest(pairs)
        total, ssum, sxor = total + n, (ssum + d[1]) % (1 << 64), sxor ^ d[2]
        bits.append(engines[r].last_plan()["optimistic_pass1"])
    assert (total, ssum, sxor) == exp
    if hot:
        assert sorted(bits) == [0, 12]        # the rank that got the hot value fell back, the other one did not
    else:
        assert bits == [12, 12]
    for e in engines:
        e.close()


# ---- pipelined host join (large inputs: chunked probe side, H2D / compute / D2H overlapped) ----------------
@pytest.mark.parametrize("nR,nS,dom", [(30000, 100000, 20000), (100000, 30000, 1 << 40), (4000, 50000, 700), (200000, 200000, 150000),
                                       (5000, 2001, 3)])
def test_pipelined_host_join_equals_oracle(nR, nS, dom, monkeypatch):
    """RHJ_HOST_CHUNK forces the chunked path at test sizes: build side partitioned once, probe side in
    chunks of 7000 tuples through double-buffered upload / result buffers; all pairs, exactly once."""
    from radixhashjoin_b200 import RadixHashJoin
    monkeypatch.setenv("RHJ_HOST_CHUNK", "7000")
    eng = RadixHashJoin(0)
    rng = np.random.default_rng(nR + nS)
    R, S = rand_rel(rng, nR, dom), rand_rel(rng, nS, dom, 1 << 34)
    exp = O.sort_pairs(O.oracle_join(R, S))
    for _ in range(2):                      # second call reuses every buffer
        got = eng.join_host(R, S)
        assert len(got) == len(exp)
        assert np.array_equal(O.sort_pairs(got), exp)
    eng.close()


# ---- optimistic pass 1 (no histogram when a sampled histogram says the partitions are balanced) ----------
def test_optimistic_pass1_taken_and_correct(engine):
    w = W.uniform_unique(23, DEV)
    out, n = engine.join_device(w.R, w.S, capacity=w.expected[0], emit=EMIT_FUSED)
    assert (n,) + engine.pairs_digest(out)[1:] == tuple(w.expected)
    assert engine.last_plan()["optimistic_pass1"] == 7     # both relations in pass 1 (1 | 2) and pass 2 (4)
    z = W.zipf_probe(23, DEV)                      # a hot probe key: the sample keeps the histogram for the probe side only
    out, n = engine.join_device(z.R, z.S, capacity=z.expected[0], emit=EMIT_FUSED)
    assert (n,) + engine.pairs_digest(out)[1:] == tuple(z.expected)
    assert engine.last_plan()["optimistic_pass1"] == 1     # build relation only


@pytest.mark.parametrize("emit", [EMIT_FUSED, EMIT_COUNT_THEN_WRITE])
def test_optimistic_overflow_falls_back_to_exact_path(emit, monkeypatch):
    """RHJ_FORCE_OPT makes skewed data take the fixed-capacity layout: a partition overflows, nothing is
    written out of bounds, and the join is re-run through the histogram path -- same result."""
    from radixhashjoin_b200 import RadixHashJoin
    monkeypatch.setenv("RHJ_FORCE_OPT", "1")
    eng = RadixHashJoin(0)
    z = W.zipf_probe(22, DEV)
    out, n = eng.join_device(z.R, z.S, capacity=z.expected[0], emit=emit)
    assert (n,) + eng.pairs_digest(out)[1:] == tuple(z.expected)
    assert eng.last_plan()["optimistic_pass1"] == 0     # the plan that produced the result is the exact one
    u = W.uniform_unique(22, DEV)
    out, n = eng.join_device(u.R, u.S, capacity=u.expected[0], emit=emit)
    assert (n,) + eng.pairs_digest(out)[1:] == tuple(u.expected)
    assert eng.last_plan()["optimistic_pass1"] == 7
    eng.close()


@pytest.mark.parametrize("emit", [EMIT_FUSED, EMIT_COUNT_THEN_WRITE])
def test_optimistic_pass2_switch_and_parity(emit, monkeypatch):
    """fixed-capacity final partitions (no pass-2 histogram) give the same pairs as the exact layout"""
    from radixhashjoin_b200 import RadixHashJoin
    u = W.uniform_unique(23, DEV)
    rng = np.random.default_rng(31)
    R, S = rand_rel(rng, 1 << 21, 1 << 62), rand_rel(rng, 3 << 20, 1 << 62)
    S["payload"][::3] = R["payload"][:1 << 20]                       # a third of the probe side matches
    exp = O.pairs_digest(O.oracle_join(R, S))
    for no_opt2, want in (("0", 7), ("1", 3)):
        monkeypatch.setenv("RHJ_NO_OPT2", no_opt2)
        eng = RadixHashJoin(0)
        out, n = eng.join_device(u.R, u.S, capacity=u.expected[0], emit=emit)
        assert (n,) + eng.pairs_digest(out)[1:] == tuple(u.expected)
        assert eng.last_plan()["optimistic_pass1"] == want
        out, n = eng.join_device(to_dev(R), to_dev(S), capacity=exp[0], emit=emit)
        assert (n,) + eng.pairs_digest(out)[1:] == exp    # whichever layout the sample chose
        eng.close()


@pytest.mark.parametrize("emit", [EMIT_FUSED, EMIT_COUNT_THEN_WRITE])
def test_optimistic_pass2_overflow_falls_back_to_exact_path(emit, monkeypatch):
    """every build key 64 times: balanced at pass 1, but the final partitions are not Poisson-sized.  Unforced, the
    sample's dispersion keeps the pass-2 histogram; forced, a final partition overflows its region and the join is
    re-run through the exact path.  Same pairs either way."""
    from radixhashjoin_b200 import RadixHashJoin
    n = 1 << 21
    perm = np.random.default_rng(64).permutation(n).astype(np.uint64)   # shuffled rows: the 1/64 sample sees the repeats
    R = O.as_tuples(np.arange(n, dtype=np.uint64), perm // np.uint64(64) * np.uint64(7919))
    S = O.as_tuples(np.arange(n, dtype=np.uint64) + np.uint64(1 << 33), np.arange(n, dtype=np.uint64) * np.uint64(7919 * 16))
    exp = O.pairs_digest(O.oracle_join(R, S))
    assert exp[0] == (n // 64 // 16) * 64
    eng = RadixHashJoin(0)
    out, cnt = eng.join_device(to_dev(R), to_dev(S), capacity=exp[0], emit=emit)
    assert (cnt,) + eng.pairs_digest(out)[1:] == exp
    assert eng.last_plan()["optimistic_pass1"] & 4 == 0      # clumped build side: pass 2 keeps its histogram
    eng.close()
    monkeypatch.setenv("RHJ_FORCE_OPT", "1")
    eng = RadixHashJoin(0)
    out, cnt = eng.join_device(to_dev(R), to_dev(S), capacity=exp[0], emit=emit)
    assert (cnt,) + eng.pairs_digest(out)[1:] == exp
    assert eng.last_plan()["optimistic_pass1"] == 0          # the plan that produced the result is the exact one
    eng.close()


# ---- multi-GPU, pipelined exchange (rhj_pipe_*): emulated ranks on one GPU -------------------------------
def _pipe_emulated_steps(world, n_local, chunks, make_data, steps=3, wire_bytes=16):
    """Drives rhj_pipe_* for `world` contexts on one GPU.  Plain tensors stand in for the symmetric blocks (every
    context sees every block); the phases run rank after rank, so every device-side wait finds its flag already set
    (kernels that wait on one another must not share a GPU).  Yields (expected sorted pairs, [per-rank (pairs, status)])
    per step; consecutive steps alternate the parity of the double-buffered receive side."""
    from radixhashjoin_b200 import RadixHashJoin
    engines = [RadixHashJoin(0) for _ in range(world)]
    try:
        plan = engines[0].shard_plan(world * n_local, world * n_local, world)
        nbytes = engines[0].pipe_sym_bytes(plan, engines[0].pipe_cfg(plan, 0, chunks, n_local, n_local, wire_bytes=wire_bytes))
        syms = [torch.zeros(nbytes // 8, dtype=torch.int64, device=DEV) for _ in range(world)]
        ptrs = [s.data_ptr() for s in syms]
        for r in range(world):
            engines[r].pipe_open(plan, engines[r].pipe_cfg(plan, r, chunks, n_local, n_local, ptrs, wire_bytes=wire_bytes))
        rows = (n_local + chunks - 1) // chunks
        for step in range(1, steps + 1):
            Rg, Sg = make_data(step)
            expect = O.sort_pairs(O.oracle_join(Rg, Sg))
            shards = [(to_dev(Rg[r * n_local:(r + 1) * n_local]), to_dev(Sg[r * n_local:(r + 1) * n_local])) for r in range(world)]
            for r in range(world):
                engines[r].pipe_begin(step)
                for rel in (0, 1):
                    for c in range(chunks):
                        engines[r].pipe_pass1(rel, c, shards[r][rel][c * rows:(c + 1) * rows])
            for r in range(world):
                for rel in (0, 1):
                    for c in range(chunks):
                        engines[r].pipe_ship(rel, c)
            for r in range(world):
                for rel in (1, 0):
                    for c in range(chunks):
                        engines[r].pipe_pass2(rel, c)
                engines[r].pipe_post()
            res = []
            for r in range(world):
                out = torch.empty((max(len(expect), 1), 2), dtype=torch.int64, device=DEV)
                pairs, n, status = engines[r].pipe_join(out)
                res.append((pairs_np(pairs), status))
            yield expect, res
    finally:
        for e in engines:
            e.close()


@pytest.mark.parametrize("world,n_local,dom,chunks", [(1, 50000, 20000, 2), (2, 60000, 1 << 40, 4), (4, 40000, 1 << 40, 3),
                                                      (8, 300000, 1 << 30, 4), (4, 500, 1 << 20, 1), (2, 3001, 1 << 33, 8),
                                                      (8, 70000, 300000, 2)])
@pytest.mark.parametrize("wire_bytes", [16, 12])
def test_pipe_shard_join_emulated_ranks(world, n_local, dom, chunks, wire_bytes):
    """rhj_pipe_*: histogram-free chunked pass 1 into fixed-capacity regions, the copy kernel (plain, or repacking to
    12-byte {value, u32 row id} records), device-side arrival + appended pass 2, join.  Union over ranks == oracle,
    three steps in a row (both parities).  Row ids reach 2^32 - 1 in the 12-byte runs."""
    rng = np.random.default_rng(world * 131 + n_local)
    base = (1 << 35) if wire_bytes == 16 else (1 << 32) - world * n_local

    def make(step):
        return rand_rel(rng, world * n_local, dom), rand_rel(rng, world * n_local, dom, base)

    for expect, res in _pipe_emulated_steps(world, n_local, chunks, make, wire_bytes=wire_bytes):
        assert all(st == 0 for _, st in res), [st for _, st in res]
        got = np.concatenate([p for p, _ in res])
        assert len(got) == len(expect)
        assert np.array_equal(O.sort_pairs(got), expect)


def test_pipe_shard_join_overflow_reaches_every_rank():
    """A single hot value overflows one (destination, sub-digit) region on every sender: every rank must see
    RHJ_PIPE_OVERFLOW (so that all ranks redo the step exactly), and the next, benign step must be clean."""
    world, n_local = 4, 50000
    rng = np.random.default_rng(5)

    def make(step):
        if step == 2:
            R = rand_rel(rng, world * n_local, 1 << 40)
            S = O.as_tuples(np.arange(world * n_local, dtype=np.uint64), np.full(world * n_local, 12345, dtype=np.uint64))
            return R, S
        return rand_rel(rng, world * n_local, 1 << 40), rand_rel(rng, world * n_local, 1 << 40, 1 << 35)

    for step, (expect, res) in enumerate(_pipe_emulated_steps(world, n_local, 2, make), start=1):
        if step == 2:
            assert all(st == 1 for _, st in res), [st for _, st in res]
        else:
            assert all(st == 0 for _, st in res), [st for _, st in res]
            got = np.concatenate([p for p, _ in res])
            assert np.array_equal(O.sort_pairs(got), expect)


def test_pipe_shard_join_12_byte_wire_reports_wide_row_ids():
    """Row ids >= 2^32 cannot travel as 12-byte records: every rank must see RHJ_PIPE_WIDE (8)."""
    world, n_local = 2, 30000
    rng = np.random.default_rng(9)

    def make(step):
        return rand_rel(rng, world * n_local, 1 << 40), rand_rel(rng, world * n_local, 1 << 40, (1 << 32) - 5)

    for expect, res in _pipe_emulated_steps(world, n_local, 2, make, steps=1, wire_bytes=12):
        assert all(st & 8 for _, st in res), [st for _, st in res]


# ---- the device-resident query path (rhj_query_execute) -------------------------------
@pytest.mark.parametrize("n,n_rows", [(0, 10), (1, 1), (5000, 64), (100000, 100000), (300000, 70001), (10, 1 << 22)])
def test_unique_rowids_equals_numpy_unique(engine, n, n_rows):
    """create_relation's de-duplication (structs.cpp:238-241) as a bitmap kernel: == np.unique (ascending)."""
    rng = np.random.default_rng(n + n_rows)
    ids = rng.integers(0, n_rows, n, dtype=np.uint64)
    got = engine.unique_rowids(torch.from_numpy(ids.view(np.int64)).to(DEV), n_rows).cpu().numpy().view(np.uint64)
    assert np.array_equal(got, np.unique(ids))


def test_unique_rowids_rejects_out_of_range_ids(engine):
    ids = torch.tensor([1, 2, 99], dtype=torch.int64, device=DEV)
    with pytest.raises(RhjError):
        engine.unique_rowids(ids, 50)


def _device_query_lines(engine, queries, rels, reorder=False):
    lines, stats = [], []
    for q in queries:
        filters = [(b, c, op, k) for (b, c, op, k) in q.filter]
        sums, st = engine.query_execute(q.table, filters, q.join, q.proj, rels, reorder_joins=reorder)
        lines.append(" ".join("NULL" for _ in q.proj) if sums is None else " ".join(str(s) for s in sums))
        stats.append(st)
    return lines, stats


def test_small_workload_through_device_query_path(engine, small_dir):
    """small.work through rhj_query_execute -- filters, create_relation, 94 joins, update_intermediate and the checksums
    all on the device: the 50 lines of small/small.result; the second pass moves no column over PCIe any more."""
    rels = Q.load_workload(small_dir)
    rels = [[np.ascontiguousarray(c) for c in r] for r in rels]
    queries = Q.parse_work(os.path.join(small_dir, "small.work"))
    expected = open(os.path.join(small_dir, "small.result")).read().split("\n")
    engine.column_cache_clear()
    lines, stats = _device_query_lines(engine, queries, rels)
    assert lines == expected[:50]
    assert sum(s["joins"] for s in stats) > 0 and sum(s["h2d_bytes"] for s in stats) > 0
    lines, stats = _device_query_lines(engine, queries, rels)
    assert lines == expected[:50]
    assert sum(s["h2d_bytes"] for s in stats) == 0          # columns are resident
    assert max(s["d2h_bytes"] for s in stats) < 4096        # counts + checksums only
    # cheapest-first join order: same checksums, and some query really is reordered
    lines, stats = _device_query_lines(engine, queries, rels, reorder=True)
    assert lines == expected[:50]
    assert sum(s["joins_reordered"] for s in stats) > 0
    engine.column_cache_clear()


def test_edge_workload_through_device_query_path(engine, edge_dir):
    """edge.work (same binding twice, a same-binding predicate, empty joins, out-of-range filter, values >= 2^32, a query
    without joins, a third join over a 288000-row intermediate) through rhj_query_execute: the 13 lines the unmodified
    reference printed."""
    rels = Q.load_workload(edge_dir, "edge.init")
    rels = [[np.ascontiguousarray(c) for c in r] for r in rels]
    queries = Q.parse_work(os.path.join(edge_dir, "edge.work"))
    expected = open(os.path.join(edge_dir, "edge.result")).read().split("\n")
    engine.column_cache_clear()
    lines, _ = _device_query_lines(engine, queries, rels)
    assert lines == expected[:13]
    lines, _ = _device_query_lines(engine, queries, rels, reorder=True)
    assert lines == expected[:13]
    engine.column_cache_clear()


def test_device_query_path_random_queries_equal_query_oracle(engine):
    """random relations and random 1-3 join queries (chains, stars, a join between two already joined bindings, duplicate
    heavy columns): rhj_query_execute == the numpy query oracle line by line."""
    rng = np.random.default_rng(77)
    rels = []
    for n, cols, dom in [(3000, 3, 50), (5000, 4, 400), (800, 2, 30), (12000, 3, 2000)]:
        rels.append([np.ascontiguousarray(rng.integers(0, dom, n, dtype=np.uint64)) for _ in range(cols)])
    work = ["0 1|0.0=1.1&0.1>10|0.2 1.0", "0 1 2|0.0=1.0&1.1=2.0&0.1<40|0.0 1.2 2.1", "1 3|0.2=1.1&1.0>5|0.0 1.2",
            "0 1 2|0.0=1.0&1.1=2.0&0.1=2.1|0.2 2.0", "0 2 1 3|0.0=1.0&1.1=2.0&2.2=3.0&3.1<1000|3.2 0.1",
            "2 2|0.0=1.1&0.1>3|0.0 1.0", "0 1|0.0=1.0&1.1=0.1&0.2>25|1.3 0.0", "3|0.0>100|0.1", "0 1|0.0=1.3&0.0>48&1.3<1|0.1"]
    engine.column_cache_clear()
    for line in work:
        q = Q.Query(line)
        want = Q.execute(q, rels, lambda R, S: O.oracle_join(R, S))
        got, _ = _device_query_lines(engine, [q], rels)
        assert got[0] == want, line
        got, _ = _device_query_lines(engine, [q], rels, reorder=True)
        assert got[0] == want, line
    engine.column_cache_clear()


@pytest.mark.skipif(not os.path.exists(os.path.join(HOST_BIN, "join_b200_query")),
                    reason="drop-in binaries are built in the dev container (need the reference sources)")
@pytest.mark.parametrize("which", ["small", "edge"])
def test_reference_program_with_device_query_path(which, small_dir, edge_dir):
    """The reference PROGRAM (join.cpp, parser, schedulers, printing unmodified) with Query::execute replaced by
    host/Query_execute.cpp -> rhj_query_execute: output identical to the unmodified reference's."""
    import subprocess
    d = small_dir if which == "small" else edge_dir
    data = open(os.path.join(d, which + ".init"), "rb").read() + open(os.path.join(d, which + ".work"), "rb").read()
    out = subprocess.run([os.path.join(HOST_BIN, "join_b200_query")], input=data, cwd=os.path.dirname(d), capture_output=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    assert out.stdout.decode() == open(os.path.join(d, which + ".result")).read()


# ---- the north-star kernel variants that are not the default: warp-aggregated histogram counters, TMA bulk-store scatter ----
@pytest.mark.parametrize("env", [{"RHJ_HIST_AGG": "1"}, {"RHJ_SCATTER_MODE": "1"}, {"RHJ_HIST_AGG": "1", "RHJ_SCATTER_MODE": "1"}])
def test_histogram_aggregation_and_bulk_store_scatter_variants(env, monkeypatch):
    """RHJ_HIST_AGG=1 (k_hist<AGG>: match.any-combined shared-memory counters) and RHJ_SCATTER_MODE=1 (k_scatter<kWriteBulk>:
    one TMA bulk store per run) are read at rhj_create: histogram, partition and the exact-path join with them == oracle,
    on uniform and on one-hot-digit inputs."""
    from radixhashjoin_b200 import RadixHashJoin
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    monkeypatch.setenv("RHJ_NO_OPT", "1")      # the exact path runs the histograms
    e = RadixHashJoin(0)
    try:
        rng = np.random.default_rng(11)
        n = 300000
        for vals in (rng.integers(0, 1 << 40, n, dtype=np.uint64), np.full(n, 0x1234, dtype=np.uint64)):
            T = O.as_tuples(np.arange(n, dtype=np.uint64), vals)
            hist = e.histogram(to_dev(T), 8, 0, DIGIT_RAW).cpu().numpy().astype(np.uint64)
            assert np.array_equal(hist, np.bincount((vals & np.uint64(0xFF)).astype(np.int64), minlength=256).astype(np.uint64))
            out, off = e.partition(to_dev(T), 8, 0, DIGIT_RAW)
            got, off = tuples_np(out), off.cpu().numpy()
            assert np.array_equal(off[1:] - off[:-1], hist.astype(np.int64))
            for d in np.flatnonzero(hist)[:8]:
                seg = got[off[d]:off[d + 1]]
                assert ((seg["payload"] & np.uint64(0xFF)) == d).all()
                assert np.array_equal(np.sort(seg["key"]), np.sort(T["key"][(vals & np.uint64(0xFF)) == d]))
        R, S = rand_rel(rng, 1 << 21, 1 << 20), rand_rel(rng, 1 << 21, 1 << 20, 1 << 33)
        for emit in (EMIT_FUSED, EMIT_COUNT_THEN_WRITE):
            plan = check_join(e, R, S, emit)
            assert plan["bits_pass2"] > 0 and plan["optimistic_pass1"] == 0
    finally:
        e.close()


def test_sample_free_shortcut_after_agreeing_joins_and_its_overflow(monkeypatch):
    """After two sampled joins of one shape fitted their histogram-free layouts, the next joins of that shape skip the sample
    (one kernel launch less); a skewed input of the SAME shape then overflows, is redone exactly and correctly, and brings the
    sample back.  RHJ_NO_TRUST=1 samples every time."""
    from radixhashjoin_b200 import RadixHashJoin
    e = RadixHashJoin(0)
    try:
        w = W.uniform_unique(25, DEV)
        out = torch.empty((1 << 25, 2), dtype=torch.int64, device=DEV)
        launches = []
        for _ in range(5):
            pairs, n = e.join_device(w.R, w.S, out=out)
            assert (n,) + e.pairs_digest(pairs)[1:] == tuple(w.expected)
            launches.append(e.last_plan()["kernel_launches"])
            assert e.last_plan()["optimistic_pass1"] & 3 == 3
        assert launches[0] == launches[1] and launches[2] == launches[1] - 1 and launches[4] == launches[2]
        # same shape, but every probe tuple carries one value: the trusted verdict is wrong, the scatter overflows
        S_hot = w.S.clone()
        S_hot[:, 1] = int(w.R[12345, 1])
        pairs, n = e.join_device(w.R, S_hot, out=out)
        assert n == 1 << 25
        p = pairs_np(pairs)
        assert (p["keyR"] == 12345).all() and np.array_equal(np.sort(p["keyS"]), np.arange(1 << 25, dtype=np.uint64))
        assert e.last_plan()["optimistic_pass1"] & 2 == 0          # redone with the probe side on the exact path
        pairs, n = e.join_device(w.R, w.S, out=out)                # ... and the next join samples again
        assert (n,) + e.pairs_digest(pairs)[1:] == tuple(w.expected)
        assert e.last_plan()["kernel_launches"] == launches[0]
    finally:
        e.close()
    monkeypatch.setenv("RHJ_NO_TRUST", "1")
    e = RadixHashJoin(0)
    try:
        ls = []
        for _ in range(4):
            e.join_device(w.R, w.S, out=out)
            ls.append(e.last_plan()["kernel_launches"])
        assert len(set(ls)) == 1
    finally:
        e.close()


# ---- the positional emitter: the variants of k_join_pos (RHJ_JOIN_POS_V) and the r02 kernel k_join<FUSED, POS> ----
@pytest.mark.parametrize("env", [{"RHJ_JOIN_POS_V": "0"}, {"RHJ_JOIN_POS_V": "1"}, {"RHJ_JOIN_POS_V": "3"}, {"RHJ_JOIN_LEAN": "0"}])
def test_positional_emitter_kernels(env, monkeypatch):
    """The fused emitter takes the positional path whenever the output buffer has a slot per probe tuple.  Shapes that
    reach every branch of k_join_pos and of the leftover launch behind it: every probe tuple matches (no holes), some do
    not (holes closed on the host side), duplicate build keys in a few / in all partitions (items handed to the ranked
    kernel, their reserved slots become holes), a build partition of several chunks, probe partitions of many rounds,
    either side as the build side, a result larger than the buffer (redone by the ranked emitter, which reports the
    need).  Sorted pairs == oracle for every kernel."""
    from radixhashjoin_b200 import RadixHashJoin
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    e = RadixHashJoin(0)
    rng = np.random.default_rng(2024)

    def run(R, S, cap=None):
        exp = O.oracle_join(R, S)
        cap = max(len(R), len(S), len(exp), 1) if cap is None else cap
        out, n = e.join_device(to_dev(R), to_dev(S), capacity=cap, emit=EMIT_FUSED)
        assert n == len(exp)
        if n <= 300000:
            assert np.array_equal(O.sort_pairs(pairs_np(out)), O.sort_pairs(exp))
        else:                                            # large results: count + multiset digest (sum and xor of mixed pairs)
            assert O.pairs_digest(pairs_np(out)) == O.pairs_digest(exp)
        return len(exp)

    def uniq(n, id_base=0):
        return O.as_tuples(rng.permutation(n).astype(np.uint64) + np.uint64(id_base), rng.permutation(1 << 22)[:n].astype(np.uint64))

    try:
        for nB, nP in ((1, 1), (5, 3000), (2000, 2000), (2560, 70000), (2561, 70000), (40000, 40000), (1 << 19, 1 << 21)):
            B = uniq(nB)
            # (a) foreign-key style: every probe tuple has exactly one partner
            P = O.as_tuples(np.arange(nP, dtype=np.uint64) + np.uint64(1 << 40), B["payload"][rng.integers(0, nB, nP)])
            assert run(B, P) == nP
            assert run(P, B) == nP                       # the probe side given as R: pairs still come out R-first
            # (b) a third of the probe tuples have no partner: holes
            Pm = P.copy()
            Pm["payload"][::3] += np.uint64(1 << 50)
            run(B, Pm)
            # (c) duplicate build keys in a few partitions: those items go to the ranked kernel
            Bd = B.copy()
            Bd["payload"][: max(1, nB // 50)] = Bd["payload"][-max(1, nB // 50):]
            run(Bd, P)
        # (d) duplicate build keys everywhere (every item is handed over), and one build key repeated beyond a table's capacity
        R, S = rand_rel(rng, 60000, 9000), rand_rel(rng, 50000, 9000, 1 << 33)
        run(R, S)
        hot = O.as_tuples(np.arange(9000, dtype=np.uint64), np.concatenate([np.full(6000, 77, dtype=np.uint64),
                                                                            np.arange(1000, 4000, dtype=np.uint64)]))
        probe = O.as_tuples(np.arange(20000, dtype=np.uint64) + np.uint64(1 << 36), rng.integers(0, 4000, 20000, dtype=np.uint64))
        run(hot, probe)
        # (e) a buffer with a slot per probe tuple but too small for the result: the error carries the need
        exp = O.oracle_join(R, S)
        with pytest.raises(RhjError) as ei:
            e.join_device(to_dev(R), to_dev(S), capacity=max(len(R), len(S)), emit=EMIT_FUSED)
        assert ei.value.code == 4 and ei.value.needed == len(exp)
    finally:
        e.close()
