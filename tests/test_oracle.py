"""CPU tests that PIN the oracle (oracle/rhj_oracle.c, oracle/query_oracle.py):

1. against the unmodified reference compiled into oracle/_ref (exact page-walk ORDER, not only the
   multiset) whenever that build is present (dev container and, via the snapshot, the GPU box);
2. against the committed golden vectors the reference produced (tests/golden/): small.result (50
   checksum lines) and small_joins.txt (94 per-join digests);
3. against closed-form expected results of the synthetic workloads.
"""
import os

import numpy as np
import pytest

import _oracle as O
import query_oracle as Q
from radixhashjoin_b200 import workloads as W

needs_ref = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


def _rand_rel(rng, n, dom, id_base=0):
    return O.as_tuples(rng.permutation(n).astype(np.uint64) + np.uint64(id_base),
                       rng.integers(0, dom, n, dtype=np.uint64))


CASES = [(0, 0, 10), (0, 9, 10), (9, 0, 10), (1, 1, 1), (2, 3, 1), (3, 2, 1), (1000, 1000, 100),
         (30000, 50000, 5000), (43000, 43100, 500), (100000, 100000, 1 << 40), (20000, 3, 7), (8191, 8191, 1),
         (300, 8192 * 3 // 300 + 1, 1)]


@needs_ref
@pytest.mark.parametrize("nR,nS,dom", CASES)
def test_oracle_equals_reference_in_order(nR, nS, dom):
    """Same pairs in the same page-walk order as Result::multiRadixHashJoin (Result.cpp:90-124)."""
    rng = np.random.default_rng(nR * 7919 + nS * 31 + dom)
    R, S = _rand_rel(rng, nR, dom), _rand_rel(rng, nS, dom, 1000000)
    a = O.oracle_join(R, S)
    b, _ = O.reference_join(R, S)
    assert len(a) == len(b)
    assert np.array_equal(a, b)


@needs_ref
def test_oracle_equals_reference_u64_extremes():
    """Row ids and values >= 2^32 and at the u64 limits (unpinned by small.work, pinned here)."""
    rng = np.random.default_rng(5)
    vals = np.array([0, 1, 2**32, 2**63, 2**64 - 1, 2**64 - 256, 255, 256], dtype=np.uint64)
    R = O.as_tuples(rng.integers(2**40, 2**64 - 1, 4000, dtype=np.uint64), vals[rng.integers(0, len(vals), 4000)])
    S = O.as_tuples(rng.integers(2**40, 2**64 - 1, 3000, dtype=np.uint64), vals[rng.integers(0, len(vals), 3000)])
    a = O.oracle_join(R, S)
    b, _ = O.reference_join(R, S)
    assert np.array_equal(a, b)


@pytest.mark.skipif(not O.have_ref(16), reason="oracle/_ref/libref_rhj_t16.so not built")
def test_sixteen_worker_build_of_the_reference_joins_like_the_shipped_one():
    """bench.py's reference arm may report the 16-worker build (NUM_OF_THREADS overridden at build time, oracle/Makefile): same
    sources, same multiset of pairs as the 8-worker build the oracle is pinned against."""
    rng = np.random.default_rng(16)
    for nR, nS, dom in ((0, 7, 3), (1, 1, 1), (300, 5000, 40), (70000, 90000, 30000), (1 << 19, 1 << 19, 1 << 18)):
        R = O.as_tuples(rng.permutation(nR).astype(np.uint64), rng.integers(0, dom, nR, dtype=np.uint64))
        S = O.as_tuples(rng.permutation(nS).astype(np.uint64) + np.uint64(1 << 35), rng.integers(0, dom, nS, dtype=np.uint64))
        a, _ = O.reference_join(R, S, threads=8)
        b, _ = O.reference_join(R, S, threads=16)
        assert len(a) == len(b) and np.array_equal(O.sort_pairs(a), O.sort_pairs(b))


def test_next_prime_known_answers():
    """auxFun.cpp:4-22 incl. its special cases (2 -> 5 because 3 is skipped by the %3 test)."""
    got = [O.liborc().orc_next_prime(i) for i in range(0, 20)]
    assert got == [2, 2, 5, 5, 5, 7, 7, 11, 11, 11, 11, 13, 13, 17, 17, 17, 17, 19, 19, 23]


def test_hash_relation_is_stable_radix_partition():
    """structs.cpp:144-204: stable partition on payload & 0xFF + histogram."""
    rng = np.random.default_rng(3)
    T = _rand_rel(rng, 50000, 1 << 20)
    out, hist = O.oracle_partition(T, 256)
    order = np.argsort(T["payload"] & np.uint64(255), kind="stable")
    assert np.array_equal(out, T[order])
    assert np.array_equal(hist, np.bincount((T["payload"] & np.uint64(255)).astype(np.int64), minlength=256))


@pytest.mark.parametrize("w", [lambda: W.uniform_unique(15), lambda: W.foreign_key(10, 16), lambda: W.zipf_probe(15)])
def test_closed_form_digests(w):
    """The synthetic generators' closed-form expected digests equal what the oracle computes."""
    wl = w()
    p = O.oracle_join(W.to_numpy_tuples(wl.R), W.to_numpy_tuples(wl.S))
    assert O.pairs_digest(p) == tuple(wl.expected)


def test_digest_helpers_agree():
    rng = np.random.default_rng(9)
    p = np.empty(1000, dtype=O.PAIR)
    p["keyR"] = rng.integers(0, 2**63, 1000, dtype=np.uint64)
    p["keyS"] = rng.integers(0, 2**63, 1000, dtype=np.uint64)
    import ctypes
    s, x = ctypes.c_uint64(), ctypes.c_uint64()
    O.liborc().orc_pairs_digest(p.ctypes.data, len(p), ctypes.byref(s), ctypes.byref(x))
    assert (len(p), s.value, x.value) == O.pairs_digest(p)


def test_small_workload_golden(small_dir, small_joins_golden):
    """small.work through the query oracle with the C oracle's join: all 50 lines of small.result and
    all 94 per-join digests logged by the unmodified reference."""
    rels = Q.load_workload(small_dir)
    queries = Q.parse_work(os.path.join(small_dir, "small.work"))
    expected = open(os.path.join(small_dir, "small.result")).read().split("\n")
    trace = []
    lines = [Q.execute(q, rels, O.oracle_join, trace) for q in queries]
    assert len(lines) == 50
    assert lines == expected[:50]
    assert sum(1 for l in lines if l.startswith("NULL")) == 5
    assert sorted(Q.join_trace_record(*t) for t in trace) == small_joins_golden


def test_edge_workload_golden(edge_dir, edge_joins_golden):
    """edge.work (what small.work never exercises: one relation bound twice, a same-binding predicate, an empty join,
    a filter past the column range, two filters on one binding, values >= 2^32, a one-row relation, a one-value join
    column, a query without joins, a join between two joined bindings over a 288000-row intermediate) through the query
    oracle: every line the unmodified reference printed, and every per-join record it logged."""
    rels = Q.load_workload(edge_dir, "edge.init")
    queries = Q.parse_work(os.path.join(edge_dir, "edge.work"))
    expected = open(os.path.join(edge_dir, "edge.result")).read().split("\n")
    trace = []
    lines = [Q.execute(q, rels, O.oracle_join, trace) for q in queries]
    assert len(lines) == 13
    assert lines == expected[:13]
    assert lines[2] == "NULL NULL" and lines[3] == "NULL" and lines[10] == "0"
    assert sorted(Q.join_trace_record(*t) for t in trace) == edge_joins_golden
    assert len(edge_joins_golden) == 14


def test_filter_gather_sum_oracles():
    rng = np.random.default_rng(11)
    col = rng.integers(0, 1000, 10000, dtype=np.uint64)
    out = np.empty(10000, dtype=np.uint64)
    for op, fn in ((">", lambda v: v > 500), ("<", lambda v: v < 500), ("=", lambda v: v == 500)):
        k = O.liborc().orc_filter(col.ctypes.data, len(col), ord(op), 500, out.ctypes.data)
        assert np.array_equal(out[:k], np.nonzero(fn(col))[0].astype(np.uint64))
    rows = rng.integers(0, 10000, 5000, dtype=np.uint64)
    t = np.empty(5000, dtype=O.TUPLE)
    O.liborc().orc_gather_tuples(col.ctypes.data, rows.ctypes.data, 5000, t.ctypes.data)
    assert np.array_equal(t["key"], rows) and np.array_equal(t["payload"], col[rows.astype(np.int64)])
    assert O.liborc().orc_column_sum(col.ctypes.data, rows.ctypes.data, 5000) == int(col[rows.astype(np.int64)].sum())
