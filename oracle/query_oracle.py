"""query_oracle.py -- numpy restatement of the reference's query path AROUND the join.

TEST INFRASTRUCTURE ONLY (same rule as rhj_oracle.c: only tests/, smoke() and bench.py's CPU legs may
import it).  It exists so that the join -- the oracle's or the CUDA library's, passed in as `join_fn`
-- can be exercised exactly the way Query::execute (Query.cpp:204-211) drives it on the contest
workload, and the printed checksums compared with small/small.result.

Parity status: PINNED.  tests/test_oracle.py runs small.work through this file with the C oracle's
join and requires all 50 output lines == tests/golden/small.result and the 94 per-join digests ==
tests/golden/small_joins.txt (both produced by the unmodified reference, tests/golden/make_small_joins.sh).

Restated (file:line in the reference):
  relList loader                      structs.cpp:17-39   (header u64 n, u64 cols, column-major u64)
  query line parser                   Query.cpp:10-63,237-242
  Query::run_filters                  Query.cpp:81-158    (statistics early-outs are observably
                                                          "filter result is empty -> NULL")
  relation::create_relation / foo     structs.cpp:217-243 (row ids de-duplicated when taken from an
                                                          intermediate, 238-241)
  update_intermediate (3 cases)       intermediate.cpp:52-183
  column_proj                         Query.cpp:66-74
  Query::print                        Query.cpp:226-235
"""
import os

import numpy as np

TUPLE = np.dtype([("key", "<u8"), ("payload", "<u8")])
PAIR = np.dtype([("keyR", "<u8"), ("keyS", "<u8")])


def load_relation(path):
    """structs.cpp:17-39 -> list of column arrays (u64)."""
    raw = np.fromfile(path, dtype="<u8")
    n, cols = int(raw[0]), int(raw[1])
    assert raw.size == n * cols + 2
    return [raw[2 + c * n: 2 + (c + 1) * n] for c in range(cols)]


def load_workload(directory, init="small.init"):
    rels = []
    with open(os.path.join(directory, init)) as f:
        for line in f:
            line = line.strip()
            if line == "Done" or not line:
                break
            rels.append(load_relation(os.path.join(directory, os.path.basename(line))))
    return rels


class Query:
    """Query.cpp:237-242 -- `tables|predicates|projections`."""

    def __init__(self, line):
        tables, preds, projs = line.strip().split("|")
        self.table = [int(t) for t in tables.split()]
        self.join, self.filter = [], []
        for p in preds.split("&"):
            for op in "=<>":
                if op in p:
                    lhs, rhs = p.split(op)
                    break
            t1, c1 = (int(x) for x in lhs.split("."))
            if "." in rhs:                       # Query.cpp:47-49: a '.' after the number => join
                t2, c2 = (int(x) for x in rhs.split("."))
                self.join.append((t1, c1, t2, c2))
            else:
                self.filter.append((t1, c1, op, int(rhs)))
        self.proj = [tuple(int(x) for x in p.split(".")) for p in projs.split()]


def parse_work(path):
    """join.cpp:27-40 -- batches separated by 'F' lines; returns the flat list of queries."""
    out = []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if line and line != "F":
                out.append(Query(line))
    return out


def run_filters(q, relations):
    """Query.cpp:81-158.  Returns (filtered_out, {binding: ascending row ids})."""
    filtered = {i: np.arange(len(relations[t][0]), dtype=np.uint64) for i, t in enumerate(q.table)}
    for (b, col, op, number) in q.filter:
        rows = filtered[b]
        vals = relations[q.table[b]][col][rows.astype(np.int64)]
        keep = vals > number if op == ">" else vals < number if op == "<" else vals == number
        filtered[b] = rows[keep]
        if filtered[b].size == 0:          # Query.cpp:104-106,124-126,140-142 (and the stats early-outs 95-97,115-117)
            return True, filtered
    return False, filtered


def create_relation(column, filtered_rows, inter_col):
    """structs.cpp:230-243: rows from the filter set, or the DISTINCT row ids of the intermediate."""
    rows = filtered_rows if inter_col is None else np.unique(inter_col)
    t = np.empty(rows.size, dtype=TUPLE)
    t["key"] = rows
    t["payload"] = column[rows.astype(np.int64)]
    return t


def _expand(existing_key, pair_key):
    """All (pair index, element index) with existing_key[e] == pair_key[p], pair-major then element
    ascending -- the visiting order of change_intermediate (intermediate.cpp:52-66,108-125)."""
    order = np.argsort(existing_key, kind="stable")
    ks = existing_key[order]
    lo = np.searchsorted(ks, pair_key, "left")
    hi = np.searchsorted(ks, pair_key, "right")
    cnt = (hi - lo).astype(np.int64)
    total = int(cnt.sum())
    pair_idx = np.repeat(np.arange(pair_key.size, dtype=np.int64), cnt)
    start = np.repeat(lo.astype(np.int64), cnt)
    first = np.repeat(np.cumsum(cnt) - cnt, cnt)
    within = np.arange(total, dtype=np.int64) - first
    elem_idx = order[start + within]
    return pair_idx, elem_idx


def update_intermediate(inter, pairs, t1, t2):
    """intermediate.cpp:146-183.  inter: list of u64 arrays or None (= empty vector)."""
    e1, e2 = inter[t1] is None, inter[t2] is None
    new = [None] * len(inter)
    if e1 and e2:                                        # case 1, 92-103 / 153-161
        new = list(inter)
        new[t1] = pairs["keyR"].copy()
        new[t2] = pairs["keyS"].copy()
        return new
    if e1 or e2:                                         # case 2, 108-125 / 162-170
        if e1:
            full, empty, match_key, new_val = t2, t1, pairs["keyS"], pairs["keyR"]
        else:
            full, empty, match_key, new_val = t1, t2, pairs["keyR"], pairs["keyS"]
        pair_idx, elem_idx = _expand(inter[full], match_key)
        for i, col in enumerate(inter):
            if col is not None:
                new[i] = col[elem_idx]
        new[empty] = new_val[pair_idx]
        return new
    # case 3, 72-87 / 130-138 / 171-180: rows whose (t1, t2) row-id pair is in the result
    assert max(int(inter[t1].max()), int(inter[t2].max())) < 2**32
    ek = (inter[t1] << np.uint64(32)) | inter[t2]
    pk = (pairs["keyR"] << np.uint64(32)) | pairs["keyS"]
    _, elem_idx = _expand(ek, pk)
    for i, col in enumerate(inter):
        if col is not None:
            new[i] = col[elem_idx]
    return new


def execute(q, relations, join_fn, trace=None, update_fn=None):
    """Query::execute (Query.cpp:204-211) + run_joins (164-201).  join_fn(R, S) -> PAIR array.
    Returns the printed line (Query.cpp:226-235).  `trace` collects (R, S, pairs) per executed join.
    `update_fn(inter, pairs, t1, t2)` replaces update_intermediate (the tests pass the CUDA one)."""
    if update_fn is None:
        update_fn = update_intermediate
    filtered_out, filtered = run_filters(q, relations)
    inter = [None] * len(q.table)
    if not filtered_out:
        for (t1, c1, t2, c2) in q.join:
            if t1 == t2:
                # parse_table (intermediate.cpp:11-44): a predicate between two columns of ONE binding.  Only its first
                # branch is defined behaviour in the reference (the binding has not been joined yet: keep the filtered
                # rows whose two values are equal, 18-26); the second branch dereferences end() (27-43).
                if inter[t1] is not None:
                    raise NotImplementedError("same-binding predicate on an already joined binding: undefined behaviour "
                                              "in the reference (intermediate.cpp:27-43)")
                cols = relations[q.table[t1]]
                rows = filtered[t1]
                keep = rows[cols[c1][rows.astype(np.int64)] == cols[c2][rows.astype(np.int64)]]
                inter[t1] = keep if keep.size else None      # an empty vector reads as "not joined yet" (structs.cpp:230)
                continue
            R = create_relation(relations[q.table[t1]][c1], filtered[t1], inter[t1])
            S = create_relation(relations[q.table[t2]][c2], filtered[t2], inter[t2])
            pairs = join_fn(R, S)
            if trace is not None:
                trace.append((R, S, pairs))
            if len(pairs) == 0:                      # Query.cpp:188-191
                filtered_out = True
                break
            inter = update_fn(inter, pairs, t1, t2)
    if filtered_out:
        return " ".join("NULL" for _ in q.proj)
    sums = []
    for (b, col) in q.proj:                          # Query.cpp:198-200 -> 66-74
        rows = inter[b]
        if rows is None:
            sums.append(0)
        else:
            with np.errstate(over="ignore"):
                sums.append(int(np.add.reduce(relations[q.table[b]][col][rows.astype(np.int64)], dtype=np.uint64)))
    return " ".join(str(s) for s in sums)


def mix64(x):
    x = np.asarray(x, dtype=np.uint64).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xbf58476d1ce4e5b9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94d049bb133111eb)
        x ^= x >> np.uint64(31)
    return x


def join_trace_record(R, S, pairs):
    """The line oracle/ref_trace_hook.cpp logs for one join: nR nS count in_digest out_sum out_xor."""
    with np.errstate(over="ignore"):
        def rel_digest(t, salt):
            h = mix64(mix64(t["key"] + np.uint64(salt)) ^ t["payload"])
            return int(np.add.reduce(h, dtype=np.uint64)) if h.size else 0
        din = (rel_digest(R, 1) * 31 + rel_digest(S, 2)) & (2**64 - 1)
        if len(pairs):
            h = mix64(pairs["keyR"] * np.uint64(0x100000001b3) + pairs["keyS"])
            osum = int(np.add.reduce(h, dtype=np.uint64))
            oxor = int(np.bitwise_xor.reduce(h))
        else:
            osum = oxor = 0
    return (len(R), len(S), len(pairs), din, osum, oxor)
