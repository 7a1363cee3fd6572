#!/usr/bin/env python3
"""Generates the `edge` workload and pins it with the REFERENCE ITSELF (dev container only: needs
/root/reference, compiled by oracle/Makefile into oracle/_ref/join_ref and join_ref_trace).

small.work never exercises: one relation bound twice, a same-binding predicate (parse_table's first
branch, intermediate.cpp:18-26), an empty join result, a filter past the column range, two filters on one
binding, '=' filters on a joined column, values >= 2^32, a one-row relation, a one-value (hot) join column,
a query without joins, and a third join between two already-joined bindings over a large intermediate.
This script builds six small relations that do, runs the unmodified reference program on them, and writes

    tests/golden/edge_relations.tar.xz   the relations (binary, structs.cpp:17-39 format)
    tests/golden/edge.init / edge.work   the input the reference read
    tests/golden/edge.result             what the reference printed
    tests/golden/edge_joins.txt          one record per executed join (oracle/ref_trace_hook.cpp)

Not pinned, because the reference's behaviour is undefined there: a same-binding predicate AFTER that binding
was joined (intermediate.cpp:27-43 dereferences end() and erases through foreign iterators), and columns whose
value RANGE does not fit a vector<bool> (structs.cpp:52).
"""
import os
import subprocess
import sys
import tarfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")
BIG = np.uint64(1 << 40)


def relations():
    rng = np.random.default_rng(20261018)
    u = lambda a: np.asarray(a, dtype=np.uint64)
    r0c1 = u(rng.integers(0, 50, 2000))
    r0 = [u(np.arange(2000)), r0c1, np.where(rng.random(2000) < 0.3, r0c1, u(rng.integers(0, 50, 2000))),
          u(rng.integers(0, 2000, 2000))]
    r1 = [u(np.arange(3000)), BIG + u(rng.integers(0, 400, 3000)), u(rng.integers(0, 50, 3000))]
    r2 = [u(np.arange(1500)), BIG + u(rng.integers(0, 400, 1500)), u(np.full(1500, 7))]
    r3 = [u([0]), u([7])]
    r4 = [u(np.arange(500)), u(100000 + np.arange(500))]
    r5 = [u(np.arange(4000)), u(rng.integers(0, 2000, 4000)), u(rng.integers(0, 50, 4000))]
    return [r0, r1, r2, r3, r4, r5]


QUERIES = """\
0 0|0.1=1.1&0.0<300|0.0 1.0 1.3
0 1|0.1=0.2&0.1=1.2&1.0>100|0.0 1.0
0 4|0.3=1.1&0.0>5|0.0 1.0
0 1|0.1=1.2&0.0>99999|0.0
1 2|0.1=1.1&0.0<2000|0.0 1.0 0.1
F
2 2|0.2=1.2&0.0<40|0.0 1.0
3 0|0.1=1.1&1.0>0|0.0 1.0 1.1
0 1 5|0.1=1.2&1.2=2.2&0.1=2.2&0.0<60|0.0 1.0 2.0
0 5|0.3=1.1&0.1=7|1.0 0.0
5 0|0.1=1.3&0.2<25&0.0>1000|0.0 1.0
F
0|0.0<100|0.0
4 0|0.0=0.1&0.0=1.3&1.0<1000|0.0 1.0
1 5 2|0.2=1.2&0.1=2.1&1.0<500&0.0<1500|0.1 1.1 2.0
F
"""


def main():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    d = os.path.join(REF, "edge")
    os.makedirs(d, exist_ok=True)
    rels = relations()
    for i, cols in enumerate(rels):
        n = len(cols[0])
        raw = np.concatenate([np.array([n, len(cols)], dtype="<u8")] + [c.astype("<u8") for c in cols])
        raw.tofile(os.path.join(d, f"r{i}"))
    init = "".join(f"./edge/r{i}\n" for i in range(len(rels))) + "Done\n"
    open(os.path.join(d, "edge.init"), "w").write(init)
    open(os.path.join(d, "edge.work"), "w").write(QUERIES)
    stdin = (init + QUERIES).encode()
    out = subprocess.run([os.path.join(REF, "join_ref")], input=stdin, cwd=REF, stdout=subprocess.PIPE, check=True).stdout
    trace = os.path.join(d, "trace.txt")
    if os.path.exists(trace):
        os.remove(trace)
    out2 = subprocess.run([os.path.join(REF, "join_ref_trace")], input=stdin, cwd=REF, stdout=subprocess.PIPE, check=True,
                          env=dict(os.environ, RHJ_TRACE_FILE=trace)).stdout
    assert out == out2, "traced and untraced reference disagree"
    open(os.path.join(HERE, "edge.result"), "wb").write(out)
    lines = sorted(open(trace).read().split("\n"), key=lambda l: [int(v) for v in l.split()] if l.strip() else [])
    open(os.path.join(HERE, "edge_joins.txt"), "w").write("\n".join(l for l in lines if l.strip()) + "\n")
    open(os.path.join(HERE, "edge.init"), "w").write(init)
    open(os.path.join(HERE, "edge.work"), "w").write(QUERIES)
    with tarfile.open(os.path.join(HERE, "edge_relations.tar.xz"), "w:xz") as tf:
        for i in range(len(rels)):
            tf.add(os.path.join(d, f"r{i}"), arcname=f"r{i}")
    sys.stdout.write(out.decode())
    print("wrote edge.result (%d queries), edge_joins.txt (%d joins)" % (out.count(b"\n"), len([l for l in lines if l.strip()])))


if __name__ == "__main__":
    main()
