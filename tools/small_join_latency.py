"""Per-call latency of rhj_join_host on contest-sized joins (steady state, one thread)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from radixhashjoin_b200 import RadixHashJoin, workloads as W
eng = RadixHashJoin(0)
for (nR, nS, dom) in [(1500, 3000, 500), (43000, 43100, 9000), (43000, 43100, 1600), (300000, 300000, 200000)]:
    w = W.duplicates(nR, nS, dom)
    R, S = W.to_numpy_tuples(w.R), W.to_numpy_tuples(w.S)
    for _ in range(3): v, n = eng.join_host_view(R, S)
    t0 = time.perf_counter()
    for _ in range(20): v, n = eng.join_host_view(R, S)
    dt = (time.perf_counter() - t0) / 20
    print(f"join_host {nR} x {nS} -> {n} pairs: {dt*1e3:.3f} ms per call, plan {eng.last_plan()}")
