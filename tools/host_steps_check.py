"""One-GPU check of HostResidentSteps (bench.py's e2e at N > 1) with the single-GPU join as the step: pinned allocation,
H2D of both shards, join, D2H of the pairs, host-side result verified by count + digest.  The multi-rank agreement logic is
covered on CPU by tests/test_distributed_cpu.py (gloo)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from radixhashjoin_b200 import RadixHashJoin
from radixhashjoin_b200 import workloads as W
from radixhashjoin_b200.distributed import HostResidentSteps

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
dev = "cuda:0"
eng = RadixHashJoin(0)
w = W.uniform_unique(log2n, dev)
out = torch.empty((w.S.shape[0], 2), dtype=torch.int64, device=dev)
eng.reserve(w.R.shape[0], w.S.shape[0])
R, S = w.R.clone(), w.S.clone()
hs = HostResidentSteps(lambda: eng.join_device(R, S, out=out), R, S, out, 1)
assert hs.ok, hs.why
R.zero_()
S.zero_()                      # every step must bring the shards back from the host
dt, cnt, d2h = hs.run(3, warmup=1)
back = hs.hout[:cnt].to(dev)
ok = (cnt,) + tuple(eng.pairs_digest(back)[1:]) == tuple(w.expected)
print(json.dumps({"host_resident_steps": "ok" if ok else "WRONG RESULT", "log2n": log2n, "ms_per_step": dt * 1e3, "count": cnt,
                  "h2d_bytes": hs.h2d_bytes, "d2h_bytes": d2h, "pinned": bool(hs.hR.is_pinned())}))
sys.exit(0 if ok else 1)
