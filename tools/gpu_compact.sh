#!/bin/bash
# 12-byte shipping check: emulated-rank tests, single-GPU regression check, then N-rank bench with 16 and 12 bytes on the wire.
# usage: tools/gpu_compact.sh N
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "dma_shard or fused_shard" > gpurun_out/compact_tests.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/compact_tests.log
tools/gpu_quick.sh "single X=1"
for b in 16 12 12 16; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 --ship-bytes $b --no-e2e --no-cpu --no-small-work > gpurun_out/ship${b}_$N.json 2> gpurun_out/ship${b}_$N.err
  echo "ship $b exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ship${b}_$N.json").read().strip().splitlines()[-1])
    print("ship $b ms/step", d["ms_per_step"], "value", d["value"], "verified", d["verified"], d.get("shard_timeline_ms"))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/ship${b}_$N.err").read()[-1500:])
PY
done
