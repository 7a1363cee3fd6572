#!/bin/bash
# N-rank bench with 12 and 16 bytes per tuple on the wire.  usage: tools/gpu_compact8.sh N "12 16"
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$1
for b in $2; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
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
