#!/bin/bash
# A/B of the 56-register scatter build (default) against the 64-register one, single GPU and pipelined N-GPU
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$1
one() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-small-work > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "== $name exit $?"; python - <<PY
import json
d=json.load(open("gpurun_out/$name.json"))
print(d["ms_per_step"], d["verified"], d["phase_ms"])
PY
}
one g1_r56 RHJ_X=1
one g1_r64 RHJ_LIB=$PWD/radixhashjoin_b200/librhj_r64.so
run() { name=$1; shift
  env $ENVV timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@" \
    > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "== $name exit $?"; python - <<PY
import json
d=json.load(open("gpurun_out/$name.json"))
print(d["ms_per_step"], d["verified"]); print(d["shard_timeline_ms"])
PY
}
ENVV="RHJ_X=1" run g${N}_r56_48 bench.py --gpus $N --steps 10 --warmup 3
ENVV="RHJ_PIPE_SHIP_CTAS=96 RHJ_PIPE_STAGES=4" run g${N}_r56_96 bench.py --gpus $N --steps 10 --warmup 3
ENVV="RHJ_LIB=$PWD/radixhashjoin_b200/librhj_r64.so RHJ_PIPE_SHIP_CTAS=96 RHJ_PIPE_STAGES=4" run g${N}_r64_96 bench.py --gpus $N --steps 10 --warmup 3
