#!/bin/bash
# usage: tools/gpu_pipe_sweep.sh N  -- copy-kernel parameter sweep + chunk-count sweep of the pipelined exchange
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$1
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 -k "pipe" > gpurun_out/pipe_tests.log 2>&1
rc=$?; echo "pytest exit $rc" | tee -a gpurun_out/pipe_tests.log; tail -5 gpurun_out/pipe_tests.log
[ $rc -ne 0 ] && exit 1
run() { # name, args...
  name=$1; shift
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@" \
    > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "== $name exit $?"; cat gpurun_out/$name.json | cut -c1-2500; tail -3 gpurun_out/$name.err | cut -c1-300
}
run s_${N}_a2a tools/a2a_bench.py --log2n 26 --sweep "${SWEEP:-48:8:8,96:8:8,144:8:8,48:16:4,48:4:16,24:8:16,24:16:8,96:8:4}"
for c in ${CHUNKS:-4 2 8}; do
  run s_${N}_pipe_c$c bench.py --gpus $N --steps 10 --warmup 3 --chunks $c
done
