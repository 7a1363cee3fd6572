#!/bin/bash
# Pipelined exchange on N GPUs: emulated-rank parity tests first, then bench (pipe vs dma) and the copy-kernel microbench.
# usage: tools/gpu_pipe.sh N [tests|notests]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$1
if [ "${2:-tests}" = "tests" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 -k "pipe or dma_shard_join_emulated" > gpurun_out/pipe_tests.log 2>&1
  rc=$?; echo "pytest exit $rc" | tee -a gpurun_out/pipe_tests.log; tail -12 gpurun_out/pipe_tests.log
  [ $rc -ne 0 ] && exit 1
fi
run() { # name, args...
  name=$1; shift
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@" \
    > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "== $name exit $?"; cat gpurun_out/$name.json; tail -4 gpurun_out/$name.err | cut -c1-400
}
run p_${N}_pipe bench.py --gpus $N --steps 10 --warmup 3
run p_${N}_a2a tools/a2a_bench.py
run p_${N}_dma bench.py --gpus $N --steps 10 --warmup 3 --shuffle dma
