#!/bin/bash
# multi-GPU bench: N ranks on one box over NCCL.  usage: tools/gpu_multi.sh N [extra bench args]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$1; shift
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/multi_$N.json 2> gpurun_out/multi_$N.err
echo "exit $?"
cat gpurun_out/multi_$N.json; tail -5 gpurun_out/multi_$N.err
