#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$1
run() { name=$1; shift
  env ${ENVV:-X=1} timeout ${TMO:-200} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@" \
    > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "== $name exit $?"; grep '^{' gpurun_out/$name.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    print(d['config']['workload'], 'ms', round(d['ms_per_step'],3), 'verified', d['verified'], 'tuples/s %.3e' % d['value'], d.get('nvlink') and round(d['nvlink']['achieved_GBps_per_direction']))
    print(' '.join(f'{k}={v}' for k,v in (d.get('shard_timeline_ms') or [])))
"; grep -v "^\*\*\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/$name.err | tail -4 | cut -c1-300
}
ENVV="RHJ_PIPE_SHIP_CTAS=128" run w${N}_12_128 bench.py --gpus $N --steps 10 --warmup 3 --wire-bytes 12
ENVV="RHJ_PIPE_SHIP_CTAS=96" run w${N}_12_96 bench.py --gpus $N --steps 10 --warmup 3 --wire-bytes 12
ENVV="RHJ_PIPE_SHIP_CTAS=148" run w${N}_12_148 bench.py --gpus $N --steps 10 --warmup 3 --wire-bytes 12
