#!/bin/bash
# usage: tools/gpu_multi_all.sh N [quick]  -- the N-GPU lines of the round: uniform (pipe), copy-kernel microbench, fk, zipf, config 5
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$1
run() { name=$1; shift
  timeout ${TMO:-200} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@" \
    > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "== $name exit $?"; grep '^{' gpurun_out/$name.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    if 'bench' in d: print(d['ship_ctas'], d['stage_kb'], d['stages'], round(d['GBps_per_gpu_per_direction'],1), 'GB/s', d['ms_all']); continue
    print(d['config']['workload'], 'ms', round(d['ms_per_step'],3), 'verified', d['verified'], 'tuples/s %.3e' % d['value'], d.get('nvlink') and round(d['nvlink']['achieved_GBps_per_direction']))
    print(' '.join(f'{k}={v}' for k,v in (d.get('shard_timeline_ms') or [])))
"; grep -v "^\*\*\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/$name.err | tail -4 | cut -c1-300
}
run m${N}_pipe bench.py --gpus $N --steps 10 --warmup 3
run m${N}_a2a tools/a2a_bench.py --log2n 26 --sweep "48:8:8,96:8:4"
run m${N}_fk bench.py --gpus $N --steps 5 --warmup 3 --workload fk --fk-probe-log2 30
run m${N}_zipf bench.py --gpus $N --steps 5 --warmup 3 --workload zipf --log2n 25
if [ "${2:-}" != "quick" ]; then
  run m${N}_pipe_c8 bench.py --gpus $N --steps 10 --warmup 3 --chunks 8
  run m${N}_cfg5 bench.py --gpus $N --steps 5 --warmup 3 --log2n 28
fi
