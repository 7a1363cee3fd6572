#!/bin/bash
# usage: tools/gpu_multi_all.sh N [variants]  -- the N-GPU lines of the round: uniform (pipelined exchange, the driver's default
# command), copy-kernel microbench, BASELINE config 3 (fk 2^24 x 2^30, strong scaling), config 4 shape (Zipf, sharded), and at
# N = 8 config 5 (2^31 x 2^31).  Lines land in gpurun_out/m<N>_*.json.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$1
run() { name=$1; shift
  env ${ENVV:-X=1} timeout ${TMO:-240} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@" \
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
if [ "$N" = "1" ]; then
  for a in "--workload fk --fk-probe-log2 30 --log2n 30" "--workload zipf --log2n 28"; do
    n=$(echo $a | cut -d' ' -f2)
    timeout 400 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-small-work --no-target $a > gpurun_out/m1_$n.json 2> gpurun_out/m1_$n.err
    echo "== m1_$n exit $?"; python -c "
import json; d=json.load(open('gpurun_out/m1_$n.json')); print(d['config']['workload'], 'ms', round(d['ms_per_step'],3), 'verified', d['verified'], 'tuples/s %.3e' % d['value'], d['phase_ms'])"
  done
  exit 0
fi
run m${N}_pipe bench.py --gpus $N --steps 20 --warmup 5
if [ "${2:-}" = "variants" ]; then
  run m${N}_pipe_c6 bench.py --gpus $N --steps 10 --warmup 3 --chunks 6
  ENVV="RHJ_PIPE_SHIP_CTAS=64" run m${N}_pipe_cta64 bench.py --gpus $N --steps 10 --warmup 3
fi
run m${N}_fk bench.py --gpus $N --steps 5 --warmup 3 --workload fk --fk-probe-log2 30
run m${N}_zipf bench.py --gpus $N --steps 5 --warmup 3 --workload zipf --log2n 25
if [ "$N" = "8" ]; then
  run m${N}_cfg5 bench.py --gpus $N --steps 5 --warmup 3 --log2n 28
fi
