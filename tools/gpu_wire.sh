#!/bin/bash
# usage: tools/gpu_wire.sh N  -- 12-byte vs 16-byte wire format of the pipelined exchange (tests first)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$1
if [ "${2:-tests}" = "tests" ]; then
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 -k "pipe" > gpurun_out/pipe_tests.log 2>&1
rc=$?; echo "pytest exit $rc" | tee -a gpurun_out/pipe_tests.log; tail -5 gpurun_out/pipe_tests.log
[ $rc -ne 0 ] && exit 1
fi
run() { name=$1; shift
  timeout ${TMO:-200} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@" \
    > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "== $name exit $?"; grep '^{' gpurun_out/$name.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    print(d['config']['workload'], 'ms', round(d['ms_per_step'],3), 'verified', d['verified'], 'tuples/s %.3e' % d['value'], d.get('nvlink') and round(d['nvlink']['achieved_GBps_per_direction']))
    print(' '.join(f'{k}={v}' for k,v in (d.get('shard_timeline_ms') or [])))
"; grep -v "^\*\*\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/$name.err | tail -4 | cut -c1-300
}
run w${N}_12 bench.py --gpus $N --steps 10 --warmup 3 --wire-bytes 12
run w${N}_16 bench.py --gpus $N --steps 10 --warmup 3 --wire-bytes 16
for extra in "$@"; do :; done
