#!/bin/bash
# usage: tools/gpu_multi_var.sh N name [ENV=VAL ...] -- prints a one-line summary
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$1; name=$2; shift 2
env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/m_${N}_$name.json 2> gpurun_out/m_${N}_$name.err
python - "$N" "$name" <<'PY'
import json,sys
N,n=sys.argv[1],sys.argv[2]
try:
    d=json.loads([l for l in open(f"gpurun_out/m_{N}_{n}.json").read().splitlines() if l.startswith("{")][-1])
    print(f"N={N} {n:14s} ms/step {d['ms_per_step']:.3f} value {d['value']:.3e} verified {d['verified']} bits {d['config']['radix_bits']} phases {d['phase_ms']}")
    if d.get('shard_timeline_ms'): print("   timeline", d['shard_timeline_ms'])
except Exception as e:
    print(N, n, "FAILED", e); print(open(f"gpurun_out/m_{N}_{n}.err").read()[-1500:])
PY
