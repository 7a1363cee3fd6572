#!/bin/bash
# quick A/B of env variants on the single-GPU bench: tools/gpu_quick.sh "NAME ENV=VAL ..." ...
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for spec in "$@"; do
  set -- $spec; name=$1; shift
  env "$@" X=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-small-work ${BENCH_ARGS:-} > gpurun_out/q_$name.json 2> gpurun_out/q_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/q_{n}.json").read().strip().splitlines()[-1])
    print(f"{n:12s} ms/step {d['ms_per_step']:.3f} verified {d['verified']} frac {d['step_roofline']['frac_of_measured_hbm']:.3f} phases {d['phase_ms']}")
except Exception as e:
    print(n, "FAILED", e); print(open(f"gpurun_out/q_{n}.err").read()[-800:])
PY
done
