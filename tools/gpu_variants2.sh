#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-small-work ${BENCH_ARGS:-} > gpurun_out/v_$name.json 2> gpurun_out/v_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/v_{n}.json").read().strip().splitlines()[-1])
    print(f"{n:14s} ms/step {d['ms_per_step']:.3f} verified {d['verified']} frac {d['step_roofline']['frac_of_measured_hbm']:.3f} phases {d['phase_ms']}")
except Exception as e:
    print(n, "FAILED", e); print(open(f"gpurun_out/v_{n}.err").read()[-800:])
PY
}
L=$PWD/radixhashjoin_b200
run v2 X=1
run v1 RHJ_JOIN_V=1
run v1early RHJ_JOIN_V=11
run v2_8k RHJ_LIB=$L/librhj_j8k.so
run v2_nopf RHJ_LIB=$L/librhj_jnp.so
run v2_8k_nopf RHJ_LIB=$L/librhj_j8knp.so
BENCH_ARGS="--emit count_then_write" run v1early_ctw RHJ_JOIN_V=11
BENCH_ARGS="--emit count_then_write" run v2_8k_ctw RHJ_LIB=$L/librhj_j8k.so
