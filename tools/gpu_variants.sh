#!/bin/bash
# Tuning round: parity tests on the current build, then bench phase times for build/env variants.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
run() {  # name, env...
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
run base X=1
run direct RHJ_SCATTER_MODE=2
run bulk RHJ_SCATTER_MODE=1
run t256 RHJ_LIB=$PWD/radixhashjoin_b200/librhj_t256.so
run t256x16 RHJ_LIB=$PWD/radixhashjoin_b200/librhj_t256x16.so
run t512x4 RHJ_LIB=$PWD/radixhashjoin_b200/librhj_t512x4.so
run t256_direct RHJ_LIB=$PWD/radixhashjoin_b200/librhj_t256.so RHJ_SCATTER_MODE=2
BENCH_ARGS="--emit count_then_write" run base_ctw X=1
BENCH_ARGS="--log2n 28" run base_28 X=1
BENCH_ARGS="--workload zipf" run base_zipf X=1
BENCH_ARGS="--workload fk --fk-build-log2 24 --log2n 28" run base_fk X=1
