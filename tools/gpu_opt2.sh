#!/bin/bash
# optimistic pass 2: targeted tests, then A/B against RHJ_NO_OPT2=1 at 2^27 and 2^28
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "optimistic or join_equals_oracle or two_pass or closed_form or zipf or pipelined" > gpurun_out/opt2_tests.log 2>&1
echo "tests exit $?"; tail -5 gpurun_out/opt2_tests.log
tools/gpu_quick.sh "opt2_27 X=1" "exact2_27 RHJ_NO_OPT2=1"
BENCH_ARGS="--log2n 28" tools/gpu_quick.sh "opt2_28 X=1" "exact2_28 RHJ_NO_OPT2=1"
BENCH_ARGS="--emit count_then_write" tools/gpu_quick.sh "opt2_ctw X=1"
