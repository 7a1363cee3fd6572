#!/bin/bash
# ncu evidence for one bench step: launch list (device time per launch) + full capture of the
# partition and join kernels.  Outputs in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-small-work ${BENCH_ARGS:-}"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_" -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_join|k_scatter|k_hist" -s ${NCU_SKIP:-15} -c ${NCU_COUNT:-5} -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/
