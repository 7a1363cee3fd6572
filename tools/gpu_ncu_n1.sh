#!/bin/bash
# ncu evidence of the final single-GPU build: launch list of one join step + `--set full` of the three bandwidth kernels
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-small-work --no-target"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_" -c 200 --csv --log-file gpurun_out/n1_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_join|k_scatter" -s 9 -c 3 -o gpurun_out/n1_full -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "set full exit $?"; ls -la gpurun_out/n1_full.ncu-rep
