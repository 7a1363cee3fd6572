#!/bin/bash
# ncu evidence for the sharded (pipelined exchange) kernels, all ranks emulated on ONE GPU (tools/pipe_emulated_step.py)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python tools/pipe_emulated_step.py --world 8 --log2n ${LOG2N:-26} --wire ${WIRE:-12} --steps 1"
python tools/pipe_emulated_step.py --world 8 --log2n ${LOG2N:-26} --wire ${WIRE:-12} --steps 2 > gpurun_out/pipe_emu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/pipe_emu_plain.log; exit 1; }
cat gpurun_out/pipe_emu_plain.log
$CMD > gpurun_out/pipe_emu_plain1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/pipe_emu_launches.csv $CMD > gpurun_out/pipe_emu_ncu1.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:k_scatter -c 2 -o gpurun_out/pipe_p1 -f $CMD > gpurun_out/pipe_emu_ncu2.log 2>&1; echo "p1 exit $?"
ncu --set full --clock-control none --import-source on -k regex:k_scatter -s 64 -c 2 -o gpurun_out/pipe_p2 -f $CMD > gpurun_out/pipe_emu_ncu3.log 2>&1; echo "p2 exit $?"
ncu --set full --clock-control none --import-source on -k regex:k_pipe_ship -c 2 -o gpurun_out/pipe_ship -f $CMD > gpurun_out/pipe_emu_ncu4.log 2>&1; echo "ship exit $?"
ls -la gpurun_out/*.ncu-rep
