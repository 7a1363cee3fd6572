#!/bin/bash
# One gpurun call for the k_join_pos change: parity of everything the change touches, A/B/C bench lines (k_join_pos<3, PF>,
# k_join_pos<4>, the r02 kernel), one `ncu --set full` capture of the new kernel, the one-GPU HostResidentSteps check, then the
# rest of the GPU suite.  Every step has its own timeout; outputs land in gpurun_out/.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
# the whole script ends inside BUDGET seconds (default 395): a step gets min(its own limit, what is left), or is skipped
BUDGET=${BUDGET:-395}
lim() { local left=$(( BUDGET - ( $(date +%s) - T0 ) )); [ $left -lt 8 ] && left=1; [ $left -lt $1 ] && echo $left || echo $1; }
K='join or closed_form or baseline_config or shard or optimistic or sample_free or positional or pipelined_host'
timeout $(lim 200) python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K" > $O/pos_pytest_touched.log 2>&1; el "pytest (touched) exit $?"; tail -3 $O/pos_pytest_touched.log
B="--steps 10 --warmup 3 --no-e2e --no-cpu --no-small-work --no-target"
timeout $(lim 60) python bench.py $B > $O/pos_bench_lean3.json 2> $O/pos_bench_lean3.err; el "bench lean3 exit $?"
RHJ_JOIN_LEAN=0 timeout $(lim 60) python bench.py $B > $O/pos_bench_r02.json 2> $O/pos_bench_r02.err; el "bench r02 exit $?"
RHJ_JOIN_POS_ITEMS=4 timeout $(lim 60) python bench.py $B > $O/pos_bench_lean4.json 2> $O/pos_bench_lean4.err; el "bench lean4 exit $?"
python - <<'PY'
import json
for v in ("lean3", "r02", "lean4"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/pos_bench_{v}.json") if l.startswith("{")][-1])
        print(v, "ms/step %.3f" % d["ms_per_step"], "verified", d["verified"], d["phase_ms"], "launches", d["gpu_launches"])
    except Exception as ex:
        print(v, "no line:", ex)
PY
timeout $(lim 90) ncu --set full --clock-control none --import-source on -k regex:k_join_pos --launch-skip 4 --launch-count 1 -f -o $O/pos_join_full \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-small-work --no-target > $O/pos_ncu_full.log 2>&1; el "ncu full exit $?"
timeout $(lim 60) python tools/host_steps_check.py 25 > $O/pos_host_steps.json 2> $O/pos_host_steps.err; el "host steps exit $?"; cat $O/pos_host_steps.json
timeout $(lim 150) python -m pytest tests/ -x -q -m gpu -k "not ($K)" > $O/pos_pytest_rest.log 2>&1; el "pytest (rest) exit $?"; tail -3 $O/pos_pytest_rest.log
timeout $(lim 60) ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/pos_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-small-work --no-target > $O/pos_ncu_list.log 2>&1; el "ncu list exit $?"
ls -la $O | tail -15
