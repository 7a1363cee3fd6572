#!/bin/bash
# Second (and last) gpurun call for k_join_pos: bench lines of its variants (RHJ_JOIN_POS_V = 0: the kernel of call 95, 1: hash kept
# in a register, 3: 1 + next item claimed early), then the join-related GPU tests with the fastest verified variant as the default,
# then one `ncu --set full` capture of it.  Ends inside BUDGET seconds.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
BUDGET=${BUDGET:-150}
lim() { local left=$(( BUDGET - ( $(date +%s) - T0 ) )); [ $left -lt 8 ] && left=1; [ $left -lt $1 ] && echo $left || echo $1; }
B="--steps 10 --warmup 3 --no-e2e --no-cpu --no-small-work --no-target"
for v in 0 1 3; do
    RHJ_JOIN_POS_V=$v timeout $(lim 80) python bench.py $B > $O/pos2_bench_v$v.json 2> $O/pos2_bench_v$v.err; el "bench V=$v exit $?"
done
best=$(python - <<'PY'
import json
best, bt = 0, 1e9
for v in (0, 1, 3):
    try:
        d = json.loads([l for l in open(f"gpurun_out/pos2_bench_v{v}.json") if l.startswith("{")][-1])
        print(f"# V={v} ms/step {d['ms_per_step']:.3f} join {d['phase_ms'].get('join')} verified {d['verified']}", file=__import__('sys').stderr)
        if d["verified"] and d["phase_ms"]["join"] < bt:
            best, bt = v, d["phase_ms"]["join"]
    except Exception as ex:
        print(f"# V={v} no line: {ex}", file=__import__('sys').stderr)
print(best)
PY
)
el "fastest verified variant: V=$best"
echo $best > $O/pos2_best.txt
RHJ_JOIN_POS_V=$best timeout $(lim 60) python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k positional > $O/pos2_pytest_positional.log 2>&1; el "pytest positional (all variants) exit $?"; tail -2 $O/pos2_pytest_positional.log
K='join_equals_oracle or u64_extremes or overflow_single or two_pass or zipf_probe or needed_capacity or closed_form or pipe_shard_join_emulated'
RHJ_JOIN_POS_V=$best timeout $(lim 90) python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K" > $O/pos2_pytest.log 2>&1; el "pytest joins (V=$best default) exit $?"; tail -2 $O/pos2_pytest.log
RHJ_JOIN_POS_V=$best timeout $(lim 60) ncu --set full --clock-control none --import-source on -k regex:k_join_pos --launch-skip 4 --launch-count 1 -f -o $O/pos2_join_full \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-small-work --no-target > $O/pos2_ncu_full.log 2>&1; el "ncu full exit $?"
ls -la $O | grep pos2
