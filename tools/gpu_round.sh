#!/bin/bash
# One GPU-box round: parity tests, smoke, bench (+ variants).  Everything lands in gpurun_out/.
# usage: tools/gpu_round.sh [quick]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" 
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest.log
tail -15 gpurun_out/pytest.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "${1:-}" != "quick" ]; then
  for v in "RHJ_NO_OPT2=1" "RHJ_NO_OPT=1" "RHJ_SCATTER_MODE=1"; do
    echo "== bench $v"
    env $v timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-small-work > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
    cat gpurun_out/bench_$v.json; tail -3 gpurun_out/bench_$v.err
  done
  echo "== bench count_then_write"
  timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-small-work --emit count_then_write > gpurun_out/bench_ctw.json 2> gpurun_out/bench_ctw.err
  cat gpurun_out/bench_ctw.json
  echo "== bench 2^28"
  timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-small-work --log2n 28 > gpurun_out/bench_28.json 2> gpurun_out/bench_28.err
  cat gpurun_out/bench_28.json; tail -3 gpurun_out/bench_28.err
fi
