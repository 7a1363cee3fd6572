#!/bin/bash
# One GPU-box round on ONE GPU: the whole -m gpu suite, smoke, the default bench line (as the driver runs it) and the
# reference arm.  Everything lands in gpurun_out/.   usage: tools/gpu_round.sh [notests]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm --format=csv > gpurun_out/gpu.txt 2>&1
if [ "${1:-}" != "notests" ]; then
  echo "== pytest -m gpu"
  timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest.log 2>&1
  echo "pytest exit $?" | tee -a gpurun_out/pytest.log
  tail -8 gpurun_out/pytest.log
  echo "== smoke"
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
  tail -4 gpurun_out/smoke.log
fi
echo "== bench"
timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== reference arm (3 steps)"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
