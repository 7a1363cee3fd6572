#!/bin/bash
# single-GPU A/B of alternative builds of librhj.so: usage tools/gpu_ab.sh [tests] name=path ...   (name "default" = the in-tree library)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ "${1:-}" = "tests" ]; then
  shift
  timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 -k "${TESTS:-join}" > gpurun_out/ab_tests.log 2>&1
  rc=$?; echo "pytest exit $rc" | tee -a gpurun_out/ab_tests.log; tail -5 gpurun_out/ab_tests.log
  [ $rc -ne 0 ] && exit 1
fi
for spec in "$@"; do
  name=${spec%%=*}; lib=${spec#*=}
  for wl in ${WORKLOADS:-uniform}; do
    if [ "$lib" = "default" ]; then e="RHJ_X=1"; else e="RHJ_LIB=$PWD/$lib"; fi
    env $e timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-small-work --no-target --workload $wl ${BENCH_ARGS:-} > gpurun_out/ab_${name}_$wl.json 2> gpurun_out/ab_${name}_$wl.err
    echo "== $name $wl exit $?"; python - <<PY
import json
d=json.load(open("gpurun_out/ab_${name}_$wl.json"))
print(round(d["ms_per_step"],4), d["verified"], d["phase_ms"])
PY
  done
done
