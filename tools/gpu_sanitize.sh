#!/bin/bash
# usage: tools/gpu_sanitize.sh memcheck|racecheck   -- ONE compute-sanitizer tool per GPU call (B200_PROFILING.md) on smoke() and
# on a small emulated pipelined-exchange step; with `memcheck` also the ThreadSanitizer run of the host program on edge.work.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOL=$1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_plain_smoke.log 2>&1 || { echo "plain smoke failed"; exit 1; }
python tools/pipe_emulated_step.py --world 4 --log2n 16 --chunks 2 --wire 12 --steps 2 > gpurun_out/san_plain_pipe.log 2>&1 || { echo "plain pipe failed"; exit 1; }
compute-sanitizer --tool $TOOL --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_${TOOL}_smoke.log 2>&1
echo "$TOOL smoke exit $?"; tail -4 gpurun_out/sanitizer_${TOOL}_smoke.log
compute-sanitizer --tool $TOOL --error-exitcode 9 python tools/pipe_emulated_step.py --world 4 --log2n 16 --chunks 2 --wire 12 --steps 2 > gpurun_out/sanitizer_${TOOL}_pipe.log 2>&1
echo "$TOOL pipe exit $?"; tail -4 gpurun_out/sanitizer_${TOOL}_pipe.log
if [ "$TOOL" = "memcheck" ] && [ -x radixhashjoin_b200/host/_build/join_b200_query_tsan ]; then
  W=$(mktemp -d); mkdir $W/edge; tar -xJf tests/golden/edge_relations.tar.xz -C $W/edge
  cat tests/golden/edge.init tests/golden/edge.work > $W/in.txt
  (cd $W && TSAN_OPTIONS="exitcode=0 second_deadlock_stack=1" $OLDPWD/radixhashjoin_b200/host/_build/join_b200_query_tsan < in.txt > out.txt 2> tsan.log)
  echo "tsan exit $?"; cmp $W/out.txt tests/golden/edge.result && echo "tsan run: output identical to edge.result"
  cp $W/tsan.log gpurun_out/tsan_join_b200_query_edge.log; grep -c "WARNING: ThreadSanitizer" gpurun_out/tsan_join_b200_query_edge.log; head -40 gpurun_out/tsan_join_b200_query_edge.log
fi
