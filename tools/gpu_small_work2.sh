#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/sw && rm -rf /tmp/sw/small && mkdir -p /tmp/sw/small
tar xJf tests/golden/small_relations.tar.xz -C /tmp/sw/small
cp tests/golden/small.init tests/golden/small.work tests/golden/small.result /tmp/sw/small/
H=$PWD/radixhashjoin_b200/host/_build
O=$PWD/gpurun_out
cd /tmp/sw
for i in 1 2 3; do
  s=$(date +%s%N)
  cat small/small.init small/small.work | RHJ_HOST_TIMING=1 timeout 900 $H/join_b200_full > out.txt 2> err.txt
  e=$(date +%s%N)
  if diff -q out.txt small/small.result > /dev/null; then ok=IDENTICAL; else ok=DIFFERENT; fi
  echo "join_b200_full run $i wall=$(( (e - s) / 1000000 )) ms output=$ok" | tee -a $O/small_work2.txt
  cat err.txt | tee -a $O/small_work2.txt
done
