#!/bin/bash
# small.work wall time (BASELINE.json metric, config 1): the reference program with our drop-in
# translation units vs the unmodified reference program, on the GPU box's host.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/sw && rm -rf /tmp/sw/small && mkdir -p /tmp/sw/small
tar xJf tests/golden/small_relations.tar.xz -C /tmp/sw/small
cp tests/golden/small.init tests/golden/small.work tests/golden/small.result /tmp/sw/small/
H=$PWD/radixhashjoin_b200/host/_build
cd /tmp/sw
nproc > $OLDPWD/gpurun_out/small_work.txt; grep -m1 "model name" /proc/cpuinfo >> $OLDPWD/gpurun_out/small_work.txt
for bin in join_b200_full join_b200_full join_b200; do
  s=$(date +%s%N)
  cat small/small.init small/small.work | timeout 900 $H/$bin > /tmp/sw/out_$bin.txt 2> /tmp/sw/err_$bin.txt
  rc=$?
  e=$(date +%s%N)
  if diff -q /tmp/sw/out_$bin.txt small/small.result > /dev/null; then ok=IDENTICAL; else ok=DIFFERENT; fi
  echo "$bin rc=$rc wall=$(( (e - s) / 1000000 )) ms output=$ok" | tee -a $OLDPWD/gpurun_out/small_work.txt
  tail -2 /tmp/sw/err_$bin.txt
done
if [ "${1:-}" = "ref" ]; then
  R=$OLDPWD/oracle/_ref/join_ref
  s=$(date +%s%N)
  cat small/small.init small/small.work | timeout 1200 $R > /tmp/sw/out_ref.txt
  e=$(date +%s%N)
  if diff -q /tmp/sw/out_ref.txt small/small.result > /dev/null; then ok=IDENTICAL; else ok=DIFFERENT; fi
  echo "join_ref (unmodified reference, -Ofast -march=x86-64-v3) wall=$(( (e - s) / 1000000 )) ms output=$ok" | tee -a $OLDPWD/gpurun_out/small_work.txt
fi
