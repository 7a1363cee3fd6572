#!/bin/sh
# Regenerates tests/golden/small_joins.txt from the REFERENCE ITSELF (dev container only: needs
# /root/reference).  oracle/Makefile builds oracle/_ref/join_ref_trace = the unmodified reference
# program with a logging hook around Result::multiRadixHashJoin (oracle/ref_trace_*.cpp).
# One line per executed join of small.work (94 lines):
#     nR nS count in_digest out_sum out_xor
# Takes ~4 minutes (the reference's update_intermediate dominates).
set -e
here=$(cd "$(dirname "$0")" && pwd)
root=$(cd "$here/../.." && pwd)
make -s -C "$root/oracle" ref
cd "$root/oracle/_ref"
rm -f /tmp/rhj_trace.txt
cat small/small.init small/small.work | RHJ_TRACE_FILE=/tmp/rhj_trace.txt ./join_ref_trace > /tmp/rhj_trace_out.txt
diff /tmp/rhj_trace_out.txt small/small.result
sort -n /tmp/rhj_trace.txt > "$here/small_joins.txt"
cp small/small.init small/small.work small/small.result "$here/"
echo "wrote $here/small_joins.txt ($(wc -l < "$here/small_joins.txt") joins)"
