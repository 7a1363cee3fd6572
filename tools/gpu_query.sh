#!/bin/bash
# device-resident query path: parity tests + wall time of the reference program with each drop-in level
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 -k "unique or device_query or dropin" > gpurun_out/query_tests.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/query_tests.log; tail -6 gpurun_out/query_tests.log
python - <<'PY' 2>&1 | tee gpurun_out/small_work_query_path.txt
import os, subprocess, tarfile, tempfile, time
root = os.getcwd()
with tempfile.TemporaryDirectory() as d:
    os.mkdir(d + "/small")
    tarfile.open("tests/golden/small_relations.tar.xz").extractall(d + "/small")
    data = open("tests/golden/small.init", "rb").read() + open("tests/golden/small.work", "rb").read()
    want = open("tests/golden/small.result", "rb").read()
    for b in ("join_b200_query", "join_b200_full", "join_b200"):
        for i in range(3):
            t0 = time.perf_counter()
            out = subprocess.run([f"{root}/radixhashjoin_b200/host/_build/{b}"], input=data, cwd=d, capture_output=True,
                                 env=dict(os.environ, RHJ_HOST_TIMING="1"), timeout=900)
            dt = time.perf_counter() - t0
            print(f"{b}: wall {dt:.3f} s, output identical: {out.stdout == want}")
            if i == 2:
                print(out.stderr.decode()[-1500:])
PY
