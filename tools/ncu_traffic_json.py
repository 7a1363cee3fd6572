#!/usr/bin/env python
"""profiles/r02_final_ncu_traffic.json from an `ncu --set full` report of one join step (run here, no GPU needed):
    python tools/ncu_traffic_json.py gpurun_out/n1_full.ncu-rep <commit> > profiles/r02_final_ncu_traffic.json"""
import csv
import json
import subprocess
import sys

rep, commit = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]


def col(name):
    return hdr.index(name)


def to_bytes(v, unit):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def to_ms(v, unit):
    v = float(v)
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}[unit]


phases = {}
names = {0: "scatter1", 1: "scatter2", 2: "join"}
for i, r in enumerate(rows[2:5]):
    rd, wr, t = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum")
    b_r, b_w = to_bytes(r[rd], units[rd]), to_bytes(r[wr], units[wr])
    phases[names[i]] = {"kernel": r[col("Kernel Name")], "dram_bytes_read": b_r, "dram_bytes_write": b_w,
                        "ncu_time_ms": to_ms(r[t], units[t]), "traffic_bytes": b_r + b_w,
                        "l1tex_pct": float(r[col("l1tex__throughput.avg.pct_of_peak_sustained_active")]),
                        "dram_pct": float(r[col("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")]),
                        "warps_active_pct": float(r[col("sm__warps_active.avg.pct_of_peak_sustained_active")]),
                        "registers": int(float(r[col("launch__registers_per_thread")])),
                        "inst_executed": float(r[col("smsp__inst_executed.sum")])}
print(json.dumps({"workload": "uniform_unique_2^27x2^27", "emitter": "fused", "build": commit,
                  "source": "ncu --set full --clock-control none of one join step of `bench.py --steps 2 --warmup 3` "
                            "(tools/gpu_ncu_n1.sh); a snapshot of that build, not of the running one",
                  "phases": phases}, indent=1))
