"""bench.py contract checks that need no GPU: the reference arm prints one well-formed JSON line, the B200 arm refuses
to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_rhj.so")


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True, timeout=timeout)


def test_reference_arm_prints_the_contract_line():
    out = _run("--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-log2n", "16")
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    lines = [l for l in out.stdout.decode().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "radixhashjoin_input_tuples_per_s" and d["unit"] == "tuples/s"
    assert d["steps"] == 2 and d["n_gpus"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == ("reference" if os.path.exists(REF_SO) else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["config"]["workload"] == "uniform_unique_2^27x2^27"


def test_reference_arm_reports_the_faster_of_its_builds():
    """on a host with more than 8 CPUs (forced here) the 16-worker build of the same sources is timed too and the faster one
    is the value; both figures are in the line"""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_rhj_t16.so")):
        return
    env = dict(os.environ, RHJ_REF_ALL_BUILDS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--ref-log2n", "18"], cwd=ROOT, capture_output=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    d = json.loads([l for l in out.stdout.decode().splitlines() if l.startswith("{")][0])
    b = d["cpu_baseline"]["builds_tuples_per_s"]
    assert set(b) == {"NUM_OF_THREADS=8", "NUM_OF_THREADS=16"}
    assert d["cpu_baseline"]["cores"] in (8, 16) and d["value"] == max(b.values()) == b[f"NUM_OF_THREADS={d['cpu_baseline']['cores']}"]


def test_b200_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return          # on the GPU box the arm is exercised by the driver itself
    out = _run("--steps", "1", "--warmup", "1", "--no-e2e", "--no-cpu", "--no-small-work", timeout=300)
    assert out.returncode != 0
    assert b"no CPU fallback" in out.stderr or b"CUDA" in out.stderr
