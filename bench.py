#!/usr/bin/env python
"""bench.py -- radixHashJoin input tuples/s on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps 20 --warmup 5          # our arm (CUDA, librhj.so)
    python bench.py --impl reference --steps 5 --warmup 3   # the reference's CPU path (oracle/_ref)
    torchrun --nproc-per-node N ... bench.py --gpus N ...   # N > 1: one rank per GPU, NCCL

A "step" is one complete radix hash join (histogram -> prefix sum -> partition scatter (x2) ->
build/probe -> emit) of the workload; inputs are resident in HBM before the timed region, the
result stays in HBM.  N = 1 runs BASELINE.json configs[1] (2^27 x 2^27 unique uniform u64 keys).
N > 1 is weak scaling: every rank owns 2^27 + 2^27 tuples of a global 2^(27+log2 N) pair of relations;
a step = the pipelined exchange (histogram-free chunked pass 1 on (rank | sub-digit), our TMA copy kernel over NVLink,
appended pass 2, one join; radixhashjoin_b200/distributed.py).  --shuffle dma|nccl select the exact exchange and the
NCCL all-to-all baseline; --workload fk runs config 3 (broadcast build side, strong scaling).

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "radixhashjoin_input_tuples_per_s"
UNIT = "tuples/s"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ------------------------------------------------------------------------------------------------
# clocks: sample SM clock + throttle reasons DURING the timed region (B200_PROFILING.md)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nvml = None

    def _loop(self):
        nv = self._nvml
        names = {}
        for n in ("HwSlowdown", "HwThermalSlowdown", "SwThermalSlowdown", "SwPowerCap", "HwPowerBrakeSlowdown"):
            v = getattr(nv, "nvmlClocksThrottleReason" + n, None) or getattr(nv, "nvmlClocksEventReason" + n, None)
            if v is not None:
                names[v] = n
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, n in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self._nvml is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        s = sorted(self.samples)
        snake = {"HwSlowdown": "hw_slowdown", "HwThermalSlowdown": "hw_thermal_slowdown",
                 "SwThermalSlowdown": "sw_thermal_slowdown", "SwPowerCap": "sw_power_cap",
                 "HwPowerBrakeSlowdown": "hw_power_brake_slowdown"}
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(snake[r] for r in self.reasons), "samples": len(s)}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


# ------------------------------------------------------------------------------------------------
# the reference arm / cpu_baseline: the reference's own multiRadixHashJoin on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(log2n, steps, warmup, budget_s=None):
    """Times oracle/_ref (the unmodified reference) -- or, if that build is absent, the single-threaded C port -- on the
    2^log2n x 2^log2n uniform workload.  The reference hard-wires NUM_OF_THREADS = 8 workers per join (JobScheduler.h:11); on a
    host with more CPUs the same sources built with 16 workers (oracle/Makefile: libref_rhj_t16.so) are timed as well and the
    FASTER build is the one reported, so that the baseline uses the host threads it can use.
    With budget_s the loop stops early (after at least one timed step per build) once that many seconds of joins have run, so
    that a 6-second-per-join configuration still ends within a few minutes.
    Returns (tuples_per_s, seconds_per_step, info, timed_steps)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as O
    from radixhashjoin_b200 import workloads as W
    w = W.uniform_unique(log2n, "cpu")
    R, S = W.to_numpy_tuples(w.R), W.to_numpy_tuples(w.S)
    n_in = len(R) + len(S)

    def timed(one, budget, warm):
        times, spent = [], 0.0
        for i in range(warm + steps):
            sec = one()
            spent += sec
            if i >= warm:
                times.append(sec)
            elif budget is not None and spent > budget / 4:
                warm = i + 1     # the warm-up already used its share: the next runs are timed ones
            if budget is not None and times and spent > budget:
                break
        return sum(times) / len(times), len(times)

    if O.have_ref():
        # RHJ_REF_ALL_BUILDS=1 times the 16-worker build on a small host too (tests)
        many = (os.cpu_count() or 1) > 8 or os.environ.get("RHJ_REF_ALL_BUILDS") == "1"
        builds = [8] + ([16] if O.have_ref(16) and many else [])
        res = {}
        for t in builds:
            def one(t=t):
                cnt, sec = O.reference_join(R, S, want_pairs=False, threads=t)
                assert cnt == len(S)
                return sec
            res[t] = timed(one, None if budget_s is None else budget_s / len(builds), warmup)
        best = min(res, key=lambda t: res[t][0])
        sec, ntimed = res[best]
        kind, cores = "reference", best
        extra = {"builds_tuples_per_s": {f"NUM_OF_THREADS={t}": n_in / res[t][0] for t in builds},
                 "build": f"unmodified sources, NUM_OF_THREADS = {best}" + (" (as shipped)" if best == 8 else
                          " (JobScheduler.h:11 overridden at build time, oracle/Makefile; the as-shipped 8-worker build was slower)")}
    else:
        def one():
            t0 = time.perf_counter()
            p = O.oracle_join(R, S)
            sec = time.perf_counter() - t0
            assert len(p) == len(S)
            return sec
        sec, ntimed = timed(one, budget_s, warmup)
        kind, cores, extra = "port", 1, {}
    info = dict({"kind": kind, "cores": cores, "host_cpus": os.cpu_count(),
                 "sample": f"uniform_unique 2^{log2n} x 2^{log2n} (same generator as the GPU workload), "
                           f"{ntimed} timed run(s) of Result::multiRadixHashJoin alone, inputs in host RAM"}, **extra)
    return n_in / sec, sec, info, ntimed


def small_work_wall(with_reference=False):
    """BASELINE.json config 1 / metric part 2: wall time of `cat small.init small.work | ./join` for the reference PROGRAM
    with the CUDA path dropped in (binaries built in the dev container by radixhashjoin_b200/host/Makefile), output
    diffed against small.result:
      join_b200_query  Query::execute -> rhj_query_execute: the whole query path device resident (the product);
      join_b200_full   Result.cpp + intermediate.cpp replaced, filters / relations / intermediates on the host (round 1).
    A process pays ~1.5-2.5 s of CUDA context creation before its first query; the steady-state figures are the program's own
    clock from the last context creation to the last call (`query_phase_*`, RHJ_HOST_TIMING) and the extra wall time per
    additional pass when the workload runs three times inside ONE process (noisy: context creation varies by +-0.5 s).  With --small-work-ref the
    unmodified reference program (oracle/_ref/join_ref) is timed the same way (takes minutes)."""
    import re
    import subprocess
    import tarfile
    import tempfile
    host = os.path.join(ROOT, "radixhashjoin_b200", "host", "_build")
    if not os.path.exists(os.path.join(host, "join_b200_query")):
        return {"unavailable": "radixhashjoin_b200/host/_build/join_b200_query not built (needs the reference sources)"}
    golden = os.path.join(ROOT, "tests", "golden")
    with tempfile.TemporaryDirectory() as d:
        os.mkdir(os.path.join(d, "small"))
        with tarfile.open(os.path.join(golden, "small_relations.tar.xz")) as tf:
            tf.extractall(os.path.join(d, "small"))
        init = open(os.path.join(golden, "small.init"), "rb").read()
        work = open(os.path.join(golden, "small.work"), "rb").read()
        expect = open(os.path.join(golden, "small.result"), "rb").read()

        def run(binary, reps, times=1):
            walls, same, err = [], True, b""
            for _ in range(reps):
                t0 = time.perf_counter()
                out = subprocess.run([binary], input=init + work * times, cwd=d, capture_output=True, timeout=1800,
                                     env=dict(os.environ, RHJ_HOST_TIMING="1"))
                walls.append(time.perf_counter() - t0)
                same = same and out.returncode == 0 and out.stdout == expect * times
                err = out.stderr
            return walls, same, err.decode(errors="replace")

        res = {"query_threads": 8}
        for name, what in (("join_b200_query", "reference join.cpp / parser / schedulers / printing + host/Query_execute.cpp -> "
                                               "rhj_query_execute (device-resident query path) + librhj.so"),
                           ("join_b200_full", "reference program + host/Result.cpp + host/intermediate.cpp + librhj.so")):
            b = os.path.join(host, name)
            if not os.path.exists(b):
                continue
            w1, s1, err = run(b, 3)
            w3, s3, err3 = run(b, 2, times=3)
            r = {"program": what, "wall_s": min(w1), "wall_s_all": [round(w, 3) for w in w1], "output_identical_to_small_result": s1 and s3,
                 "wall_s_three_passes": min(w3), "extra_wall_s_per_additional_pass": (min(w3) - min(w1)) / 2}
            m = re.search(r"(\d+) queries on the device: (\d+) joins, (\d+) kernel launches, H2D (\d+) bytes .* D2H (\d+) bytes", err)
            if m:
                r["pcie"] = {"queries": int(m.group(1)), "joins": int(m.group(2)), "kernel_launches": int(m.group(3)),
                             "h2d_bytes_total_columns_once": int(m.group(4)), "d2h_bytes_total": int(m.group(5)),
                             "d2h_bytes_per_query": round(int(m.group(5)) / max(int(m.group(1)), 1), 1)}
            m = re.search(r"context creation thread-time ([0-9.]+) ms", err)
            if m:
                r["cuda_context_creation_thread_ms"] = float(m.group(1))
            # the program's own clock: from the moment the last query thread had its CUDA context to the end of the last call
            ph1 = re.search(r"query phase after the last context was created: ([0-9.]+) ms", err)
            ph3 = re.search(r"query phase after the last context was created: ([0-9.]+) ms", err3)
            if ph1 and ph3:
                r["query_phase_s_one_pass"] = float(ph1.group(1)) / 1e3
                r["query_phase_s_per_pass_of_three"] = float(ph3.group(1)) / 3e3
            res[name] = r
        res["wall_s"] = res["join_b200_query"]["wall_s"]
        res["output_identical_to_small_result"] = all(v["output_identical_to_small_result"] for v in res.values() if isinstance(v, dict))
        ref_bin = os.path.join(ROOT, "oracle", "_ref", "join_ref")
        if with_reference and os.path.exists(ref_bin):
            rw, rsame, _ = run(ref_bin, 1)
            res["reference_wall_s"] = rw[0]
            res["reference_output_identical"] = rsame
            res["reference_build"] = "unmodified reference, g++ -Ofast -march=x86-64-v3 -funroll-loops -pthread"
        else:
            res["reference_wall_s_builder_measured"] = {"value": 94.0, "source": "profiles/r01_call9_small_work_wall.txt (same build, "
                                                        "8 threads, GPU box host); re-time with --small-work-ref"}
        return res


def bytes_moved(nR, nS, m, plan):
    """HBM bytes one join actually moves: two scatters (32 B per tuple each), the join's read, the pairs,
    plus the histograms that RAN -- plan['optimistic_pass1'] bit 0 / 1: the build / probe relation skipped its pass-1
    histogram, bit 2: both skipped the pass-2 one.  The canonical 96n + 16m of SURVEY 8d counts one histogram read
    (16n) regardless."""
    nB, nP = min(nR, nS), max(nR, nS)
    n = nB + nP
    o = plan["optimistic_pass1"]
    passes = (1 if plan["bits_pass1"] else 0) + (1 if plan["bits_pass2"] else 0)
    b = 32 * n * passes + 16 * n + 16 * m
    if plan["bits_pass1"]:
        b += (0 if o & 1 else 16 * nB) + (0 if o & 2 else 16 * nP)
    if plan["bits_pass2"] and not o & 4:
        b += 16 * n
    return b


def target_2p28(eng, dev, emit, peak):
    """The size the north star quotes its bar for (2^28 x 2^28 uniform, >= 50 % of the HBM roofline on one B200),
    measured like the headline: inputs resident, CUDA events around 5 joins after 3 warm-ups, count + digest checked."""
    import torch
    from radixhashjoin_b200 import workloads as W
    w = W.uniform_unique(28, dev)
    n = w.R.shape[0]
    out = torch.empty((n, 2), dtype=torch.int64, device=dev)
    eng.reserve(n, n)
    for _ in range(3):
        pairs, count = eng.join_device(w.R, w.S, out=out, emit=emit)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        pairs, count = eng.join_device(w.R, w.S, out=out, emit=emit)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    plan = eng.last_plan()
    ok = tuple(eng.pairs_digest(pairs)) == tuple(w.expected)
    canon, moved = 96 * 2 * n + 16 * count, bytes_moved(n, n, count, plan)
    return {"workload": w.name, "ms_per_step": ms, "steps": 5, "warmup": 3, "value": 2 * n / (ms * 1e-3), "unit": UNIT,
            "verified": ok, "radix_bits": [plan["bits_pass1"], plan["bits_pass2"]], "optimistic_mask": plan["optimistic_pass1"],
            "canonical_bytes": canon, "frac_canonical_of_measured_hbm": canon / (ms * 1e-3) / 1e9 / peak,
            "bytes_moved": moved, "frac_moved_of_measured_hbm": moved / (ms * 1e-3) / 1e9 / peak,
            "workspace_GiB": round(eng.workspace_bytes() / 2**30, 2)}


def _workload_name(log2n, world):
    """config.workload of both arms: the single-GPU configuration, or the global relation pair sharded over N ranks"""
    if world == 1:
        return f"uniform_unique_2^{log2n}x2^{log2n}"
    gbits = log2n + (world.bit_length() - 1)
    return f"uniform_unique_global_2^{gbits}x2^{gbits}_sharded_over_{world}"


def run_reference(args, rank):
    """The reference's own CPU path on the stated configuration: at N = 1 every step is one complete
    2^log2n x 2^log2n join (BASELINE configs[1] by default: 2^27 x 2^27, ~6 s per join, ~14 GiB of host RAM); the
    number of timed steps is capped by --ref-budget-s.  At N > 1 (config = a global relation pair sharded over N GPUs,
    which does not fit one host's reference run) rank 0 times the per-GPU share, 2^log2n x 2^log2n, and says so."""
    if rank != 0:
        return
    log2n = args.ref_log2n if args.ref_log2n else args.log2n
    val, sec, info, timed = cpu_reference_run(log2n, args.steps, min(args.warmup, 1), budget_s=args.ref_budget_s)
    same = args.gpus == 1 and log2n == args.log2n
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": timed, "steps_requested": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": _workload_name(args.log2n, args.gpus), "tuples_per_gpu": 2 << args.log2n,
                       "tuple_bytes": 16,
                       "sample_per_step": (f"the whole configuration: one 2^{log2n} x 2^{log2n} join per step" if same else
                                           f"2^{log2n} x 2^{log2n} tuples of the same generator per step (one GPU's share of the "
                                           "sharded configuration; the reference is a single-process program)"),
                       "timed_steps_cap": f"{args.ref_budget_s} s of joins (--ref-budget-s)"},
            "cpu_baseline": dict(info, value=val, unit=UNIT),
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def _device(local_rank):
    """The rank's device.  (tests/test_bench_flow.py replaces this, the engine and the exchange classes to walk through
    the orchestration below -- verification, line assembly, the N > 1 e2e and its watchdog -- without a GPU.)"""
    return f"cuda:{local_rank}"


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from radixhashjoin_b200 import RadixHashJoin, EMIT_FUSED, EMIT_COUNT_THEN_WRITE
    from radixhashjoin_b200 import workloads as W

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = _device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    emit = EMIT_FUSED if args.emit == "fused" else EMIT_COUNT_THEN_WRITE
    eng = RadixHashJoin(local_rank)
    log2n = args.log2n
    n_local = 1 << log2n

    # ---- inputs, resident in HBM before the timed region ----
    if world == 1:
        if args.workload == "uniform":
            w = W.uniform_unique(log2n, dev)
        elif args.workload == "zipf":
            w = W.zipf_probe(log2n, dev)
        elif args.workload == "fk":
            w = W.foreign_key(args.fk_build_log2, args.fk_probe_log2 or log2n, dev)
        else:
            raise SystemExit("unknown workload")
        R, S, expected = w.R, w.S, w.expected
        wname = w.name
    else:
        lw = world.bit_length() - 1
        assert (1 << lw) == world, "world size must be a power of two"
        gbits = log2n + lw
        if args.workload == "uniform":   # weak scaling: 2^log2n + 2^log2n tuples per GPU of a global 2^gbits pair (config 5 shape)
            w = W.uniform_unique(log2n, dev, row_offset=rank * n_local, log2_global=gbits)
            expected = W.uniform_unique_global_digest(gbits, dev, lo=rank * n_local, hi=(rank + 1) * n_local)
            wname = _workload_name(log2n, world)
        elif args.workload == "zipf":    # weak scaling: rank's rows of a global Zipf 2^gbits pair (config 4 shape)
            w = W.zipf_probe(gbits, dev, rank=rank, world=world)
            expected = w.expected
            wname = f"zipf_global_2^{gbits}x2^{gbits}_sharded_over_{world}"
        else:                            # fk, STRONG scaling: the global sizes of config 3 are fixed, every rank owns 1 / world
            pbits = args.fk_probe_log2 if args.fk_probe_log2 else gbits
            w = W.foreign_key(args.fk_build_log2, pbits, dev, rank=rank, world=world)
            expected = w.expected
            wname = f"fk_2^{args.fk_build_log2}x2^{pbits}_sharded_over_{world}"
        R, S = w.R, w.S
    nR, nS = R.shape[0], S.shape[0]
    n_in_local = nR + nS
    compact = False
    strategy = None

    if world == 1:
        cap = nS if args.workload != "dup" else nS
        out = torch.empty((cap, 2), dtype=torch.int64, device=dev)
        eng.reserve(nR, nS)

        def step():
            return eng.join_device(R, S, out=out, emit=emit)
    else:
        from radixhashjoin_b200.distributed import (BroadcastShardedJoin, DmaShardedJoin, PipeShardedJoin, ShardedJoin,
                                                    broadcast_is_cheaper)
        n_max = max(nR, nS)
        # what a rank receives / emits: ~ its share for hashed distinct keys; the rank that owns the hot Zipf keys gets more
        slack = int(n_max * (2.0 if args.workload == "zipf" else 1.05)) + 4096
        out = torch.empty((slack, 2), dtype=torch.int64, device=dev)
        strategy = args.shuffle
        if broadcast_is_cheaper(nR * world, nS * world, world) and args.shuffle == "pipe":
            strategy = "broadcast"
        if strategy not in ("pipe", "broadcast"):
            recvR = torch.empty((slack, 2), dtype=torch.int64, device=dev)
            recvS = torch.empty((slack, 2), dtype=torch.int64, device=dev)
            eng.reserve(slack, slack)
        if strategy == "broadcast":
            # small build side (foreign-key join): all-gather it, never shuffle the probe side, join locally
            bj = BroadcastShardedJoin(eng, world, rank, nR * world, nS * world, min(nR, nS))
            eng.reserve(nR * world if nR < nS else nR, nS * world if nS <= nR else nS)

            def step():
                pairs, count, _ = bj.step(R, S, out)
                return pairs, count
        elif strategy == "pipe":
            # histogram-free chunked pass 1 -> hand-written TMA copy kernel over NVLink -> appended pass 2 -> join;
            # no collective call and one host synchronisation per step (radixhashjoin_b200/distributed.py)
            # 12-byte records on the wire when every row id fits 32 bits (checked once here, and by the copy kernel in every
            # step) and the exchange is wire-bound (4+ ranks); 16-byte tuples otherwise
            mx = torch.stack([R[:, 0].max(), S[:, 0].max()]).max()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            wire = args.wire_bytes or (12 if world >= 4 and int(mx.item()) < (1 << 32) else 16)
            compact = wire == 12
            pj = PipeShardedJoin(eng, world, rank, nR * world, nS * world, nR, nS, chunks=args.chunks,
                                 exact_recv_capacity=slack, wire_bytes=wire)

            def step():
                pairs, count, _ = pj.step(R, S, out)
                return pairs, count

            def timeline():
                marks = []
                pj.step(R, S, out, marks)
                torch.cuda.synchronize()
                return pj.timeline(marks)
        elif strategy == "dma":
            # the exact exchange (what the pipelined one falls back to): pass 1 with histograms into staging, the copy engines
            # ship one chunk per peer while the SMs partition the other relation, pass 2 on the received pieces
            dj = DmaShardedJoin(eng, world, rank, nR * world, nS * world, n_max, slack)
            del recvR, recvS

            def step():
                pairs, count, _ = dj.step(R, S, out)
                return pairs, count

            def timeline():
                marks = []
                dj.step(R, S, out, marks)
                torch.cuda.synchronize()
                return dj.timeline(marks)
        else:
            sj = ShardedJoin(world, rank, lambda T: eng.shuffle_partition(T, world),
                             lambda a, b: eng.join_device(a, b, out=out, emit=emit))

            def step():
                # rank partition (our kernels) -> NCCL all-to-all -> local join
                pairs, count, _ = sj.step(R, S, recvR, recvS)
                return pairs, count

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then EXACTLY K timed steps between barrier + synchronize ----
    for _ in range(max(args.warmup, 3)):
        pairs, count = step()
    sampler = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    ev0.record()
    for _ in range(args.steps):
        pairs, count = step()
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    plan = eng.last_plan()
    launches_per_step = plan["kernel_launches"] + (0 if world == 1 else (6 if strategy == "nccl" else 0))
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = n_in_local * world / (ms_step * 1e-3)

    # ---- verification of the timed configuration (outside the timed region) ----
    def verify(dig):
        """count + multiset digest of this rank's pairs against the closed form (N > 1: reduced over the ranks)"""
        if world == 1:
            return tuple(dig) == tuple(expected)

        def _i64(v):
            return v - (1 << 64) if v >= (1 << 63) else v
        # every rank knows the closed-form digest of what ITS probe rows must produce; digests add up (count, sum) / xor
        tot = torch.tensor([dig[0], _i64(dig[1]), expected[0], _i64(expected[1])], dtype=torch.int64, device=dev)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        xr = [None] * world
        dist.all_gather_object(xr, (dig[2], expected[2]))
        x = e = 0
        for a, b in xr:
            x ^= a
            e ^= b
        m64 = (1 << 64) - 1
        return (int(tot[0].item()), int(tot[1].item()) & m64, x) == (int(tot[2].item()), int(tot[3].item()) & m64, e)

    verified = verify(eng.pairs_digest(pairs))
    m_local = count
    # how evenly the result (= the work of the local joins) is spread over the ranks: equal values meet on one rank, so skewed
    # probe keys load their owner more ("accept imbalance and report it"; no hot-key replication here)
    shard_balance = None
    if world > 1:
        cs = [None] * world
        dist.all_gather_object(cs, int(count))
        mean = sum(cs) / world
        shard_balance = {"pairs_per_rank": cs, "max_over_mean": (max(cs) / mean) if mean else None}

    shard_timeline = None
    if world > 1 and strategy in ("dma", "pipe"):
        shard_timeline = timeline()
    # ---- per-phase device times -> roofline of the dominant kernel (separate, untimed passes) ----
    eng.set_profiling(True)
    acc = {}
    reps = 5
    for _ in range(reps):
        step()
        torch.cuda.synchronize()
        for k, v in eng.last_phase_ms().items():
            acc[k] = acc.get(k, 0.0) + v / reps
    eng.set_profiling(False)
    nvlink = None
    if shard_timeline and strategy == "pipe" and any(k.startswith("ship_") for k, _ in shard_timeline):
        # the pipelined exchange has no library phase marks: the join kernel is what runs between the last pass 2 and the
        # end of the step; the wire is busy from the first chunk's pass 1 to the last chunk's copy kernel
        tl = dict(shard_timeline)
        p2 = [v for k, v in shard_timeline if k.startswith("pass2_")]
        p1 = [v for k, v in shard_timeline if k.startswith("pass1_")]
        sent = [v for k, v in shard_timeline if k.startswith("ship_")]
        acc = {"join": tl["join_done"] - max(p2), "scatter1": max(p1) - tl["start"]}
        wire_ms = max(sent) - min(p1)
        out_bytes = (12 if compact else 16) * n_in_local * (world - 1) / world
        nvlink = {"bytes_out_per_gpu": out_bytes, "wire_busy_ms": wire_ms, "achieved_GBps_per_direction": out_bytes / wire_ms / 1e6,
                  "peak_GBps_per_direction": 770.0, "peak_source": "B200_PROFILING.md measured peer copy",
                  "frac": out_bytes / wire_ms / 1e6 / 770.0,
                  "note": "copy kernel (k_pipe_ship) running next to the partitioning kernels, one profiled step, this rank"}
    n_join = n_in_local  # at N > 1: ~ balanced, what this rank joins after the exchange
    alg_bytes = {"hist1": 16 * n_join, "scatter1": 32 * n_join, "hist2": 16 * n_join, "scatter2": 32 * n_join,
                 "join": 16 * n_join + 16 * m_local, "join_write": 16 * n_join + 16 * m_local}
    dom = max((k for k in alg_bytes if acc.get(k, 0) > 0), key=lambda k: acc[k], default=None)
    peak, peak_src = measured_hbm_peak()
    roofline = None
    traffic = None
    traffic_src = None
    for cand in ("r02_final_ncu_traffic.json", "r01_final_ncu_traffic.json"):
        # dram bytes of the same kernel from a committed `ncu --set full` capture of this workload: a SNAPSHOT of the build named
        # in the file, not a measurement of the running one (ncu cannot run inside the timed bench)
        try:
            with open(os.path.join(ROOT, "profiles", cand)) as f:
                tj = json.load(f)
            if world == 1 and tj["workload"] == wname and tj["emitter"] == args.emit and dom in tj["phases"]:
                traffic = tj["phases"][dom]["traffic_bytes"]
                traffic_src = f"profiles/{cand} (ncu snapshot of build {tj.get('build', 'r01 final')})"
                break
        except Exception:
            continue
    if dom:
        ach = alg_bytes[dom] / (acc[dom] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": peak_src, "kernel_ms": acc[dom],
                    "algorithmic_bytes_per_launch": alg_bytes[dom],
                    # the same figure for every bandwidth kernel of the step (the dominant one is the headline above)
                    "by_kernel": {k: {"kernel_ms": round(acc[k], 4), "algorithmic_bytes_per_launch": alg_bytes[k],
                                      "frac": alg_bytes[k] / (acc[k] * 1e-3) / 1e9 / peak}
                                  for k in alg_bytes if acc.get(k, 0) > 0}}
    b_alg_step = 96 * n_join + 16 * m_local     # canonical 2-pass plan
    step_gbs = b_alg_step / (ms_step * 1e-3) / 1e9
    moved_step = bytes_moved(nR, nS, m_local, plan) if world == 1 else None

    # ---- end to end through the host entry point (pinned host buffers, H2D + D2H inside) ----
    e2e = None
    if world == 1 and not args.no_e2e:
        hR = torch.empty((nR, 2), dtype=torch.int64).pin_memory()
        hS = torch.empty((nS, 2), dtype=torch.int64).pin_memory()
        hR.copy_(R)
        hS.copy_(S)
        torch.cuda.synchronize()
        for _ in range(2):
            view, cnt = eng.join_host_view(hR, hS)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        k = max(1, min(args.steps, args.e2e_steps))
        for _ in range(k):
            view, cnt = eng.join_host_view(hR, hS)
            first = int(view["keyR"][0])  # the host reads the result
        dt = (time.perf_counter() - t0) / k
        # the host result of the last step, checked like the device one (count + multiset digest)
        import numpy as np
        dres = torch.from_numpy(view.view(np.int64).reshape(-1, 2)).to(dev)
        e2e_ok = (cnt,) + tuple(eng.pairs_digest(dres)[1:]) == tuple(expected) if expected else None
        del dres
        e2e = {"value": n_in_local / dt, "unit": UNIT, "h2d_bytes_per_step": 16 * n_in_local,
               "d2h_bytes_per_step": 16 * cnt, "ms_per_step": dt * 1e3, "steps": k, "verified": e2e_ok,
               "api": "rhj_join_host (count-then-write emitter, pinned host inputs, pinned host result; probe side streamed in "
                      "2^24-tuple chunks so H2D, compute and D2H overlap)"}
        # the same call with the arrays the drop-in host/Result.cpp passes: the reference's relation::tuples are
        # `new tuple[]` (structs.cpp:217-243), i.e. PAGEABLE host memory
        pR, pS = hR.numpy().copy(), hS.numpy().copy()
        del hR, hS
        for _ in range(1):
            view, cnt = eng.join_host_view(pR, pS)
        t0 = time.perf_counter()
        kp = max(1, min(3, k))
        for _ in range(kp):
            view, cnt = eng.join_host_view(pR, pS)
            first = int(view["keyR"][0])
        dtp = (time.perf_counter() - t0) / kp
        dres = torch.from_numpy(view.view(np.int64).reshape(-1, 2)).to(dev)
        ok_p = (cnt,) + tuple(eng.pairs_digest(dres)[1:]) == tuple(expected) if expected else None
        del dres, pR, pS
        e2e["pageable"] = {"value": n_in_local / dtp, "unit": UNIT, "ms_per_step": dtp * 1e3, "steps": kp, "verified": ok_p,
                           "note": "inputs in pageable (malloc'd) host arrays, as host/Result.cpp passes them"}
    elif world > 1:
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "note": "--no-e2e"}

    def sharded_e2e():
        """e2e at N > 1: the shards live in pinned HOST memory; every step copies this rank's R and S to the device, runs the
        sharded join and copies the pairs back (radixhashjoin_b200/distributed.py: HostResidentSteps); wall clock, max over ranks."""
        from radixhashjoin_b200.distributed import HostResidentSteps
        none = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        try:
            hs = HostResidentSteps(step, R, S, out, world, dist=dist, local_world=env_int("LOCAL_WORLD_SIZE", world))
            if not hs.ok:   # agreed between the ranks before the first step: nobody is left waiting
                return dict(none, note=hs.why)
            k = max(1, min(args.steps, args.e2e_steps))
            dt, cnt, d2h = hs.run(k, warmup=1)
            out[:cnt].copy_(hs.hout[:cnt])   # the HOST copy of the last result is what gets checked
            e2e_ok = verify(eng.pairs_digest(out[:cnt]))
            by = torch.tensor([hs.h2d_bytes, d2h], dtype=torch.int64, device=dev)
            dist.all_reduce(by, op=dist.ReduceOp.SUM)
            return {"value": n_in_local * world / dt, "unit": UNIT, "h2d_bytes_per_step": int(by[0].item()),
                    "d2h_bytes_per_step": int(by[1].item()), "ms_per_step": dt * 1e3, "steps": k, "verified": e2e_ok,
                    "api": f"{strategy} sharded join with host-resident shards: per step and rank H2D of both shards from "
                           "pinned host memory -> exchange + join -> D2H of the pairs into pinned host memory, read by the "
                           "host; bytes are summed over the ranks, the time is the slowest rank's wall clock"}
        except Exception as ex:  # an error every rank hits alike must not cost the headline line
            return dict(none, note="e2e failed: " + repr(ex)[:160])

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only, bounded sample) ----
    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu:
        del out
        torch.cuda.empty_cache()
        v, sec, info, _ = cpu_reference_run(args.cpu_log2n, 1, 0)
        cpu = dict(info, value=v, unit=UNIT, seconds=sec)

    target = None
    if world == 1 and rank == 0 and args.workload == "uniform" and log2n == 27 and not args.no_target:
        try:
            del R, S, w
            torch.cuda.empty_cache()
            target = target_2p28(eng, dev, emit, peak)
        except Exception as ex:  # never lose the headline line to the extra block
            target = {"unavailable": repr(ex)[:200]}

    def make_line(e2e_v):
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong" if world > 1 and args.workload == "fk" else "weak",
                "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": {"workload": wname, "tuples_per_gpu": n_in_local, "tuple_bytes": 16,
                           "matches_per_gpu": int(m_local), "emitter": args.emit,
                           "cache": "inputs (2 x %.1f GiB per GPU) are far larger than the 126 MB L2; no flush needed"
                                    % (nR * 16 / 2**30),
                           "radix_bits": [plan["bits_pass1"], plan["bits_pass2"]],
                           "parallelism": "1 GPU" if world == 1 else (
                               f"{world} ranks: the build side is all-gathered ({min(nR, nS) * world} tuples), the probe side is never "
                               "shuffled, every rank joins its probe shard locally" if strategy == "broadcast" else
                               f"{world} ranks: {args.chunks} row chunks per relation; histogram-free pass 1 on (rank | sub-digit) into "
                               f"fixed-capacity regions, TMA copy kernel ships them ({12 if compact else 16} B per tuple) over NVLink while "
                               "the next chunk is partitioned, "
                               "pass 2 appends each arrived chunk to fixed-capacity final partitions, one join; no collective in the "
                               f"step (exact-path steps among the timed ones: {pj.exact_steps})" if strategy == "pipe" else
                               f"{world} ranks: exact exchange -- pass 1 with histograms on (rank | sub-digit), copy engines ship one chunk per "
                               "peer over NVLink overlapped with the other relation's passes, then local pass 2 + join"
                               if strategy == "dma" else
                               f"{world} ranks: rank-radix partition + NCCL all-to-all + local join")},
                "verified": verified, "clocks": clocks, "gpu_launches": launches_per_step * args.steps,
                "phase_ms": {k: round(v, 4) for k, v in acc.items() if v > 0},
                "roofline": roofline,
                "step_roofline": {"algorithmic_bytes": b_alg_step, "formula": "96*n + 16*m (SURVEY 8d)",
                                  "achieved_GBps": step_gbs, "frac_of_measured_hbm": step_gbs / peak,
                                  "bytes_moved": moved_step,
                                  "frac_moved_of_measured_hbm": moved_step / (ms_step * 1e-3) / 1e9 / peak if moved_step else None,
                                  "note": "canonical bytes credit the histograms the histogram-free passes skip; bytes_moved "
                                          "counts only what ran (plan optimistic mask %d)" % plan["optimistic_pass1"]},
                "e2e": e2e_v, "cpu_baseline": cpu}
        if target:
            line["target_2p28"] = target
        if shard_balance:
            line["shard_balance"] = shard_balance
        if shard_timeline:
            line["shard_timeline_ms"] = shard_timeline
        if nvlink:
            line["nvlink"] = nvlink
        if world == 1 and not args.no_small_work:
            line["small_work"] = small_work_wall(args.small_work_ref)
        return line

    if world > 1 and not args.no_e2e:
        # The host-resident e2e runs LAST and under a watchdog: whatever happens to it (a rank that dies in the pinned
        # allocation, an exchange left waiting), rank 0 still prints the line of the measurements above.
        done = threading.Lock()

        def bail():
            if done.acquire(blocking=False):
                if rank == 0:
                    print(json.dumps(make_line({"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                                                "note": f"e2e did not finish within {args.e2e_timeout_s} s"})), flush=True)
                os._exit(0)
        wd = threading.Timer(args.e2e_timeout_s, bail)
        wd.daemon = True
        wd.start()
        e2e = sharded_e2e()
        wd.cancel()
        if not done.acquire(blocking=False):
            time.sleep(600)   # the watchdog is printing the line and ends the process
    if rank == 0:
        print(json.dumps(make_line(e2e)), flush=True)
    if world > 1:
        # the line is out; a teardown that waits for a rank that is gone must not keep the job alive
        td = threading.Timer(60.0, lambda: os._exit(0))
        td.daemon = True
        td.start()
        dist.destroy_process_group()
        td.cancel()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2n", type=int, default=27, help="tuples per relation per GPU = 2^log2n")
    ap.add_argument("--workload", default="uniform", choices=["uniform", "zipf", "fk"])
    ap.add_argument("--fk-build-log2", type=int, default=24)
    ap.add_argument("--fk-probe-log2", type=int, default=0, help="fk: GLOBAL probe size (default: 2^log2n per GPU)")
    ap.add_argument("--emit", default="fused", choices=["fused", "count_then_write"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-timeout-s", type=float, default=240.0, help="N > 1: give up on the host-resident e2e after this long")
    ap.add_argument("--cpu-log2n", type=int, default=26, help="cpu_baseline sample size (2^k x 2^k)")
    ap.add_argument("--ref-log2n", type=int, default=0, help="--impl reference: 2^k x 2^k per step (default: --log2n, the stated config)")
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="--impl reference: stop timing after this many seconds of joins")
    ap.add_argument("--shuffle", default="pipe", choices=["pipe", "dma", "nccl"],
                    help="multi-GPU exchange: pipelined histogram-free chunks shipped by our copy kernel (default), the exact exchange "
                         "(pass-1 chunks shipped by the copy engines after histograms), or rank partition + NCCL all-to-all")
    ap.add_argument("--chunks", type=int, default=4, help="pipe shuffle: row chunks per relation")
    ap.add_argument("--wire-bytes", type=int, default=0, choices=[0, 12, 16],
                    help="pipe shuffle: bytes per tuple on the wire (0 = 12 when row ids fit 32 bits and N >= 4, else 16)")
    ap.add_argument("--no-small-work", action="store_true")
    ap.add_argument("--small-work-ref", action="store_true", help="also time the unmodified reference program (minutes)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-target", action="store_true", help="skip the extra 2^28 x 2^28 block of the N=1 line")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        raise SystemExit("launch N > 1 with torchrun (one rank per GPU); see the module docstring")
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
