"""bench.py's orchestration walked through WITHOUT a GPU: the engine is replaced by the oracle join and the pipelined
exchange by the gloo all-to-all of test_distributed_cpu.py, so that what runs is bench.py's own control flow -- workload
set-up, timed loop, verification against the closed forms (reduced over the ranks at N = 2), per-phase / roofline
bookkeeping, the host-resident e2e at N > 1 with its agreement between the ranks and its watchdog, and the assembly of the
one JSON line.  The driver runs bench.py at N = 1, 2, 4, 8 at the end of a round, when nothing can be fixed any more: a
NameError in a branch only N > 1 takes must show up here.  Numbers printed by these runs mean nothing."""
import io
import json
import os
import socket
import sys
import time

import numpy as np
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Event:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3 + 1e-3


def _worker(rank, world, port, argv, q, e2e_hangs=False, fake_pinned=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      LOCAL_WORLD_SIZE=str(world))
    import torch.distributed as dist
    import _oracle as O
    import bench
    import radixhashjoin_b200
    import radixhashjoin_b200.distributed as D
    from radixhashjoin_b200 import workloads as W

    # ---- no GPU: CUDA plumbing replaced by no-ops, the device is the CPU ----
    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda d: None
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.empty_cache = lambda: None
    torch.cuda.Event = _Event
    bench._device = lambda local_rank: "cpu"
    real_init = dist.init_process_group
    dist.init_process_group = lambda backend, **kw: real_init("gloo", rank=rank, world_size=world)

    def to_pairs_tensor(p):
        return torch.from_numpy(np.ascontiguousarray(p).view(np.uint64).reshape(-1, 2).view(np.int64).copy())

    class FakeEngine:
        """the oracle join behind the few engine calls bench.py makes"""
        device = 0

        def __init__(self, device=0):
            pass

        def reserve(self, nR, nS):
            pass

        def join_device(self, R, S, out=None, emit=0):
            p = O.oracle_join(W.to_numpy_tuples(R), W.to_numpy_tuples(S))
            out[:len(p)] = to_pairs_tensor(p)
            return out[:len(p)], len(p)

        def last_plan(self):
            return {"kernel_launches": 11, "bits_pass1": 8, "bits_pass2": 8, "optimistic_pass1": 7}

        def pairs_digest(self, pairs):
            return O.pairs_digest(pairs.numpy().view(np.uint64).reshape(-1, 2).copy().view(O.PAIR).reshape(-1))

        def set_profiling(self, on):
            pass

        def last_phase_ms(self):
            return {"scatter1": 1.0, "scatter2": 1.1, "join": 0.9, "plan": 0.01}

    class FakePipe:
        """PipeShardedJoin's surface over the gloo all-to-all exchange"""
        exact_steps = 0

        def __init__(self, engine, world_, rank_, *a, **kw):
            self.sj = D.ShardedJoin(world_, rank_, lambda T: D.cpu_partition_by_rank(T, world_), None)
            self.engine = engine

        def step(self, R, S, out, marks=None):
            if e2e_hangs and getattr(self, "in_e2e", False) and rank == 1:
                time.sleep(3600)          # a rank that never comes back from the exchange
            self.sj.join_fn = lambda a, b: self.engine.join_device(a, b, out=out)
            pairs, count, _ = self.sj.step(R, S)
            if marks is not None:
                for name in ("start", "pass1_0.0_done", "ship_0.0_sent", "pass2_0.0_done", "join_done"):
                    ev = _Event()
                    ev.record()
                    marks.append((name, ev))
                    time.sleep(0.001)
            return pairs, count, None

        @staticmethod
        def timeline(marks):
            return [(n, round(marks[0][1].elapsed_time(ev), 3)) for n, ev in marks]

    radixhashjoin_b200.RadixHashJoin = FakeEngine
    D.PipeShardedJoin = FakePipe
    if e2e_hangs or fake_pinned:
        # pinned memory "works" (plain host tensors); with e2e_hangs rank 1 then hangs in the first e2e step
        real_hs = D.HostResidentSteps

        class HS(real_hs):
            def __init__(self, step_fn, R, S, out, world_, **kw):
                pipe = step_fn.__closure__ and [c.cell_contents for c in step_fn.__closure__ if isinstance(c.cell_contents, FakePipe)]
                for p in pipe or []:
                    p.in_e2e = e2e_hangs
                kw.update(alloc_host=lambda shape: torch.empty(shape, dtype=torch.int64), sync=lambda: None, avail_fn=lambda: 1 << 40)
                super().__init__(step_fn, R, S, out, world_, **kw)
        D.HostResidentSteps = HS

    buf = io.StringIO()
    real_stdout, sys.stdout = sys.stdout, buf
    real_exit = os._exit

    def fake_exit(code):          # the watchdog ends the process: hand over what was printed first
        q.put((rank, buf.getvalue()))
        time.sleep(0.5)
        real_exit(code)
    os._exit = fake_exit
    sys.argv = ["bench.py"] + argv
    try:
        bench.main()
    finally:
        sys.stdout = real_stdout
    q.put((rank, buf.getvalue()))


def _run(world, argv, e2e_hangs=False, timeout=300, fake_pinned=False):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, argv, q, e2e_hangs, fake_pinned)) for r in range(world)]
    for p in procs:
        p.start()
    outs = dict(q.get(timeout=timeout) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        if p.is_alive():
            p.kill()
    lines = [l for l in outs[0].splitlines() if l.startswith("{")]
    assert len(lines) == 1, outs
    for r in range(1, world):
        assert not [l for l in outs[r].splitlines() if l.startswith("{")]     # only rank 0 prints
    return json.loads(lines[0])


CONTRACT = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "verified", "phase_ms", "step_roofline")


def test_single_gpu_flow_prints_the_contract_line():
    d = _run(1, ["--gpus", "1", "--steps", "3", "--warmup", "3", "--log2n", "12", "--no-e2e", "--no-cpu", "--no-small-work", "--no-target"])
    for k in CONTRACT:
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["verified"] is True and d["gpu_launches"] == 33
    assert d["config"]["workload"] == "uniform_unique_2^12x2^12" and d["scaling"] == "weak"
    assert d["roofline"]["kernel"] == "scatter2" and d["roofline"]["bound"] == "hbm" and 0 < d["roofline"]["frac"]
    assert set(d["roofline"]["by_kernel"]) == {"scatter1", "scatter2", "join"}
    assert d["roofline"]["by_kernel"]["scatter2"]["frac"] == d["roofline"]["frac"]
    assert d["step_roofline"]["bytes_moved"] == 80 * (2 << 12) + 16 * (1 << 12)     # both histogram-free passes ran (mask 7)


def test_two_rank_flow_verifies_across_ranks_and_agrees_on_the_e2e():
    """N = 2 over gloo: digests reduce to the global closed form; pinned host memory cannot be allocated here, which every rank
    learns through HostResidentSteps' agreement -- the line carries e2e.value null with the reason instead of a hang."""
    d = _run(2, ["--gpus", "2", "--steps", "2", "--warmup", "3", "--log2n", "11"])
    for k in CONTRACT:
        assert k in d, k
    assert d["n_gpus"] == 2 and d["verified"] is True and d["value"] > 0
    assert d["config"]["workload"] == "uniform_unique_global_2^12x2^12_sharded_over_2"
    assert "no collective in the step" in d["config"]["parallelism"]
    assert d["e2e"]["value"] is None and "pinned host allocation failed" in d["e2e"]["note"] or "another rank" in d["e2e"]["note"]
    assert d["shard_timeline_ms"][0][0] == "start" and d["nvlink"]["bytes_out_per_gpu"] > 0
    assert sum(d["shard_balance"]["pairs_per_rank"]) == 1 << 12 and 1.0 <= d["shard_balance"]["max_over_mean"] < 1.2
    assert d["roofline"]["kernel"] in ("scatter1", "join")


def test_two_rank_flow_host_resident_e2e():
    """the same with host buffers that can be allocated (plain tensors stand in for pinned memory): the e2e steps run, the HOST
    copy of the result verifies against the closed form, bytes are summed over the ranks"""
    d = _run(2, ["--gpus", "2", "--steps", "2", "--warmup", "3", "--log2n", "11", "--e2e-steps", "2"], fake_pinned=True)
    n = 1 << 11
    assert d["verified"] is True
    assert d["e2e"]["value"] > 0 and d["e2e"]["verified"] is True and d["e2e"]["steps"] == 2
    assert d["e2e"]["h2d_bytes_per_step"] == 2 * 2 * n * 16 and d["e2e"]["d2h_bytes_per_step"] == 2 * n * 16
    assert "host-resident shards" in d["e2e"]["api"]


def test_two_rank_e2e_that_hangs_still_prints_the_line():
    """rank 1 never returns from the first e2e step: the watchdog prints the line of the device-timed measurements and ends
    both processes"""
    t0 = time.time()
    d = _run(2, ["--gpus", "2", "--steps", "2", "--warmup", "3", "--log2n", "11", "--e2e-timeout-s", "8"], e2e_hangs=True)
    assert time.time() - t0 < 120
    assert d["n_gpus"] == 2 and d["verified"] is True and d["value"] > 0
    assert d["e2e"]["value"] is None and "did not finish" in d["e2e"]["note"]
