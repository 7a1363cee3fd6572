"""Multi-GPU radix hash join: one process per GPU, one exchange step.

The reference is single-process (join.cpp:42-50); equi-join partitions are independent, so the path
shards like this:

  1. every rank groups its local R and S tuples by DESTINATION RANK = top log2(world) bits of the
     high hash word (``rhj_shuffle_partition_device``: our histogram / scan / scatter kernels);
  2. the per-destination counts are exchanged (all_to_all of 2*world int64);
  3. the tuples are exchanged with ``all_to_all_single`` (NCCL over NVLink / NVSwitch; gloo in the
     CPU tests) -- equal join values always land on the same rank;
  4. every rank joins what it received with the single-GPU path (``rhj_join_device``), whose own
     partition bits come from the LOW hash word, so local partitions stay balanced;
  5. results stay sharded (row ids are global); only counts/digests are reduced for verification.

``partition_fn`` and ``join_fn`` are injectable so that the CPU (gloo) tests can drive exactly this
exchange logic with a numpy partitioner and the oracle join; the product path passes the engine's
CUDA entry points.
"""
import numpy as np

_M64 = (1 << 64) - 1


def hash64_np(v):
    """numpy restatement of the device hash (csrc/rhj_device.cuh: hash64) -- used by the CPU tests
    of the exchange logic and to check that equal values share a destination rank."""
    v = np.asarray(v, dtype=np.uint64).copy()
    c = np.uint64(0xd6e8feb86659fd93)
    with np.errstate(over="ignore"):
        v ^= v >> np.uint64(32)
        v *= c
        v ^= v >> np.uint64(32)
        v *= c
        v ^= v >> np.uint64(32)
    return v


def rank_of_values(values, world):
    """Destination rank of each join value: top log2(world) bits of the high hash word."""
    bits = world.bit_length() - 1
    assert (1 << bits) == world, "world size must be a power of two"
    if bits == 0:
        return np.zeros(len(values), dtype=np.int64)
    hi = (hash64_np(values) >> np.uint64(32)).astype(np.uint64)
    return (hi >> np.uint64(32 - bits)).astype(np.int64)


class ShardedJoin:
    """Drives steps 1-4 for one rank.  Tensors are (n, 2) int64 relations (row id, value)."""

    def __init__(self, world, rank, partition_fn, join_fn, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.world = world
        self.rank = rank
        self.partition_fn = partition_fn   # T -> (grouped T, [count per rank])
        self.join_fn = join_fn             # (R, S) -> (pairs, count)
        self.group = group
        self._device = None

    def exchange_counts(self, cR, cS):
        import torch
        send = torch.tensor([v for pair in zip(cR, cS) for v in pair], dtype=torch.int64)
        recv = torch.empty_like(send)
        dev = self._device
        if dev is not None:
            send, recv = send.to(dev), recv.to(dev)
        self.dist.all_to_all_single(recv, send, group=self.group)
        r = recv.cpu().tolist()
        return r[0::2], r[1::2]

    def exchange_tuples(self, grouped, send_counts, recv_counts, recv_buf=None):
        import torch
        total = sum(recv_counts)
        if recv_buf is None:
            recv_buf = torch.empty((max(total, 1), 2), dtype=torch.int64, device=grouped.device)
        assert total <= recv_buf.shape[0], "receive buffer too small for this rank's share"
        out = recv_buf[:total]
        self.dist.all_to_all_single(out, grouped, output_split_sizes=list(recv_counts),
                                    input_split_sizes=list(send_counts), group=self.group)
        return out

    def step(self, R_local, S_local, recvR=None, recvS=None):
        """One sharded join of the local shards; returns (pairs, count, received (nR, nS))."""
        self._device = R_local.device if R_local.is_cuda else None
        gR, cR = self.partition_fn(R_local)
        gS, cS = self.partition_fn(S_local)
        rR, rS = self.exchange_counts(cR, cS)
        myR = self.exchange_tuples(gR, cR, rR, recvR)
        myS = self.exchange_tuples(gS, cS, rS, recvS)
        pairs, count = self.join_fn(myR, myS)
        return pairs, count, (sum(rR), sum(rS))


def cpu_partition_by_rank(T, world):
    """numpy stand-in for rhj_shuffle_partition_device (CPU tests of the exchange logic only)."""
    import torch
    a = T.numpy().view(np.uint64).reshape(-1, 2)
    ranks = rank_of_values(a[:, 1], world)
    order = np.argsort(ranks, kind="stable")
    counts = np.bincount(ranks, minlength=world).tolist()
    return torch.from_numpy(a[order].view(np.int64).copy()), counts


def broadcast_is_cheaper(nR_global, nS_global, world):
    """Exchange strategy of a sharded join (skew / small-build caveat): replicating the build side
    costs every rank nB * (world - 1) / world tuples in, shuffling both sides (nB + nP) * (world - 1) / world^2;
    with the redundant build-side partitioning on top, broadcasting pays once nB * world <= nP."""
    nB, nP = min(nR_global, nS_global), max(nR_global, nS_global)
    return world > 1 and nB * world <= nP


def host_memory_available():
    """Bytes of host RAM this process group may still take: the smaller of the machine's available memory and what is
    left under the cgroup limit (a container that is OOM-killed cannot report anything)."""
    import psutil
    avail = psutil.virtual_memory().available
    try:
        with open("/sys/fs/cgroup/memory.max") as f:
            lim = f.read().strip()
        if lim != "max":
            with open("/sys/fs/cgroup/memory.current") as f:
                avail = min(avail, int(lim) - int(f.read().strip()))
    except (OSError, ValueError):
        pass
    return max(0, int(avail))


class HostResidentSteps:
    """End-to-end driver of a sharded join whose shards live in HOST memory (bench.py `e2e` at N > 1): every timed step
    copies this rank's R and S shards from pinned host memory into the device tensors the join reads, runs one sharded
    join (`step_fn`, any of the classes below) and copies the result pairs back into a pinned host buffer, which the host
    then reads.  Nothing is overlapped across steps: a step is H2D -> exchange + join -> D2H.

    Whether the pinned buffers could be allocated is AGREED between the ranks before the first step (one all-reduce), so
    that a rank which is out of host memory makes every rank skip the measurement instead of leaving its peers waiting in
    the exchange.  `alloc_host` / `sync` / `avail_fn` are injectable for the CPU (gloo) test of this logic."""

    SAFETY = 1.5      # pinned bytes of all local ranks x this must fit the host's available memory

    def __init__(self, step_fn, R, S, out, world, dist=None, group=None, alloc_host=None, sync=None, avail_fn=None,
                 local_world=None):
        import torch
        self.torch = torch
        self.step_fn, self.R, self.S, self.out = step_fn, R, S, out
        self.world, self.dist, self.group = world, dist, group
        self.sync = sync if sync is not None else (lambda: torch.cuda.synchronize(R.device))
        alloc_host = alloc_host if alloc_host is not None else (lambda shape: torch.empty(shape, dtype=torch.int64, pin_memory=True))
        avail_fn = avail_fn if avail_fn is not None else host_memory_available
        self.h2d_bytes = 8 * (R.numel() + S.numel())
        need = self.h2d_bytes + 8 * out.numel()
        self.hR = self.hS = self.hout = None
        self.why = None
        # phase 1: every rank looks at the host's free memory BEFORE anyone allocates (the ranks of one box share it)
        ok = self._all(need * (local_world or world) * self.SAFETY <= avail_fn())
        if not ok:
            self.why = "not enough host memory for pinned copies of every rank's shards and result"
        else:
            # phase 2: allocate, then agree again
            try:
                self.hR, self.hS, self.hout = alloc_host(tuple(R.shape)), alloc_host(tuple(S.shape)), alloc_host(tuple(out.shape))
                self.hR.copy_(R)
                self.hS.copy_(S)
                self.sync()
                mine = True
            except (RuntimeError, MemoryError) as ex:
                mine, self.why = False, f"pinned host allocation failed: {str(ex)[:120]}"
            ok = self._all(mine)
            if not ok and self.why is None:
                self.why = "another rank could not allocate its pinned host buffers"
        self.ok = ok
        if not self.ok:
            self.hR = self.hS = self.hout = None

    def _all(self, flag):
        if self.dist is None or self.world == 1:
            return bool(flag)
        t = self.torch.tensor([1 if flag else 0], dtype=self.torch.int64, device=self.R.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)
        return bool(int(t.item()))

    def one(self):
        """H2D of both shards -> one sharded join -> D2H of the pairs; returns (host pairs view, count)."""
        self.R.copy_(self.hR, non_blocking=True)
        self.S.copy_(self.hS, non_blocking=True)
        pairs, count = self.step_fn()
        if count:
            self.hout[:count].copy_(pairs, non_blocking=True)
        self.sync()
        return self.hout[:count], count

    def run(self, steps, warmup=1):
        """(seconds per step = max over ranks, count of the last step, D2H bytes of the last step on this rank)"""
        import time
        assert self.ok, self.why
        for _ in range(warmup):
            self.one()
        self._barrier()
        t0 = time.perf_counter()
        count = 0
        first = 0
        for _ in range(steps):
            view, count = self.one()
            first ^= int(view[0, 0]) if count else 0     # the host reads the result
        dt = (time.perf_counter() - t0) / steps
        self._barrier()
        if self.dist is not None and self.world > 1:
            t = self.torch.tensor([dt], dtype=self.torch.float64, device=self.R.device)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
            dt = float(t.item())
        return dt, count, 16 * count

    def _barrier(self):
        self.sync()
        if self.dist is not None and self.world > 1:
            self.dist.barrier(group=self.group)
        self.sync()


class BroadcastShardedJoin:
    """Multi-GPU join for a small build side (foreign-key joins, BASELINE config 3): every rank gathers the
    whole build relation (one all-gather of nB tuples; the probe relation is never shuffled) and joins it with
    its local probe shard through the single-GPU path.  Results stay sharded by probe row."""

    def __init__(self, engine, world, rank, nR_global, nS_global, n_build_local, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.engine, self.world, self.rank = engine, world, rank
        self.group = group if group is not None else dist.group.WORLD
        self.build_is_S = nS_global < nR_global
        dev = torch.device("cuda", engine.device)
        self.n_build_local = int(n_build_local)   # equal on every rank (the all-gather is unpadded)
        self.full = torch.empty((world * self.n_build_local, 2), dtype=torch.int64, device=dev)

    def step(self, R_local, S_local, out, marks=None):
        B, P = (S_local, R_local) if self.build_is_S else (R_local, S_local)
        if B.shape[0] != self.n_build_local:
            raise ValueError("the build shards must have the size given at construction on every rank")
        self.dist.all_gather_into_tensor(self.full, B, group=self.group)
        R, S = (P, self.full) if self.build_is_S else (self.full, P)
        pairs, count = self.engine.join_device(R, S, out=out)
        return pairs, count, None


class PipeShardedJoin:
    """Multi-GPU join over the pipelined exchange (rhj_pipe_* in include/rhj.h): the default at N > 1.

    Every rank cuts each local relation into `chunks` row chunks.  Pass 1 of a chunk scatters on
    (destination rank | sub-digit) into fixed-capacity regions -- no histogram, no all-gather, no host
    round trip; a hand-written copy kernel on a second, high-priority stream ships the filled part of every
    region into the same region of the destination's symmetric receive buffer (TMA bulk copies through a
    shared-memory ring, peer stores over NVLink / NVSwitch) while the SMs partition the next chunk; the
    destination appends a chunk to its fixed-capacity final partitions as soon as the chunk's flags have
    arrived from all ranks (a device-side wait in front of pass 2); one build/probe/emit pass follows.
    A step is ~40 kernel launches on two streams, one host synchronisation (the result count) and no
    collective call; torch.distributed only provides the symmetric-memory rendezvous at construction.

    Skewed or duplicate-heavy inputs overflow a region somewhere; every rank learns it through the flags
    and all ranks redo the step through the exact (histogram) exchange, DmaShardedJoin, and stay on it for
    the next 16 steps.
    """

    OVERFLOW, TIMEOUT, BAD, WIDE = 1, 2, 4, 8

    def __init__(self, engine, world, rank, nR_global, nS_global, nR_local_max, nS_local_max, chunks=4, group=None,
                 ship_ctas=0, exact_recv_capacity=None, wire_bytes=16):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.torch, self.dist = torch, dist
        self.engine, self.world, self.rank = engine, world, rank
        self.group = group if group is not None else dist.group.WORLD
        self.plan = engine.shard_plan(nR_global, nS_global, world)
        self.build_rel = 1 if self.plan.build_is_S else 0
        self.order = (self.build_rel, 1 - self.build_rel)
        self.chunks = int(chunks)
        self.nmax = (int(nR_local_max), int(nS_local_max))
        self.chunk_rows = tuple((n + self.chunks - 1) // self.chunks for n in self.nmax)
        self._globals = (nR_global, nS_global)
        self._exact_cap = exact_recv_capacity
        dev = torch.device("cuda", engine.device)
        self.wire_bytes = int(wire_bytes)   # 12: the copy kernel repacks to {u64 value, u32 row id}; row ids must fit 32 bits
        cfg = engine.pipe_cfg(self.plan, rank, self.chunks, self.nmax[0], self.nmax[1], ship_ctas=ship_ctas,
                              wire_bytes=self.wire_bytes)
        self.sym_bytes = engine.pipe_sym_bytes(self.plan, cfg)
        self.sym = symm_mem.empty((self.sym_bytes // 8,), dtype=torch.int64, device=dev)
        self.sym.zero_()
        self.hdl = symm_mem.rendezvous(self.sym, self.group)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        engine.pipe_open(self.plan, engine.pipe_cfg(self.plan, rank, self.chunks, self.nmax[0], self.nmax[1], ptrs, ship_ctas,
                                                    self.wire_bytes))
        torch.cuda.synchronize(dev)
        dist.barrier(group=self.group)       # every block is zero before anyone ships into it
        self.copy_stream = torch.cuda.Stream(device=dev, priority=-1)   # the copy kernel's CTAs go first when an SM frees up
        self.epoch = 0
        self.exact_left = 0
        self.backoff = 0
        self.exact_steps = 0
        self._exact = None

    def _mark(self, stream=None):
        ev = self.torch.cuda.Event(enable_timing=True)
        ev.record(stream if stream is not None else self.torch.cuda.current_stream())
        return ev

    def _exact_step(self, R_local, S_local, out, marks=None):
        if self._exact is None:
            n_local_max = max(self.nmax)
            cap = self._exact_cap or int(n_local_max * 1.3) + 8192
            self._exact = DmaShardedJoin(self.engine, self.world, self.rank, self._globals[0], self._globals[1], n_local_max,
                                         cap, group=self.group)
        self.exact_steps += 1
        if marks is not None:
            del marks[:]   # the timeline of a redone step is the exact exchange's
        return self._exact.step(R_local, S_local, out, marks)

    def step(self, R_local, S_local, out, marks=None):
        """One sharded join; returns (pairs, count, None).  `marks` (a list) collects (label, CUDA event)
        pairs for a timeline of the step."""
        torch, eng = self.torch, self.engine
        if self.exact_left > 0:
            self.exact_left -= 1
            return self._exact_step(R_local, S_local, out, marks)
        rels = (R_local, S_local)
        for rel in (0, 1):
            if rels[rel].shape[0] > self.nmax[rel]:
                raise ValueError("local shard larger than the nR_local_max / nS_local_max given at construction")
        self.epoch += 1
        cur = torch.cuda.current_stream()
        cs = self.copy_stream
        if marks is not None:
            marks.append(("start", self._mark()))
        eng.pipe_begin(self.epoch)
        for rel in self.order:
            T, rows = rels[rel], self.chunk_rows[rel]
            n = T.shape[0]
            for c in range(self.chunks):
                lo = min(n, c * rows)
                hi = min(n, lo + rows)
                eng.pipe_pass1(rel, c, T[lo:hi])
                ev = torch.cuda.Event(enable_timing=marks is not None)
                ev.record(cur)
                if marks is not None:
                    marks.append((f"pass1_{rel}.{c}_done", ev))
                cs.wait_event(ev)
                eng.pipe_ship(rel, c, stream=cs)
                if marks is not None:
                    marks.append((f"ship_{rel}.{c}_sent", self._mark(cs)))
        for rel in self.order:
            for c in range(self.chunks):
                eng.pipe_pass2(rel, c)
                if marks is not None:
                    marks.append((f"pass2_{rel}.{c}_done", self._mark()))
        eng.pipe_post()
        pairs, count, status = eng.pipe_join(out)
        if marks is not None:
            marks.append(("join_done", self._mark()))
        cur.wait_stream(cs)
        if status & self.WIDE:
            raise RuntimeError(f"rank {self.rank}: a row id does not fit 32 bits -- construct PipeShardedJoin with wire_bytes=16")
        if status & (self.TIMEOUT | self.BAD):
            raise RuntimeError(f"rank {self.rank}: pipelined exchange failed (status {status}: "
                               f"{'timeout ' if status & self.TIMEOUT else ''}{'bad region end' if status & self.BAD else ''})")
        if status & self.OVERFLOW:
            # every rank saw the same status: all redo this step exactly, and stay exact for a while -- twice as long each time
            # the pipelined attempt fails again (a skewed workload should not pay a wasted step every 17 joins)
            self.backoff = min(1024, 2 * self.backoff) if self.backoff else 16
            self.exact_left = self.backoff
            return self._exact_step(R_local, S_local, out, marks)
        self.backoff = 0
        return pairs, count, None

    @staticmethod
    def timeline(marks):
        t0 = marks[0][1]
        return [(name, round(t0.elapsed_time(ev), 3)) for name, ev in marks]


class DmaShardedJoin:
    """Multi-GPU join over the EXACT exchange (rhj_shardx_* in include/rhj.h): what PipeShardedJoin falls back to when a
    fixed-capacity region overflows (skewed or duplicate-heavy inputs), and the r01 default.

    Pass 1 of the join partitions each local shard on (destination rank | sub-digit) into a local staging buffer after an
    exact histogram; the chunk for every destination is contiguous and already pass-1 partitioned, so the copy engines
    ship it with ONE peer copy per destination over NVLink / NVSwitch straight into the destination's symmetric receive
    buffer.  The build relation is partitioned and shipped first; the probe relation is partitioned while it is in flight
    and shipped next; pass 2 of a relation runs as soon as it has landed.  torch.distributed supplies the plumbing: one
    small all-gather per relation (histograms), symmetric-memory rendezvous for the peer buffers, device-side barriers.
    """

    def __init__(self, engine, world, rank, nR_global, nS_global, n_local_max, recv_capacity, group=None):
        import os
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.torch, self.dist = torch, dist
        self.engine, self.world, self.rank = engine, world, rank
        self.group = group if group is not None else dist.group.WORLD
        self.plan = engine.shard_plan(nR_global, nS_global, world)
        self.build_rel = 1 if self.plan.build_is_S else 0
        self.probe_rel = 1 - self.build_rel
        self.slots = [self.build_rel, self.probe_rel]      # shipping order
        dev = torch.device("cuda", engine.device)
        self.capacity = {s: recv_capacity for s in self.slots}
        self.recv = {s: symm_mem.empty((recv_capacity, 2), dtype=torch.int64, device=dev) for s in self.slots}
        self.hdl = {s: symm_mem.rendezvous(self.recv[s], self.group) for s in self.slots}
        self.peer = {s: [self.hdl[s].get_buffer(p, (recv_capacity, 2), torch.int64) for p in range(world)] for s in self.slots}
        self.stage = {s: torch.empty((n_local_max, 2), dtype=torch.int64, device=dev) for s in self.slots}
        ndig = world << self.plan.bits_pass1
        self.hist = {s: torch.empty(ndig, dtype=torch.int64, device=dev) for s in self.slots}
        self.all_hist = {s: torch.empty((world, ndig), dtype=torch.int64, device=dev) for s in self.slots}
        self.copy_stream = torch.cuda.Stream(device=dev)
        # a few copy lanes: the peer copies of a relation run on different copy engines (measured on 8 x B200: 1 lane
        # 10.6 ms/join, 2 lanes 10.7, 4 lanes 13.9, 8 lanes 14.6 -- too many concurrent flows through the switch collapse)
        lanes = int(os.environ.get("RHJ_COPY_LANES", "2"))
        self.peer_streams = [torch.cuda.Stream(device=dev) for _ in range(max(1, min(world, lanes)))]

    def _ship(self, slot, lay, marks=None):
        """peer copies of one relation on the copy stream, fenced by device-side barriers"""
        torch = self.torch
        send_off, send_cnt, dst_off = lay[:3]
        cs = self.copy_stream
        cs.wait_stream(torch.cuda.current_stream())        # the staging buffer is complete
        with torch.cuda.stream(cs):
            self.hdl[slot].barrier(channel=0)                # every rank is done with this receive buffer
            if marks is not None:
                marks.append((f"dma{slot}_start", self._mark(cs)))
            fork = torch.cuda.Event()
            fork.record(cs)
            for k in range(1, self.world + 1):
                d = (self.rank + k) % self.world             # remote chunks first, staggered; own chunk last
                if send_cnt[d]:
                    ps = self.peer_streams[k % len(self.peer_streams)]
                    ps.wait_event(fork)
                    a, b, n = dst_off[d], send_off[d], send_cnt[d]
                    with torch.cuda.stream(ps):
                        self.peer[slot][d][a:a + n].copy_(self.stage[slot][b:b + n], non_blocking=True)
                    cs.wait_stream(ps)
            if marks is not None:
                marks.append((f"dma{slot}_sent", self._mark(cs)))
            self.hdl[slot].barrier(channel=1)                # every rank's copies have landed
            done = torch.cuda.Event(enable_timing=marks is not None)
            done.record(cs)
            if marks is not None:
                marks.append((f"dma{slot}_landed", done))
        return done

    def _mark(self, stream=None):
        ev = self.torch.cuda.Event(enable_timing=True)
        ev.record(stream if stream is not None else self.torch.cuda.current_stream())
        return ev

    def step(self, R_local, S_local, out, marks=None):
        """One sharded join; returns (pairs, count, (received build, received probe)).  `marks` (a list)
        collects (label, CUDA event) pairs for a timeline of the step."""
        torch, eng, plan = self.torch, self.engine, self.plan
        src = {0: R_local, 1: S_local}
        if marks is not None:
            marks.append(("start", self._mark()))
        eng.shardx_begin(plan)
        lay, landed = {}, {}
        for s in self.slots:
            eng.shardx_pass1(plan, s, src[s], self.stage[s], self.hist[s])
            if marks is not None:
                marks.append((f"pass1_{s}_done", self._mark()))
            self.dist.all_gather_into_tensor(self.all_hist[s], self.hist[s], group=self.group)
            lay[s] = eng.shardx_layout(plan, self.rank, s, self.all_hist[s])
            if lay[s][4] > self.capacity[s]:
                # lay[s][4] = the largest share any rank receives, the same number everywhere: all ranks raise together
                # (a rank that stopped alone would leave its peers waiting in the device-side barriers of _ship)
                raise RuntimeError(f"rank {self.rank}: receive buffers of relation {s} too small on some rank "
                                   f"({lay[s][4]} > {self.capacity[s]} tuples; this rank receives {lay[s][3]})")
            landed[s] = self._ship(s, lay[s], marks)         # in flight while the next relation is partitioned
        for s in self.slots:
            torch.cuda.current_stream().wait_event(landed[s])
            eng.shardx_pass2(plan, s, self.recv[s][:lay[s][3]])   # overlaps the transfer of the relation behind it
            if marks is not None:
                marks.append((f"pass2_{s}_done", self._mark()))
        pairs, count = eng.shardx_join(plan, out)
        if marks is not None:
            marks.append(("join_done", self._mark()))
        return pairs, count, (lay[self.build_rel][3], lay[self.probe_rel][3])

    @staticmethod
    def timeline(marks):
        """[(label, ms since 'start')] from the events a profiled step collected (after a synchronize)."""
        t0 = marks[0][1]
        return [(name, round(t0.elapsed_time(ev), 3)) for name, ev in marks]
