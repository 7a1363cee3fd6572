"""Host-side mirror of the reference's join surface over the C ABI (include/rhj.h).

Reference surface (all citations are file:line in the reference tree):

* ``struct tuple {key = row id, payload = join value}``            structs.h:33-36
* ``struct relation {tuple *tuples; uint64_t num_tuples}``         structs.h:38-49
* ``Result::multiRadixHashJoin(js, relR, relS)``                    Result.h:30, Result.cpp:90-124
* ``Result::isEmpty()``, page list of 8191 ``key_tuple`` pairs      Result.h:9-38, Result.cpp:10-35

Here a relation on the device is a ``torch.int64`` CUDA tensor of shape ``(n, 2)`` whose rows are
``(row id, value)`` bit patterns of the reference's u64 fields; a result is a ``(count, 2)`` tensor of
``(rowidR, rowidS)``.  On the host the same layouts are numpy structured arrays (``TUPLE_DTYPE``,
``PAIR_DTYPE``).  PyTorch only provides memory and streams; every operation calls ``librhj.so``.
"""
import ctypes

import numpy as np

from . import _lib

TUPLE_DTYPE = np.dtype([("key", "<u8"), ("payload", "<u8")])
PAIR_DTYPE = np.dtype([("keyR", "<u8"), ("keyS", "<u8")])

EMIT_FUSED = 0
EMIT_COUNT_THEN_WRITE = 1
DIGIT_RAW = 0
DIGIT_HASH = 1

PAGE_CAPACITY = (128 * 1024 - 8) // 16  # Result.cpp:7,11 -> 8191 pairs per page

_STATUS = {1: "RHJ_ERR_CUDA", 2: "RHJ_ERR_ARG", 3: "RHJ_ERR_NOMEM", 4: "RHJ_ERR_CAPACITY", 5: "RHJ_ERR_STATE",
           6: "RHJ_ERR_NO_DEVICE"}


class RhjError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{_STATUS.get(code, code)}: {msg}")
        self.code = code


def _torch():
    import torch
    return torch


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() else ctypes.c_void_p(0)


def _check_rel(t):
    torch = _torch()
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.int64 and t.dim() == 2 and t.shape[1] == 2
            and t.is_contiguous()):
        raise TypeError("a device relation is a contiguous CUDA int64 tensor of shape (n, 2): (row id, value)")
    return t.shape[0]


class RadixHashJoin:
    """One join context (``rhj_ctx``) on one GPU.  Not thread-safe: one per query thread."""

    def __init__(self, device=0):
        self._lib = _lib.load()
        self._ctx = ctypes.c_void_p()
        self.device = int(device)
        rc = self._lib.rhj_create(self.device, ctypes.byref(self._ctx))
        if rc != 0:
            self._ctx = ctypes.c_void_p()
            raise RhjError(rc, "rhj_create failed (librhj.so needs an sm_100 GPU; there is no CPU fallback)")

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._lib.rhj_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers -------------------------------------------------------------------------------
    def _ck(self, rc, ok=(0,)):
        if rc not in ok:
            raise RhjError(rc, self._lib.rhj_last_error(self._ctx).decode())
        return rc

    def _stream(self, stream):
        torch = _torch()
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        return ctypes.c_void_p(stream.cuda_stream)

    def reserve(self, nR, nS):
        self._ck(self._lib.rhj_reserve(self._ctx, nR, nS))

    def workspace_bytes(self):
        return int(self._lib.rhj_workspace_bytes(self._ctx))

    def last_plan(self):
        info = _lib.PlanInfo()
        self._ck(self._lib.rhj_last_plan(self._ctx, ctypes.byref(info)))
        return {f: int(getattr(info, f)) for f, _ in _lib.PlanInfo._fields_}

    PHASES = ("hist1", "scan1", "scatter1", "hist2", "scan2", "scatter2", "plan", "join", "scan_items", "join_write")

    def set_profiling(self, on):
        self._ck(self._lib.rhj_set_profiling(self._ctx, 1 if on else 0))

    def last_phase_ms(self):
        """{phase: device milliseconds} of the last join (needs set_profiling(True))."""
        ms = (ctypes.c_float * len(self.PHASES))()
        self._ck(self._lib.rhj_last_phase_ms(self._ctx, ms))
        return {n: float(ms[i]) for i, n in enumerate(self.PHASES)}

    # ---- the join ------------------------------------------------------------------------------
    def join_device(self, R, S, out=None, capacity=None, emit=EMIT_FUSED, stream=None):
        """``Result::multiRadixHashJoin`` on device-resident relations.

        Returns ``(pairs, count)`` where ``pairs`` is ``out[:count]`` (a view).  With the fused emitter
        ``out`` (or ``capacity``) must be able to hold the result; RhjError(RHJ_ERR_CAPACITY) carries
        the needed count in ``.needed`` otherwise.
        """
        torch = _torch()
        nR, nS = _check_rel(R), _check_rel(S)
        if out is None:
            if capacity is None:
                capacity = max(nR, nS)
            out = torch.empty((max(int(capacity), 1), 2), dtype=torch.int64, device=R.device)
        capacity = out.shape[0]
        cnt = ctypes.c_uint64()
        rc = self._lib.rhj_join_device(self._ctx, _ptr(R), nR, _ptr(S), nS, _ptr(out), capacity, ctypes.byref(cnt),
                                       int(emit), self._stream(stream))
        if rc == 4:
            e = RhjError(rc, self._lib.rhj_last_error(self._ctx).decode())
            e.needed = int(cnt.value)
            raise e
        self._ck(rc)
        return out[:cnt.value], int(cnt.value)

    def join_count_device(self, R, S, stream=None):
        """Partition + count pass of the two-pass emitter; returns the exact result size."""
        nR, nS = _check_rel(R), _check_rel(S)
        cnt = ctypes.c_uint64()
        self._ck(self._lib.rhj_join_count_device(self._ctx, _ptr(R), nR, _ptr(S), nS, ctypes.byref(cnt),
                                                 self._stream(stream)))
        return int(cnt.value)

    def join_write_device(self, out, stream=None):
        """Write pass of the two-pass emitter into ``out`` (sized from join_count_device)."""
        self._ck(self._lib.rhj_join_write_device(self._ctx, _ptr(out), out.shape[0], self._stream(stream)))
        return out

    def join_host(self, R, S):
        """Host relations (numpy TUPLE_DTYPE, or pinned int64 torch tensors of shape (n, 2)) in,
        numpy PAIR_DTYPE array out -- the work ``Result::multiRadixHashJoin`` does for a query thread,
        H2D and D2H copies included.  The returned array is a copy of the context's pinned buffer
        unless ``copy=False`` semantics are wanted (see join_host_view)."""
        view, n = self.join_host_view(R, S)
        return view.copy() if n else np.empty(0, dtype=PAIR_DTYPE)

    def join_host_view(self, R, S):
        """Like join_host but returns a zero-copy numpy view of the context-owned pinned result
        (valid until the next call on this context) and the pair count."""
        pr, nr, keep_r = _host_ptr(R)
        ps, ns, keep_s = _host_ptr(S)
        out = ctypes.c_void_p()
        cnt = ctypes.c_uint64()
        self._ck(self._lib.rhj_join_host(self._ctx, pr, nr, ps, ns, ctypes.byref(out), ctypes.byref(cnt)))
        n = int(cnt.value)
        if n == 0:
            return np.empty(0, dtype=PAIR_DTYPE), 0
        buf = (ctypes.c_uint64 * (2 * n)).from_address(out.value)
        return np.frombuffer(buf, dtype=PAIR_DTYPE), n

    # ---- the steps -----------------------------------------------------------------------------
    def histogram(self, T, bits, shift=0, kind=DIGIT_RAW, stream=None):
        """HistogramJob::run + global sum (JobScheduler.cpp:149-155, structs.cpp:168-173)."""
        torch = _torch()
        n = _check_rel(T)
        hist = torch.empty(1 << bits, dtype=torch.int64, device=T.device)
        self._ck(self._lib.rhj_histogram_device(self._ctx, _ptr(T), n, bits, shift, kind, _ptr(hist),
                                                self._stream(stream)))
        return hist

    def partition(self, T, bits, shift=0, kind=DIGIT_RAW, stream=None):
        """relation_info::hash_relation (structs.cpp:144-204): (partitioned tuples, offsets[2^bits+1])."""
        torch = _torch()
        n = _check_rel(T)
        out = torch.empty_like(T)
        off = torch.empty((1 << bits) + 1, dtype=torch.int64, device=T.device)
        self._ck(self._lib.rhj_partition_device(self._ctx, _ptr(T), n, bits, shift, kind, _ptr(out), _ptr(off),
                                                self._stream(stream)))
        return out, off

    def shuffle_partition(self, T, world, stream=None):
        """Groups tuples by destination rank for the multi-GPU exchange: (grouped tuples, counts list)."""
        torch = _torch()
        n = _check_rel(T)
        out = torch.empty_like(T)
        counts = (ctypes.c_uint64 * world)()
        self._ck(self._lib.rhj_shuffle_partition_device(self._ctx, _ptr(T), n, world, _ptr(out), counts,
                                                        self._stream(stream)))
        return out, [int(c) for c in counts]

    # ---- multi-GPU: radix plan + the exact (histogram) exchange (include/rhj.h, rhj_shard_plan_make, rhj_shardx_*) ----
    def shard_plan(self, nR_global, nS_global, world):
        plan = _lib.ShardPlan()
        self._ck(self._lib.rhj_shard_plan_make(nR_global, nS_global, world, ctypes.byref(plan)))
        return plan

    def shardx_begin(self, plan, stream=None):
        self._ck(self._lib.rhj_shardx_begin(self._ctx, ctypes.byref(plan), self._stream(stream)))

    def shardx_pass1(self, plan, rel, T, stage, hist, stream=None):
        """pass 1 of relation rel into the local staging tensor (ordered by destination rank, partition)"""
        n = _check_rel(T)
        self._ck(self._lib.rhj_shardx_pass1_device(self._ctx, ctypes.byref(plan), rel, _ptr(T), n, _ptr(stage), _ptr(hist),
                                                   self._stream(stream)))

    def shardx_layout(self, plan, rank, rel, all_hist, stream=None):
        """(send_off[d], send_cnt[d], dst_off[d], recv_total, recv_max) from the all-gathered histograms; recv_max = the
        largest recv_total of any rank (the same number on every rank)"""
        W = plan.world
        so, sc, do = (ctypes.c_uint64 * W)(), (ctypes.c_uint64 * W)(), (ctypes.c_uint64 * W)()
        tot, worst = ctypes.c_uint64(), ctypes.c_uint64()
        self._ck(self._lib.rhj_shardx_layout_device(self._ctx, ctypes.byref(plan), rank, rel, _ptr(all_hist), so, sc, do,
                                                    ctypes.byref(tot), ctypes.byref(worst), self._stream(stream)))
        return [int(v) for v in so], [int(v) for v in sc], [int(v) for v in do], int(tot.value), int(worst.value)

    def shardx_pass2(self, plan, rel, recv, stream=None):
        n = _check_rel(recv)
        self._ck(self._lib.rhj_shardx_pass2_device(self._ctx, ctypes.byref(plan), rel, _ptr(recv), n, self._stream(stream)))

    def shardx_join(self, plan, out, stream=None):
        cnt = ctypes.c_uint64()
        rc = self._lib.rhj_shardx_join_device(self._ctx, ctypes.byref(plan), _ptr(out), out.shape[0], ctypes.byref(cnt),
                                              self._stream(stream))
        if rc == 4:
            e = RhjError(rc, self._lib.rhj_last_error(self._ctx).decode())
            e.needed = int(cnt.value)
            raise e
        self._ck(rc)
        return out[:cnt.value], int(cnt.value)

    # ---- multi-GPU: pipelined exchange (include/rhj.h, rhj_pipe_*) --------------------------------
    def pipe_cfg(self, plan, rank, chunks, nR_local_max, nS_local_max, sym_ptrs=None, ship_ctas=0, wire_bytes=16):
        cfg = _lib.PipeCfg()
        cfg.world, cfg.rank, cfg.chunks, cfg.ship_ctas, cfg.wire_bytes = plan.world, rank, chunks, ship_ctas, wire_bytes
        cfg.nR_local_max, cfg.nS_local_max = nR_local_max, nS_local_max
        for i, p in enumerate(sym_ptrs or []):
            cfg.sym[i] = p
        return cfg

    def pipe_sym_bytes(self, plan, cfg):
        n = int(self._lib.rhj_pipe_sym_bytes(ctypes.byref(plan), ctypes.byref(cfg)))
        if n == 0:
            raise RhjError(2, "rhj_pipe_sym_bytes: bad plan / configuration")
        return n

    def pipe_open(self, plan, cfg):
        self._ck(self._lib.rhj_pipe_open(self._ctx, ctypes.byref(plan), ctypes.byref(cfg)))

    def pipe_begin(self, epoch, stream=None):
        self._ck(self._lib.rhj_pipe_begin(self._ctx, epoch, self._stream(stream)))

    def pipe_pass1(self, rel, chunk, rows, stream=None):
        n = _check_rel(rows)
        self._ck(self._lib.rhj_pipe_pass1_device(self._ctx, rel, chunk, _ptr(rows), n, self._stream(stream)))

    def pipe_ship(self, rel, chunk, stream=None):
        self._ck(self._lib.rhj_pipe_ship_device(self._ctx, rel, chunk, self._stream(stream)))

    def pipe_pass2(self, rel, chunk, stream=None):
        self._ck(self._lib.rhj_pipe_pass2_device(self._ctx, rel, chunk, self._stream(stream)))

    def pipe_post(self, stream=None):
        self._ck(self._lib.rhj_pipe_post_device(self._ctx, self._stream(stream)))

    def pipe_join(self, out, stream=None):
        """(pairs, count, status): status != 0 (RHJ_PIPE_* bits) means the step must be redone exactly"""
        cnt, status = ctypes.c_uint64(), ctypes.c_uint32()
        rc = self._lib.rhj_pipe_join_device(self._ctx, _ptr(out), out.shape[0], ctypes.byref(cnt), ctypes.byref(status),
                                            self._stream(stream))
        if rc == 4:
            e = RhjError(rc, self._lib.rhj_last_error(self._ctx).decode())
            e.needed = int(cnt.value)
            raise e
        self._ck(rc)
        return out[:cnt.value], int(cnt.value), int(status.value)

    # ---- neighbours on the query path ----------------------------------------------------------
    def filter(self, col, op, constant, rowids=None, stream=None):
        """Query::run_filters predicate (Query.cpp:94-146): surviving row ids, input order kept."""
        torch = _torch()
        n_in = rowids.numel() if rowids is not None else col.numel()
        out = torch.empty(max(n_in, 1), dtype=torch.int64, device=col.device)
        cnt = ctypes.c_uint64()
        if isinstance(op, str):
            op = ord(op)
        self._ck(self._lib.rhj_filter_u64_device(self._ctx, _ptr(col), _ptr(rowids) if rowids is not None else None,
                                                 n_in, op, ctypes.c_uint64(int(constant) & (2**64 - 1)), _ptr(out),
                                                 ctypes.byref(cnt), self._stream(stream)))
        return out[:cnt.value]

    def gather_tuples(self, col, rowids, stream=None):
        """relation::foo (structs.cpp:217-226): tuples[i] = (rowids[i], col[rowids[i]])."""
        torch = _torch()
        n = rowids.numel()
        out = torch.empty((n, 2), dtype=torch.int64, device=col.device)
        self._ck(self._lib.rhj_gather_tuples_device(self._ctx, _ptr(col), _ptr(rowids), n, _ptr(out),
                                                    self._stream(stream)))
        return out

    def gather_sum(self, col, rowids, stream=None):
        """column_proj (Query.cpp:66-74): sum of col[rowids] mod 2^64 (python int)."""
        s = ctypes.c_uint64()
        self._ck(self._lib.rhj_gather_sum_u64_device(self._ctx, _ptr(col), _ptr(rowids), rowids.numel(),
                                                     ctypes.byref(s), self._stream(stream)))
        return int(s.value)

    def unique_rowids(self, rowids, n_rows, stream=None):
        """create_relation's de-duplication (structs.cpp:238-241): the distinct row ids of a device column, ascending."""
        torch = _torch()
        n = rowids.numel()
        out = torch.empty(max(min(n, n_rows), 1), dtype=torch.int64, device=rowids.device)
        cnt = ctypes.c_uint64()
        self._ck(self._lib.rhj_unique_rowids_device(self._ctx, _ptr(rowids), n, n_rows, _ptr(out), ctypes.byref(cnt),
                                                    self._stream(stream)))
        return out[:cnt.value]

    def query_execute(self, tables, filters, joins, projs, relations, reorder_joins=False):
        """Query::execute (Query.cpp:204-211) on the device.  `relations` = list of relLists (each a list of contiguous
        numpy u64 host columns, which must stay alive: they are uploaded once and cached by address); `tables` = the
        relList index of every binding; filters (binding, column, op, constant), joins (b1, c1, b2, c2), projs (binding,
        column) as in Query.h:8-33.  Returns (sums or None when the result is empty, stats dict)."""
        keep = []
        binds = (_lib.QRelation * len(tables))()
        for i, t in enumerate(tables):
            cols = relations[t]
            arr = (ctypes.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
            keep.append(arr)
            binds[i].columns = ctypes.cast(arr, ctypes.POINTER(ctypes.c_void_p))
            binds[i].num_tuples = len(cols[0])
            binds[i].num_columns = len(cols)
        fl = (_lib.QFilter * max(len(filters), 1))()
        for i, (b, c, op, k) in enumerate(filters):
            fl[i].binding, fl[i].column, fl[i].op, fl[i].constant = b, c, ord(op) if isinstance(op, str) else op, int(k)
        jn = (_lib.QJoin * max(len(joins), 1))()
        for i, (b1, c1, b2, c2) in enumerate(joins):
            jn[i].binding1, jn[i].column1, jn[i].binding2, jn[i].column2 = b1, c1, b2, c2
        pj = (_lib.QProj * max(len(projs), 1))()
        for i, (b, c) in enumerate(projs):
            pj[i].binding, pj[i].column = b, c
        d = _lib.QueryDesc(len(tables), len(filters), len(joins), len(projs), 1 if reorder_joins else 0, 0, binds, fl, jn, pj)
        sums = (ctypes.c_uint64 * max(len(projs), 1))()
        empty = ctypes.c_int()
        st = _lib.QueryStats()
        self._ck(self._lib.rhj_query_execute(self._ctx, ctypes.byref(d), sums, ctypes.byref(empty), ctypes.byref(st)))
        stats = {f: int(getattr(st, f)) for f, _ in _lib.QueryStats._fields_}
        return (None if empty.value else [int(sums[i]) for i in range(len(projs))]), stats

    def column_cache_clear(self):
        self._lib.rhj_column_cache_clear()

    def intermediate_expand_host(self, match_col, pairs, match_on_S, cols):
        """update_intermediate case 2 (intermediate.cpp:108-125,162-170) as join + gather on the GPU.
        match_col / cols: numpy u64 host columns of the old intermediate; pairs: PAIR_DTYPE join result.
        Returns (carried columns, new binding's column) as numpy copies."""
        match_col = np.ascontiguousarray(match_col, dtype=np.uint64)
        pairs = np.ascontiguousarray(pairs, dtype=PAIR_DTYPE)
        cols = [np.ascontiguousarray(c, dtype=np.uint64) for c in cols]
        ptrs = (ctypes.c_void_p * max(len(cols), 1))(*[c.ctypes.data for c in cols])
        outs = (ctypes.c_void_p * (len(cols) + 1))()
        rows = ctypes.c_uint64()
        self._ck(self._lib.rhj_intermediate_expand_host(self._ctx, match_col.ctypes.data, len(match_col),
                                                        pairs.ctypes.data, len(pairs), 1 if match_on_S else 0, ptrs,
                                                        len(cols), outs, ctypes.byref(rows)))
        m = int(rows.value)
        take = lambda p: np.frombuffer((ctypes.c_uint64 * m).from_address(p), dtype=np.uint64).copy() if m else \
            np.empty(0, dtype=np.uint64)
        return [take(outs[i]) for i in range(len(cols))], take(outs[len(cols)])

    def intermediate_filter_host(self, col1, col2, pairs, cols):
        """update_intermediate case 3 (intermediate.cpp:130-138,171-180): rows whose (col1, col2) row-id
        pair is in the join result.  Returns the surviving rows of every carried column."""
        col1 = np.ascontiguousarray(col1, dtype=np.uint64)
        col2 = np.ascontiguousarray(col2, dtype=np.uint64)
        pairs = np.ascontiguousarray(pairs, dtype=PAIR_DTYPE)
        cols = [np.ascontiguousarray(c, dtype=np.uint64) for c in cols]
        ptrs = (ctypes.c_void_p * max(len(cols), 1))(*[c.ctypes.data for c in cols])
        outs = (ctypes.c_void_p * max(len(cols), 1))()
        rows = ctypes.c_uint64()
        self._ck(self._lib.rhj_intermediate_filter_host(self._ctx, col1.ctypes.data, col2.ctypes.data, len(col1),
                                                        pairs.ctypes.data, len(pairs), ptrs, len(cols), outs,
                                                        ctypes.byref(rows)))
        m = int(rows.value)
        take = lambda p: np.frombuffer((ctypes.c_uint64 * m).from_address(p), dtype=np.uint64).copy() if m else \
            np.empty(0, dtype=np.uint64)
        return [take(outs[i]) for i in range(len(cols))]

    def pairs_digest(self, pairs, stream=None):
        """(count, sum, xor) of mix64(keyR*0x100000001b3 + keyS): order-independent multiset digest."""
        s, x = ctypes.c_uint64(), ctypes.c_uint64()
        n = pairs.shape[0]
        self._ck(self._lib.rhj_pairs_digest_device(self._ctx, _ptr(pairs), n, ctypes.byref(s), ctypes.byref(x),
                                                   self._stream(stream)))
        return n, int(s.value), int(x.value)


def _host_ptr(a):
    """(void*, n, keepalive) of a host relation: numpy TUPLE_DTYPE / (n,2) u64|i64, or a CPU torch tensor."""
    if isinstance(a, np.ndarray):
        if a.dtype != TUPLE_DTYPE:
            if a.ndim == 2 and a.shape[1] == 2 and a.dtype in (np.uint64, np.int64):
                pass
            else:
                raise TypeError("host relation must be TUPLE_DTYPE or an (n, 2) 64-bit integer array")
        a = np.ascontiguousarray(a)
        return ctypes.c_void_p(a.ctypes.data), a.shape[0], a
    torch = _torch()
    if isinstance(a, torch.Tensor) and not a.is_cuda and a.dtype == torch.int64 and a.dim() == 2 and a.shape[1] == 2:
        a = a.contiguous()
        return ctypes.c_void_p(a.data_ptr()), a.shape[0], a
    raise TypeError("unsupported host relation type")


# ---- reference-shaped objects ---------------------------------------------------------------------
class Relation:
    """``struct relation`` (structs.h:38-49): host tuples {key = row id, payload = value}."""

    def __init__(self, tuples):
        self.tuples = np.ascontiguousarray(tuples, dtype=TUPLE_DTYPE)

    @property
    def num_tuples(self):
        return len(self.tuples)

    @classmethod
    def create_relation(cls, column, rowids):
        """relation::create_relation / foo (structs.cpp:217-243) for an already de-duplicated row-id list."""
        rowids = np.asarray(rowids, dtype=np.uint64)
        t = np.empty(len(rowids), dtype=TUPLE_DTYPE)
        t["key"] = rowids
        t["payload"] = np.asarray(column, dtype=np.uint64)[rowids.astype(np.int64)]
        return cls(t)


class Result:
    """``struct Result`` (Result.h:19-38) with the join done on the GPU.

    ``multiRadixHashJoin`` has the reference's signature minus the JobScheduler (its pthread fan-out
    is what the CUDA kernels replace).  ``pages()`` yields the result the way a consumer walks the
    reference's page list: newest page first, only the head page partial (intermediate.cpp:151-160).
    """

    capacity = PAGE_CAPACITY

    def __init__(self, engine=None):
        self._engine = engine
        self.pairs = np.empty(0, dtype=PAIR_DTYPE)

    def multiRadixHashJoin(self, relR, relS):
        if self._engine is None:
            self._engine = RadixHashJoin(0)
        self.pairs = self._engine.join_host(relR.tuples, relS.tuples)
        return self

    def isEmpty(self):
        return len(self.pairs) == 0  # head == nullptr, Result.cpp:16-18

    @property
    def size(self):
        """Pairs in the head page (Result.size); == capacity when empty (Result.cpp:12)."""
        n = len(self.pairs)
        if n == 0:
            return self.capacity
        r = n % self.capacity
        return r if r else self.capacity

    def pages(self):
        n = len(self.pairs)
        npages = (n + self.capacity - 1) // self.capacity
        for pg in range(npages - 1, -1, -1):
            yield self.pairs[pg * self.capacity:min(n, (pg + 1) * self.capacity)]
