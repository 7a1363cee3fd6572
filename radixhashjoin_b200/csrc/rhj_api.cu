// rhj_api.cu -- host side of librhj.so: the C ABI declared in include/rhj.h.
//
// Replaces, on the join path, the HistogramJob / PartitionJob / JoinJob fan-out that
// Result::multiRadixHashJoin (Result.cpp:90-124) and relation_info::hash_relation
// (structs.cpp:144-204) schedule on the reference's pthread pool with a fixed sequence of kernel
// launches on one CUDA stream.  No host synchronisation happens between the launches of one
// join; sizes that depend on the data (partition sizes, work items, match counts) stay on the
// device.  There is no CPU fallback anywhere in this file.
#include "rhj_ctx.cuh"
#include "rhj_kernels.cuh"

namespace {

struct Plan {
    u64 nB, nP;
    int build_is_S;
    int bits, b1, b2;
    u32 nparts;
};

Plan make_plan(u64 nR, u64 nS) {
    Plan p{};
    p.build_is_S = nS < nR;  // build on the smaller relation (JobScheduler.cpp:187-190 does it per bucket)
    p.nB = p.build_is_S ? nS : nR;
    p.nP = p.build_is_S ? nR : nS;
    int bits = 0;
    if (p.nB > kBuildCap) {
        while (bits < 2 * kMaxBitsPerPass && (p.nB >> bits) > kTargetBuildPerPart) ++bits;
    }
    p.bits = bits;
    p.b1 = bits <= kMaxBitsPerPass ? bits : (bits + 1) / 2;
    p.b2 = bits - p.b1;
    p.nparts = 1u << bits;
    return p;
}

inline u32 tiles_of(u64 n) { return (u32) ((n + kTile - 1) / kTile); }

// ---- kernel launch helpers -----------------------------------------------------------------------
template <typename K>
cudaError_t set_smem(K k, size_t bytes) {
    return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes);
}

constexpr size_t kScatterSmem = (size_t) kTile * sizeof(Tup);
constexpr size_t kJoinSmem = kJoinSmemBytes;

int launch_hist(rhj_ctx *ctx, cudaStream_t st, const PartArgs &a, int kind, bool seg) {
    u32 total = a.rel[0].ntiles + a.rel[1].ntiles;
    if (!total) return RHJ_OK;
    u32 grid = std::min<u32>(total, (u32) ctx->num_sms * 4);
    bool agg = ctx->hist_agg;
#define HIST(K, S)                                                            \
    do {                                                                      \
        if (agg) k_hist<K, S, true><<<grid, kPartThreads, 0, st>>>(a);        \
        else k_hist<K, S, false><<<grid, kPartThreads, 0, st>>>(a);           \
    } while (0)
    if (kind == kDigitRaw) { if (seg) HIST(kDigitRaw, true); else HIST(kDigitRaw, false); }
    else if (kind == kDigitHash) { if (seg) HIST(kDigitHash, true); else HIST(kDigitHash, false); }
    else { if (seg) HIST(kDigitRank, true); else HIST(kDigitRank, false); }
#undef HIST
    CK(cudaGetLastError());
    ctx->info.kernel_launches++;
    return RHJ_OK;
}

template <int K, bool S, int W>
cudaError_t launch_scatter_t(cudaStream_t st, const PartArgs &a, u32 grid) {
    const size_t smem = kScatterSmem;
    cudaError_t e = set_smem(k_scatter<K, S, W>, smem);
    if (e != cudaSuccess) return e;
    k_scatter<K, S, W><<<grid, kPartThreads, smem, st>>>(a);
    return cudaGetLastError();
}

int launch_scatter(rhj_ctx *ctx, cudaStream_t st, const PartArgs &a, int kind, bool seg) {
    u32 grid = a.rel[0].ntiles + a.rel[1].ntiles;
    if (!grid) return RHJ_OK;
    const int w = ctx->scatter_mode;
    cudaError_t e;
#define SC(K, S) (w == 1 ? launch_scatter_t<K, S, kWriteBulk>(st, a, grid) : launch_scatter_t<K, S, kWriteStaged>(st, a, grid))
    if (kind == kDigitRaw) e = seg ? SC(kDigitRaw, true) : SC(kDigitRaw, false);
    else if (kind == kDigitHash) e = seg ? SC(kDigitHash, true) : SC(kDigitHash, false);
    else e = seg ? SC(kDigitRank, true) : SC(kDigitRank, false);
#undef SC
    CK(e);
    ctx->info.kernel_launches++;
    return RHJ_OK;
}

template <int MODE>
int launch_join(rhj_ctx *ctx, cudaStream_t st, const JoinArgs &a, u32 item_cap) {
    u32 grid = std::min<u32>(std::max<u32>(item_cap, 1), (u32) ctx->num_sms * 2);
    CK(set_smem(k_join<MODE>, kJoinSmem));
    k_join<MODE><<<grid, kJoinThreads, kJoinSmem, st>>>(a);
    CK(cudaGetLastError());
    ctx->info.kernel_launches++;
    return RHJ_OK;
}

// Layout of the two metadata buffers for a plan.
struct Meta {
    u64 *hist1[2], *hist2[2], *scalars;                  // zeroed
    u64 *off1[2], *cur1[2], *off2[2], *cur2[2];          // written by the scans
    u32 *tile0[2];
    size_t zero_bytes;
};

int layout_meta(rhj_ctx *ctx, u32 nparts, Meta &m) {
    size_t zero_u64 = 2 * (size_t) kMaxDigits + 2 * (size_t) nparts + kScCount;
    size_t meta_u64 = 2 * (size_t) (kMaxDigits + 1) + 2 * (size_t) kMaxDigits + 2 * (size_t) (nparts + 1) +
                      2 * (size_t) nparts + (size_t) (kMaxDigits + 2);
    int rc;
    if ((rc = ensure(ctx, ctx->zero, zero_u64 * 8))) return rc;
    if ((rc = ensure(ctx, ctx->meta, meta_u64 * 8))) return rc;
    u64 *z = (u64 *) ctx->zero.p;
    m.hist1[0] = z; z += kMaxDigits;
    m.hist1[1] = z; z += kMaxDigits;
    m.hist2[0] = z; z += nparts;
    m.hist2[1] = z; z += nparts;
    m.scalars = z;
    m.zero_bytes = zero_u64 * 8;
    u64 *q = (u64 *) ctx->meta.p;
    m.off1[0] = q; q += kMaxDigits + 1;
    m.off1[1] = q; q += kMaxDigits + 1;
    m.cur1[0] = q; q += kMaxDigits;
    m.cur1[1] = q; q += kMaxDigits;
    m.off2[0] = q; q += nparts + 1;
    m.off2[1] = q; q += nparts + 1;
    m.cur2[0] = q; q += nparts;
    m.cur2[1] = q; q += nparts;
    m.tile0[0] = (u32 *) q;
    m.tile0[1] = m.tile0[0] + (kMaxDigits + 1);
    return RHJ_OK;
}

u64 *scalars_of(rhj_ctx *ctx, u32 nparts) {
    return (u64 *) ctx->zero.p + 2 * (size_t) kMaxDigits + 2 * (size_t) nparts;
}

// Partition both relations on `bits` hash bits (one or two passes) and build the work-item list.
// Leaves ctx->cur describing the partitioned relations.  Enqueues only; no host sync.
int partition_and_plan(rhj_ctx *ctx, cudaStream_t st, const Tup *dR, u64 nR, const Tup *dS, u64 nS) {
    Plan pl = make_plan(nR, nS);
    const Tup *inB = pl.build_is_S ? dS : dR;
    const Tup *inP = pl.build_is_S ? dR : dS;
    ctx->info = rhj_plan_info{};
    ctx->info.bits_total = pl.bits;
    ctx->info.bits_pass1 = pl.b1;
    ctx->info.bits_pass2 = pl.b2;
    ctx->info.build_is_S = pl.build_is_S;
    ctx->info.n_partitions = pl.nparts;
    ctx->cur.valid = false;
    ctx->cur.counted = false;
    ctx->nmarks = 0;

    Meta m;
    int rc;
    if ((rc = layout_meta(ctx, pl.nparts, m))) return rc;
    CK(cudaMemsetAsync(ctx->zero.p, 0, m.zero_bytes, st));

    const u64 ntot = pl.nB + pl.nP;
    const Tup *finB, *finP;
    const u64 *offB, *offP;
    bool planned = false;
    u32 item_cap = 0;

    if (pl.bits == 0) {
        finB = inB;
        finP = inP;
        k_set_single_part<<<1, 1, 0, st>>>(m.off2[0], pl.nB, m.off2[1], pl.nP);
        CK(cudaGetLastError());
        ctx->info.kernel_launches++;
        offB = m.off2[0];
        offP = m.off2[1];
    } else {
        if ((rc = ensure(ctx, ctx->bufA, ntot * sizeof(Tup)))) return rc;
        Tup *A = (Tup *) ctx->bufA.p;
        // ---- pass 1: top b1 bits of hash32 ----
        PartArgs a{};
        a.shift = 32 - pl.b1;
        a.mask = (1u << pl.b1) - 1;
        a.ndig = 1u << pl.b1;
        a.rel[0] = PartRel{inB, A, pl.nB, m.hist1[0], m.cur1[0], nullptr, nullptr, 1, tiles_of(pl.nB)};
        a.rel[1] = PartRel{inP, A + pl.nB, pl.nP, m.hist1[1], m.cur1[1], nullptr, nullptr, 1, tiles_of(pl.nP)};
        mark(ctx, st, RHJ_PHASE_HIST1);
        if ((rc = launch_hist(ctx, st, a, kDigitHash, false))) return rc;
        mark(ctx, st, RHJ_PHASE_SCAN1);
        ScanDigitsArgs sd{};
        for (int i = 0; i < 2; ++i) {
            sd.hist[i] = m.hist1[i];
            sd.off[i] = m.off1[i];
            sd.cursor[i] = m.cur1[i];
            sd.tile0[i] = pl.b2 ? m.tile0[i] : nullptr;
        }
        sd.ndig = a.ndig;
        k_scan_digits<<<2, kMaxDigits, 0, st>>>(sd);
        CK(cudaGetLastError());
        ctx->info.kernel_launches++;
        mark(ctx, st, RHJ_PHASE_SCATTER1);
        if ((rc = launch_scatter(ctx, st, a, kDigitHash, false))) return rc;

        if (pl.b2 == 0) {
            finB = A;
            finP = A + pl.nB;
            offB = m.off1[0];
            offP = m.off1[1];
        } else {
            // ---- pass 2: next b2 bits, inside every pass-1 partition ----
            if ((rc = ensure(ctx, ctx->bufB, ntot * sizeof(Tup)))) return rc;
            Tup *B = (Tup *) ctx->bufB.p;
            PartArgs b{};
            b.shift = 32 - pl.bits;
            b.mask = (1u << pl.b2) - 1;
            b.ndig = 1u << pl.b2;
            const u32 nseg = 1u << pl.b1;
            b.rel[0] = PartRel{A, B, pl.nB, m.hist2[0], m.cur2[0], m.off1[0], m.tile0[0], nseg, tiles_of(pl.nB) + nseg};
            b.rel[1] = PartRel{A + pl.nB, B + pl.nB, pl.nP, m.hist2[1], m.cur2[1], m.off1[1], m.tile0[1], nseg,
                               tiles_of(pl.nP) + nseg};
            mark(ctx, st, RHJ_PHASE_HIST2);
            if ((rc = launch_hist(ctx, st, b, kDigitHash, true))) return rc;
            mark(ctx, st, RHJ_PHASE_SCAN2);
            // offsets of all 2^bits sub-partitions + the work-item list, one CTA per pass-1 partition
            u64 cap64 = (u64) pl.nparts + pl.nP / kProbeChunk + 2;
            if (cap64 > 0x7fffffffull) return fail(ctx, RHJ_ERR_ARG, "relation too large for the work-item table");
            item_cap = (u32) cap64;
            if ((rc = ensure(ctx, ctx->items, (size_t) item_cap * sizeof(Item)))) return rc;
            ScanPlanArgs sp{};
            for (int i = 0; i < 2; ++i) {
                sp.hist2[i] = m.hist2[i];
                sp.off1[i] = m.off1[i];
                sp.off2[i] = m.off2[i];
                sp.cursor2[i] = m.cur2[i];
            }
            sp.nseg = nseg;
            sp.ndig = b.ndig;
            sp.items = (Item *) ctx->items.p;
            sp.item_cap = item_cap;
            sp.nitems = (u32 *) (m.scalars + kScNItems);
            sp.err = (u32 *) (m.scalars + kScErr);
            k_scan_parts_plan<<<nseg, kMaxDigits, 0, st>>>(sp);
            CK(cudaGetLastError());
            ctx->info.kernel_launches++;
            planned = true;
            // segment offsets inside A are relative to each relation's base: rel.in already points there
            mark(ctx, st, RHJ_PHASE_SCATTER2);
            if ((rc = launch_scatter(ctx, st, b, kDigitHash, true))) return rc;
            finB = B;
            finP = B + pl.nB;
            offB = m.off2[0];
            offP = m.off2[1];
        }
    }

    // ---- plan: work items (two-pass plans were planned by k_scan_parts_plan) ----
    if (!planned) {
        mark(ctx, st, RHJ_PHASE_PLAN);
        u64 cap64 = (u64) pl.nparts + pl.nP / kProbeChunk + 2;
        if (cap64 > 0x7fffffffull) return fail(ctx, RHJ_ERR_ARG, "relation too large for the work-item table");
        item_cap = (u32) cap64;
        if ((rc = ensure(ctx, ctx->items, (size_t) item_cap * sizeof(Item)))) return rc;
        PlanArgs pa{};
        pa.offB = offB;
        pa.offP = offP;
        pa.nparts = pl.nparts;
        pa.items = (Item *) ctx->items.p;
        pa.item_cap = item_cap;
        pa.nitems = (u32 *) (m.scalars + kScNItems);
        pa.err = (u32 *) (m.scalars + kScErr);
        k_plan<<<1, 1024, 0, st>>>(pa);
        CK(cudaGetLastError());
        ctx->info.kernel_launches++;
    }

    ctx->cur.valid = true;
    ctx->cur.build = finB;
    ctx->cur.probe = finP;
    ctx->cur.offB = offB;
    ctx->cur.offP = offP;
    ctx->cur.nparts = pl.nparts;
    ctx->cur.item_cap = item_cap;
    ctx->cur.build_is_S = pl.build_is_S;
    return RHJ_OK;
}

JoinArgs join_args(rhj_ctx *ctx, int work_slot) {
    u64 *sc = scalars_of(ctx, ctx->cur.nparts);
    JoinArgs j{};
    j.build = ctx->cur.build;
    j.probe = ctx->cur.probe;
    j.offB = ctx->cur.offB;
    j.offP = ctx->cur.offP;
    j.items = (const Item *) ctx->items.p;
    j.nitems = (const u32 *) (sc + kScNItems);
    j.work_counter = (u32 *) (sc + work_slot);
    j.out_cursor = sc + kScCursor;
    j.build_is_S = ctx->cur.build_is_S;
    return j;
}

int read_scalars(rhj_ctx *ctx, cudaStream_t st) {
    u64 *sc = scalars_of(ctx, ctx->cur.nparts);
    CK(cudaMemcpyAsync(ctx->h_scalars, sc, kScCount * sizeof(u64), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (ctx->h_scalars[kScErr]) return fail(ctx, RHJ_ERR_STATE, "device-side planning error (work-item table overflow)");
    ctx->info.n_items = (u32) ctx->h_scalars[kScNItems];
    return RHJ_OK;
}

int count_phase(rhj_ctx *ctx, cudaStream_t st) {
    int rc;
    if ((rc = ensure(ctx, ctx->item_cnt, (size_t) ctx->cur.item_cap * 8))) return rc;
    if ((rc = ensure(ctx, ctx->item_off, (size_t) ctx->cur.item_cap * 8))) return rc;
    JoinArgs j = join_args(ctx, kScWork0);
    j.item_cnt = (u64 *) ctx->item_cnt.p;
    mark(ctx, st, RHJ_PHASE_JOIN);
    if ((rc = launch_join<kJoinCount>(ctx, st, j, ctx->cur.item_cap))) return rc;
    mark(ctx, st, RHJ_PHASE_SCAN_ITEMS);
    u64 *sc = scalars_of(ctx, ctx->cur.nparts);
    k_scan_items<<<1, 1024, 0, st>>>((const u64 *) ctx->item_cnt.p, (const u32 *) (sc + kScNItems),
                                     (u64 *) ctx->item_off.p, sc + kScTotal);
    CK(cudaGetLastError());
    ctx->info.kernel_launches++;
    mark(ctx, st, -1);
    if ((rc = read_scalars(ctx, st))) return rc;
    ctx->cur.count = ctx->h_scalars[kScTotal];
    ctx->cur.counted = true;
    return RHJ_OK;
}

int write_phase(rhj_ctx *ctx, cudaStream_t st, Pair *d_out, u64 capacity) {
    if (!ctx->cur.valid || !ctx->cur.counted) return fail(ctx, RHJ_ERR_STATE, "write pass without a count pass");
    if (capacity < ctx->cur.count) return fail(ctx, RHJ_ERR_CAPACITY, "output buffer smaller than the counted result");
    if (ctx->cur.count == 0) return RHJ_OK;
    JoinArgs j = join_args(ctx, kScWork1);
    j.item_off = (const u64 *) ctx->item_off.p;
    j.out = d_out;
    j.capacity = capacity;
    if (ctx->nmarks && ctx->mark_phase[ctx->nmarks - 1] == -1) ctx->nmarks--;  // reopen the mark list
    mark(ctx, st, RHJ_PHASE_JOIN_WRITE);
    int rc = launch_join<kJoinWrite>(ctx, st, j, ctx->cur.item_cap);
    mark(ctx, st, -1);
    return rc;
}

}  // namespace

// =================================================================================================
extern "C" {

const char *rhj_version(void) { return RHJ_VERSION; }

int rhj_create(int device, rhj_ctx **out) {
    if (!out) return RHJ_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) {
        cudaGetLastError();
        return RHJ_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return RHJ_ERR_NO_DEVICE;
    if (prop.major != 10) return RHJ_ERR_NO_DEVICE;  // sm_100a cubin only
    rhj_ctx *ctx = new rhj_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    const char *e;
    if ((e = getenv("RHJ_HIST_AGG"))) ctx->hist_agg = atoi(e) != 0;
    if ((e = getenv("RHJ_SCATTER_MODE"))) ctx->scatter_mode = atoi(e);
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaHostAlloc((void **) &ctx->h_scalars, kScCount * sizeof(u64), cudaHostAllocDefault) != cudaSuccess) {
        delete ctx;
        return RHJ_ERR_CUDA;
    }
    *out = ctx;
    return RHJ_OK;
}

int rhj_destroy(rhj_ctx *ctx) {
    if (!ctx) return RHJ_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for_each_buf(ctx, [](DevBuf &b) { if (b.p) cudaFree(b.p); });
    if (ctx->h_out) cudaFreeHost(ctx->h_out);
    if (ctx->h_iu) cudaFreeHost(ctx->h_iu);
    if (ctx->h_scalars) cudaFreeHost(ctx->h_scalars);
    for (cudaEvent_t e : ctx->ev)
        if (e) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return RHJ_OK;
}

const char *rhj_last_error(const rhj_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

uint64_t rhj_workspace_bytes(const rhj_ctx *ctx) {
    if (!ctx) return 0;
    uint64_t s = 0;
    for_each_buf(const_cast<rhj_ctx *>(ctx), [&s](DevBuf &b) { s += b.cap; });
    return s;
}

int rhj_reserve(rhj_ctx *ctx, uint64_t nR, uint64_t nS) {
    if (!ctx) return RHJ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    Plan pl = make_plan(nR, nS);
    Meta m;
    int rc;
    if ((rc = layout_meta(ctx, pl.nparts, m))) return rc;
    if (pl.bits > 0 && (rc = ensure(ctx, ctx->bufA, (pl.nB + pl.nP) * sizeof(Tup)))) return rc;
    if (pl.b2 > 0 && (rc = ensure(ctx, ctx->bufB, (pl.nB + pl.nP) * sizeof(Tup)))) return rc;
    u64 cap = (u64) pl.nparts + pl.nP / kProbeChunk + 2;
    if ((rc = ensure(ctx, ctx->items, cap * sizeof(Item)))) return rc;
    if ((rc = ensure(ctx, ctx->item_cnt, cap * 8))) return rc;
    if ((rc = ensure(ctx, ctx->item_off, cap * 8))) return rc;
    return RHJ_OK;
}

int rhj_set_profiling(rhj_ctx *ctx, int on) {
    if (!ctx) return RHJ_ERR_ARG;
    ctx->profiling = on != 0;
    ctx->nmarks = 0;
    return RHJ_OK;
}

int rhj_last_phase_ms(rhj_ctx *ctx, float *ms) {
    if (!ctx || !ms) return RHJ_ERR_ARG;
    for (int i = 0; i < RHJ_NUM_PHASES; ++i) ms[i] = 0.f;
    for (int i = 0; i + 1 < ctx->nmarks; ++i) {
        int ph = ctx->mark_phase[i];
        if (ph < 0 || ph >= RHJ_NUM_PHASES) continue;
        float t = 0.f;
        if (cudaEventElapsedTime(&t, ctx->ev[i], ctx->ev[i + 1]) != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, RHJ_ERR_STATE, "phase events not complete (synchronise the stream first)");
        }
        ms[ph] += t;
    }
    return RHJ_OK;
}

int rhj_last_plan(const rhj_ctx *ctx, rhj_plan_info *info) {
    if (!ctx || !info) return RHJ_ERR_ARG;
    *info = ctx->info;
    return RHJ_OK;
}

// ---- join -------------------------------------------------------------------------------------------

int rhj_join_count_device(rhj_ctx *ctx, const rhj_tuple *dR, uint64_t nR, const rhj_tuple *dS, uint64_t nS,
                          uint64_t *count, void *stream) {
    if (!ctx || !count) return RHJ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    *count = 0;
    ctx->cur.valid = false;
    if (nR == 0 || nS == 0) {  // nothing can match; the reference returns head == nullptr
        ctx->info = rhj_plan_info{};
        ctx->cur.valid = true;
        ctx->cur.counted = true;
        ctx->cur.count = 0;
        return RHJ_OK;
    }
    if (!dR || !dS) return fail(ctx, RHJ_ERR_ARG, "null relation pointer");
    int rc;
    if ((rc = partition_and_plan(ctx, st, (const Tup *) dR, nR, (const Tup *) dS, nS))) return rc;
    if ((rc = count_phase(ctx, st))) return rc;
    *count = ctx->cur.count;
    return RHJ_OK;
}

int rhj_join_write_device(rhj_ctx *ctx, rhj_pair *d_out, uint64_t capacity, void *stream) {
    if (!ctx) return RHJ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    if (ctx->cur.valid && ctx->cur.counted && ctx->cur.count == 0) return RHJ_OK;
    if (!d_out) return fail(ctx, RHJ_ERR_ARG, "null output pointer");
    int rc = write_phase(ctx, st, (Pair *) d_out, capacity);
    if (rc) return rc;
    CK(cudaStreamSynchronize(st));
    return RHJ_OK;
}

int rhj_join_device(rhj_ctx *ctx, const rhj_tuple *dR, uint64_t nR, const rhj_tuple *dS, uint64_t nS,
                    rhj_pair *d_out, uint64_t capacity, uint64_t *count, int emit, void *stream) {
    if (!ctx || !count) return RHJ_ERR_ARG;
    if (emit == RHJ_EMIT_COUNT_THEN_WRITE) {
        int rc = rhj_join_count_device(ctx, dR, nR, dS, nS, count, stream);
        if (rc) return rc;
        if (*count > capacity) return fail(ctx, RHJ_ERR_CAPACITY, "output buffer smaller than the counted result");
        return rhj_join_write_device(ctx, d_out, capacity, stream);
    }
    if (emit != RHJ_EMIT_FUSED) return fail(ctx, RHJ_ERR_ARG, "unknown emitter");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    *count = 0;
    ctx->cur.valid = false;
    if (nR == 0 || nS == 0) {
        ctx->info = rhj_plan_info{};
        return RHJ_OK;
    }
    if (!dR || !dS || (!d_out && capacity)) return fail(ctx, RHJ_ERR_ARG, "null pointer");
    int rc;
    if ((rc = partition_and_plan(ctx, st, (const Tup *) dR, nR, (const Tup *) dS, nS))) return rc;
    JoinArgs j = join_args(ctx, kScWork0);
    j.out = (Pair *) d_out;
    j.capacity = capacity;
    mark(ctx, st, RHJ_PHASE_JOIN);
    if ((rc = launch_join<kJoinFused>(ctx, st, j, ctx->cur.item_cap))) return rc;
    mark(ctx, st, -1);
    if ((rc = read_scalars(ctx, st))) return rc;
    *count = ctx->h_scalars[kScCursor];
    if (*count > capacity) return fail(ctx, RHJ_ERR_CAPACITY, "output buffer too small for the fused emitter");
    return RHJ_OK;
}

int rhj_join_host(rhj_ctx *ctx, const rhj_tuple *R, uint64_t nR, const rhj_tuple *S, uint64_t nS,
                  const rhj_pair **out, uint64_t *count) {
    if (!ctx || !out || !count) return RHJ_ERR_ARG;
    *out = nullptr;
    *count = 0;
    if (nR == 0 || nS == 0) return RHJ_OK;
    if (!R || !S) return fail(ctx, RHJ_ERR_ARG, "null relation pointer");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int rc;
    if ((rc = ensure(ctx, ctx->inR, nR * sizeof(Tup)))) return rc;
    if ((rc = ensure(ctx, ctx->inS, nS * sizeof(Tup)))) return rc;
    CK(cudaMemcpyAsync(ctx->inR.p, R, nR * sizeof(Tup), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->inS.p, S, nS * sizeof(Tup), cudaMemcpyHostToDevice, st));
    uint64_t n = 0;
    if ((rc = rhj_join_count_device(ctx, (const rhj_tuple *) ctx->inR.p, nR, (const rhj_tuple *) ctx->inS.p, nS, &n, st)))
        return rc;
    if (n == 0) return RHJ_OK;
    if ((rc = ensure(ctx, ctx->outP, n * sizeof(Pair)))) return rc;
    if ((rc = ensure_pinned(ctx, &ctx->h_out, &ctx->h_out_cap, n * sizeof(Pair)))) return rc;
    if ((rc = write_phase(ctx, st, (Pair *) ctx->outP.p, n))) return rc;
    CK(cudaMemcpyAsync(ctx->h_out, ctx->outP.p, n * sizeof(Pair), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *out = (const rhj_pair *) ctx->h_out;
    *count = n;
    return RHJ_OK;
}

// Result page list, Result.cpp:21-35: pages are filled in append order, the newest page is the
// head and the only partial one.
void *rhj_pairs_to_pages(const rhj_pair *pairs, uint64_t count, uint64_t *head_size) {
    const uint64_t page_bytes = 128 * 1024;                       // BUCKET_SIZE, Result.cpp:7
    const uint64_t cap = (page_bytes - sizeof(void *)) / sizeof(rhj_pair);  // 8191, Result.cpp:11
    if (head_size) *head_size = cap;                              // empty Result: size == capacity, Result.cpp:12
    void *head = nullptr;
    for (uint64_t at = 0; at < count; at += cap) {
        uint64_t k = std::min(cap, count - at);
        char *page = (char *) malloc(page_bytes);
        if (!page) abort();
        *(void **) page = head;
        memcpy(page + sizeof(void *), pairs + at, k * sizeof(rhj_pair));
        head = page;
        if (head_size) *head_size = k;
    }
    return head;
}

// ---- the steps, individually ----------------------------------------------------------------------

int rhj_histogram_device(rhj_ctx *ctx, const rhj_tuple *d_in, uint64_t n, int bits, int shift, int digit_kind,
                         uint64_t *d_hist, void *stream) {
    if (!ctx || !d_hist || bits < 0 || bits > kMaxBitsPerPass || shift < 0 || shift > 63) return RHJ_ERR_ARG;
    if (digit_kind != RHJ_DIGIT_RAW && digit_kind != RHJ_DIGIT_HASH) return RHJ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    u32 ndig = 1u << bits;
    CK(cudaMemsetAsync(d_hist, 0, ndig * sizeof(u64), st));
    if (n) {
        PartArgs a{};
        a.shift = shift;
        a.mask = ndig - 1;
        a.ndig = ndig;
        a.rel[0] = PartRel{(const Tup *) d_in, nullptr, n, (u64 *) d_hist, nullptr, nullptr, nullptr, 1, tiles_of(n)};
        int rc = launch_hist(ctx, st, a, digit_kind, false);
        if (rc) return rc;
    }
    CK(cudaStreamSynchronize(st));
    return RHJ_OK;
}

static int partition_one(rhj_ctx *ctx, cudaStream_t st, const Tup *in, u64 n, int bits, int shift, int kind, Tup *out,
                         u64 *d_offsets) {
    u32 ndig = 1u << bits;
    Meta m;
    int rc;
    ctx->cur.valid = false;  // the metadata buffers are shared with the join
    if ((rc = layout_meta(ctx, 1, m))) return rc;
    CK(cudaMemsetAsync(ctx->zero.p, 0, m.zero_bytes, st));
    PartArgs a{};
    a.shift = shift;
    a.mask = ndig - 1;
    a.ndig = ndig;
    a.rel[0] = PartRel{in, out, n, m.hist1[0], m.cur1[0], nullptr, nullptr, 1, tiles_of(n)};
    if ((rc = launch_hist(ctx, st, a, kind, false))) return rc;
    ScanDigitsArgs sd{};
    sd.hist[0] = sd.hist[1] = m.hist1[0];
    sd.off[0] = sd.off[1] = m.off1[0];
    sd.cursor[0] = sd.cursor[1] = m.cur1[0];
    sd.tile0[0] = sd.tile0[1] = nullptr;
    sd.ndig = ndig;
    k_scan_digits<<<1, kMaxDigits, 0, st>>>(sd);
    CK(cudaGetLastError());
    if ((rc = launch_scatter(ctx, st, a, kind, false))) return rc;
    if (d_offsets)
        CK(cudaMemcpyAsync(d_offsets, m.off1[0], (ndig + 1) * sizeof(u64), cudaMemcpyDeviceToDevice, st));
    return RHJ_OK;
}

int rhj_partition_device(rhj_ctx *ctx, const rhj_tuple *d_in, uint64_t n, int bits, int shift, int digit_kind,
                         rhj_tuple *d_out, uint64_t *d_offsets, void *stream) {
    if (!ctx || bits < 0 || bits > kMaxBitsPerPass || shift < 0 || shift > 63) return RHJ_ERR_ARG;
    if (digit_kind != RHJ_DIGIT_RAW && digit_kind != RHJ_DIGIT_HASH) return RHJ_ERR_ARG;
    if (n && (!d_in || !d_out)) return fail(ctx, RHJ_ERR_ARG, "null pointer");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    int rc = partition_one(ctx, st, (const Tup *) d_in, n, bits, shift, digit_kind, (Tup *) d_out, (u64 *) d_offsets);
    if (rc) return rc;
    CK(cudaStreamSynchronize(st));
    return RHJ_OK;
}

int rhj_shuffle_partition_device(rhj_ctx *ctx, const rhj_tuple *d_in, uint64_t n, int world, rhj_tuple *d_out,
                                 uint64_t *h_counts, void *stream) {
    if (!ctx || !h_counts || world < 1 || world > 256 || (world & (world - 1))) return RHJ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    int bits = 0;
    while ((1 << bits) < world) ++bits;
    if (bits == 0) {
        if (n) CK(cudaMemcpyAsync(d_out, d_in, n * sizeof(Tup), cudaMemcpyDeviceToDevice, st));
        CK(cudaStreamSynchronize(st));
        h_counts[0] = n;
        return RHJ_OK;
    }
    int rc = partition_one(ctx, st, (const Tup *) d_in, n, bits, 32 - bits, kDigitRank, (Tup *) d_out, nullptr);
    if (rc) return rc;
    u64 off[257];
    CK(cudaMemcpyAsync(off, ((u64 *) ctx->meta.p), (world + 1) * sizeof(u64), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int r = 0; r < world; ++r) h_counts[r] = off[r + 1] - off[r];
    return RHJ_OK;
}

}  // extern "C"
