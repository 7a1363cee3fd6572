// rhj_api.cu -- host side of librhj.so: the C ABI declared in include/rhj.h.
//
// Replaces, on the join path, the HistogramJob / PartitionJob / JoinJob fan-out that
// Result::multiRadixHashJoin (Result.cpp:90-124) and relation_info::hash_relation
// (structs.cpp:144-204) schedule on the reference's pthread pool with a fixed sequence of kernel
// launches on one CUDA stream.  No host synchronisation happens between the launches of one
// join; sizes that depend on the data (partition sizes, work items, match counts) stay on the
// device.  There is no CPU fallback anywhere in this file.
#include <future>
#include <thread>
#include <vector>

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
        while (bits < 2 * kPlanBitsPerPass && (p.nB >> bits) > kTargetBuildPerPart) ++bits;
    }
    p.bits = bits;
    p.b1 = bits <= kPlanBitsPerPass ? bits : (bits + 1) / 2;
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
    else if (kind == kDigitShard) HIST(kDigitShard, false);
    else { if (seg) HIST(kDigitRank, true); else HIST(kDigitRank, false); }
#undef HIST
    CK(cudaGetLastError());
    ctx->info.kernel_launches++;
    return RHJ_OK;
}

template <int K, bool S, int W, int IO = kIoAos>
cudaError_t launch_scatter_t(cudaStream_t st, const PartArgs &a, u32 grid) {
    const size_t smem = kScatterSmem;
    if (a.ndig > 512) {  // 1024-digit passes (sharded plans) pay for the larger counter arrays, 512-digit ones do not
        cudaError_t e = set_smem(k_scatter<K, S, W, kMaxDigits, false, IO>, smem);
        if (e != cudaSuccess) return e;
        k_scatter<K, S, W, kMaxDigits, false, IO><<<grid, kPartThreads, smem, st>>>(a);
        return cudaGetLastError();
    }
    cudaError_t e = set_smem(k_scatter<K, S, W, 512, false, IO>, smem);
    if (e != cudaSuccess) return e;
    k_scatter<K, S, W, 512, false, IO><<<grid, kPartThreads, smem, st>>>(a);
    return cudaGetLastError();
}

int launch_scatter(rhj_ctx *ctx, cudaStream_t st, const PartArgs &a, int kind, bool seg, bool limit = false) {
    u32 grid = a.rel[0].ntiles + a.rel[1].ntiles;
    if (!grid) return RHJ_OK;
    if (limit && kind == kDigitShard) {  // pipelined exchange, pass 1: (rank | sub-digit) regions of fixed capacity, per-digit output base
        if (seg || a.shard_local) return fail(ctx, RHJ_ERR_STATE, "bounds-checked shard scatter: unsegmented, peer_out bases");
        if (a.ndig > 512) {
            // 1024 digits: 1024 threads on 8192-tuple tiles, one CTA per SM (see k_scatter); one relation per launch
            if (a.rel[1].ntiles) return fail(ctx, RHJ_ERR_STATE, "bounds-checked 1024-digit shard scatter: one relation per launch");
            constexpr int kBigThreads = 1024, kBigTile = kBigThreads * kPartItems;
            PartArgs b = a;
            b.rel[0].ntiles = (u32) ((a.rel[0].n + kBigTile - 1) / kBigTile);
            auto kern = k_scatter<kDigitShard, false, kWriteStaged, kMaxDigits, true, kIoAos, kBigThreads>;
            CK(set_smem(kern, (size_t) kBigTile * sizeof(Tup)));
            kern<<<b.rel[0].ntiles, kBigThreads, (size_t) kBigTile * sizeof(Tup), st>>>(b);
        } else {
            CK(set_smem(k_scatter<kDigitShard, false, kWriteStaged, 512, true>, kScatterSmem));
            k_scatter<kDigitShard, false, kWriteStaged, 512, true><<<grid, kPartThreads, kScatterSmem, st>>>(a);
        }
        CK(cudaGetLastError());
        ctx->info.kernel_launches++;
        return RHJ_OK;
    }
    if (limit) {  // optimistic passes: fixed-capacity regions, bounds-checked staged stores
        if (kind != kDigitHash) return fail(ctx, RHJ_ERR_STATE, "bounds-checked scatter: hash digits only");
        if (!a.rel[0].in && a.rel[0].in_packed) {  // packed 12-byte records in (pipelined exchange, pass 2), 16-byte tuples out
            if (!seg) return fail(ctx, RHJ_ERR_STATE, "packed input is only wired for the segmented pass");
            if (a.ndig > 512) {
                CK(set_smem(k_scatter<kDigitHash, true, kWriteStaged, kMaxDigits, true, kIoPacked12In>, kScatterSmem));
                k_scatter<kDigitHash, true, kWriteStaged, kMaxDigits, true, kIoPacked12In><<<grid, kPartThreads, kScatterSmem, st>>>(a);
            } else {
                CK(set_smem(k_scatter<kDigitHash, true, kWriteStaged, 512, true, kIoPacked12In>, kScatterSmem));
                k_scatter<kDigitHash, true, kWriteStaged, 512, true, kIoPacked12In><<<grid, kPartThreads, kScatterSmem, st>>>(a);
            }
            CK(cudaGetLastError());
            ctx->info.kernel_launches++;
            return RHJ_OK;
        }
        if (a.ndig > 512) {  // 1024-digit second pass: only the sharded plans of very large relations (2^28 tuples per rank) need it
            if (!seg) return fail(ctx, RHJ_ERR_STATE, "bounds-checked scatter: > 512 digits only in the segmented pass");
            CK(set_smem(k_scatter<kDigitHash, true, kWriteStaged, kMaxDigits, true>, kScatterSmem));
            k_scatter<kDigitHash, true, kWriteStaged, kMaxDigits, true><<<grid, kPartThreads, kScatterSmem, st>>>(a);
        } else if (seg) {
            CK(set_smem(k_scatter<kDigitHash, true, kWriteStaged, 512, true>, kScatterSmem));
            k_scatter<kDigitHash, true, kWriteStaged, 512, true><<<grid, kPartThreads, kScatterSmem, st>>>(a);
        } else {
            CK(set_smem(k_scatter<kDigitHash, false, kWriteStaged, 512, true>, kScatterSmem));
            k_scatter<kDigitHash, false, kWriteStaged, 512, true><<<grid, kPartThreads, kScatterSmem, st>>>(a);
        }
        CK(cudaGetLastError());
        ctx->info.kernel_launches++;
        return RHJ_OK;
    }
    const int w = kind == kDigitShard ? ctx->shard_scatter_mode : ctx->scatter_mode;
    cudaError_t e;
#define SC(K, S) (w == 1 ? launch_scatter_t<K, S, kWriteBulk>(st, a, grid) : launch_scatter_t<K, S, kWriteStaged>(st, a, grid))
    if (kind == kDigitRaw) e = seg ? SC(kDigitRaw, true) : SC(kDigitRaw, false);
    else if (kind == kDigitHash) e = seg ? SC(kDigitHash, true) : SC(kDigitHash, false);
    else if (kind == kDigitShard) e = SC(kDigitShard, false);
    else e = seg ? SC(kDigitRank, true) : SC(kDigitRank, false);
#undef SC
    CK(e);
    ctx->info.kernel_launches++;
    return RHJ_OK;
}

template <int MODE, bool POS = false>
int launch_join(rhj_ctx *ctx, cudaStream_t st, const JoinArgs &a, u32 item_cap) {
    u32 grid = std::min<u32>(std::max<u32>(item_cap, 1), (u32) ctx->num_sms * RHJ_JOIN_MINBLOCKS);
    CK(set_smem(k_join<MODE, POS>, kJoinSmem));
    k_join<MODE, POS><<<grid, kJoinThreads, kJoinSmem, st>>>(a);
    CK(cudaGetLastError());
    ctx->info.kernel_launches++;
    return RHJ_OK;
}

// Closes the holes a positional emit left in d_out[0, cursor): the `holes` unmatched slots hold RHJ_HOLE pairs; the valid
// pairs behind F = cursor - holes move into the holes before F.  *ok = false when the sentinel count does not add up (a
// genuine pair equal to the sentinel): the caller then redoes the join with the ranked emitter.
int close_holes(rhj_ctx *ctx, cudaStream_t st, Pair *d_out, u64 cursor, u64 holes, bool *ok) {
    *ok = true;
    const u64 F = cursor - holes;
    const u64 ntile64 = (cursor + kHoleTile - 1) / kHoleTile;
    if (ntile64 > 0x7fffffffull) return fail(ctx, RHJ_ERR_ARG, "result too large");
    const u32 ntiles = (u32) ntile64;
    int rc;
    if ((rc = ensure(ctx, ctx->filt_cnt, (size_t) 2 * ntiles * 4))) return rc;
    if ((rc = ensure(ctx, ctx->filt_off, (size_t) 2 * ntiles * 8 + 16))) return rc;
    const u64 kmax = std::min(holes, F) + 1;   // a hole before F needs a valid pair behind it: at most min(holes, F) moves
    if ((rc = ensure(ctx, ctx->filt_tmp, (size_t) 2 * kmax * 8))) return rc;
    u32 *cnt = (u32 *) ctx->filt_cnt.p;
    u64 *off = (u64 *) ctx->filt_off.p, *totals = off + 2 * (size_t) ntiles;
    u64 *lh = (u64 *) ctx->filt_tmp.p, *lt = lh + kmax;
    k_holes_count<<<ntiles, 256, 0, st>>>(d_out, cursor, F, ntiles, cnt);
    k_holes_scan<<<2, 1024, 0, st>>>(cnt, ntiles, off, totals);
    CK(cudaGetLastError());
    ctx->info.kernel_launches += 2;
    CK(cudaMemcpyAsync(ctx->h_scalars, totals, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const u64 kh = ctx->h_scalars[0], kt = ctx->h_scalars[1];
    if (kh != kt || kh >= kmax) {
        *ok = false;
        return RHJ_OK;
    }
    if (kh) {
        k_holes_list<<<ntiles, 256, 0, st>>>(d_out, cursor, F, ntiles, off, lh, lt);
        k_holes_fill<<<(u32) std::min<u64>((kh + 255) / 256, (u64) ctx->num_sms * 8), 256, 0, st>>>(d_out, lh, lt, kh);
        CK(cudaGetLastError());
        ctx->info.kernel_launches += 2;
        CK(cudaStreamSynchronize(st));
    }
    return RHJ_OK;
}

// Layout of the two metadata buffers for a plan.
struct Meta {
    u64 *hist1[2], *hist2[2], *scalars;                  // zeroed
    u64 *off1[2], *cur1[2], *off2[2], *cur2[2];          // written by the scans
    u32 *tile0[2];
    u64 *segb[2], *sege[2];                              // optimistic pass 1: where each pass-1 partition lies
    size_t zero_bytes;
};

int layout_meta(rhj_ctx *ctx, u32 nparts, Meta &m) {
    size_t zero_u64 = 2 * (size_t) kMaxDigits + 2 * (size_t) nparts + kScCount;
    size_t meta_u64 = 2 * (size_t) (kMaxDigits + 1) + 2 * (size_t) kMaxDigits + 2 * (size_t) (nparts + 1) +
                      2 * (size_t) nparts + (size_t) (kMaxDigits + 2) + 4 * (size_t) kMaxDigits;
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
    q += kMaxDigits + 2;
    for (int i = 0; i < 2; ++i) { m.segb[i] = q; q += kMaxDigits; }
    for (int i = 0; i < 2; ++i) { m.sege[i] = q; q += kMaxDigits; }
    return RHJ_OK;
}

u64 *scalars_of(rhj_ctx *ctx, u32 nparts) {
    return (u64 *) ctx->zero.p + 2 * (size_t) kMaxDigits + 2 * (size_t) nparts;
}

// The positional emitter.  Default: k_join_pos (build partitions of one chunk with unique keys; instruction-level-parallel
// loops) followed by k_join<FUSED> over the items it left (several build chunks, duplicate build keys -- usually none).
// RHJ_JOIN_LEAN=0: the r02 kernel k_join<FUSED, POS>, which handles everything in one launch.
int launch_join_positional(rhj_ctx *ctx, cudaStream_t st, const JoinArgs &a, u32 item_cap) {
    if (!ctx->join_lean) return launch_join<kJoinFused, true>(ctx, st, a, item_cap);
    int rc;
    if ((rc = ensure(ctx, ctx->items_left, (size_t) std::max<u32>(item_cap, 1) * sizeof(Item)))) return rc;
    u64 *sc = scalars_of(ctx, ctx->cur.nparts);
    Item *left = (Item *) ctx->items_left.p;
    u32 *nleft = (u32 *) (sc + kScLeft);
    const u32 grid = std::min<u32>(std::max<u32>(item_cap, 1), (u32) ctx->num_sms * RHJ_JOIN_MINBLOCKS);
    // three probe tuples per thread and round, the next round's in flight (four without prefetch: measured slower, 1.77 vs
    // 1.55 ms); RHJ_JOIN_POS_V selects the variant of the kernel (see k_join_pos)
    switch (ctx->join_pos_v) {
    case 1:
        CK(set_smem(k_join_pos<3, 1>, kJoinSmem));
        k_join_pos<3, 1><<<grid, kJoinThreads, kJoinSmem, st>>>(a, left, nleft);
        break;
    case 3:
        CK(set_smem(k_join_pos<3, 3>, kJoinSmem));
        k_join_pos<3, 3><<<grid, kJoinThreads, kJoinSmem, st>>>(a, left, nleft);
        break;
    default:
        CK(set_smem(k_join_pos<3, 0>, kJoinSmem));
        k_join_pos<3, 0><<<grid, kJoinThreads, kJoinSmem, st>>>(a, left, nleft);
    }
    CK(cudaGetLastError());
    ctx->info.kernel_launches++;
    JoinArgs r = a;   // the ranked emitter over the leftover list, appending behind the same output cursor
    r.items = (const Item *) ctx->items_left.p;
    r.nitems = (const u32 *) (sc + kScLeft);
    r.work_counter = (u32 *) (sc + kScWork1);
    return launch_join<kJoinFused>(ctx, st, r, item_cap);
}


// Builds the per-tile descriptor tables of a segmented pass (one 16-byte entry per tile) and points
// the relations at them.  `slot` 0/1 selects the half of ctx->tiles a lone relation uses.
int build_tile_tables(rhj_ctx *ctx, cudaStream_t st, PartArgs &b, int nrel, int slot = 0) {
    size_t need[2] = {0, 0}, total = 0;
    for (int i = 0; i < nrel; ++i) {
        need[i] = ((size_t) b.rel[i].ntiles + 1) * sizeof(TileDesc);
        total += need[i];
    }
    // lone relations (sharded join: the slots arrive separately) use fixed thirds so they never alias
    size_t half = 0;
    if (nrel == 1) {
        half = need[0];
        total = 3 * half;
    }
    if (total > ctx->tiles.cap) {
        // growing would free a table a previously launched kernel may still read: finish that work first
        CK(cudaStreamSynchronize(st));
        int rc = ensure(ctx, ctx->tiles, std::max(total, 2 * ctx->tiles.cap));
        if (rc) return rc;
        if (nrel == 1) half = ctx->tiles.cap / 3 / sizeof(TileDesc) * sizeof(TileDesc);
    } else if (nrel == 1) {
        half = ctx->tiles.cap / 3 / sizeof(TileDesc) * sizeof(TileDesc);
    }
    char *base = (char *) ctx->tiles.p + (nrel == 1 ? (size_t) slot * half : 0);
    for (int i = 0; i < nrel; ++i) {
        TileDesc *t = (TileDesc *) base;
        k_tile_table<<<b.rel[i].nseg + 1, 256, 0, st>>>(b.rel[i].seg_off, b.rel[i].seg_end, b.rel[i].seg_tile0, b.rel[i].nseg,
                                                        b.rel[i].ntiles, t);
        CK(cudaGetLastError());
        ctx->info.kernel_launches++;
        b.rel[i].tiles = t;
        base += need[i];
    }
    return RHJ_OK;
}

// Per-partition capacity of the optimistic pass-2 layout: a final partition of hashed, mostly distinct keys holds
// Poisson(mean) tuples; mean + 16 sigma + 64 leaves room for keys repeated a few times.  Multiple of 8 tuples so
// every region starts on a 128-byte line.
inline u64 fixed_cap2(u64 n, u32 nparts) {
    const double mean = (double) n / nparts;
    return ((u64) (mean + 16.0 * std::sqrt(mean) + 64.0) + 8) & ~(u64) 7;
}

// Second radix pass (if the plan has one) over pass-1-partitioned relations inX[0] (build) and
// inX[1] (probe) whose pass-1 offsets / first-tile tables are off1X / tile0X, then the work-item
// plan.  Leaves ctx->cur describing the final partitions.  Enqueues only; no host sync.
int second_pass_and_plan(rhj_ctx *ctx, cudaStream_t st, const Plan &pl, const Meta &m, const Tup *const inX[2],
                         const u64 *const off1X[2], const u32 *const tile0X[2], const u64 *const seg_begX[2] = nullptr,
                         const u64 *const seg_endX[2] = nullptr, bool fixed2 = false) {
    int rc;
    const Tup *finB, *finP;
    const u64 *offB, *offP, *endB = nullptr, *endP = nullptr;
    bool planned = false;
    u32 item_cap = 0;
    const u64 cap[2] = {fixed_cap2(pl.nB, pl.nparts), fixed_cap2(pl.nP, pl.nparts)};
    const size_t regB = (size_t) pl.nparts * cap[0], regP = (size_t) pl.nparts * cap[1];
    // the padded layout needs ~1.4x the memory of the packed one: when that does not fit, pack (histogram path)
    if (fixed2 && pl.b2 > 0 && ensure(ctx, ctx->bufB, (regB + regP + kTile) * sizeof(Tup)) != RHJ_OK) {
        fixed2 = false;
        ctx->err.clear();  // not an error: the packed layout is tried next
    }
    if (pl.b2 == 0) {
        finB = inX[0];
        finP = inX[1];
        offB = off1X[0];
        offP = off1X[1];
    } else if (fixed2) {
        // ---- optimistic pass 2: no histogram; final partition p scatters into the fixed region [p * cap, (p + 1) * cap) ----
        Tup *B = (Tup *) ctx->bufB.p;
        PartArgs b{};
        b.shift = 32 - pl.bits;
        b.mask = (1u << pl.b2) - 1;
        b.ndig = 1u << pl.b2;
        b.overflow = (u32 *) (m.scalars + kScOverflow);
        const u32 nseg = 1u << pl.b1;
        b.rel[0] = PartRel{inX[0], B, pl.nB, m.hist2[0], m.cur2[0], off1X[0], tile0X[0], nseg, tiles_of(pl.nB) + nseg};
        b.rel[1] = PartRel{inX[1], B + regB, pl.nP, m.hist2[1], m.cur2[1], off1X[1], tile0X[1], nseg, tiles_of(pl.nP) + nseg};
        b.rel[0].dump = regB + regP;  // one dump tile behind both relations (indices are relative to each `out`)
        b.rel[1].dump = regP;
        for (int i = 0; i < 2; ++i) {
            b.rel[i].limit_cap = cap[i];
            if (seg_begX) {
                b.rel[i].seg_off = seg_begX[i];
                b.rel[i].seg_end = seg_endX[i];
            }
        }
        if ((rc = build_tile_tables(ctx, st, b, 2))) return rc;
        u64 cap64 = (u64) pl.nparts + pl.nP / kProbeChunk + 2;
        if (cap64 > 0x7fffffffull) return fail(ctx, RHJ_ERR_ARG, "relation too large for the work-item table");
        item_cap = (u32) cap64;
        if ((rc = ensure(ctx, ctx->items, (size_t) item_cap * sizeof(Item)))) return rc;
        PlanFixedArgs pf{};
        for (int i = 0; i < 2; ++i) {
            pf.end[i] = m.cur2[i];
            pf.beg[i] = m.off2[i];
            pf.cap[i] = cap[i];
        }
        pf.nseg = nseg;
        pf.ndig = b.ndig;
        pf.items = (Item *) ctx->items.p;
        pf.item_cap = item_cap;
        pf.nitems = (u32 *) (m.scalars + kScNItems);
        pf.err = (u32 *) (m.scalars + kScErr);
        pf.overflow = b.overflow;
        k_fixed_cursors2<<<dim3((pl.nparts + 255) / 256, 2), 256, 0, st>>>(pf);
        CK(cudaGetLastError());
        ctx->info.kernel_launches++;
        mark(ctx, st, RHJ_PHASE_SCATTER2);
        if ((rc = launch_scatter(ctx, st, b, kDigitHash, true, true))) return rc;
        mark(ctx, st, RHJ_PHASE_PLAN);
        k_plan_fixed<<<nseg, kMaxDigits, 0, st>>>(pf);
        CK(cudaGetLastError());
        ctx->info.kernel_launches++;
        planned = true;
        finB = B;
        finP = B + regB;
        offB = m.off2[0];
        offP = m.off2[1];
        endB = m.cur2[0];
        endP = m.cur2[1];
        ctx->info.optimistic_pass1 |= 4u;
    } else {
        // ---- pass 2: next b2 bits, inside every pass-1 partition ----
        if ((rc = ensure(ctx, ctx->bufB, (pl.nB + pl.nP) * sizeof(Tup)))) return rc;
        Tup *B = (Tup *) ctx->bufB.p;
        PartArgs b{};
        b.shift = 32 - pl.bits;
        b.mask = (1u << pl.b2) - 1;
        b.ndig = 1u << pl.b2;
        const u32 nseg = 1u << pl.b1;
        b.rel[0] = PartRel{inX[0], B, pl.nB, m.hist2[0], m.cur2[0], off1X[0], tile0X[0], nseg, tiles_of(pl.nB) + nseg};
        b.rel[1] = PartRel{inX[1], B + pl.nB, pl.nP, m.hist2[1], m.cur2[1], off1X[1], tile0X[1], nseg,
                           tiles_of(pl.nP) + nseg};
        if (seg_begX) {  // pass 1 used fixed-capacity regions: where a segment lies is not where pass 2 packs it
            for (int i = 0; i < 2; ++i) {
                b.rel[i].seg_off = seg_begX[i];
                b.rel[i].seg_end = seg_endX[i];
            }
        }
        if ((rc = build_tile_tables(ctx, st, b, 2))) return rc;
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
            sp.off1[i] = off1X[i];
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
        mark(ctx, st, RHJ_PHASE_SCATTER2);
        if ((rc = launch_scatter(ctx, st, b, kDigitHash, true))) return rc;
        finB = B;
        finP = B + pl.nB;
        offB = m.off2[0];
        offP = m.off2[1];
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
    ctx->cur.endB = endB ? endB : offB + 1;
    ctx->cur.endP = endP ? endP : offP + 1;
    ctx->cur.nparts = pl.nparts;
    ctx->cur.item_cap = item_cap;
    ctx->cur.build_is_S = pl.build_is_S;
    return RHJ_OK;
}

// Partition both relations on `bits` hash bits (one or two passes) and build the work-item list.
// Leaves ctx->cur describing the partitioned relations.  Enqueues only; no host sync.
// Per-partition capacity of the optimistic pass-1 layout: expected size + 12.5 % + 8192 tuples.
inline u64 fixed_cap(u64 n, u32 ndig) { return n / ndig + n / ndig / 8 + 8192; }

// Samples 1/64 of both relations and decides whether every pass-1 partition will fit its fixed
// region with room to spare.  One small kernel + a 4 KiB D2H + a stream synchronise.
int sample_says_balanced(rhj_ctx *ctx, cudaStream_t st, const Plan &pl, const Tup *inB, const Tup *inP, bool ok[2],
                         bool poisson[2]) {
    const u32 ndig = 1u << pl.b1;
    int rc;
    if ((rc = ensure(ctx, ctx->sample, 2 * (size_t) ndig * sizeof(u32)))) return rc;
    CK(cudaMemsetAsync(ctx->sample.p, 0, 2 * (size_t) ndig * sizeof(u32), st));
    SampleArgs sa{};
    sa.in[0] = inB;
    sa.in[1] = inP;
    sa.n[0] = pl.nB;
    sa.n[1] = pl.nP;
    sa.shift = 32 - pl.b1;
    sa.mask = ndig - 1;
    sa.ndig = ndig;
    sa.hist = (u32 *) ctx->sample.p;
    const u64 nmax = std::max(pl.nB, pl.nP);
    const u32 grid = (u32) ((nmax / 64 + 8 + 511) / 512);
    k_sample_hist<<<dim3(grid, 2), 512, 0, st>>>(sa);
    CK(cudaGetLastError());
    ctx->info.kernel_launches++;
    static thread_local u32 h[2 * kMaxDigits];
    CK(cudaMemcpyAsync(h, ctx->sample.p, 2 * (size_t) ndig * sizeof(u32), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const u64 ns[2] = {pl.nB, pl.nP};
    for (int r = 0; r < 2; ++r) {
        u32 mx = 0;
        for (u32 d = 0; d < ndig; ++d) mx = std::max(mx, h[r * ndig + d]);
        // sampled count s ~ true/64 with standard error sqrt(s): leave 5 sigma
        const double est = 64.0 * (mx + 5.0 * std::sqrt((double) mx + 1.0));
        ok[r] = est <= (double) fixed_cap(ns[r], ndig);
        // Index of dispersion of the sampled counts: 1 for hashed distinct keys; a key repeated f times moves as one
        // clump and raises it to 1 + (f - 1) / 64.  Fixed-capacity FINAL partitions (pass 2) are sized for Poisson
        // counts, so they are only tried when the sample shows no clumping (f < ~20; smaller f fits or is retried).
        double sum = 0, sq = 0;
        for (u32 d = 0; d < ndig; ++d) {
            sum += h[r * ndig + d];
            sq += (double) h[r * ndig + d] * h[r * ndig + d];
        }
        const double mean = sum / ndig, var = sq / ndig - mean * mean;
        poisson[r] = mean >= 64.0 && var <= 1.3 * mean;
    }
    return RHJ_OK;
}

int partition_and_plan(rhj_ctx *ctx, cudaStream_t st, const Tup *dR, u64 nR, const Tup *dS, u64 nS,
                       bool allow_optimistic = true) {
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
    ctx->pending.nB = 0;

    Meta m;
    int rc;
    if ((rc = layout_meta(ctx, pl.nparts, m))) return rc;
    CK(cudaMemsetAsync(ctx->zero.p, 0, m.zero_bytes, st));

    const u64 ntot = pl.nB + pl.nP;
    const Tup *finB = nullptr, *finP = nullptr;
    const u64 *offB = nullptr, *offP = nullptr;

    if (pl.bits == 0) {
        finB = inB;
        finP = inP;
        k_set_single_part<<<1, 1, 0, st>>>(m.off2[0], pl.nB, m.off2[1], pl.nP);
        CK(cudaGetLastError());
        ctx->info.kernel_launches++;
        offB = m.off2[0];
        offP = m.off2[1];
    } else {
        // ---- optimistic pass 1 (two-pass plans, large inputs): no histogram, fixed-capacity regions ----
        bool opt[2] = {false, false};  // per relation (0 = build, 1 = probe): a skewed probe side does not cost the build side its shortcut
        bool poisson[2] = {false, false};
        if (!allow_optimistic) ctx->opt2_skip = 16;  // an optimistic layout just overflowed: stay exact in pass 2 for a while
        if (allow_optimistic && ctx->optimistic && pl.b2 > 0 && ntot >= ((u64) 1 << 22)) {
            auto &tr = ctx->trust;
            if (ctx->trust_sample && !ctx->force_optimistic && tr.streak >= 2 && tr.age < 64 && tr.nB == pl.nB && tr.nP == pl.nP) {
                for (int i = 0; i < 2; ++i) {  // same shape as the last sampled joins, which all fitted: reuse their verdict
                    opt[i] = tr.opt[i];
                    poisson[i] = tr.poisson[i];
                }
                tr.age++;
            } else {
                if ((rc = sample_says_balanced(ctx, st, pl, inB, inP, opt, poisson))) return rc;
                ctx->pending.nB = pl.nB;
                ctx->pending.nP = pl.nP;
                for (int i = 0; i < 2; ++i) {
                    ctx->pending.opt[i] = opt[i];
                    ctx->pending.poisson[i] = poisson[i];
                }
            }
            if (ctx->force_optimistic) opt[0] = opt[1] = poisson[0] = poisson[1] = true;  // test hook: exercise the overflow -> exact retry
        }
        bool fixed2 = opt[0] && opt[1] && poisson[0] && poisson[1] && ctx->optimistic2 && pl.b2 <= 9;
        if (fixed2 && ctx->opt2_skip > 0 && !ctx->force_optimistic) {
            ctx->opt2_skip--;
            fixed2 = false;
        }
        if (opt[0] || opt[1]) {
            const u32 nd1 = 1u << pl.b1;
            const size_t regB = opt[0] ? (size_t) nd1 * fixed_cap(pl.nB, nd1) : pl.nB;
            const size_t regP = opt[1] ? (size_t) nd1 * fixed_cap(pl.nP, nd1) : pl.nP;
            if (ensure(ctx, ctx->bufA, (regB + regP + kTile) * sizeof(Tup)) != RHJ_OK) {
                // the padded layout does not fit where the packed one might: take the histogram path for both relations
                ctx->err.clear();
                opt[0] = opt[1] = false;
                fixed2 = false;
            }
        }
        if (opt[0] || opt[1]) {
            const u32 nd1 = 1u << pl.b1;
            const u64 cap[2] = {fixed_cap(pl.nB, nd1), fixed_cap(pl.nP, nd1)};
            const u64 nX[2] = {pl.nB, pl.nP};
            const Tup *inX0[2] = {inB, inP};
            const size_t regB = opt[0] ? (size_t) nd1 * cap[0] : pl.nB, regP = opt[1] ? (size_t) nd1 * cap[1] : pl.nP;
            Tup *A = (Tup *) ctx->bufA.p;
            Tup *outX[2] = {A, A + regB};
            PartArgs a{};
            a.shift = 32 - pl.b1;
            a.mask = nd1 - 1;
            a.ndig = nd1;
            a.overflow = (u32 *) (m.scalars + kScOverflow);
            for (int i = 0; i < 2; ++i) {
                a.rel[i] = PartRel{inX0[i], outX[i], nX[i], m.hist1[i], m.cur1[i], nullptr, nullptr, 1, tiles_of(nX[i])};
                a.rel[i].limit_cap = opt[i] ? cap[i] : ((u64) 1 << 50);  // no bound for a relation with exact offsets
                a.rel[i].dump = i == 0 ? regB + regP : regP;             // one dump tile behind both relations
            }
            FixedArgs fa{};
            for (int i = 0; i < 2; ++i) {
                fa.cursor[i] = m.cur1[i];
                fa.seg_beg[i] = m.segb[i];
                fa.seg_end[i] = m.sege[i];
                fa.off1[i] = m.off1[i];
                fa.tile0[i] = m.tile0[i];
                fa.cap[i] = cap[i];
            }
            fa.ndig = nd1;
            fa.overflow = a.overflow;
            fa.rel_mask = (opt[0] ? 1u : 0u) | (opt[1] ? 2u : 0u);
            if (!(opt[0] && opt[1])) {
                // the skewed relation keeps its histogram + prefix sum (only its tiles are launched)
                const int x = opt[0] ? 1 : 0;
                PartArgs ah = a;
                ah.rel[x ^ 1].ntiles = 0;
                mark(ctx, st, RHJ_PHASE_HIST1);
                if ((rc = launch_hist(ctx, st, ah, kDigitHash, false))) return rc;
                ScanDigitsArgs sd{};
                sd.hist[0] = sd.hist[1] = m.hist1[x];
                sd.off[0] = sd.off[1] = m.off1[x];
                sd.cursor[0] = sd.cursor[1] = m.cur1[x];
                sd.tile0[0] = sd.tile0[1] = m.tile0[x];
                sd.ndig = nd1;
                mark(ctx, st, RHJ_PHASE_SCAN1);
                k_scan_digits<<<1, kMaxDigits, 0, st>>>(sd);
                CK(cudaGetLastError());
                ctx->info.kernel_launches++;
            }
            k_fixed_cursors<<<2, 256, 0, st>>>(fa);
            CK(cudaGetLastError());
            ctx->info.kernel_launches++;
            mark(ctx, st, RHJ_PHASE_SCATTER1);
            if ((rc = launch_scatter(ctx, st, a, kDigitHash, false, true))) return rc;
            mark(ctx, st, RHJ_PHASE_SCAN1);
            k_fixed_finish<<<2, kMaxDigits, 0, st>>>(fa);
            CK(cudaGetLastError());
            ctx->info.kernel_launches++;
            const Tup *inX[2] = {outX[0], outX[1]};
            const u64 *off1X[2] = {m.off1[0], m.off1[1]};
            const u32 *tile0X[2] = {m.tile0[0], m.tile0[1]};
            // a relation with exact offsets lies where pass 2 will pack it from: [off1[s], off1[s + 1])
            const u64 *segbX[2] = {opt[0] ? m.segb[0] : m.off1[0], opt[1] ? m.segb[1] : m.off1[1]};
            const u64 *segeX[2] = {opt[0] ? m.sege[0] : m.off1[0] + 1, opt[1] ? m.sege[1] : m.off1[1] + 1};
            ctx->info.optimistic_pass1 = fa.rel_mask;
            return second_pass_and_plan(ctx, st, pl, m, inX, off1X, tile0X, segbX, segeX, fixed2);
        }
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

        const Tup *inX[2] = {A, A + pl.nB};
        const u64 *off1X[2] = {m.off1[0], m.off1[1]};
        const u32 *tile0X[2] = {m.tile0[0], m.tile0[1]};
        return second_pass_and_plan(ctx, st, pl, m, inX, off1X, tile0X);
    }
    const Tup *inX[2] = {finB, finP};
    const u64 *off1X[2] = {offB, offP};
    const u32 *tile0X[2] = {nullptr, nullptr};
    return second_pass_and_plan(ctx, st, pl, m, inX, off1X, tile0X);
}

JoinArgs join_args(rhj_ctx *ctx, int work_slot) {
    u64 *sc = scalars_of(ctx, ctx->cur.nparts);
    JoinArgs j{};
    j.build = ctx->cur.build;
    j.probe = ctx->cur.probe;
    j.offB = ctx->cur.offB;
    j.offP = ctx->cur.offP;
    j.endB = ctx->cur.endB;
    j.endP = ctx->cur.endP;
    j.items = (const Item *) ctx->items.p;
    j.nitems = (const u32 *) (sc + kScNItems);
    j.work_counter = (u32 *) (sc + work_slot);
    j.out_cursor = sc + kScCursor;
    j.build_is_S = ctx->cur.build_is_S;
    return j;
}

constexpr int kRetryExact = 1000;  // internal: re-run the partition phase with the exact histogram path

// Book-keeping of the sample-free shortcut (rhj_ctx::trust): called with the outcome of the FIRST attempt of a join.
void settle_trust(rhj_ctx *ctx, bool overflowed) {
    auto &tr = ctx->trust;
    if (overflowed) {
        tr.streak = 0;
        return;
    }
    const auto &pe = ctx->pending;
    if (!pe.nB) return;  // this join did not sample (reused a verdict, or is too small for the optimistic passes)
    const bool same = tr.streak > 0 && tr.nB == pe.nB && tr.nP == pe.nP && tr.opt[0] == pe.opt[0] && tr.opt[1] == pe.opt[1] &&
                      tr.poisson[0] == pe.poisson[0] && tr.poisson[1] == pe.poisson[1];
    const int streak = same ? tr.streak + 1 : 1;
    tr = pe;
    tr.streak = streak;
    tr.age = 0;
}

int read_scalars(rhj_ctx *ctx, cudaStream_t st) {
    u64 *sc = scalars_of(ctx, ctx->cur.nparts);
    CK(cudaMemcpyAsync(ctx->h_scalars, sc, kScCount * sizeof(u64), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (ctx->h_scalars[kScOverflow]) return kRetryExact;  // the optimistic pass-1 layout was too small
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
    // a second write pass after one count (a retry into another buffer) must not find every item already taken
    CK(cudaMemsetAsync(scalars_of(ctx, ctx->cur.nparts) + kScWork1, 0, 8, st));
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

namespace {

// One radix pass (histogram -> prefix sum -> scatter) of ONE relation occupying meta slot `slot`.
// SEG = second pass inside the pass-1 partitions.  Enqueues only.
int one_relation_pass(rhj_ctx *ctx, cudaStream_t st, const Plan &pl, const Meta &m, int slot, bool second, const Tup *in,
                      Tup *out, u64 n) {
    int rc;
    PartArgs a{};
    if (!second) {
        a.shift = 32 - pl.b1;
        a.mask = (1u << pl.b1) - 1;
        a.ndig = 1u << pl.b1;
        a.rel[0] = PartRel{in, out, n, m.hist1[slot], m.cur1[slot], nullptr, nullptr, 1, tiles_of(n)};
        if ((rc = launch_hist(ctx, st, a, kDigitHash, false))) return rc;
        ScanDigitsArgs sd{};
        sd.hist[0] = sd.hist[1] = m.hist1[slot];
        sd.off[0] = sd.off[1] = m.off1[slot];
        sd.cursor[0] = sd.cursor[1] = m.cur1[slot];
        sd.tile0[0] = sd.tile0[1] = pl.b2 ? m.tile0[slot] : nullptr;
        sd.ndig = a.ndig;
        k_scan_digits<<<1, kMaxDigits, 0, st>>>(sd);
        CK(cudaGetLastError());
        ctx->info.kernel_launches++;
        return launch_scatter(ctx, st, a, kDigitHash, false);
    }
    const u32 nseg = 1u << pl.b1;
    a.shift = 32 - pl.bits;
    a.mask = (1u << pl.b2) - 1;
    a.ndig = 1u << pl.b2;
    a.rel[0] = PartRel{in, out, n, m.hist2[slot], m.cur2[slot], m.off1[slot], m.tile0[slot], nseg, tiles_of(n) + nseg};
    if ((rc = build_tile_tables(ctx, st, a, 1, slot))) return rc;
    if ((rc = launch_hist(ctx, st, a, kDigitHash, true))) return rc;
    ScanPartsRelArgs sr{};
    sr.hist2 = m.hist2[slot];
    sr.off1 = m.off1[slot];
    sr.off2 = m.off2[slot];
    sr.cursor2 = m.cur2[slot];
    sr.nseg = nseg;
    sr.ndig = a.ndig;
    k_scan_parts_rel<<<nseg, kMaxDigits, 0, st>>>(sr);
    CK(cudaGetLastError());
    ctx->info.kernel_launches++;
    return launch_scatter(ctx, st, a, kDigitHash, true);
}

// Partitions one relation completely (0, 1 or 2 passes).  On return *fin / *off describe its final
// partitions (meta slot `slot`).  tmpA / tmpB are the pass outputs (n tuples each).
int partition_relation(rhj_ctx *ctx, cudaStream_t st, const Plan &pl, const Meta &m, int slot, const Tup *in, u64 n,
                       Tup *tmpA, Tup *tmpB, const Tup **fin, const u64 **off) {
    int rc;
    if (pl.bits == 0) {
        k_set_single_part<<<1, 1, 0, st>>>(m.off2[slot], n, m.off2[slot], n);
        CK(cudaGetLastError());
        ctx->info.kernel_launches++;
        *fin = in;
        *off = m.off2[slot];
        return RHJ_OK;
    }
    if ((rc = one_relation_pass(ctx, st, pl, m, slot, false, in, tmpA, n))) return rc;
    if (pl.b2 == 0) {
        *fin = tmpA;
        *off = m.off1[slot];
        return RHJ_OK;
    }
    if ((rc = one_relation_pass(ctx, st, pl, m, slot, true, tmpA, tmpB, n))) return rc;
    *fin = tmpB;
    *off = m.off2[slot];
    return RHJ_OK;
}

// Host -> device copy of a caller's relation on stream `st`.  Pinned (or registered) sources go straight to the copy
// engine.  PAGEABLE sources -- what the reference's Query::run_joins passes: relation::tuples is `new tuple[]`,
// structs.cpp:217-243 -- would make cudaMemcpyAsync stage them through the driver's single-threaded bounce buffer at
// ~13 GB/s and block the calling thread; here the host fills a pinned ring (4 x 32 MiB) with a few memcpy threads (RHJ_STAGE_THREADS, default 8)
// while the copy engine drains the slot before (measured on the 2^27 x 2^27 end-to-end join: 416 -> see DESIGN.md).
// Returns when the last slice has been ENQUEUED; the copies complete in stream order.
int upload_host(rhj_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, cudaStream_t st) {
    if (!bytes) return RHJ_OK;
    cudaPointerAttributes at{};
    const bool pageable = cudaPointerGetAttributes(&at, h_src) != cudaSuccess || at.type == cudaMemoryTypeUnregistered;
    cudaGetLastError();
    if (!pageable || bytes < ((size_t) 8 << 20)) {
        CK(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
        return RHJ_OK;
    }
    if (!ctx->stage_pin) {
        cudaError_t e = cudaHostAlloc(&ctx->stage_pin, rhj_ctx::kStageSlot * rhj_ctx::kStageSlots, cudaHostAllocDefault);
        if (e != cudaSuccess) {
            cudaGetLastError();
            ctx->stage_pin = nullptr;
            CK(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));  // no pinned memory left: the driver's path
            return RHJ_OK;
        }
        for (auto &ev : ctx->stage_ev) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    }
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const int nthr = (int) std::min<unsigned>(ctx->stage_threads, std::max(1u, hw / 2));
    for (size_t off = 0; off < bytes; off += rhj_ctx::kStageSlot) {
        const size_t len = std::min(rhj_ctx::kStageSlot, bytes - off);
        const int slot = ctx->stage_next;
        ctx->stage_next = (slot + 1) % rhj_ctx::kStageSlots;
        char *pin = (char *) ctx->stage_pin + (size_t) slot * rhj_ctx::kStageSlot;
        CK(cudaEventSynchronize(ctx->stage_ev[slot]));  // the copy engine is done with this slot (a fresh event is complete)
        const char *src = (const char *) h_src + off;
        if (nthr > 1 && len >= ((size_t) 4 << 20)) {
            const size_t per = ((len / nthr) + 4095) & ~(size_t) 4095;
            std::vector<std::thread> helpers;
            for (int t = 1; t < nthr; ++t) {
                const size_t b = std::min(len, (size_t) t * per), e2 = std::min(len, (size_t) (t + 1) * per);
                if (e2 > b) helpers.emplace_back([=] { memcpy(pin + b, src + b, e2 - b); });
            }
            memcpy(pin, src, std::min(len, per));
            for (auto &h : helpers) h.join();
        } else {
            memcpy(pin, src, len);
        }
        CK(cudaMemcpyAsync((char *) d_dst + off, pin, len, cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(ctx->stage_ev[slot], st));
    }
    return RHJ_OK;
}

// Pipelined host join for large inputs: the build side is uploaded and partitioned once; the probe
// side streams through in chunks -- H2D of chunk c+1, partition + count + write of chunk c and D2H
// of chunk c-1's pairs run concurrently (three streams, double-buffered chunk and result buffers),
// so the PCIe link is busy in both directions instead of H2D, compute and D2H taking turns.
// Every chunk re-builds the shared-memory tables from the (resident) build partitions.
int join_host_pipelined(rhj_ctx *ctx, const Tup *hR, u64 nR, const Tup *hS, u64 nS, const rhj_pair **out, uint64_t *count) {
    cudaStream_t st = ctx->stream;
    int rc;
    if (!ctx->s_in) {
        CK(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CK(cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->ev_cmp[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming));
        }
    }
    Plan pl = make_plan(nR, nS);
    const Tup *hB = pl.build_is_S ? hS : hR, *hP = pl.build_is_S ? hR : hS;
    ctx->info = rhj_plan_info{};
    ctx->info.bits_total = pl.bits;
    ctx->info.bits_pass1 = pl.b1;
    ctx->info.bits_pass2 = pl.b2;
    ctx->info.build_is_S = pl.build_is_S;
    ctx->info.n_partitions = pl.nparts;
    ctx->cur.valid = false;
    ctx->nmarks = 0;
    Meta m;
    if ((rc = layout_meta(ctx, pl.nparts, m))) return rc;
    const u64 chunk = std::min<u64>(ctx->host_chunk, pl.nP);
    const u64 nchunks = (pl.nP + chunk - 1) / chunk;

    // ---- build side: upload, partition, keep ----
    if ((rc = ensure(ctx, ctx->inR, pl.nB * sizeof(Tup)))) return rc;
    if ((rc = ensure(ctx, ctx->bufA, pl.nB * sizeof(Tup)))) return rc;
    if ((rc = ensure(ctx, ctx->bufB, pl.nB * sizeof(Tup)))) return rc;
    for (int i = 0; i < 2; ++i)
        if ((rc = ensure(ctx, ctx->pin[i], chunk * sizeof(Tup)))) return rc;
    if ((rc = ensure(ctx, ctx->pA, chunk * sizeof(Tup)))) return rc;
    if ((rc = ensure(ctx, ctx->pB, chunk * sizeof(Tup)))) return rc;
    u64 cap64 = (u64) pl.nparts + chunk / kProbeChunk + 2;
    if (cap64 > 0x7fffffffull) return fail(ctx, RHJ_ERR_ARG, "relation too large for the work-item table");
    const u32 item_cap = (u32) cap64;
    if ((rc = ensure(ctx, ctx->items, (size_t) item_cap * sizeof(Item)))) return rc;
    if ((rc = ensure(ctx, ctx->item_cnt, (size_t) item_cap * 8))) return rc;
    if ((rc = ensure(ctx, ctx->item_off, (size_t) item_cap * 8))) return rc;

    if ((rc = upload_host(ctx, ctx->inR.p, hB, pl.nB * sizeof(Tup), st))) return rc;
    // Probe chunk c goes up on s_in from a helper thread (pageable sources are staged through the pinned ring there), so
    // that the staging of chunk c + 1 overlaps the kernels AND the host-side count read-back of chunk c.
    std::future<int> up;
    auto start_upload = [&](u64 c) {
        const int nb = (int) (c & 1);
        const u64 n_up = std::min(chunk, pl.nP - c * chunk);
        up = std::async(std::launch::async, [=]() -> int {
            CK(cudaSetDevice(ctx->device));
            if (c >= 2) CK(cudaStreamWaitEvent(ctx->s_in, ctx->ev_cmp[nb], 0));  // the buffer was last read by chunk c - 2's kernels
            int rc2 = upload_host(ctx, ctx->pin[nb].p, hP + c * chunk, n_up * sizeof(Tup), ctx->s_in);
            if (rc2) return rc2;
            CK(cudaEventRecord(ctx->ev_in[nb], ctx->s_in));
            return RHJ_OK;
        });
    };
    start_upload(0);  // the first probe chunk goes up while the build side is partitioned
    CK(cudaMemsetAsync(ctx->zero.p, 0, m.zero_bytes, st));
    const Tup *finB;
    const u64 *offB;
    if ((rc = partition_relation(ctx, st, pl, m, 0, (const Tup *) ctx->inR.p, pl.nB, (Tup *) ctx->bufA.p, (Tup *) ctx->bufB.p,
                                 &finB, &offB)))
        return rc;

    u64 total = 0;
    u64 *sc = m.scalars;
    for (u64 c = 0; c < nchunks; ++c) {
        const int b = (int) (c & 1);
        const u64 n_c = std::min(chunk, pl.nP - c * chunk);
        if ((rc = up.get())) return rc;             // chunk c is on its way (ev_in[b] recorded)
        if (c + 1 < nchunks) start_upload(c + 1);   // staged while this chunk is computed
        CK(cudaStreamWaitEvent(st, ctx->ev_in[b], 0));
        // counters of the probe side and the per-join scalars start from zero for every chunk
        CK(cudaMemsetAsync(m.hist1[1], 0, (size_t) kMaxDigits * 8, st));
        CK(cudaMemsetAsync(m.hist2[1], 0, ((size_t) pl.nparts + kScCount) * 8, st));  // hist2[1] | scalars are contiguous
        const Tup *finP;
        const u64 *offP;
        if ((rc = partition_relation(ctx, st, pl, m, 1, (const Tup *) ctx->pin[b].p, n_c, (Tup *) ctx->pA.p, (Tup *) ctx->pB.p,
                                     &finP, &offP)))
            return rc;
        PlanPartsArgs pa{};
        pa.offB = offB;
        pa.offP = offP;
        pa.ndig = std::min<u32>(pl.nparts, kMaxDigits);
        pa.items = (Item *) ctx->items.p;
        pa.item_cap = item_cap;
        pa.nitems = (u32 *) (sc + kScNItems);
        pa.err = (u32 *) (sc + kScErr);
        k_plan_parts<<<pl.nparts / pa.ndig, kMaxDigits, 0, st>>>(pa);
        CK(cudaGetLastError());
        ctx->info.kernel_launches++;
        ctx->cur.valid = true;
        ctx->cur.counted = false;
        ctx->cur.build = finB;
        ctx->cur.probe = finP;
        ctx->cur.offB = offB;
        ctx->cur.offP = offP;
        ctx->cur.endB = offB + 1;
        ctx->cur.endP = offP + 1;
        ctx->cur.nparts = pl.nparts;
        ctx->cur.item_cap = item_cap;
        ctx->cur.build_is_S = pl.build_is_S;
        if ((rc = count_phase(ctx, st))) return rc;  // host sync: this chunk's exact result size
        const u64 n_out = ctx->cur.count;
        if (n_out) {
            if ((total + n_out) * sizeof(Pair) > ctx->h_out_cap) {
                // grow the pinned result, keeping what earlier chunks delivered (their copies must have landed)
                CK(cudaStreamSynchronize(ctx->s_out));
                const u64 guess = std::max<u64>(total + n_out, (total + n_out) / (c + 1) * nchunks + (total + n_out) / 8);
                void *nu = nullptr;
                cudaError_t e = cudaHostAlloc(&nu, guess * sizeof(Pair), cudaHostAllocDefault);
                if (e != cudaSuccess) {
                    cudaGetLastError();
                    return fail(ctx, RHJ_ERR_NOMEM, "cudaHostAlloc result", e);
                }
                if (total) memcpy(nu, ctx->h_out, total * sizeof(Pair));
                if (ctx->h_out) cudaFreeHost(ctx->h_out);
                ctx->h_out = nu;
                ctx->h_out_cap = guess * sizeof(Pair);
            }
            if (c >= 2) CK(cudaStreamWaitEvent(st, ctx->ev_out[b], 0));  // result buffer b is free again
            if (n_out * sizeof(Pair) > ctx->pout[b].cap) CK(cudaStreamSynchronize(ctx->s_out));  // ... before it is regrown
            if ((rc = ensure(ctx, ctx->pout[b], n_out * sizeof(Pair)))) return rc;
            if ((rc = write_phase(ctx, st, (Pair *) ctx->pout[b].p, n_out))) return rc;
        }
        CK(cudaEventRecord(ctx->ev_cmp[b], st));
        if (n_out) {
            CK(cudaStreamWaitEvent(ctx->s_out, ctx->ev_cmp[b], 0));
            CK(cudaMemcpyAsync((Pair *) ctx->h_out + total, ctx->pout[b].p, n_out * sizeof(Pair), cudaMemcpyDeviceToHost,
                               ctx->s_out));
            CK(cudaEventRecord(ctx->ev_out[b], ctx->s_out));
        }
        total += n_out;
    }
    CK(cudaStreamSynchronize(ctx->s_out));
    CK(cudaStreamSynchronize(st));
    ctx->cur.valid = false;
    *count = total;
    *out = total ? (const rhj_pair *) ctx->h_out : nullptr;
    return RHJ_OK;
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
    if ((e = getenv("RHJ_NO_OPT"))) ctx->optimistic = atoi(e) == 0;
    if ((e = getenv("RHJ_FORCE_OPT"))) ctx->force_optimistic = atoi(e) != 0;
    if ((e = getenv("RHJ_NO_OPT2"))) ctx->optimistic2 = atoi(e) == 0;
    if ((e = getenv("RHJ_NO_TRUST"))) ctx->trust_sample = atoi(e) == 0;
    if ((e = getenv("RHJ_NO_POS"))) ctx->positional = atoi(e) == 0;
    if ((e = getenv("RHJ_JOIN_LEAN"))) ctx->join_lean = atoi(e) != 0;
    if ((e = getenv("RHJ_JOIN_POS_V"))) ctx->join_pos_v = atoi(e);
    if ((e = getenv("RHJ_NO_SHARD_OPT2"))) ctx->shard_optimistic2 = atoi(e) == 0;
    if ((e = getenv("RHJ_SHARD_OPT2_WORLD"))) ctx->shard_opt2_world = (u32) std::max(0, atoi(e));
    if ((e = getenv("RHJ_HOST_CHUNK"))) ctx->host_chunk = std::max<long long>(1, atoll(e));
    if ((e = getenv("RHJ_STAGE_THREADS"))) ctx->stage_threads = (unsigned) std::max(1, std::min(32, atoi(e)));
    if ((e = getenv("RHJ_SHARD_SCATTER_MODE"))) ctx->shard_scatter_mode = atoi(e);
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
    if (ctx->stage_pin) cudaFreeHost(ctx->stage_pin);
    for (cudaEvent_t e : ctx->stage_ev)
        if (e) cudaEventDestroy(e);
    if (ctx->h_iu) cudaFreeHost(ctx->h_iu);
    if (ctx->h_scalars) cudaFreeHost(ctx->h_scalars);
    for (cudaEvent_t e : ctx->ev)
        if (e) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) {
        if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]);
        if (ctx->ev_cmp[i]) cudaEventDestroy(ctx->ev_cmp[i]);
        if (ctx->ev_out[i]) cudaEventDestroy(ctx->ev_out[i]);
    }
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
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
    size_t a_tuples = pl.nB + pl.nP;
    if (pl.b2 > 0 && ctx->optimistic)  // the optimistic pass-1 layout gives every partition a fixed region with headroom
        a_tuples = std::max<size_t>(a_tuples, ((size_t) 1 << pl.b1) * (fixed_cap(pl.nB, 1u << pl.b1) + fixed_cap(pl.nP, 1u << pl.b1)) + kTile);
    if (pl.bits > 0 && (rc = ensure(ctx, ctx->bufA, a_tuples * sizeof(Tup)))) return rc;
    size_t b_tuples = pl.nB + pl.nP;
    if (pl.b2 > 0 && ctx->optimistic && ctx->optimistic2)  // fixed-capacity final partitions
        b_tuples = std::max<size_t>(b_tuples, (size_t) pl.nparts * (fixed_cap2(pl.nB, pl.nparts) + fixed_cap2(pl.nP, pl.nparts)) + kTile);
    if (pl.b2 > 0 && (rc = ensure(ctx, ctx->bufB, b_tuples * sizeof(Tup)))) return rc;
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
    for (int attempt = 0; attempt < 2; ++attempt) {  // attempt 1 = exact histogram path after an optimistic overflow
        if ((rc = partition_and_plan(ctx, st, (const Tup *) dR, nR, (const Tup *) dS, nS, attempt == 0))) return rc;
        rc = count_phase(ctx, st);
        if (attempt == 0 && (rc == RHJ_OK || rc == kRetryExact)) settle_trust(ctx, rc == kRetryExact);
        if (rc != kRetryExact) break;
    }
    if (rc) return rc;
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
    // positional emit (one output slot per probe tuple, holes closed afterwards) needs room for a slot per probe tuple
    bool pos = ctx->positional && capacity >= std::max(nR, nS);
    if (pos && ctx->pos_skip > 0) {
        ctx->pos_skip--;
        pos = false;
    }
    bool allow_opt = true;
    for (int attempt = 0; attempt < 4; ++attempt) {
        if ((rc = partition_and_plan(ctx, st, (const Tup *) dR, nR, (const Tup *) dS, nS, allow_opt))) return rc;
        JoinArgs j = join_args(ctx, kScWork0);
        j.out = (Pair *) d_out;
        j.capacity = capacity;
        j.holes = scalars_of(ctx, ctx->cur.nparts) + kScHoles;
        mark(ctx, st, RHJ_PHASE_JOIN);
        if (pos) rc = launch_join_positional(ctx, st, j, ctx->cur.item_cap);
        else rc = launch_join<kJoinFused>(ctx, st, j, ctx->cur.item_cap);
        if (rc) return rc;
        mark(ctx, st, -1);
        rc = read_scalars(ctx, st);
        if (attempt == 0 && (rc == RHJ_OK || rc == kRetryExact)) settle_trust(ctx, rc == kRetryExact);
        if (rc == kRetryExact && allow_opt) {  // an optimistic layout overflowed: exact histogram path
            allow_opt = false;
            continue;
        }
        if (rc || !pos) break;
        const u64 cursor = ctx->h_scalars[kScCursor], holes = ctx->h_scalars[kScHoles];
        if (cursor > capacity) {  // slots + ranked items did not fit: the ranked emitter alone tells the exact need
            pos = false;
            continue;
        }
        if (holes) {
            bool ok = false;
            if ((rc = close_holes(ctx, st, (Pair *) d_out, cursor, holes, &ok))) return rc;
            if (!ok) {
                pos = false;
                continue;
            }
            ctx->h_scalars[kScCursor] = cursor - holes;
            if (holes * 64 > cursor) ctx->pos_skip = 16;  // many probe tuples without a match: the ranked emitter is cheaper
        }
        break;
    }
    if (rc == kRetryExact) return fail(ctx, RHJ_ERR_STATE, "overflow flag set on the exact path");
    if (rc) return rc;
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
    // large probe sides stream through in chunks so that H2D, compute and D2H overlap
    if (std::max(nR, nS) >= 2 * ctx->host_chunk)
        return join_host_pipelined(ctx, (const Tup *) R, nR, (const Tup *) S, nS, out, count);
    if ((rc = ensure(ctx, ctx->inR, nR * sizeof(Tup)))) return rc;
    if ((rc = ensure(ctx, ctx->inS, nS * sizeof(Tup)))) return rc;
    if ((rc = upload_host(ctx, ctx->inR.p, R, nR * sizeof(Tup), st))) return rc;
    if ((rc = upload_host(ctx, ctx->inS.p, S, nS * sizeof(Tup), st))) return rc;
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
        if (!page) {  // out of host memory: free what was built and report it (count > 0 with a NULL head and *head_size == 0)
            while (head) {
                void *next = *(void **) head;
                free(head);
                head = next;
            }
            if (head_size) *head_size = 0;
            return nullptr;
        }
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
    if (!ctx || !d_hist || bits < 0 || bits > kPlanBitsPerPass || shift < 0 || shift > 63) return RHJ_ERR_ARG;
    if (digit_kind != RHJ_DIGIT_RAW && digit_kind != RHJ_DIGIT_HASH) return RHJ_ERR_ARG;
    // the hashed digit is taken from a 32-bit word, the raw one from the 64-bit value
    if (shift + bits > (digit_kind == RHJ_DIGIT_HASH ? 32 : 64)) return RHJ_ERR_ARG;
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
    if (!ctx || bits < 0 || bits > kPlanBitsPerPass || shift < 0 || shift > 63) return RHJ_ERR_ARG;
    if (digit_kind != RHJ_DIGIT_RAW && digit_kind != RHJ_DIGIT_HASH) return RHJ_ERR_ARG;
    if (shift + bits > (digit_kind == RHJ_DIGIT_HASH ? 32 : 64)) return RHJ_ERR_ARG;
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

// ---- multi-GPU -------------------------------------------------------------------------

int rhj_shard_plan_make(uint64_t nR_global, uint64_t nS_global, int world, rhj_shard_plan *plan) {
    if (!plan || world < 1 || world > kMaxPeers || (world & (world - 1))) return RHJ_ERR_ARG;
    int rb = 0;
    while ((1 << rb) < world) ++rb;
    const u64 nB = (std::min(nR_global, nS_global) + world - 1) / world;  // expected build tuples per rank
    int bits = 0;
    if (nB > kBuildCap)
        while (bits < 2 * kPlanBitsPerPass && (nB >> bits) > kTargetBuildPerPart) ++bits;
    int b1 = std::min(kMaxBitsPerPass - rb, (bits + 1) / 2);  // balanced: the received data always gets a second pass
    // ... but a 1024-digit scatter costs ~50 % more per tuple than a 512-digit one (4-tuple runs): when the second pass can
    // take the bit and still fit 512 digits, keep the first pass at rank bits + sub-digit bits <= 9 (4 ranks: 2 + 7 | 9)
    if (rb + b1 > kPlanBitsPerPass && bits - (kPlanBitsPerPass - rb) <= kPlanBitsPerPass) b1 = std::max(0, kPlanBitsPerPass - rb);
    if (const char *e = getenv("RHJ_SHARD_SUBBITS")) b1 = std::max(0, std::min(std::min(atoi(e), bits), kMaxBitsPerPass - rb));
    if (bits - b1 > kMaxBitsPerPass) bits = b1 + kMaxBitsPerPass;
    plan->world = world;
    plan->rank_bits = rb;
    plan->bits_total = bits;
    plan->bits_pass1 = b1;
    plan->bits_pass2 = bits - b1;
    plan->build_is_S = nS_global < nR_global;
    return RHJ_OK;
}

static PartArgs shard_args(const rhj_shard_plan *sp) {
    PartArgs a{};
    a.rank_bits = sp->rank_bits;
    a.sub_bits = sp->bits_pass1;
    a.ndig = (u32) sp->world << sp->bits_pass1;
    a.mask = a.ndig - 1;
    a.shift = 0;
    return a;
}

// ---- multi-GPU, DMA-shipped variant: pass 1 partitions locally on (destination rank | sub-digit) into
// a staging buffer, the copy engines ship one contiguous chunk per destination over NVLink while the
// SMs work on the other relation, and pass 2 consumes the received chunks as (source, partition) pieces.

namespace {

struct ShardMeta {
    u64 *loc_off[2];   // [kMaxDigits + 1] source side: offsets of the (dest | p1) digits inside the staging buffer
    u64 *seg_off[2];   // [kMaxDigits + 1] destination side: piece boundaries inside the receive buffer
    u64 *tot;          // [kMaxPeers * kMaxPeers] ship matrix of the relation being laid out
};

int layout_shard_meta(rhj_ctx *ctx, ShardMeta &sm) {
    size_t n = 4 * (size_t) (kMaxDigits + 1) + (size_t) kMaxPeers * kMaxPeers;
    int rc = ensure(ctx, ctx->shard_meta, n * 8);
    if (rc) return rc;
    u64 *q = (u64 *) ctx->shard_meta.p;
    for (int i = 0; i < 2; ++i) { sm.loc_off[i] = q; q += kMaxDigits + 1; }
    for (int i = 0; i < 2; ++i) { sm.seg_off[i] = q; q += kMaxDigits + 1; }
    sm.tot = q;
    return RHJ_OK;
}

// Per-slot view of the sharded join's metadata: slots 0 / 1 are relations R / S (the arrays of Meta / ShardMeta).
struct SlotArrays {
    u64 *cur1, *loc_off, *seg_off, *off1, *hist2, *cur2, *off2;
    u32 *tile0;
    DevBuf *out;
};
int slot_arrays(rhj_ctx *ctx, u32 nparts, int slot, SlotArrays &a) {
    Meta m;
    ShardMeta sm;
    int rc;
    if ((rc = layout_meta(ctx, nparts, m))) return rc;
    if ((rc = layout_shard_meta(ctx, sm))) return rc;
    if (slot < 0 || slot > 1) return RHJ_ERR_ARG;
    a = SlotArrays{m.cur1[slot], sm.loc_off[slot], sm.seg_off[slot], m.off1[slot], m.hist2[slot], m.cur2[slot], m.off2[slot],
                   m.tile0[slot], slot ? &ctx->bufB2 : &ctx->bufB};
    return RHJ_OK;
}

Plan shard_local_plan(const rhj_shard_plan *sp, u64 nR, u64 nS) {
    Plan pl{};
    pl.build_is_S = sp->build_is_S;
    pl.nB = pl.build_is_S ? nS : nR;
    pl.nP = pl.build_is_S ? nR : nS;
    pl.bits = sp->bits_total;
    pl.b1 = sp->bits_pass1;
    pl.b2 = sp->bits_pass2;
    pl.nparts = 1u << pl.bits;
    return pl;
}

}  // namespace

// Starts a sharded join on this context: sizes the metadata, zeroes the counters.  Enqueues only.
int rhj_shardx_begin(rhj_ctx *ctx, const rhj_shard_plan *sp, void *stream) {
    if (!ctx || !sp) return RHJ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    ctx->cur.valid = false;
    ctx->cur.counted = false;
    ctx->nmarks = 0;
    ctx->info = rhj_plan_info{};
    ctx->info.bits_total = sp->bits_total;
    ctx->info.bits_pass1 = sp->bits_pass1;
    ctx->info.bits_pass2 = sp->bits_pass2;
    ctx->info.build_is_S = sp->build_is_S;
    ctx->info.n_partitions = 1u << sp->bits_total;
    Meta m;
    int rc;
    if ((rc = layout_meta(ctx, 1u << sp->bits_total, m))) return rc;
    CK(cudaMemsetAsync(ctx->zero.p, 0, m.zero_bytes, st));
    ctx->shard_n[0] = ctx->shard_n[1] = ctx->shard_n[2] = 0;
    ctx->shard_cap[0] = ctx->shard_cap[1] = ctx->shard_cap[2] = 0;
    ctx->shard_poisson[0] = ctx->shard_poisson[1] = ctx->shard_poisson[2] = false;
    ctx->shard_count = 0;
    return RHJ_OK;
}

// Pass 1 of relation `rel` (0 = R, 1 = S): histogram on
// (destination rank | sub-digit) into d_hist[world << bits_pass1] (the caller all-gathers it), prefix
// sum, scatter into the local staging buffer d_stage[n], which ends up ordered by destination rank,
// then by pass-1 partition.  Enqueues only.
int rhj_shardx_pass1_device(rhj_ctx *ctx, const rhj_shard_plan *sp, int rel, const rhj_tuple *d_in, uint64_t n,
                            rhj_tuple *d_stage, uint64_t *d_hist, void *stream) {
    if (!ctx || !sp || !d_hist || rel < 0 || rel > 1 || (n && (!d_in || !d_stage))) return RHJ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    SlotArrays sl;
    int rc;
    if ((rc = slot_arrays(ctx, 1u << sp->bits_total, rel, sl))) return rc;
    PartArgs a = shard_args(sp);
    a.shard_local = 1;
    CK(cudaMemsetAsync(d_hist, 0, (size_t) a.ndig * sizeof(u64), st));
    a.rel[0] = PartRel{(const Tup *) d_in, (Tup *) d_stage, n, (u64 *) d_hist, sl.cur1, nullptr, nullptr, 1, tiles_of(n)};
    if (rel == 0) mark(ctx, st, RHJ_PHASE_HIST1);
    if ((rc = launch_hist(ctx, st, a, kDigitShard, false))) return rc;
    ScanDigitsArgs sd{};
    sd.hist[0] = sd.hist[1] = (const u64 *) d_hist;
    sd.off[0] = sd.off[1] = sl.loc_off;
    sd.cursor[0] = sd.cursor[1] = sl.cur1;
    sd.tile0[0] = sd.tile0[1] = nullptr;
    sd.ndig = a.ndig;
    k_scan_digits<<<1, kMaxDigits, 0, st>>>(sd);
    CK(cudaGetLastError());
    ctx->info.kernel_launches++;
    return launch_scatter(ctx, st, a, kDigitShard, false);
}

// Layout of slot `rel` from the all-gathered histograms d_all_hist[world][world << bits_pass1]:
// device side, the piece tables pass 2 needs; host side, what to ship where:
//   send_off[d], send_cnt[d]  this rank's chunk for destination d inside its staging buffer (tuples)
//   dst_off[d]                where that chunk starts inside destination d's receive buffer
//   *recv_total               tuples this rank receives
//   *recv_max                 (optional) the largest recv_total of any rank: the same number on every rank, so a receive
//                             buffer that is too small SOMEWHERE is an error every rank can raise together (a rank that
//                             stops alone leaves its peers waiting in the exchange)
// Synchronises the stream.
int rhj_shardx_layout_device(rhj_ctx *ctx, const rhj_shard_plan *sp, int rank, int rel, const uint64_t *d_all_hist,
                             uint64_t *send_off, uint64_t *send_cnt, uint64_t *dst_off, uint64_t *recv_total,
                             uint64_t *recv_max, void *stream) {
    if (!ctx || !sp || !d_all_hist || !send_off || !send_cnt || !dst_off || !recv_total || rel < 0 || rel > 1 || rank < 0 ||
        rank >= (int) sp->world)
        return RHJ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    SlotArrays sl;
    ShardMeta sm;
    int rc;
    if ((rc = slot_arrays(ctx, 1u << sp->bits_total, rel, sl))) return rc;
    if ((rc = layout_shard_meta(ctx, sm))) return rc;
    ShardLayoutArgs a{};
    a.all_hist = (const u64 *) d_all_hist;
    a.seg_off = sl.seg_off;
    a.seg_tile0 = sl.tile0;
    a.off1 = sl.off1;
    a.tot = sm.tot;
    a.world = sp->world;
    a.rank = (u32) rank;
    a.sub_bits = sp->bits_pass1;
    k_shard_layout<<<1, 1024, 0, st>>>(a);
    CK(cudaGetLastError());
    ctx->info.kernel_launches++;
    const u32 W = sp->world, nd1 = 1u << sp->bits_pass1;
    static thread_local u64 h_tot[kMaxPeers * kMaxPeers], h_loc[kMaxDigits + 1], h_off1[kMaxDigits + 1];
    CK(cudaMemcpyAsync(h_tot, sm.tot, (size_t) W * W * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_loc, sl.loc_off, ((size_t) (W << sp->bits_pass1) + 1) * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_off1, sl.off1, ((size_t) nd1 + 1) * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    {
        // Index of dispersion of the pass-1 partition sizes this rank receives (exact counts): 1 for hashed distinct keys,
        // f for keys repeated f times.  Only Poisson-like data gets the histogram-free second pass.
        double sum = 0, sq = 0;
        for (u32 i = 0; i < nd1; ++i) {
            const double c = (double) (h_off1[i + 1] - h_off1[i]);
            sum += c;
            sq += c * c;
        }
        const double mean = sum / nd1, var = sq / nd1 - mean * mean;
        ctx->shard_poisson[rel] = nd1 >= 16 && mean >= 4096.0 && var <= 1.75 * mean;
    }
    u64 total = 0;
    for (u32 d = 0; d < W; ++d) {
        send_off[d] = h_loc[(size_t) d * nd1];
        send_cnt[d] = h_loc[(size_t) (d + 1) * nd1] - h_loc[(size_t) d * nd1];
        u64 before = 0;
        for (u32 s2 = 0; s2 < (u32) rank; ++s2) before += h_tot[s2 * W + d];
        dst_off[d] = before;
        total += h_tot[d * W + rank];
    }
    *recv_total = total;
    if (recv_max) {
        u64 worst = 0;
        for (u32 d = 0; d < W; ++d) {
            u64 in = 0;
            for (u32 s2 = 0; s2 < W; ++s2) in += h_tot[s2 * W + d];
            worst = std::max(worst, in);
        }
        *recv_max = worst;
    }
    ctx->shard_n[rel] = total;
    return RHJ_OK;
}

// Pass 2 of slot `rel` over what this rank received (d_recv[n_recv], world << bits_pass1 pieces):
// histogram, per-partition offsets, scatter into the slot's final partition buffer.  Enqueues only.
static int shardx_pass2_impl(rhj_ctx *ctx, const rhj_shard_plan *sp, int rel, const rhj_tuple *d_recv, uint64_t n_recv, void *stream,
                             bool allow_fixed = true) {
    if (!ctx || !sp || rel < 0 || rel > 1 || (n_recv && !d_recv)) return RHJ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    SlotArrays sl;
    int rc;
    const u32 nparts = 1u << sp->bits_total;
    if ((rc = slot_arrays(ctx, nparts, rel, sl))) return rc;
    ctx->shard_recv[rel] = (const Tup *) d_recv;
    // histogram-free second pass (fixed-capacity final partitions) when the received sizes look Poisson.  On by default
    // only where it was measured to help: 8.29 -> 7.49 ms per join on 2 GPUs; the one run on 8 GPUs that the round's GPU
    // budget allowed gave 12.2 instead of 10.3-10.6 ms, 4 GPUs are unmeasured (profiles/r01_multi_gpu_notes.md).
    // RHJ_SHARD_OPT2_WORLD=<n> raises the limit for measurements.
    bool fixed = allow_fixed && ctx->optimistic && ctx->shard_optimistic2 && d_recv && sp->bits_pass2 > 0 && sp->bits_pass2 <= 9 &&
                 (sp->world <= ctx->shard_opt2_world || ctx->force_optimistic) &&
                 n_recv >= ((u64) 1 << 16) && (ctx->shard_poisson[rel] || ctx->force_optimistic);
    if (fixed && ctx->opt2_skip > 0 && !ctx->force_optimistic) {
        ctx->opt2_skip--;
        fixed = false;
    }
    const u64 cap = fixed ? fixed_cap2(n_recv, nparts) : 0;
    ctx->shard_cap[rel] = cap;
    if ((rc = ensure(ctx, *sl.out, (fixed ? (u64) nparts * cap + kTile : std::max<u64>(n_recv, 1)) * sizeof(Tup)))) return rc;
    const u32 nd1 = 1u << sp->bits_pass1, npieces = sp->world << sp->bits_pass1;
    PartArgs b{};
    // bits_pass2 == 0 (tiny relations): a one-digit pass that only merges the pieces of a partition
    b.shift = std::min(31, 32 - (int) sp->bits_total);
    b.mask = (1u << sp->bits_pass2) - 1;
    b.ndig = 1u << sp->bits_pass2;
    b.rel[0] = PartRel{(const Tup *) d_recv, (Tup *) sl.out->p, n_recv, sl.hist2, sl.cur2, sl.seg_off, sl.tile0,
                       npieces, tiles_of(n_recv) + npieces, nd1 - 1};
    if (nd1 == 1) b.rel[0].group_mask = 0x80000000u;  // every piece is partition 0: (seg & mask) == 0
    if ((rc = build_tile_tables(ctx, st, b, 1, rel))) return rc;
    if (fixed) {
        Meta m;
        if ((rc = layout_meta(ctx, nparts, m))) return rc;
        b.overflow = (u32 *) (m.scalars + kScOverflow);
        b.rel[0].limit_cap = cap;
        b.rel[0].dump = (u64) nparts * cap;
        PlanFixedArgs pf{};
        pf.end[0] = sl.cur2;
        pf.beg[0] = sl.off2;
        pf.cap[0] = cap;
        pf.nseg = nd1;
        pf.ndig = b.ndig;
        k_fixed_cursors2<<<dim3((nparts + 255) / 256, 1), 256, 0, st>>>(pf);
        CK(cudaGetLastError());
        ctx->info.kernel_launches++;
        if (rel == 0) mark(ctx, st, RHJ_PHASE_SCATTER2);
        return launch_scatter(ctx, st, b, kDigitHash, true, true);
    }
    if (rel == 0) mark(ctx, st, RHJ_PHASE_HIST2);
    if ((rc = launch_hist(ctx, st, b, kDigitHash, true))) return rc;
    ScanPartsRelArgs sr{};
    sr.hist2 = sl.hist2;
    sr.off1 = sl.off1;
    sr.off2 = sl.off2;
    sr.cursor2 = sl.cur2;
    sr.nseg = nd1;
    sr.ndig = b.ndig;
    k_scan_parts_rel<<<nd1, kMaxDigits, 0, st>>>(sr);
    CK(cudaGetLastError());
    ctx->info.kernel_launches++;
    if ((rc = launch_scatter(ctx, st, b, kDigitHash, true))) return rc;
    return RHJ_OK;
}

int rhj_shardx_pass2_device(rhj_ctx *ctx, const rhj_shard_plan *sp, int rel, const rhj_tuple *d_recv, uint64_t n_recv,
                            void *stream) {
    return shardx_pass2_impl(ctx, sp, rel, d_recv, n_recv, stream);
}

// Work-item plan + build/probe + fused emit over the final partitions of both relations (slots 0 = R, 1 = S).
int rhj_shardx_join_device(rhj_ctx *ctx, const rhj_shard_plan *sp, rhj_pair *d_out, uint64_t capacity, uint64_t *count,
                           void *stream) {
    if (!ctx || !sp || !count) return RHJ_ERR_ARG;
    const int build_slot = sp->build_is_S ? 1 : 0, probe_slot = build_slot ^ 1;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    const u64 nB = ctx->shard_n[build_slot], nP = ctx->shard_n[probe_slot];
    const u32 nparts = 1u << sp->bits_total;
    Meta m;
    SlotArrays sb, spb;
    int rc;
    if ((rc = layout_meta(ctx, nparts, m))) return rc;
    if ((rc = slot_arrays(ctx, nparts, build_slot, sb))) return rc;
    if ((rc = slot_arrays(ctx, nparts, probe_slot, spb))) return rc;
    ctx->shard_count = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {  // attempt 1 = after a fixed-capacity second pass overflowed
        // per-launch scalars start from zero; the output cursor restarts where the previous join of this step ended
        CK(cudaMemsetAsync(m.scalars + kScWork0, 0, 8, st));
        CK(cudaMemsetAsync(m.scalars + kScNItems, 0, 8, st));
        ctx->h_scalars[kScCursor] = ctx->shard_count;
        CK(cudaMemcpyAsync(m.scalars + kScCursor, ctx->h_scalars + kScCursor, 8, cudaMemcpyHostToDevice, st));
        if (nB && nP) {
            u64 cap64 = (u64) nparts + nP / kProbeChunk + 2;
            if (cap64 > 0x7fffffffull) return fail(ctx, RHJ_ERR_ARG, "relation too large for the work-item table");
            u32 item_cap = (u32) cap64;
            if ((rc = ensure(ctx, ctx->items, (size_t) item_cap * sizeof(Item)))) return rc;
            mark(ctx, st, RHJ_PHASE_PLAN);
            const u64 capB = ctx->shard_cap[build_slot], capP = ctx->shard_cap[probe_slot];
            PlanPartsArgs pa{};
            pa.offB = sb.off2;
            pa.offP = spb.off2;
            pa.endB = capB ? sb.cur2 : nullptr;
            pa.endP = capP ? spb.cur2 : nullptr;
            pa.capB = capB;
            pa.capP = capP;
            pa.ndig = std::min<u32>(nparts, kMaxDigits);
            pa.items = (Item *) ctx->items.p;
            pa.item_cap = item_cap;
            pa.nitems = (u32 *) (m.scalars + kScNItems);
            pa.err = (u32 *) (m.scalars + kScErr);
            pa.overflow = (u32 *) (m.scalars + kScOverflow);
            k_plan_parts<<<nparts / pa.ndig, kMaxDigits, 0, st>>>(pa);
            CK(cudaGetLastError());
            ctx->info.kernel_launches++;
            ctx->cur.valid = true;
            ctx->cur.build = (const Tup *) sb.out->p;
            ctx->cur.probe = (const Tup *) spb.out->p;
            ctx->cur.offB = sb.off2;
            ctx->cur.offP = spb.off2;
            ctx->cur.endB = capB ? sb.cur2 : sb.off2 + 1;
            ctx->cur.endP = capP ? spb.cur2 : spb.off2 + 1;
            ctx->cur.nparts = nparts;
            ctx->cur.item_cap = item_cap;
            ctx->cur.build_is_S = build_slot != 0;
            ctx->info.optimistic_pass1 = (capB ? 4u : 0u) | (capP ? 8u : 0u);
            JoinArgs j = join_args(ctx, kScWork0);
            j.out = (Pair *) d_out;
            j.capacity = capacity;
            mark(ctx, st, RHJ_PHASE_JOIN);
            if ((rc = launch_join<kJoinFused>(ctx, st, j, item_cap))) return rc;
            mark(ctx, st, -1);
        }
        rc = read_scalars(ctx, st);
        if (rc != kRetryExact || attempt == 1) break;
        // A fixed-capacity region overflowed (duplicate-heavy data): redo the second pass of every slot that used one through
        // the exact histogram path -- what they received is still in place -- and join again.
        CK(cudaMemsetAsync(m.scalars + kScOverflow, 0, 8, st));
        ctx->opt2_skip = 16;
        for (int sl = 0; sl < 2; ++sl) {  // both relations: the flag does not say which one overflowed
            if (!ctx->shard_cap[sl]) continue;
            if ((rc = shardx_pass2_impl(ctx, sp, sl, (const rhj_tuple *) ctx->shard_recv[sl], ctx->shard_n[sl], stream, false))) return rc;
        }
    }
    if (rc == kRetryExact) return fail(ctx, RHJ_ERR_STATE, "sharded join: overflow flag set on the exact path");
    if (rc) return rc;
    *count = ctx->h_scalars[kScCursor];
    if (*count > capacity) return fail(ctx, RHJ_ERR_CAPACITY, "output buffer too small for the fused emitter");
    ctx->shard_count = *count;
    return RHJ_OK;
}

}  // extern "C"

#include "rhj_pipe.cuh"
