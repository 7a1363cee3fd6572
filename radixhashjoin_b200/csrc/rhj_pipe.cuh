// rhj_pipe.cuh -- host side of the pipelined multi-GPU exchange (rhj_pipe_* in include/rhj.h); part of the
// rhj_api.cu translation unit (uses its launch helpers).  Device side and design: rhj_pipe_kernels.cuh.
#pragma once
#include "rhj_pipe_kernels.cuh"

namespace {

// Capacity of one (chunk, destination, sub-digit) region: hashed digits of `rows` rows are multinomial, so
// mean + 1/16 + 16 sigma + 64 never overflows on distinct keys; multiple of 8 tuples (128-byte lines).
inline u64 pipe_cap1(u64 rows, u32 ndig) {
    const double mean = (double) rows / ndig;
    return ((u64) (mean + mean / 16.0 + 16.0 * std::sqrt(mean) + 64.0) + 8) & ~(u64) 7;
}

struct PipeLayout {
    u64 chunk_rows[2], cap1[2];
    u64 off_recv[2][2], off_end[2][2], off_flag, off_status, bytes;
    u32 wire;
};

// The same arithmetic on every rank: where everything lies inside a rank's symmetric block.
int pipe_layout(const rhj_shard_plan *sp, const rhj_pipe_cfg *cfg, PipeLayout &L) {
    if (!sp || !cfg || cfg->world != sp->world || cfg->rank >= cfg->world || cfg->chunks < 1 || cfg->chunks > (u32) kPipeMaxChunks)
        return RHJ_ERR_ARG;
    const u32 W = sp->world, ndig = W << sp->bits_pass1, C = cfg->chunks;
    if (ndig > (u32) kMaxDigits) return RHJ_ERR_ARG;
    L.wire = cfg->wire_bytes ? cfg->wire_bytes : 16;
    if (L.wire != 12 && L.wire != 16) return RHJ_ERR_ARG;
    const u64 nmax[2] = {cfg->nR_local_max, cfg->nS_local_max};
    u64 at = 0;
    auto take = [&at](u64 bytes) {
        const u64 o = at;
        at += (bytes + 255) & ~(u64) 255;
        return o;
    };
    for (int rel = 0; rel < 2; ++rel) {
        L.chunk_rows[rel] = (nmax[rel] + C - 1) / C;
        L.cap1[rel] = pipe_cap1(L.chunk_rows[rel], ndig);
        for (int q = 0; q < 2; ++q) {
            L.off_recv[rel][q] = take((u64) C * ndig * L.cap1[rel] * L.wire);
            L.off_end[rel][q] = take((u64) C * ndig * 8);
        }
    }
    L.off_flag = take((u64) 2 * 2 * kPipeMaxChunks * kMaxPeers * 8);
    L.off_status = take((u64) 2 * kMaxPeers * 8);
    L.bytes = at;
    return RHJ_OK;
}

inline u64 *pipe_flag(const rhj_ctx *ctx, u32 r, u32 q, int rel, int chunk) {
    return (u64 *) ((char *) ctx->pipe.sym[r] + ctx->pipe.off_flag) + (((u64) q * 2 + rel) * kPipeMaxChunks + chunk) * kMaxPeers;
}
inline u64 *pipe_status(const rhj_ctx *ctx, u32 r, u32 q) {
    return (u64 *) ((char *) ctx->pipe.sym[r] + ctx->pipe.off_status) + (u64) q * kMaxPeers;
}
inline Tup *pipe_recv(const rhj_ctx *ctx, u32 r, int rel, u32 q) {  // 16-byte tuples or packed 12-byte records (pipe.wire)
    return (Tup *) ((char *) ctx->pipe.sym[r] + ctx->pipe.off_recv[rel][q]);
}
inline u64 *pipe_end(const rhj_ctx *ctx, u32 r, int rel, u32 q) {
    return (u64 *) ((char *) ctx->pipe.sym[r] + ctx->pipe.off_end[rel][q]);
}
// segment tables of (rel, chunk): seg_off[nseg + 1] | seg_end[nseg + 1] | seg_tile0[nseg + 1] (u32, padded)
struct PipeSegs {
    u64 *seg_off, *seg_end;
    u32 *seg_tile0;
};
inline PipeSegs pipe_segs(const rhj_ctx *ctx, int rel, int chunk) {
    u64 *q = (u64 *) ctx->pipe.segs.p + ((u64) rel * kPipeMaxChunks + chunk) * 3 * (kMaxDigits + 2);
    return PipeSegs{q, q + (kMaxDigits + 2), (u32 *) (q + 2 * (kMaxDigits + 2))};
}

}  // namespace

extern "C" {

uint64_t rhj_pipe_sym_bytes(const rhj_shard_plan *sp, const rhj_pipe_cfg *cfg) {
    PipeLayout L;
    if (pipe_layout(sp, cfg, L)) return 0;
    return L.bytes;
}

// Wires a context into the exchange: remembers the peers' symmetric blocks, sizes the local staging, cursors,
// segment tables and the final-partition buffers.  The blocks must be zero-filled before the first step.
int rhj_pipe_open(rhj_ctx *ctx, const rhj_shard_plan *sp, const rhj_pipe_cfg *cfg) {
    if (!ctx) return RHJ_ERR_ARG;
    PipeLayout L;
    if (pipe_layout(sp, cfg, L)) return fail(ctx, RHJ_ERR_ARG, "rhj_pipe_open: bad plan / configuration");
    CK(cudaSetDevice(ctx->device));
    auto &P = ctx->pipe;
    P.open = false;
    P.plan = *sp;
    P.world = cfg->world;
    P.rank = cfg->rank;
    P.chunks = cfg->chunks;
    // the plain copy kernel saturates the link with 48 single-thread CTAs; the repacking one is bound by its ring depth
    // (4 x 8 KiB in flight per CTA) and wants more CTAs (measured on 8 GPUs: 96 CTAs 7.58 ms per join, 128 7.81, 148 7.86)
    P.ship_ctas = cfg->ship_ctas ? cfg->ship_ctas : (L.wire == 12 ? 96 : 48);
    if (const char *e = getenv("RHJ_PIPE_SHIP_CTAS")) P.ship_ctas = (u32) std::max(1, atoi(e));
    P.wire = L.wire;
    P.nmax[0] = cfg->nR_local_max;
    P.nmax[1] = cfg->nS_local_max;
    for (u32 r = 0; r < P.world; ++r) {
        if (!cfg->sym[r]) return fail(ctx, RHJ_ERR_ARG, "rhj_pipe_open: null symmetric block");
        P.sym[r] = cfg->sym[r];
    }
    for (int rel = 0; rel < 2; ++rel) {
        P.chunk_rows[rel] = L.chunk_rows[rel];
        P.cap1[rel] = L.cap1[rel];
        for (int q = 0; q < 2; ++q) {
            P.off_recv[rel][q] = L.off_recv[rel][q];
            P.off_end[rel][q] = L.off_end[rel][q];
        }
    }
    P.off_flag = L.off_flag;
    P.off_status = L.off_status;
    P.sym_bytes = L.bytes;
    const u32 ndig = P.world << sp->bits_pass1, nparts = 1u << sp->bits_total;
    int rc;
    for (int rel = 0; rel < 2; ++rel) {
        if ((rc = ensure(ctx, P.stage[rel], ((u64) P.chunks * ndig * P.cap1[rel] + 2 * kTile) * sizeof(Tup)))) return rc;  // + the dump tile (<= 8192 tuples)
        // a rank receives ~ 1 / world of the global relation; destination ranks are hashed, so +1/64 covers the imbalance
        P.cap2[rel] = fixed_cap2(P.nmax[rel] + P.nmax[rel] / 64 + 1, nparts);
    }
    if ((rc = ensure(ctx, P.cursors, (u64) 2 * kPipeMaxChunks * kMaxDigits * 8))) return rc;
    if ((rc = ensure(ctx, P.segs, (u64) 2 * kPipeMaxChunks * 3 * (kMaxDigits + 2) * 8))) return rc;
    if ((rc = ensure(ctx, P.done, 256))) return rc;
    CK(cudaMemset(P.done.p, 0, 256));
    // final partitions (slot 0 = R, slot 1 = S), the work-item table and the tile table of the largest chunk
    SlotArrays sl;
    for (int rel = 0; rel < 2; ++rel) {
        if ((rc = slot_arrays(ctx, nparts, rel, sl))) return rc;
        if ((rc = ensure(ctx, *sl.out, ((u64) nparts * P.cap2[rel] + kTile) * sizeof(Tup)))) return rc;
    }
    const u64 np_bound = (u64) P.chunks * ndig * std::max(P.cap1[0], P.cap1[1]);
    if ((rc = ensure(ctx, ctx->items, ((u64) nparts + np_bound / kProbeChunk + 2) * sizeof(Item)))) return rc;
    const u64 tiles_bound = (u64) ndig * std::max(P.cap1[0], P.cap1[1]) / kTile + ndig + 2;
    if ((rc = ensure(ctx, ctx->tiles, 3 * (tiles_bound + 1) * sizeof(TileDesc)))) return rc;
    P.stage_bytes = 8192;
    P.stages = 8;
    if (const char *e = getenv("RHJ_PIPE_STAGE_KB")) P.stage_bytes = (u32) std::max(1, std::min(64, atoi(e))) * 1024;
    if (const char *e = getenv("RHJ_PIPE_STAGES")) P.stages = (u32) std::max(2, std::min((int) kPipeMaxStages, atoi(e)));
    if ((u64) P.stages * P.stage_bytes > 200 * 1024) return fail(ctx, RHJ_ERR_ARG, "rhj_pipe_open: copy-kernel ring larger than 200 KiB");
    CK(cudaFuncSetAttribute(k_pipe_ship, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (P.stages * P.stage_bytes)));
    CK(cudaFuncSetAttribute(k_pipe_ship12, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (kPipe12Stages * kPipe12Block * 28)));
    P.epoch = 0;
    P.open = true;
    return RHJ_OK;
}

// Starts step `epoch` (1, 2, 3, ... -- the same number on every rank): zeroes the per-step counters, resets the
// pass-1 cursors of all chunks and the cursors of the final partitions.  Enqueues only.
int rhj_pipe_begin(rhj_ctx *ctx, uint64_t epoch, void *stream) {
    if (!ctx || !ctx->pipe.open || epoch == 0) return RHJ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    auto &P = ctx->pipe;
    const rhj_shard_plan *sp = &P.plan;
    int rc;
    if ((rc = rhj_shardx_begin(ctx, sp, stream))) return rc;
    P.epoch = epoch;
    const u32 ndig = P.world << sp->bits_pass1, nparts = 1u << sp->bits_total;
    PipeBeginArgs b{};
    for (int rel = 0; rel < 2; ++rel) {
        b.cursor[rel] = (u64 *) P.cursors.p + (u64) rel * kPipeMaxChunks * kMaxDigits;
        b.cap1[rel] = P.cap1[rel];
    }
    b.chunks = P.chunks;
    b.ndig = ndig;
    k_pipe_begin<<<dim3((P.chunks * ndig + 255) / 256, 2), 256, 0, st>>>(b);
    CK(cudaGetLastError());
    ctx->info.kernel_launches++;
    for (int rel = 0; rel < 2; ++rel) {
        SlotArrays sl;
        if ((rc = slot_arrays(ctx, nparts, rel, sl))) return rc;
        PlanFixedArgs pf{};
        pf.end[0] = sl.cur2;
        pf.beg[0] = sl.off2;
        pf.cap[0] = P.cap2[rel];
        pf.nseg = 1u << sp->bits_pass1;
        pf.ndig = 1u << sp->bits_pass2;
        k_fixed_cursors2<<<dim3((nparts + 255) / 256, 1), 256, 0, st>>>(pf);
        CK(cudaGetLastError());
        ctx->info.kernel_launches++;
        ctx->shard_cap[rel] = P.cap2[rel];
    }
    return RHJ_OK;
}

// Pass 1 of chunk `chunk` of relation `rel` (0 = R, 1 = S): d_rows[n] (n <= the chunk size fixed at open) are
// scattered on (destination rank | sub-digit) into the chunk's fixed-capacity regions.  Enqueues only.
int rhj_pipe_pass1_device(rhj_ctx *ctx, int rel, int chunk, const rhj_tuple *d_rows, uint64_t n, void *stream) {
    if (!ctx || !ctx->pipe.open || rel < 0 || rel > 1 || chunk < 0 || chunk >= (int) ctx->pipe.chunks || (n && !d_rows))
        return RHJ_ERR_ARG;
    auto &P = ctx->pipe;
    if (n > P.chunk_rows[rel]) return fail(ctx, RHJ_ERR_ARG, "rhj_pipe_pass1_device: chunk larger than configured at rhj_pipe_open");
    if (n == 0) return RHJ_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    const rhj_shard_plan *sp = &P.plan;
    Meta m;
    int rc;
    if ((rc = layout_meta(ctx, 1u << sp->bits_total, m))) return rc;
    const u32 ndig = P.world << sp->bits_pass1, nd1 = 1u << sp->bits_pass1, q = (u32) (P.epoch & 1);
    PartArgs a = shard_args(sp);
    a.shard_local = 0;
    a.overflow = (u32 *) (m.scalars + kScOverflow);
    u64 *cursor = (u64 *) P.cursors.p + ((u64) rel * kPipeMaxChunks + chunk) * kMaxDigits;
    a.rel[0] = PartRel{(const Tup *) d_rows, nullptr, n, nullptr, cursor, nullptr, nullptr, 1, tiles_of(n)};
    a.rel[0].limit_cap = P.cap1[rel];
    Tup *stage_c = (Tup *) P.stage[rel].p + (u64) chunk * ndig * P.cap1[rel];
    a.rel[0].dump_ptr = (Tup *) P.stage[rel].p + (u64) P.chunks * ndig * P.cap1[rel];
    // the rank's own digits go straight to where the other ranks' copies of this chunk will land around them:
    // region (chunk, source = rank, p1) of the own receive buffer
    Tup *self = pipe_recv(ctx, P.rank, rel, q) + (((u64) chunk * P.world + P.rank) * nd1) * P.cap1[rel] -
                ((u64) P.rank << sp->bits_pass1) * P.cap1[rel];
    // (12-byte wire format: the own digits are staged as well and repacked by the copy kernel like everybody else's)
    for (u32 d = 0; d < P.world; ++d) a.peer_out[0][d] = (d == P.rank && P.wire == 16) ? self : stage_c;
    return launch_scatter(ctx, st, a, kDigitShard, false, true);
}

// Ships chunk `chunk` of relation `rel`: the filled part of every remote region goes to the same region of the
// destination's receive buffer, region ends and one flag per destination follow.  Call it on a second stream,
// after an event recorded behind the chunk's pass 1.  Enqueues only.
int rhj_pipe_ship_device(rhj_ctx *ctx, int rel, int chunk, void *stream) {
    if (!ctx || !ctx->pipe.open || rel < 0 || rel > 1 || chunk < 0 || chunk >= (int) ctx->pipe.chunks) return RHJ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    auto &P = ctx->pipe;
    const rhj_shard_plan *sp = &P.plan;
    Meta m;
    int rc;
    if ((rc = layout_meta(ctx, 1u << sp->bits_total, m))) return rc;
    const u32 ndig = P.world << sp->bits_pass1, nd1 = 1u << sp->bits_pass1, q = (u32) (P.epoch & 1);
    PipeShipArgs a{};
    a.stage = (const Tup *) P.stage[rel].p + (u64) chunk * ndig * P.cap1[rel];
    a.cursor = (const u64 *) P.cursors.p + ((u64) rel * kPipeMaxChunks + chunk) * kMaxDigits;
    a.cap1 = P.cap1[rel];
    a.world = P.world;
    a.rank = P.rank;
    a.ndig = ndig;
    a.sub_bits = sp->bits_pass1;
    for (u32 r = 0; r < P.world; ++r) {
        a.peer_recv[r] = pipe_recv(ctx, r, rel, q);
        a.peer_end[r] = pipe_end(ctx, r, rel, q);
        a.peer_flag[r] = pipe_flag(ctx, r, q, rel, chunk);
    }
    a.region0 = ((u64) chunk * P.world + P.rank) * nd1;
    a.epoch = P.epoch;
    a.done = (u32 *) P.done.p;
    a.overflow = (u32 *) (m.scalars + kScOverflow);
    const u32 nremote = (P.world - 1) << sp->bits_pass1;
    const u32 grid = std::max<u32>(1, std::min<u32>(P.ship_ctas, std::max<u32>(nremote, 1)));
    a.stages = P.stages;
    a.stage_bytes = P.stage_bytes;
    if (P.wire == 12)
        k_pipe_ship12<<<std::max<u32>(1, std::min<u32>(P.ship_ctas, ndig)), kPipeShip12Threads, (size_t) kPipe12Stages * kPipe12Block * 28, st>>>(a);
    else
        k_pipe_ship<<<grid, kPipeShipThreads, (size_t) P.stages * P.stage_bytes, st>>>(a);
    CK(cudaGetLastError());
    ctx->info.kernel_launches++;
    return RHJ_OK;
}

// Pass 2 of chunk `chunk` of relation `rel` on the destination: waits (on the device) until the chunk has arrived
// from every rank, then appends it to the relation's fixed-capacity final partitions.  Enqueues only.
int rhj_pipe_pass2_device(rhj_ctx *ctx, int rel, int chunk, void *stream) {
    if (!ctx || !ctx->pipe.open || rel < 0 || rel > 1 || chunk < 0 || chunk >= (int) ctx->pipe.chunks) return RHJ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    auto &P = ctx->pipe;
    const rhj_shard_plan *sp = &P.plan;
    Meta m;
    SlotArrays sl;
    int rc;
    const u32 nparts = 1u << sp->bits_total;
    if ((rc = layout_meta(ctx, nparts, m))) return rc;
    if ((rc = slot_arrays(ctx, nparts, rel, sl))) return rc;
    const u32 nd1 = 1u << sp->bits_pass1, nseg = P.world << sp->bits_pass1, q = (u32) (P.epoch & 1);
    PipeSegs sg = pipe_segs(ctx, rel, chunk);
    PipeArriveArgs ar{};
    ar.flag = pipe_flag(ctx, P.rank, q, rel, chunk);
    ar.region0 = (u64) chunk * nseg;
    ar.region_end = pipe_end(ctx, P.rank, rel, q) + ar.region0;
    ar.cap1 = P.cap1[rel];
    ar.world = P.world;
    ar.nseg = nseg;
    ar.epoch = P.epoch;
    ar.seg_off = sg.seg_off;
    ar.seg_end = sg.seg_end;
    ar.seg_tile0 = sg.seg_tile0;
    ar.status = m.scalars + kScPipeStatus;

    PartArgs b{};
    b.shift = std::min(31, 32 - (int) sp->bits_total);
    b.mask = (1u << sp->bits_pass2) - 1;
    b.ndig = 1u << sp->bits_pass2;
    b.overflow = (u32 *) (m.scalars + kScOverflow);
    const u32 tiles_bound = (u32) ((u64) nseg * P.cap1[rel] / kTile) + nseg + 1;
    b.rel[0] = PartRel{pipe_recv(ctx, P.rank, rel, q), (Tup *) sl.out->p, (u64) nseg * P.cap1[rel], sl.hist2, sl.cur2, sg.seg_off,
                       sg.seg_tile0, nseg, tiles_bound, nd1 - 1};
    if (nd1 == 1) b.rel[0].group_mask = 0x80000000u;  // every segment is partition 0
    if (P.wire == 12) {
        b.rel[0].in = nullptr;
        b.rel[0].in_packed = (const unsigned char *) pipe_recv(ctx, P.rank, rel, q);
    }
    b.rel[0].seg_end = sg.seg_end;
    b.rel[0].limit_cap = P.cap2[rel];
    b.rel[0].dump = (u64) nparts * P.cap2[rel];
    // the tile table of what arrived is written by the arrival kernel itself (one launch instead of two)
    const size_t need = ((size_t) tiles_bound + 1) * sizeof(TileDesc);
    if (2 * need > ctx->tiles.cap) return fail(ctx, RHJ_ERR_STATE, "rhj_pipe_pass2_device: tile table smaller than sized at rhj_pipe_open");
    ar.tiles = (TileDesc *) ((char *) ctx->tiles.p + (size_t) rel * need);
    ar.ntiles = tiles_bound;
    b.rel[0].tiles = ar.tiles;
    k_pipe_arrive<<<1, 1024, 0, st>>>(ar);
    CK(cudaGetLastError());
    ctx->info.kernel_launches++;
    return launch_scatter(ctx, st, b, kDigitHash, true, true);
}

// After the last pass 2 of the step: publishes this rank's overflow verdict to every rank.  Enqueues only.
int rhj_pipe_post_device(rhj_ctx *ctx, void *stream) {
    if (!ctx || !ctx->pipe.open) return RHJ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    auto &P = ctx->pipe;
    Meta m;
    int rc;
    if ((rc = layout_meta(ctx, 1u << P.plan.bits_total, m))) return rc;
    PipePostArgs a{};
    const u32 q = (u32) (P.epoch & 1);
    for (u32 r = 0; r < P.world; ++r) a.peer_status[r] = pipe_status(ctx, r, q);
    a.overflow = (const u32 *) (m.scalars + kScOverflow);
    a.status = m.scalars + kScPipeStatus;
    a.world = P.world;
    a.rank = P.rank;
    a.epoch = P.epoch;
    k_pipe_post<<<1, 32, 0, st>>>(a);
    CK(cudaGetLastError());
    ctx->info.kernel_launches++;
    return RHJ_OK;
}

// Work-item plan + build/probe + fused emit over the final partitions, then the verdicts of all ranks.
// *status = 0: *count pairs in d_out are this rank's share of the result.  Otherwise (RHJ_PIPE_* bits, the same
// value on every rank unless RHJ_PIPE_TIMEOUT is set) the step must be redone through the exact exchange
// (rhj_shardx_*).  Synchronises the stream.
int rhj_pipe_join_device(rhj_ctx *ctx, rhj_pair *d_out, uint64_t capacity, uint64_t *count, uint32_t *status, void *stream) {
    if (!ctx || !ctx->pipe.open || !count || !status) return RHJ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    auto &P = ctx->pipe;
    const rhj_shard_plan *sp = &P.plan;
    const u32 nparts = 1u << sp->bits_total, ndig = P.world << sp->bits_pass1, q = (u32) (P.epoch & 1);
    const int bs = sp->build_is_S ? 1 : 0, ps = bs ^ 1;
    Meta m;
    SlotArrays sb, spb;
    int rc;
    *count = 0;
    *status = 0;
    if ((rc = layout_meta(ctx, nparts, m))) return rc;
    if ((rc = slot_arrays(ctx, nparts, bs, sb))) return rc;
    if ((rc = slot_arrays(ctx, nparts, ps, spb))) return rc;
    const u64 np_bound = (u64) P.chunks * ndig * P.cap1[ps];
    const u64 cap64 = (u64) nparts + np_bound / kProbeChunk + 2;
    if (cap64 > 0x7fffffffull) return fail(ctx, RHJ_ERR_ARG, "relation too large for the work-item table");
    const u32 item_cap = (u32) cap64;
    if ((rc = ensure(ctx, ctx->items, (size_t) item_cap * sizeof(Item)))) return rc;
    PlanPartsArgs pa{};
    pa.offB = sb.off2;
    pa.offP = spb.off2;
    pa.endB = sb.cur2;
    pa.endP = spb.cur2;
    pa.capB = P.cap2[bs];
    pa.capP = P.cap2[ps];
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
    ctx->cur.endB = sb.cur2;
    ctx->cur.endP = spb.cur2;
    ctx->cur.nparts = nparts;
    ctx->cur.item_cap = item_cap;
    ctx->cur.build_is_S = bs;
    ctx->info.optimistic_pass1 = 3u | 4u | 8u;
    JoinArgs j = join_args(ctx, kScWork0);
    j.out = (Pair *) d_out;
    j.capacity = capacity;
    j.holes = m.scalars + kScHoles;
    // positional emit (rhj_join.cuh): one slot per probe tuple, no ranking and no reservation latency in the probe loop
    bool pos = ctx->positional && ctx->pos_skip == 0;
    if (!pos && ctx->pos_skip > 0) ctx->pos_skip--;
    if (pos) rc = launch_join_positional(ctx, st, j, item_cap);
    else rc = launch_join<kJoinFused>(ctx, st, j, item_cap);
    if (rc) return rc;
    PipeCollectArgs ca{};
    ca.status_in = pipe_status(ctx, P.rank, q);
    ca.status = m.scalars + kScPipeStatus;
    ca.world = P.world;
    ca.epoch = P.epoch;
    k_pipe_collect<<<1, 32, 0, st>>>(ca);
    CK(cudaGetLastError());
    ctx->info.kernel_launches++;
    u64 *sc = scalars_of(ctx, nparts);
    CK(cudaMemcpyAsync(ctx->h_scalars, sc, kScCount * sizeof(u64), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    u64 bits = ctx->h_scalars[kScPipeStatus] & 0xff;
    if (ctx->h_scalars[kScOverflow] & 1) bits |= kPipeOvf;
    if (ctx->h_scalars[kScOverflow] & 2) bits |= kPipeWide;
    *status = (uint32_t) bits;
    if (bits) return RHJ_OK;  // the caller redoes the step (or reports the timeout)
    if (ctx->h_scalars[kScErr]) return fail(ctx, RHJ_ERR_STATE, "device-side planning error (work-item table overflow)");
    ctx->info.n_items = (u32) ctx->h_scalars[kScNItems];
    if (pos) {
        const u64 cursor = ctx->h_scalars[kScCursor], holes = ctx->h_scalars[kScHoles];
        bool ok = cursor <= capacity;
        if (ok && holes) {
            if ((rc = close_holes(ctx, st, (Pair *) d_out, cursor, holes, &ok))) return rc;
            if (ok) ctx->h_scalars[kScCursor] = cursor - holes;
            if (holes * 64 > cursor) ctx->pos_skip = 16;
        }
        if (!ok) {  // the slots did not fit the buffer (or a genuine pair looks like a hole): the ranked emitter over the same items
            CK(cudaMemsetAsync(sc + kScWork0, 0, 8, st));
            CK(cudaMemsetAsync(sc + kScCursor, 0, 8, st));
            if ((rc = launch_join<kJoinFused>(ctx, st, j, item_cap))) return rc;
            CK(cudaMemcpyAsync(ctx->h_scalars, sc, kScCount * sizeof(u64), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
        }
    }
    *count = ctx->h_scalars[kScCursor];
    ctx->cur.valid = false;
    if (*count > capacity) return fail(ctx, RHJ_ERR_CAPACITY, "output buffer too small for the fused emitter");
    return RHJ_OK;
}

}  // extern "C"
