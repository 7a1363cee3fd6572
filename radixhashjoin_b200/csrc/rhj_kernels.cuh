// rhj_kernels.cuh -- the sm_100a kernels of the radix hash join.
//
// north_star step            reference loop it replaces                       kernel here
//  (1) histogram             HistogramJob::run, JobScheduler.cpp:149-155      k_hist
//  (2) prefix sum            PartitionJob::run 163-169, Result.cpp:100-107    k_scan_digits, k_scan_parts_plan
//  (3) partition scatter     PartitionJob::run 170-174 + structs.cpp:183-194  k_scatter
//  (4) per-bucket build/probe Result::join_buckets, Result.cpp:43-76           k_join            (rhj_join.cuh)
//  (5) emitter               add_result/addAll, Result.cpp:21-35,78-84,111-121 k_join<COUNT|WRITE|FUSED>, k_scan_items;
//                                                                             k_join_pos + k_holes_* (positional emit)
//  filters / gathers         Query.cpp:94-146, structs.cpp:217-226, Query.cpp:66-74   k_filter_*, k_gather_*
//
// Data layout in HBM: relations stay 16-byte AoS tuples {rowid, value} end to end (the caller's
// layout, structs.h:33-36); partitions are contiguous ranges described by u64 offset arrays;
// results are 16-byte {rowidR,rowidS} pairs (Result.h:9-12) in one flat array.
#pragma once
#include "rhj_device.cuh"
#include "rhj_join.cuh"

namespace rhj {

// ---- tuning constants -------------------------------------------------------------------------
#ifndef RHJ_PART_THREADS
#define RHJ_PART_THREADS 512
#endif
#ifndef RHJ_PART_ITEMS
#define RHJ_PART_ITEMS 8
#endif
#ifndef RHJ_PART_MINBLOCKS
#define RHJ_PART_MINBLOCKS 2
#endif
#ifndef RHJ_SCATTER_MAXNREG
#define RHJ_SCATTER_MAXNREG 56
#endif
constexpr int kPartThreads = RHJ_PART_THREADS;  // threads per partition CTA
constexpr int kPartItems = RHJ_PART_ITEMS;      // tuples per thread
constexpr int kTile = kPartThreads * kPartItems;  // 4096 tuples = 64 KiB staged per CTA
constexpr int kMaxBitsPerPass = 10;               // a pass can fan out to 1024 digits (sharded pass 1: rank | sub-digit)
constexpr int kMaxDigits = 1 << kMaxBitsPerPass;
constexpr int kPlanBitsPerPass = 9;               // single-GPU plans use <= 512 digits per pass (longer runs)
constexpr int kMaxPeers = 16;                     // ranks of the fused partition + shuffle pass

// kDigitShard = (destination rank << sub_bits) | pass-1 digit: one pass both shuffles and partitions
enum DigitKind { kDigitRaw = 0, kDigitHash = 1, kDigitRank = 2, kDigitShard = 3 };

// ---- (1)+(3): descriptors of a partition pass over up to two relations at once ---------------
struct PartRel {
    const Tup *in;
    Tup *out;
    u64 n;
    u64 *hist;             // [nseg * ndig] digit counts (written by k_hist)
    u64 *cursor;           // [nseg * ndig] running output cursors (consumed by k_scatter)
    const u64 *seg_off;    // SEG: [nseg+1] segment boundaries inside `in` (pass-1 partitions)
    const u32 *seg_tile0;  // SEG: [nseg+1] first tile of each segment
    u32 nseg;
    u32 ntiles;            // tiles of this relation (SEG: host-side upper bound)
    u32 group_mask;        // SEG: segments that share counters: group = seg & group_mask (0 = every segment
                           // is its own group).  Sharded pass 2: segment = (source rank, pass-1 partition).
    const void *tiles;     // SEG: optional TileDesc[ntiles] built by k_tile_table (else binary search)
    const u64 *seg_end;    // SEG: optional [nseg] segment ends (null: seg_off[seg + 1]); fixed-capacity pass-1 layout
    u64 limit_cap;         // k_scatter<LIMIT>: digit d may only fill [d * limit_cap, (d + 1) * limit_cap) of `out`
    const unsigned char *in_packed;  // kIoPacked12In: record i at byte 12 * i
    // k_scatter<LIMIT>: a run that does not fit its partition's region is written to the kTile tuples at out[dump..]
    // instead (nothing out of bounds, no per-tuple check); the overflow flag makes the host discard the attempt.
    u64 dump;
    Tup *dump_ptr;         // same, as an absolute address, for scatters whose output base depends on the digit (peer_out)
};
// How a partitioning kernel reads tuples: 16-byte AoS (everything single-GPU, pass 1 of the sharded joins), or
// kIoPacked12In: packed 12-byte records {u64 value, u32 row id} -- the wire format of the pipelined exchange
// (rhj_pipe_kernels.cuh); a tile is fetched with one TMA bulk copy and unpacked from shared memory.
enum TupleIo { kIoAos = 0, kIoPacked12In = 1 };
// One 16-byte descriptor per pass-2 tile (k_tile_table): a CTA finds its tuple range with ONE load
// instead of a 9-step binary search over seg_tile0 -- that dependent-load chain sat in front of
// every tile's first tuple load.
struct __align__(16) TileDesc {
    u64 beg;
    u32 len;  // 0 = no such tile
    u32 seg;
};
__device__ __forceinline__ u32 seg_group(const PartRel &r, u32 seg) { return r.group_mask ? (seg & r.group_mask) : seg; }
struct PartArgs {
    PartRel rel[2];
    int shift;
    u32 mask;
    u32 ndig;
    int rank_bits;                  // kDigitShard: log2(world)
    int sub_bits;                   // kDigitShard: bits of the pass-1 digit below the rank
    Tup *peer_out[2][kMaxPeers];    // kDigitShard: per relation, the destination ranks' receive buffers (peer memory)
    int shard_local;                // kDigitShard: 1 = write to rel.out (local staging, shipped by DMA afterwards)
    u32 *overflow;                  // k_scatter<LIMIT>: set to 1 when a digit outgrows its fixed-capacity region
};

template <int KIND>
__device__ __forceinline__ u32 digit(u64 v, const PartArgs &a) {
    if (KIND == kDigitRaw) return (u32) (v >> a.shift) & a.mask;
    if (KIND == kDigitHash) return (hash32(v) >> a.shift) & a.mask;
    if (KIND == kDigitRank) return (hash_hi32(v) >> a.shift) & a.mask;
    const u64 h = hash64(v);
    const u32 rank = a.rank_bits ? (u32) (h >> (64 - a.rank_bits)) : 0u;
    const u32 sub = a.sub_bits ? ((u32) h >> (32 - a.sub_bits)) : 0u;
    return (rank << a.sub_bits) | sub;
}

// Maps a relation-local tile index to its tuple range.  SEG tiles never straddle a segment.
template <bool SEG>
__device__ __forceinline__ bool tile_range(const PartRel &r, u32 lt, u32 &seg, u64 &beg, u64 &end) {
    if (!SEG) {
        seg = 0;
        beg = (u64) lt * kTile;
        end = min(beg + (u64) kTile, r.n);
        return beg < r.n;
    }
    if (r.tiles) {
        const uint4 q = __ldg(reinterpret_cast<const uint4 *>(r.tiles) + lt);  // {beg.lo, beg.hi, len, seg}
        if (q.z == 0) return false;
        beg = ((u64) q.y << 32) | q.x;
        end = beg + q.z;
        seg = q.w;
        return true;
    }
    if (lt >= r.seg_tile0[r.nseg]) return false;
    u32 lo = 0, hi = r.nseg;  // last seg with seg_tile0[seg] <= lt
    while (hi - lo > 1) {
        u32 mid = (lo + hi) >> 1;
        if (r.seg_tile0[mid] <= lt) lo = mid; else hi = mid;
    }
    seg = lo;
    beg = r.seg_off[lo] + (u64) (lt - r.seg_tile0[lo]) * kTile;
    end = min(beg + (u64) kTile, r.seg_end ? r.seg_end[lo] : r.seg_off[lo + 1]);
    return true;
}

// Expands (seg_off, seg_tile0) into one TileDesc per tile; entries past the last tile get len 0.
// One CTA per segment (+1 that clears the tail up to the host-side bound `ntiles`).
__global__ void __launch_bounds__(256) k_tile_table(const u64 *seg_off, const u64 *seg_end, const u32 *seg_tile0, u32 nseg,
                                                    u32 ntiles, TileDesc *tiles) {
    const u32 seg = blockIdx.x;
    if (seg == nseg) {
        for (u32 t = seg_tile0[nseg] + threadIdx.x; t < ntiles; t += blockDim.x) tiles[t] = TileDesc{0, 0, 0};
        return;
    }
    const u64 b = seg_off[seg], e = seg_end ? seg_end[seg] : seg_off[seg + 1];
    const u32 t0 = seg_tile0[seg], t1 = seg_tile0[seg + 1];
    for (u32 t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
        u64 beg = b + (u64) (t - t0) * kTile;
        tiles[t] = TileDesc{beg, (u32) min((u64) kTile, e - beg), seg};
    }
}

// (1) Histogram.  Each CTA owns a contiguous range of tiles, counts digits in shared memory
// (one ATOMS per tuple; with AGG the lanes of a warp that hit the same counter are combined
// first by match.any so a skewed digit costs one atomic per warp instead of 32 serialised ones)
// and flushes to the global u64 counters only when its (relation, segment) changes.
// Algorithmic bytes: 16 per tuple read.
template <int KIND, bool SEG, bool AGG>
__global__ void __launch_bounds__(kPartThreads) k_hist(PartArgs a) {
    __shared__ u32 s_h[kMaxDigits];
    const u32 tid = threadIdx.x;
    const u32 total = a.rel[0].ntiles + a.rel[1].ntiles;
    const u32 per = (total + gridDim.x - 1) / gridDim.x;
    const u32 t0 = min(total, blockIdx.x * per), t1 = min(total, t0 + per);
    for (u32 d = tid; d < a.ndig; d += kPartThreads) s_h[d] = 0;
    __syncthreads();
    int cur_rel = -1;
    u32 cur_seg = 0;
    for (u32 t = t0; t < t1; ++t) {
        const int ri = t >= a.rel[0].ntiles;
        const PartRel &r = a.rel[ri];
        u32 seg;
        u64 beg, end;
        if (!tile_range<SEG>(r, t - (ri ? a.rel[0].ntiles : 0), seg, beg, end)) continue;
        if (cur_rel >= 0 && (ri != cur_rel || seg != cur_seg)) {
            __syncthreads();
            u64 *h = a.rel[cur_rel].hist + (u64) seg_group(a.rel[cur_rel], cur_seg) * a.ndig;
            for (u32 d = tid; d < a.ndig; d += kPartThreads) {
                u32 c = s_h[d];
                if (c) { atomicAdd(h + d, (u64) c); s_h[d] = 0; }
            }
            __syncthreads();
        }
        cur_rel = ri;
        cur_seg = seg;
        Tup v[kPartItems];
        bool ok[kPartItems];
#pragma unroll
        for (int j = 0; j < kPartItems; ++j) {
            u64 idx = beg + (u64) j * kPartThreads + tid;
            ok[j] = idx < end;
            if (ok[j]) v[j] = ld_stream(r.in + idx);
        }
#pragma unroll
        for (int j = 0; j < kPartItems; ++j) {
            u32 d = ok[j] ? digit<KIND>(v[j].val, a) : kEmpty;
            if (AGG) {
                u32 peers = __match_any_sync(0xffffffffu, d);
                if (ok[j] && lane_id() == (u32) (__ffs(peers) - 1)) atomicAdd(&s_h[d], (u32) __popc(peers));
            } else {
                if (ok[j]) atomicAdd(&s_h[d], 1u);
            }
        }
    }
    __syncthreads();
    if (cur_rel >= 0) {
        u64 *h = a.rel[cur_rel].hist + (u64) seg_group(a.rel[cur_rel], cur_seg) * a.ndig;
        for (u32 d = tid; d < a.ndig; d += kPartThreads) {
            u32 c = s_h[d];
            if (c) atomicAdd(h + d, (u64) c);
        }
    }
}

// (2a) Exclusive prefix sum over the <=512 digit counters of one pass, one CTA per relation.
// Writes offsets[ndig+1], the scatter cursors (= offsets) and, when a second pass follows, the
// first-tile table of the pass-1 partitions.
struct ScanDigitsArgs {
    const u64 *hist[2];
    u64 *off[2];
    u64 *cursor[2];
    u32 *tile0[2];  // may be null
    u32 ndig;
};
__global__ void __launch_bounds__(kMaxDigits) k_scan_digits(ScanDigitsArgs a) {
    __shared__ u64 s_w[32];
    __shared__ u32 s_wt[32];
    const int ri = blockIdx.x;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    u64 c = tid < a.ndig ? a.hist[ri][tid] : 0;
    u32 tl = (u32) ((c + kTile - 1) / kTile);
    u64 inc = warp_incl_scan64(c);
    u32 tinc = warp_incl_scan(tl);
    if (lane == 31) { s_w[warp] = inc; s_wt[warp] = tinc; }
    __syncthreads();
    if (warp == 0) {
        u64 w = lane < (kMaxDigits / 32) ? s_w[lane] : 0;
        u32 wt = lane < (kMaxDigits / 32) ? s_wt[lane] : 0;
        u64 wi = warp_incl_scan64(w);
        u32 wti = warp_incl_scan(wt);
        s_w[lane] = wi - w;
        s_wt[lane] = wti - wt;
    }
    __syncthreads();
    u64 ex = inc - c + s_w[warp];
    u32 tex = tinc - tl + s_wt[warp];
    if (tid < a.ndig) {
        a.off[ri][tid] = ex;
        a.cursor[ri][tid] = ex;
        if (a.tile0[ri]) a.tile0[ri][tid] = tex;
        if (tid == a.ndig - 1) {
            a.off[ri][a.ndig] = ex + c;
            if (a.tile0[ri]) a.tile0[ri][a.ndig] = tex + tl;
        }
    }
}

// (3) Partition scatter with shared-memory staging (software write-combining).
// One CTA = one tile of 4096 tuples:
//   load 8 tuples/thread (coalesced 16-B loads) -> rank each inside its digit with one
//   shared-memory atomic -> one global atomic per non-empty digit reserves that digit's run in
//   the output -> tuples are placed in shared memory sorted by digit -> the sorted tile is
//   written out so that every digit's run is one contiguous burst: either by all threads
//   (consecutive threads -> consecutive 16-B slots, full 32-B sectors / 128-B lines inside a run)
//   or, with BULK, by one TMA bulk store (cp.async.bulk shared->global) per run.
// Algorithmic bytes: 16 read + 16 written per tuple.
enum ScatterWrite { kWriteStaged = 0, kWriteBulk = 1 };
// THREADS: 512 (4096-tuple tiles, two CTAs per SM) everywhere except the 1024-digit pass 1 of the pipelined exchange, which
// runs 1024 threads on 8192-tuple tiles (one CTA per SM): at 4 tuples per digit and tile the runs are 64 bytes and every
// fourth tuple costs a global reservation; the larger tile doubles the runs and halves the reservations.
template <int KIND, bool SEG, int WMODE, int MAXD, bool LIMIT = false, int IO = kIoAos, int THREADS = kPartThreads>
#if RHJ_SCATTER_MAXNREG
// 56 registers (0-48 bytes of spills) instead of the 62-64 two 512-thread CTAs per SM would allow: leaves room in the
// register file for a CTA of the multi-GPU copy kernel next to two scatter CTAs (rhj_pipe_kernels.cuh)
__global__ void __maxnreg__(RHJ_SCATTER_MAXNREG) k_scatter(PartArgs a) {
#else
__global__ void __launch_bounds__(THREADS, THREADS == kPartThreads ? RHJ_PART_MINBLOCKS : 1) k_scatter(PartArgs a) {
#endif
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    Tup *s_tup = reinterpret_cast<Tup *>(dyn_smem);
    __shared__ u32 s_cnt[MAXD];
    __shared__ u32 s_off[MAXD];
    __shared__ u64 s_delta[MAXD];  // global index of sorted slot i of digit d = s_delta[d] + i
    __shared__ u32 s_w[32];

    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ri = blockIdx.x >= a.rel[0].ntiles;
    const PartRel &r = a.rel[ri];
    u32 seg;
    u64 beg, end;
    if (SEG || THREADS == kPartThreads) {
        if (!tile_range<SEG>(r, blockIdx.x - (ri ? a.rel[0].ntiles : 0), seg, beg, end)) return;
    } else {  // unsegmented pass on THREADS * kPartItems-tuple tiles (PartRel::ntiles counts those)
        seg = 0;
        beg = (u64) (blockIdx.x - (ri ? a.rel[0].ntiles : 0)) * (THREADS * kPartItems);
        end = min(beg + (u64) (THREADS * kPartItems), r.n);
        if (beg >= r.n) return;
    }
    const u32 ntile = (u32) (end - beg);

    for (u32 d = tid; d < a.ndig; d += THREADS) s_cnt[d] = 0;
    Tup v[kPartItems];
    u32 dr[kPartItems];  // digit << 16 | rank   (digit < 1024, rank < 8192)
    if (IO == kIoPacked12In) {
        // one TMA bulk copy brings the tile's 12-byte records into the staging area (which the sorted tile reuses later);
        // every thread then unpacks its 8 records with conflict-free 4-byte shared-memory loads (stride 3 words)
        __shared__ __align__(8) u64 s_bar;
        if (tid == 0) {
            mbar_init(&s_bar, 1);
            const u32 bytes = (ntile * 12 + 15) & ~15u;  // the region capacity is a multiple of 16 bytes: the slack is readable
            mbar_expect_tx(&s_bar, bytes);
            bulk_g2s(dyn_smem, r.in_packed + beg * 12, bytes, &s_bar);
        }
        __syncthreads();
        mbar_wait(&s_bar, 0);
        const u32 *w = reinterpret_cast<const u32 *>(dyn_smem);
#pragma unroll
        for (int j = 0; j < kPartItems; ++j) {
            u32 i = j * THREADS + tid;
            if (i < ntile) {
                v[j].val = (u64) w[3 * i] | ((u64) w[3 * i + 1] << 32);
                v[j].key = w[3 * i + 2];
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < kPartItems; ++j) {
            u32 i = j * THREADS + tid;
            if (i < ntile) v[j] = ld_stream(r.in + beg + i);
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kPartItems; ++j) {
        u32 i = j * THREADS + tid;
        if (i < ntile) {
            u32 d = digit<KIND>(v[j].val, a);
            u32 rk = atomicAdd(&s_cnt[d], 1u);
            dr[j] = (d << 16) | rk;
        }
    }
    __syncthreads();
    // reserve the runs (global atomics issued first so their latency overlaps the block scan);
    // thread t owns the kDigitsPerThread consecutive digits starting at t * kDigitsPerThread
    constexpr int kDigitsPerThread = (MAXD + THREADS - 1) / THREADS;
    u32 c[kDigitsPerThread];
    u64 g[kDigitsPerThread];
    u32 csum = 0;
#pragma unroll
    for (int k = 0; k < kDigitsPerThread; ++k) {
        u32 d = tid * kDigitsPerThread + k;
        c[k] = d < a.ndig ? s_cnt[d] : 0;
        g[k] = 0;
        if (c[k]) g[k] = atomicAdd(r.cursor + (u64) seg_group(r, seg) * a.ndig + d, (u64) c[k]);
        csum += c[k];
    }
    u32 inc = warp_incl_scan(csum);
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        u32 w = lane < (THREADS / 32) ? s_w[lane] : 0;
        u32 wi = warp_incl_scan(w);
        s_w[lane] = wi - w;
    }
    __syncthreads();
    {
        u32 ex = inc - csum + s_w[warp];
#pragma unroll
        for (int k = 0; k < kDigitsPerThread; ++k) {
            u32 d = tid * kDigitsPerThread + k;
            if (d < a.ndig) {
                s_off[d] = ex;
                u64 delta = g[k] - ex;
                // per-destination output bases: folded into the digit's delta here, once per digit, as an offset from
                // peer_out[ri][0] -- indexing the parameter bank by the digit in the write-out loop would serialise a
                // warp over its distinct destinations
                if (KIND == kDigitShard && !a.shard_local) delta += (u64) (a.peer_out[ri][d >> a.sub_bits] - a.peer_out[ri][0]);
                // (the reservation's result is first looked at here, after the scan, so its latency stays hidden)
                if (LIMIT && c[k] && g[k] + c[k] > ((u64) seg_group(r, seg) * a.ndig + d + 1) * r.limit_cap) {
                    *a.overflow = 1;  // the optimistic layout is too small: this run goes to the dump tile
                    delta = r.dump;
                    if (KIND == kDigitShard && !a.shard_local) delta = (u64) (r.dump_ptr - a.peer_out[ri][0]);
                }
                s_delta[d] = delta;
            }
            ex += c[k];
        }
    }
    __syncthreads();
    {
#pragma unroll
    for (int j = 0; j < kPartItems; ++j) {
        u32 i = j * THREADS + tid;
        if (i < ntile) s_tup[s_off[dr[j] >> 16] + (dr[j] & 0xffffu)] = v[j];
    }
    if (WMODE == kWriteBulk) {
        fence_async_smem();
        __syncthreads();
        for (u32 d = tid; d < a.ndig; d += THREADS) {
            u32 cd = s_cnt[d];
            Tup *ob = (KIND == kDigitShard && !a.shard_local) ? a.peer_out[ri][0] : r.out;
            if (cd) bulk_s2g(ob + s_delta[d] + s_off[d], s_tup + s_off[d], cd * (u32) sizeof(Tup));
        }
        bulk_commit();
        bulk_wait_read0();
    } else {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kPartItems; ++j) {
            u32 i = j * THREADS + tid;
            if (i < ntile) {
                Tup t = s_tup[i];
                u32 d = digit<KIND>(t.val, a);
                Tup *ob = (KIND == kDigitShard && !a.shard_local) ? a.peer_out[ri][0] : r.out;
                st_stream(ob + s_delta[d] + i, t);
            }
        }
    }
    }
}

// ---- (4)+(5): work planning ---------------------------------------------------------------------
struct PlanArgs {
    const u64 *offB;  // [nparts+1] build-side partition offsets
    const u64 *offP;  // [nparts+1] probe-side partition offsets
    u32 nparts;
    Item *items;
    u32 item_cap;
    u32 *nitems;
    u32 *err;
};
// One work item per (partition with both sides non-empty, chunk of <= kProbeChunk probe tuples):
// a skewed probe partition is split over many CTAs that each rebuild the (small) table.
__global__ void __launch_bounds__(1024) k_plan(PlanArgs a) {
    __shared__ u32 s_w[32];
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 per = (a.nparts + 1023) / 1024;
    const u32 p0 = min(a.nparts, tid * per), p1 = min(a.nparts, p0 + per);
    u32 c = 0;
    for (u32 p = p0; p < p1; ++p) {
        u64 nb = a.offB[p + 1] - a.offB[p], np = a.offP[p + 1] - a.offP[p];
        if (nb && np) c += (u32) ((np + kProbeChunk - 1) / kProbeChunk);
    }
    u32 inc = warp_incl_scan(c);
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        u32 w = s_w[lane];
        u32 wi = warp_incl_scan(w);
        s_w[lane] = wi - w;
        if (lane == 31) {
            *a.nitems = wi <= a.item_cap ? wi : 0;
            if (wi > a.item_cap) *a.err = 1;
        }
    }
    __syncthreads();
    u32 at = inc - c + s_w[warp];
    for (u32 p = p0; p < p1; ++p) {
        u64 nb = a.offB[p + 1] - a.offB[p], np = a.offP[p + 1] - a.offP[p];
        if (nb && np) {
            u32 k = (u32) ((np + kProbeChunk - 1) / kProbeChunk);
            for (u32 ch = 0; ch < k; ++ch)
                if (at + ch < a.item_cap) a.items[at + ch] = Item{p, ch};
            at += k;
        }
    }
}

// (2b')+(plan), two-pass plans: one CTA per pass-1 partition.  The 2^b2 sub-partition counters of
// a pass-1 partition only need a LOCAL prefix sum on top of that partition's pass-1 offset, so the
// 2^bits_total-entry scan and the work-item planning run fully in parallel: offsets + cursors for
// both relations, then this partition's work items appended with one global atomic per CTA
// (item order is irrelevant).
struct ScanPlanArgs {
    const u64 *hist2[2];   // [nseg * ndig] sub-partition counts (build, probe)
    const u64 *off1[2];    // [nseg + 1] pass-1 offsets
    u64 *off2[2];          // [nseg * ndig + 1]
    u64 *cursor2[2];       // [nseg * ndig]
    u32 nseg, ndig;
    Item *items;
    u32 item_cap;
    u32 *nitems;
    u32 *err;
};
__global__ void __launch_bounds__(kMaxDigits) k_scan_parts_plan(ScanPlanArgs a) {
    __shared__ u64 s_w[32];
    __shared__ u32 s_base;
    const u32 seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 p = seg * a.ndig + tid;
    u64 cnt[2];
#pragma unroll
    for (int ri = 0; ri < 2; ++ri) {
        u64 c = tid < a.ndig ? a.hist2[ri][p] : 0;
        cnt[ri] = c;
        u64 inc = warp_incl_scan64(c);
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            u64 w = lane < (kMaxDigits / 32) ? s_w[lane] : 0;
            u64 wi = warp_incl_scan64(w);
            s_w[lane] = wi - w;
        }
        __syncthreads();
        u64 off = a.off1[ri][seg] + inc - c + s_w[warp];
        if (tid < a.ndig) {
            a.off2[ri][p] = off;
            a.cursor2[ri][p] = off;
            if (seg == a.nseg - 1 && tid == a.ndig - 1) a.off2[ri][p + 1] = off + c;
        }
        __syncthreads();
    }
    // work items of this pass-1 partition
    u32 k = (cnt[0] && cnt[1]) ? (u32) ((cnt[1] + kProbeChunk - 1) / kProbeChunk) : 0;
    u32 inc = warp_incl_scan(k);
    u32 *s_w32 = reinterpret_cast<u32 *>(s_w);
    if (lane == 31) s_w32[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        u32 w = lane < (kMaxDigits / 32) ? s_w32[lane] : 0;
        u32 wi = warp_incl_scan(w);
        s_w32[lane] = wi - w;
        u32 total = __shfl_sync(0xffffffffu, wi, 31);
        if (lane == 0) s_base = total ? atomicAdd(a.nitems, total) : 0;
    }
    __syncthreads();
    u32 at = s_base + inc - k + s_w32[warp];
    for (u32 ch = 0; ch < k; ++ch) {
        if (at + ch < a.item_cap) a.items[at + ch] = Item{p, ch};
        else *a.err = 1;
    }
}

// block-wide exclusive scans for 1024-thread CTAs (one element per thread)
__device__ __forceinline__ u64 block_excl_scan64(u64 v, u64 *s_w, u64 *total) {
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 inc = warp_incl_scan64(v);
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        u64 w = s_w[lane];
        u64 wi = warp_incl_scan64(w);
        s_w[lane] = wi - w;
        if (lane == 31) s_w[32] = wi;
    }
    __syncthreads();
    u64 ex = inc - v + s_w[warp];
    if (total) *total = s_w[32];
    __syncthreads();
    return ex;
}

// Receive-side layout of the DMA-shipped sharded join, for ONE relation.  all_hist[src][dest <<
// sub_bits | p1] = pass-1 histograms of every rank.  Every source ships to this rank one
// contiguous chunk that is already partitioned by p1, so the receive buffer is [src0: p1 = 0..][src1:
// ...]: world << sub_bits contiguous PIECES.  Pass 2 treats each piece as a segment whose counters are
// shared per p1 (group = piece & (2^sub_bits - 1)).
//   seg_off / seg_tile0 [world << sub_bits | +1]   piece boundaries / first pass-2 tile
//   off1 [2^sub_bits + 1]                           pass-1 partition sizes summed over sources (prefix)
//   tot [world * world]                             tot[src * world + dest] = tuples src ships to dest
struct ShardLayoutArgs {
    const u64 *all_hist;  // [world][world << sub_bits]
    u64 *seg_off;
    u32 *seg_tile0;
    u64 *off1;
    u64 *tot;
    u32 world, rank;
    int sub_bits;
};
__global__ void __launch_bounds__(1024) k_shard_layout(ShardLayoutArgs a) {
    __shared__ u64 s_w[33];
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 nd1 = 1u << a.sub_bits, ndig = a.world << a.sub_bits;
    // tot[src][dest]
    for (u32 pr = warp; pr < a.world * a.world; pr += 32) {
        const u32 src = pr / a.world, dest = pr % a.world;
        u64 c = 0;
        for (u32 p = lane; p < nd1; p += 32) c += a.all_hist[(u64) src * ndig + (dest << a.sub_bits) + p];
        c = warp_sum64(c);
        if (lane == 0) a.tot[pr] = c;
    }
    // pieces this rank receives, in (src, p1) order
    const u32 nseg = ndig;  // world * nd1
    u64 cnt = 0;
    if (tid < nseg) {
        const u32 src = tid >> a.sub_bits, p1 = tid & (nd1 - 1);
        cnt = a.all_hist[(u64) src * ndig + (a.rank << a.sub_bits) + p1];
    }
    u64 total;
    u64 ex = block_excl_scan64(cnt, s_w, &total);
    u64 tl = (cnt + kTile - 1) / kTile, ttotal;
    u64 tex = block_excl_scan64(tl, s_w, &ttotal);
    if (tid < nseg) {
        a.seg_off[tid] = ex;
        a.seg_tile0[tid] = (u32) tex;
        if (tid == nseg - 1) {
            a.seg_off[nseg] = total;
            a.seg_tile0[nseg] = (u32) ttotal;
        }
    }
    // pass-1 partitions summed over sources
    u64 g = 0;
    if (tid < nd1)
        for (u32 src = 0; src < a.world; ++src) g += a.all_hist[(u64) src * ndig + (a.rank << a.sub_bits) + tid];
    u64 gex = block_excl_scan64(g, s_w, &total);
    if (tid < nd1) {
        a.off1[tid] = gex;
        if (tid == nd1 - 1) a.off1[nd1] = total;
    }
}

// ---- optimistic pass 1: no histogram -----------------------------------------------------------------
// The pass-1 histogram only serves to place 2^b1 partitions back to back.  When a sampled histogram says
// the partitions are balanced, pass 1 skips it: every partition gets a fixed-capacity region (expected
// size + 12.5 % + 8192), the scatter appends with its usual per-digit cursors, and the exact sizes fall
// out of the cursors afterwards.  A partition that outgrows its region raises `overflow` (nothing is
// written out of bounds) and the host re-runs the join through the exact histogram path.

// 1/64 sample: one 128-byte granule (8 tuples) out of every 64, at a position that varies per block.
struct SampleArgs {
    const Tup *in[2];
    u64 n[2];
    int shift;
    u32 mask, ndig;
    u32 *hist;  // [2][ndig]
};
__global__ void __launch_bounds__(512) k_sample_hist(SampleArgs a) {
    __shared__ u32 s_h[kMaxDigits];
    const int ri = blockIdx.y;
    for (u32 d = threadIdx.x; d < a.ndig; d += blockDim.x) s_h[d] = 0;
    __syncthreads();
    const u64 t = (u64) blockIdx.x * blockDim.x + threadIdx.x;
    const u64 k = t >> 3;
    const u64 idx = ((k << 6) + ((k * 29) & 63)) * 8 + (t & 7);
    if (idx < a.n[ri]) {
        Tup v = ld_stream(a.in[ri] + idx);
        atomicAdd(&s_h[(hash32(v.val) >> a.shift) & a.mask], 1u);
    }
    __syncthreads();
    for (u32 d = threadIdx.x; d < a.ndig; d += blockDim.x) {
        u32 c = s_h[d];
        if (c) atomicAdd(&a.hist[ri * a.ndig + d], c);
    }
}

struct FixedArgs {
    u64 *cursor[2];       // [ndig] scatter cursors, start at d * cap
    u64 cap[2];
    u32 ndig;
    // k_fixed_finish outputs, per relation
    u64 *seg_beg[2];      // [ndig] d * cap
    u64 *seg_end[2];      // [ndig] seg_beg + tuples appended
    u64 *off1[2];         // [ndig + 1] exact prefix of the partition sizes (pass 2 packs its output with it)
    u32 *tile0[2];        // [ndig + 1]
    u32 *overflow;
    u32 rel_mask;         // bit r set: relation r uses the fixed-capacity layout (the other one has a histogram)
};
__device__ __forceinline__ u64 fixed_beg1(const FixedArgs &a, int ri, u32 d) { return (u64) d * a.cap[ri]; }
__global__ void k_fixed_cursors(FixedArgs a) {
    const int ri = blockIdx.x;
    if (!((a.rel_mask >> ri) & 1u)) return;
    for (u32 d = threadIdx.x; d < a.ndig; d += blockDim.x) a.cursor[ri][d] = fixed_beg1(a, ri, d);
}
__global__ void __launch_bounds__(kMaxDigits) k_fixed_finish(FixedArgs a) {
    __shared__ u64 s_w[33];
    const int ri = blockIdx.x;
    if (!((a.rel_mask >> ri) & 1u)) return;
    const u32 tid = threadIdx.x;
    u64 beg = 0, cnt = 0;
    if (tid < a.ndig) {
        beg = fixed_beg1(a, ri, tid);
        const u64 room = (u64) (tid + 1) * a.cap[ri] - beg;
        cnt = a.cursor[ri][tid] - beg;
        if (cnt > room) {
            *a.overflow = 1;
            cnt = room;
        }
    }
    u64 total, ttotal;
    u64 ex = block_excl_scan64(cnt, s_w, &total);
    u64 tl = (cnt + kTile - 1) / kTile;
    u64 tex = block_excl_scan64(tl, s_w, &ttotal);
    if (tid < a.ndig) {
        a.seg_beg[ri][tid] = beg;
        a.seg_end[ri][tid] = beg + cnt;
        a.off1[ri][tid] = ex;
        a.tile0[ri][tid] = (u32) tex;
        if (tid == a.ndig - 1) {
            a.off1[ri][a.ndig] = total;
            a.tile0[ri][a.ndig] = (u32) ttotal;
        }
    }
}

// Optimistic pass 2: every final partition p owns the fixed region [p * cap, (p + 1) * cap) of the output, so
// the pass needs no histogram (16 bytes per tuple less HBM traffic).  k_fixed_cursors2 starts the scatter
// cursors at the region starts; after the scatter the cursors ARE the partition ends, and k_plan_fixed (one
// CTA per pass-1 partition, like k_scan_parts_plan) writes the region starts, clamps an overflowed end
// (the host then re-runs the exact path) and appends the work items.
struct PlanFixedArgs {
    u64 *end[2];      // [nparts] scatter cursors: in = region start + tuples appended, out = clamped partition end
    u64 *beg[2];      // [nparts] p * cap
    u64 cap[2];
    u32 nseg, ndig;
    Item *items;
    u32 item_cap;
    u32 *nitems;
    u32 *err;
    u32 *overflow;
};
// Region starts are multiples of 8 tuples (128-byte lines): measured 5 % faster in k_join than unaligned or
// pseudo-randomly shifted starts (profiles/r01_tuning_notes.md).
__device__ __forceinline__ u64 fixed_beg(const PlanFixedArgs &a, int ri, u32 p) { return (u64) p * a.cap[ri]; }
__global__ void k_fixed_cursors2(PlanFixedArgs a) {
    const int ri = blockIdx.y;
    const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < a.nseg * a.ndig) {
        a.end[ri][p] = fixed_beg(a, ri, p);
        a.beg[ri][p] = fixed_beg(a, ri, p);
    }
}
__global__ void __launch_bounds__(kMaxDigits) k_plan_fixed(PlanFixedArgs a) {
    __shared__ u32 s_w32[32];
    __shared__ u32 s_base;
    const u32 seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 p = seg * a.ndig + tid;
    u64 cnt[2] = {0, 0};
    if (tid < a.ndig) {
#pragma unroll
        for (int ri = 0; ri < 2; ++ri) {
            const u64 b = fixed_beg(a, ri, p), room = (u64) (p + 1) * a.cap[ri] - b;
            u64 c = a.end[ri][p] - b;
            if (c > room) {
                *a.overflow = 1;
                c = room;
                a.end[ri][p] = b + c;
            }
            a.beg[ri][p] = b;
            cnt[ri] = c;
        }
    }
    u32 k = (cnt[0] && cnt[1]) ? (u32) ((cnt[1] + kProbeChunk - 1) / kProbeChunk) : 0;
    u32 inc = warp_incl_scan(k);
    if (lane == 31) s_w32[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        u32 w = lane < (kMaxDigits / 32) ? s_w32[lane] : 0;
        u32 wi = warp_incl_scan(w);
        s_w32[lane] = wi - w;
        u32 total = __shfl_sync(0xffffffffu, wi, 31);
        if (lane == 0) s_base = total ? atomicAdd(a.nitems, total) : 0;
    }
    __syncthreads();
    u32 at = s_base + inc - k + s_w32[warp];
    for (u32 ch = 0; ch < k; ++ch) {
        if (at + ch < a.item_cap) a.items[at + ch] = Item{p, ch};
        else *a.err = 1;
    }
}

// (2b'') per-relation half of k_scan_parts_plan: offsets + cursors of one relation's 2^b2
// sub-partitions per pass-1 partition (one CTA each), so that a relation whose data has arrived can
// run its second pass while the other relation is still in flight.
struct ScanPartsRelArgs {
    const u64 *hist2;  // [nseg * ndig]
    const u64 *off1;   // [nseg + 1]
    u64 *off2;         // [nseg * ndig + 1]
    u64 *cursor2;      // [nseg * ndig]
    u32 nseg, ndig;
};
__global__ void __launch_bounds__(kMaxDigits) k_scan_parts_rel(ScanPartsRelArgs a) {
    __shared__ u64 s_w[33];
    const u32 seg = blockIdx.x, tid = threadIdx.x;
    const u32 p = seg * a.ndig + tid;
    u64 c = tid < a.ndig ? a.hist2[p] : 0;
    u64 ex = block_excl_scan64(c, s_w, nullptr);
    if (tid < a.ndig) {
        u64 off = a.off1[seg] + ex;
        a.off2[p] = off;
        a.cursor2[p] = off;
        if (seg == a.nseg - 1 && tid == a.ndig - 1) a.off2[p + 1] = off + c;
    }
}
// ... and the planning half: work items of every final partition from both relations' offsets
struct PlanPartsArgs {
    const u64 *offB;   // [nparts (+1)] partition starts
    const u64 *offP;
    u64 *endB;         // fixed-capacity layouts: [nparts] scatter cursors = partition ends (clamped here); null = packed, end = off[p + 1]
    u64 *endP;
    u64 capB, capP;    // region capacity of a fixed-capacity layout
    u32 ndig;  // partitions per CTA
    Item *items;
    u32 item_cap;
    u32 *nitems;
    u32 *err;
    u32 *overflow;
};
__device__ __forceinline__ u64 part_size(const u64 *off, u64 *end, u64 cap, u32 p, u32 *overflow) {
    if (!end) return off[p + 1] - off[p];
    u64 c = end[p] - off[p];
    if (c > cap) {  // an optimistic region overflowed: keep every range inside the buffer, the host re-runs the exact path
        *overflow = 1;
        c = cap;
        end[p] = off[p] + c;
    }
    return c;
}
__global__ void __launch_bounds__(kMaxDigits) k_plan_parts(PlanPartsArgs a) {
    __shared__ u64 s_w[33];
    __shared__ u32 s_base;
    const u32 tid = threadIdx.x;
    const u32 p = blockIdx.x * a.ndig + tid;
    u64 k = 0;
    if (tid < a.ndig) {
        u64 nb = part_size(a.offB, a.endB, a.capB, p, a.overflow), np = part_size(a.offP, a.endP, a.capP, p, a.overflow);
        if (nb && np) k = (np + kProbeChunk - 1) / kProbeChunk;
    }
    u64 total;
    u64 ex = block_excl_scan64(k, s_w, &total);
    if (tid == 0) s_base = total ? atomicAdd(a.nitems, (u32) total) : 0;
    __syncthreads();
    u32 at = s_base + (u32) ex;
    for (u32 ch = 0; ch < (u32) k; ++ch) {
        if (at + ch < a.item_cap) a.items[at + ch] = Item{p, ch};
        else *a.err = 1;
    }
}

__global__ void k_set_single_part(u64 *offB, u64 nB, u64 *offP, u64 nP) {
    offB[0] = 0; offB[1] = nB;
    offP[0] = 0; offP[1] = nP;
}

// (5) prefix sum of the per-item match counts -> per-item output offsets + total
__global__ void __launch_bounds__(1024) k_scan_items(const u64 *cnt, const u32 *nitems, u64 *off, u64 *total) {
    __shared__ u64 s_w[32];
    const u32 n = *nitems;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 per = (n + 1023) / 1024;
    const u32 i0 = min(n, tid * per), i1 = min(n, i0 + per);
    u64 c = 0;
    for (u32 i = i0; i < i1; ++i) c += cnt[i];
    u64 inc = warp_incl_scan64(c);
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        u64 w = s_w[lane];
        u64 wi = warp_incl_scan64(w);
        s_w[lane] = wi - w;
        if (lane == 31) *total = wi;
    }
    __syncthreads();
    u64 run = inc - c + s_w[warp];
    for (u32 i = i0; i < i1; ++i) {
        off[i] = run;
        run += cnt[i];
    }
}

// ---- closing the holes of a positional emit (rhj_join.cuh, k_join<FUSED, POS>) ------------------------------------
// Order of the result is free, so the final count is F = cursor - holes and the valid pairs behind F move into the holes
// before F.  k_holes_count / k_holes_scan / k_holes_list build the two position lists (holes in [0, F), valid pairs in
// [F, cursor)), k_holes_fill moves.  Only launched when a join left holes.
constexpr u32 kHoleTile = 2048;
__device__ __forceinline__ bool is_hole(const Pair &q) { return q.r == RHJ_HOLE && q.s == RHJ_HOLE; }
// cnt[tile] = holes of the tile that lie before F; cnt[ntiles + tile] = valid pairs of the tile at or behind F
__global__ void __launch_bounds__(256) k_holes_count(const Pair *out, u64 cursor, u64 F, u32 ntiles, u32 *cnt) {
    __shared__ u32 s_c[2];
    if (threadIdx.x < 2) s_c[threadIdx.x] = 0;
    __syncthreads();
    const u64 base = (u64) blockIdx.x * kHoleTile;
    u32 h = 0, v = 0;
    for (u32 k = threadIdx.x; k < kHoleTile; k += 256) {
        const u64 i = base + k;
        if (i < cursor) {
            const bool hole = is_hole(out[i]);
            h += i < F && hole;
            v += i >= F && !hole;
        }
    }
    h = (u32) warp_sum64(h);
    v = (u32) warp_sum64(v);
    if (lane_id() == 0) {
        if (h) atomicAdd(&s_c[0], h);
        if (v) atomicAdd(&s_c[1], v);
    }
    __syncthreads();
    if (threadIdx.x < 2) cnt[threadIdx.x * ntiles + blockIdx.x] = s_c[threadIdx.x];
}
// exclusive scans of both count arrays (one CTA each): off[part * ntiles + tile], totals[part]
__global__ void __launch_bounds__(1024) k_holes_scan(const u32 *cnt, u32 ntiles, u64 *off, u64 *totals) {
    __shared__ u64 s_w[33];
    const u32 part = blockIdx.x;
    const u32 per = (ntiles + 1023) / 1024;
    const u32 i0 = min(ntiles, threadIdx.x * per), i1 = min(ntiles, i0 + per);
    u64 c = 0;
    for (u32 i = i0; i < i1; ++i) c += cnt[part * ntiles + i];
    u64 total;
    u64 run = block_excl_scan64(c, s_w, &total);
    for (u32 i = i0; i < i1; ++i) {
        off[part * ntiles + i] = run;
        run += cnt[part * ntiles + i];
    }
    if (threadIdx.x == 0) totals[part] = total;
}
__global__ void __launch_bounds__(256) k_holes_list(const Pair *out, u64 cursor, u64 F, u32 ntiles, const u64 *off, u64 *list_head,
                                                    u64 *list_tail) {
    __shared__ u32 s_at[2];
    if (threadIdx.x < 2) s_at[threadIdx.x] = 0;
    __syncthreads();
    const u64 base = (u64) blockIdx.x * kHoleTile;
    for (u32 k = threadIdx.x; k < kHoleTile; k += 256) {
        const u64 i = base + k;
        if (i >= cursor) continue;
        const bool hole = is_hole(out[i]);
        if (i < F && hole) list_head[off[blockIdx.x] + atomicAdd(&s_at[0], 1u)] = i;
        if (i >= F && !hole) list_tail[off[ntiles + blockIdx.x] + atomicAdd(&s_at[1], 1u)] = i;
    }
}
__global__ void __launch_bounds__(256) k_holes_fill(Pair *out, const u64 *list_head, const u64 *list_tail, u64 k) {
    for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < k; i += (u64) gridDim.x * blockDim.x) out[list_head[i]] = out[list_tail[i]];
}

}  // namespace rhj
