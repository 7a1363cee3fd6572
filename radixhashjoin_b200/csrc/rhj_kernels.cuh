// rhj_kernels.cuh -- the sm_100a kernels of the radix hash join.
//
// north_star step            reference loop it replaces                       kernel here
//  (1) histogram             HistogramJob::run, JobScheduler.cpp:149-155      k_hist
//  (2) prefix sum            PartitionJob::run 163-169, Result.cpp:100-107    k_scan_digits, k_scan_parts
//  (3) partition scatter     PartitionJob::run 170-174 + structs.cpp:183-194  k_scatter
//  (4) per-bucket build/probe Result::join_buckets, Result.cpp:43-76           k_join            (rhj_join.cuh)
//  (5) emitter               add_result/addAll, Result.cpp:21-35,78-84,111-121 k_join<COUNT|WRITE|FUSED>, k_scan_items
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
constexpr int kPartThreads = RHJ_PART_THREADS;  // threads per partition CTA
constexpr int kPartItems = RHJ_PART_ITEMS;      // tuples per thread
constexpr int kTile = kPartThreads * kPartItems;  // 4096 tuples = 64 KiB staged per CTA
constexpr int kMaxBitsPerPass = 9;
constexpr int kMaxDigits = 1 << kMaxBitsPerPass;  // 512 digits per pass

enum DigitKind { kDigitRaw = 0, kDigitHash = 1, kDigitRank = 2 };

template <int KIND>
__device__ __forceinline__ u32 digit(u64 v, int shift, u32 mask) {
    if (KIND == kDigitRaw) return (u32) (v >> shift) & mask;
    if (KIND == kDigitHash) return (hash32(v) >> shift) & mask;
    return (hash_hi32(v) >> shift) & mask;
}

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
};
struct PartArgs {
    PartRel rel[2];
    int shift;
    u32 mask;
    u32 ndig;
};

// Maps a relation-local tile index to its tuple range.  SEG tiles never straddle a segment.
template <bool SEG>
__device__ __forceinline__ bool tile_range(const PartRel &r, u32 lt, u32 &seg, u64 &beg, u64 &end) {
    if (!SEG) {
        seg = 0;
        beg = (u64) lt * kTile;
        end = min(beg + (u64) kTile, r.n);
        return beg < r.n;
    }
    if (lt >= r.seg_tile0[r.nseg]) return false;
    u32 lo = 0, hi = r.nseg;  // last seg with seg_tile0[seg] <= lt
    while (hi - lo > 1) {
        u32 mid = (lo + hi) >> 1;
        if (r.seg_tile0[mid] <= lt) lo = mid; else hi = mid;
    }
    seg = lo;
    beg = r.seg_off[lo] + (u64) (lt - r.seg_tile0[lo]) * kTile;
    end = min(beg + (u64) kTile, r.seg_off[lo + 1]);
    return true;
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
            u64 *h = a.rel[cur_rel].hist + (u64) cur_seg * a.ndig;
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
            u32 d = ok[j] ? digit<KIND>(v[j].val, a.shift, a.mask) : kEmpty;
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
        u64 *h = a.rel[cur_rel].hist + (u64) cur_seg * a.ndig;
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

// (2b) Exclusive prefix sum over the 2^bits_total final partition counters (up to 2^18), one CTA
// per relation, each thread owning a contiguous slice.
struct ScanPartsArgs {
    const u64 *hist[2];
    u64 *off[2];
    u64 *cursor[2];
    u32 nparts;
};
__global__ void __launch_bounds__(1024) k_scan_parts(ScanPartsArgs a) {
    __shared__ u64 s_w[32];
    const int ri = blockIdx.x;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 per = (a.nparts + 1023) / 1024;
    const u32 p0 = min(a.nparts, tid * per), p1 = min(a.nparts, p0 + per);
    u64 c = 0;
    for (u32 p = p0; p < p1; ++p) c += a.hist[ri][p];
    u64 inc = warp_incl_scan64(c);
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        u64 w = s_w[lane];
        u64 wi = warp_incl_scan64(w);
        s_w[lane] = wi - w;
    }
    __syncthreads();
    u64 run = inc - c + s_w[warp];
    for (u32 p = p0; p < p1; ++p) {
        a.off[ri][p] = run;
        a.cursor[ri][p] = run;
        run += a.hist[ri][p];
    }
    if (p1 == a.nparts && p0 < p1) a.off[ri][a.nparts] = run;
    if (a.nparts == 0 && tid == 0) a.off[ri][0] = 0;
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
template <int KIND, bool SEG, int WMODE>
__global__ void __launch_bounds__(kPartThreads, RHJ_PART_MINBLOCKS) k_scatter(PartArgs a) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    Tup *s_tup = reinterpret_cast<Tup *>(dyn_smem);
    __shared__ u32 s_cnt[kMaxDigits];
    __shared__ u32 s_off[kMaxDigits];
    __shared__ u64 s_delta[kMaxDigits];  // global index of sorted slot i of digit d = s_delta[d] + i
    __shared__ u32 s_w[32];

    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ri = blockIdx.x >= a.rel[0].ntiles;
    const PartRel &r = a.rel[ri];
    u32 seg;
    u64 beg, end;
    if (!tile_range<SEG>(r, blockIdx.x - (ri ? a.rel[0].ntiles : 0), seg, beg, end)) return;
    const u32 ntile = (u32) (end - beg);

    for (u32 d = tid; d < a.ndig; d += kPartThreads) s_cnt[d] = 0;
    Tup v[kPartItems];
    u32 dr[kPartItems];  // digit << 16 | rank   (rank < 4096)
#pragma unroll
    for (int j = 0; j < kPartItems; ++j) {
        u32 i = j * kPartThreads + tid;
        if (i < ntile) v[j] = ld_stream(r.in + beg + i);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kPartItems; ++j) {
        u32 i = j * kPartThreads + tid;
        if (i < ntile) {
            u32 d = digit<KIND>(v[j].val, a.shift, a.mask);
            u32 rk = atomicAdd(&s_cnt[d], 1u);
            dr[j] = (d << 16) | rk;
        }
    }
    __syncthreads();
    // reserve the runs (global atomics issued first so their latency overlaps the block scan);
    // thread t owns the kDigitsPerThread consecutive digits starting at t * kDigitsPerThread
    constexpr int kDigitsPerThread = (kMaxDigits + kPartThreads - 1) / kPartThreads;
    u32 c[kDigitsPerThread];
    u64 g[kDigitsPerThread];
    u32 csum = 0;
#pragma unroll
    for (int k = 0; k < kDigitsPerThread; ++k) {
        u32 d = tid * kDigitsPerThread + k;
        c[k] = d < a.ndig ? s_cnt[d] : 0;
        g[k] = 0;
        if (c[k]) g[k] = atomicAdd(r.cursor + (u64) seg * a.ndig + d, (u64) c[k]);
        csum += c[k];
    }
    u32 inc = warp_incl_scan(csum);
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        u32 w = lane < (kPartThreads / 32) ? s_w[lane] : 0;
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
                s_delta[d] = g[k] - ex;
            }
            ex += c[k];
        }
    }
    __syncthreads();
    {
#pragma unroll
    for (int j = 0; j < kPartItems; ++j) {
        u32 i = j * kPartThreads + tid;
        if (i < ntile) s_tup[s_off[dr[j] >> 16] + (dr[j] & 0xffffu)] = v[j];
    }
    if (WMODE == kWriteBulk) {
        fence_async_smem();
        __syncthreads();
        for (u32 d = tid; d < a.ndig; d += kPartThreads) {
            u32 cd = s_cnt[d];
            if (cd) bulk_s2g(r.out + s_delta[d] + s_off[d], s_tup + s_off[d], cd * (u32) sizeof(Tup));
        }
        bulk_commit();
        bulk_wait_read0();
    } else {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kPartItems; ++j) {
            u32 i = j * kPartThreads + tid;
            if (i < ntile) {
                Tup t = s_tup[i];
                u32 d = digit<KIND>(t.val, a.shift, a.mask);
                st_stream(r.out + s_delta[d] + i, t);
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

}  // namespace rhj
