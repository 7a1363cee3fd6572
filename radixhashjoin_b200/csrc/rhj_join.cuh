// rhj_join.cuh -- (4)+(5): per-partition build / probe / emit, software-pipelined.
//
// Replaces JoinJob::run + Result::join_buckets (JobScheduler.cpp:186-192, Result.cpp:43-76: a
// bucket-chain index with `payload % prime` hashing, one pthread job per bucket) and the
// add_result / addAll page list (Result.cpp:21-35,78-84,111-121).
//
// Persistent CTAs, two per SM, each walking its work items (partition p, probe chunk c) with a
// fixed stride.  Per item ("stage"):
//   - the build partition's tuples are TMA-bulk-loaded verbatim into one of TWO shared-memory
//     buffers (cp.async.bulk + mbarrier); the load for stage i+1 is issued at the top of stage i,
//     so it is in flight while stage i builds and probes;
//   - the first 2048 probe tuples of stage i+1 are loaded into registers at the top of stage i as
//     well (coalesced 16-B ld.global.nc), and the descriptor of stage i+2 is fetched then too;
//   - build: every staged tuple claims a slot of an open-addressing table of u32 indices (linear
//     probing, shared-memory atomicCAS, load factor <= 0.625); passing an equal value while
//     claiming flags the chunk as "has duplicate keys";
//   - probe: unique-key chunks stop at the first hit, duplicate-key chunks count then re-walk;
//   - emit: matches of a round are ranked with ballots + one shared atomic per warp, one thread
//     reserves the round's output range (FUSED: one global atomic per round; WRITE: running offset
//     from the count pass; COUNT: nothing is written) and lanes store 16-B pairs at consecutive
//     positions.
// A build partition larger than one buffer (duplicate-heavy keys that no radix bit can split) is
// processed in further chunks of the same stage, re-reading the probe chunk per build chunk.
// Algorithmic bytes: 16 per input tuple read + 16 per result pair written.
#pragma once
#include "rhj_device.cuh"

namespace rhj {

constexpr int kJoinThreads = 512;
constexpr int kJoinItems = 4;                        // probe tuples per thread per round
constexpr int kRound = kJoinThreads * kJoinItems;    // 2048 probe tuples per round
#ifndef RHJ_JOIN_CAP
#define RHJ_JOIN_CAP 2560
#endif
#ifndef RHJ_JOIN_SLOTS
#define RHJ_JOIN_SLOTS 4096
#endif
#ifndef RHJ_JOIN_PREFETCH
#define RHJ_JOIN_PREFETCH 1
#endif
constexpr u32 kBuildCap = RHJ_JOIN_CAP;              // build tuples per staged chunk (40 KiB), x2 buffers
constexpr u32 kSlots = RHJ_JOIN_SLOTS;               // open-addressing slots (u32 index) (16 KiB)
constexpr u32 kProbeChunk = 16384;                   // probe tuples per work item
constexpr u32 kTargetBuildPerPart = 2048;            // radix bits are chosen for this average
constexpr size_t kJoinSmemBytes = 2 * (size_t) kBuildCap * sizeof(Tup) + (size_t) kSlots * sizeof(u32);

struct Item {
    u32 part;
    u32 chunk;
};

enum JoinMode { kJoinCount = 0, kJoinWrite = 1, kJoinFused = 2 };

struct JoinArgs {
    const Tup *build;    // partitioned build relation (the smaller input)
    const Tup *probe;    // partitioned probe relation
    const u64 *offB;
    const u64 *offP;
    const Item *items;
    const u32 *nitems;
    u32 *work_counter;   // v1 kernel only: dynamic item scheduler
    u64 *item_cnt;       // COUNT: out; per-item match count
    const u64 *item_off; // WRITE: per-item output offset
    u64 *out_cursor;     // FUSED: global reservation cursor (ends as the total match count)
    Pair *out;
    u64 capacity;
    int build_is_S;      // output is always (rowidR, rowidS): Result.cpp:66-69
};

__device__ __forceinline__ u32 slot_of(u64 v) { return hash32(v) & (kSlots - 1); }

struct Stage {
    u64 b0;       // first build tuple of the partition
    u64 nb;       // build tuples of the partition
    u64 p0;       // first probe tuple of this chunk
    u64 out_off;  // WRITE: output offset of the item
    u32 np;       // probe tuples of this chunk (<= kProbeChunk)
    bool valid;
};

template <int MODE>
__device__ __forceinline__ Stage load_stage(const JoinArgs &a, u32 idx, u32 nitems) {
    Stage s;
    s.valid = idx < nitems;
    s.b0 = s.nb = s.p0 = s.out_off = 0;
    s.np = 0;
    if (s.valid) {
        const Item it = a.items[idx];
        s.b0 = a.offB[it.part];
        s.nb = a.offB[it.part + 1] - s.b0;
        const u64 pbeg = a.offP[it.part], pend = a.offP[it.part + 1];
        s.p0 = pbeg + (u64) it.chunk * kProbeChunk;
        s.np = (u32) min((u64) kProbeChunk, pend - s.p0);
        if (MODE == kJoinWrite) s.out_off = a.item_off[idx];
    }
    return s;
}

template <int MODE>
__global__ void __launch_bounds__(kJoinThreads, 2) k_join(JoinArgs a) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    Tup *s_tup0 = reinterpret_cast<Tup *>(dyn_smem);
    u32 *s_slot = reinterpret_cast<u32 *>(dyn_smem + 2 * (size_t) kBuildCap * sizeof(Tup));
    __shared__ __align__(8) u64 s_bar[2];
    __shared__ u32 s_cnt[2];
    __shared__ u64 s_base[2];
    __shared__ u64 s_red[32];

    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 lt_mask = lanemask_lt();
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        s_cnt[0] = 0;
        s_cnt[1] = 0;
    }
    __syncthreads();
    const u32 nitems = *a.nitems;
    const u32 G = gridDim.x;
    u32 idx = blockIdx.x;
    u32 buf = 0, phase = 0, rr = 0;

    Stage cur = load_stage<MODE>(a, idx, nitems);
    Stage nxt = load_stage<MODE>(a, idx + G, nitems);
    Tup t[kJoinItems];
    if (cur.valid) {
        if (tid == 0) {
            const u32 bytes = (u32) min((u64) kBuildCap, cur.nb) * (u32) sizeof(Tup);
            mbar_expect_tx(&s_bar[0], bytes);
            bulk_g2s(s_tup0, a.build + cur.b0, bytes, &s_bar[0]);
        }
        if (RHJ_JOIN_PREFETCH) {
#pragma unroll
            for (int j = 0; j < kJoinItems; ++j) {
                u32 i = j * kJoinThreads + tid;
                if (i < cur.np) t[j] = ld_stream(a.probe + cur.p0 + i);
            }
        }
    }

    while (cur.valid) {
        // ---- top of the stage: everything the NEXT stage needs goes in flight now ----
        const Stage nx2 = load_stage<MODE>(a, idx + 2 * G, nitems);
        Tup tn[kJoinItems];
        if (nxt.valid) {
            if (tid == 0) {
                const u32 bytes = (u32) min((u64) kBuildCap, nxt.nb) * (u32) sizeof(Tup);
                mbar_expect_tx(&s_bar[buf ^ 1], bytes);
                bulk_g2s(s_tup0 + (buf ^ 1) * kBuildCap, a.build + nxt.b0, bytes, &s_bar[buf ^ 1]);
            }
            if (RHJ_JOIN_PREFETCH) {
#pragma unroll
                for (int j = 0; j < kJoinItems; ++j) {
                    u32 i = j * kJoinThreads + tid;
                    if (i < nxt.np) tn[j] = ld_stream(a.probe + nxt.p0 + i);
                }
            }
        }
        const Tup *s_tup = s_tup0 + buf * kBuildCap;
        u64 my_count = 0;                 // COUNT
        u64 run_base = cur.out_off;       // WRITE (used by thread 0)

        for (u64 bb = 0; bb < cur.nb; bb += kBuildCap) {
            const u32 nb = (u32) min((u64) kBuildCap, cur.nb - bb);
            if (bb) {  // overflow chunk: reload this stage's buffer synchronously
                __syncthreads();
                if (tid == 0) {
                    mbar_expect_tx(&s_bar[buf], nb * (u32) sizeof(Tup));
                    bulk_g2s(s_tup0 + buf * kBuildCap, a.build + cur.b0 + bb, nb * (u32) sizeof(Tup), &s_bar[buf]);
                }
            }
            for (u32 i = tid; i < kSlots; i += kJoinThreads) s_slot[i] = kEmpty;
            mbar_wait(&s_bar[buf], (phase >> buf) & 1u);
            phase ^= 1u << buf;
            __syncthreads();

            // ---- build ----
            int dup = 0;
            for (u32 i = tid; i < nb; i += kJoinThreads) {
                const u64 v = s_tup[i].val;
                u32 h = slot_of(v);
                while (true) {
                    u32 old = atomicCAS(&s_slot[h], kEmpty, i);
                    if (old == kEmpty) break;
                    if (s_tup[old].val == v) dup = 1;
                    h = (h + 1) & (kSlots - 1);
                }
            }
            dup = __syncthreads_or(dup);

            // ---- probe + emit, rounds of kRound probe tuples ----
            for (u32 q0 = 0; q0 < cur.np; q0 += kRound) {
                if (q0 || bb || !RHJ_JOIN_PREFETCH) {  // round 0 of the first chunk was prefetched by the previous stage
#pragma unroll
                    for (int j = 0; j < kJoinItems; ++j) {
                        u32 i = q0 + j * kJoinThreads + tid;
                        if (i < cur.np) t[j] = ld_stream(a.probe + cur.p0 + i);
                    }
                }
                if (!dup) {
                    // unique build keys: at most one match per probe tuple
                    u32 m[kJoinItems], ball[kJoinItems];
                    u32 wtotal = 0;
#pragma unroll
                    for (int j = 0; j < kJoinItems; ++j) {
                        m[j] = kEmpty;
                        if (q0 + j * kJoinThreads + tid < cur.np) {
                            u32 h = slot_of(t[j].val);
                            u32 k;
                            while ((k = s_slot[h]) != kEmpty) {
                                if (s_tup[k].val == t[j].val) { m[j] = k; break; }
                                h = (h + 1) & (kSlots - 1);
                            }
                        }
                        ball[j] = __ballot_sync(0xffffffffu, m[j] != kEmpty);
                        wtotal += __popc(ball[j]);
                    }
                    if (MODE == kJoinCount) {
                        if (lane == 0) my_count += wtotal;
                        continue;
                    }
                    u32 wbase = 0;
                    if (lane == 0 && wtotal) wbase = atomicAdd(&s_cnt[rr], wtotal);
                    wbase = __shfl_sync(0xffffffffu, wbase, 0);
                    __syncthreads();
                    if (tid == 0) {
                        u32 c = s_cnt[rr];
                        u64 base;
                        if (MODE == kJoinFused) base = c ? atomicAdd(a.out_cursor, (u64) c) : 0;
                        else { base = run_base; run_base += c; }
                        s_base[rr] = base;
                        s_cnt[rr ^ 1] = 0;
                    }
                    __syncthreads();
                    u64 pos = s_base[rr] + wbase;
#pragma unroll
                    for (int j = 0; j < kJoinItems; ++j) {
                        if (m[j] != kEmpty) {
                            u64 at = pos + __popc(ball[j] & lt_mask);
                            u64 bk = s_tup[m[j]].key;
                            if (at < a.capacity) {
                                if (a.build_is_S) st_stream(a.out + at, t[j].key, bk);
                                else st_stream(a.out + at, bk, t[j].key);
                            }
                        }
                        pos += __popc(ball[j]);
                    }
                    rr ^= 1;
                } else {
                    // duplicate build keys: count every match, reserve, then re-walk and write
                    u32 cnt[kJoinItems];
                    u32 mine = 0;
#pragma unroll
                    for (int j = 0; j < kJoinItems; ++j) {
                        cnt[j] = 0;
                        if (q0 + j * kJoinThreads + tid < cur.np) {
                            u32 h = slot_of(t[j].val);
                            u32 k;
                            while ((k = s_slot[h]) != kEmpty) {
                                if (s_tup[k].val == t[j].val) cnt[j]++;
                                h = (h + 1) & (kSlots - 1);
                            }
                        }
                        mine += cnt[j];
                    }
                    if (MODE == kJoinCount) {
                        my_count += mine;
                        continue;
                    }
                    u32 incl = warp_incl_scan(mine);
                    u32 wtotal = __shfl_sync(0xffffffffu, incl, 31);
                    u32 wbase = 0;
                    if (lane == 0 && wtotal) wbase = atomicAdd(&s_cnt[rr], wtotal);
                    wbase = __shfl_sync(0xffffffffu, wbase, 0);
                    __syncthreads();
                    if (tid == 0) {
                        u32 c = s_cnt[rr];
                        u64 base;
                        if (MODE == kJoinFused) base = c ? atomicAdd(a.out_cursor, (u64) c) : 0;
                        else { base = run_base; run_base += c; }
                        s_base[rr] = base;
                        s_cnt[rr ^ 1] = 0;
                    }
                    __syncthreads();
                    u64 at = s_base[rr] + wbase + (incl - mine);
#pragma unroll
                    for (int j = 0; j < kJoinItems; ++j) {
                        if (cnt[j]) {
                            u32 h = slot_of(t[j].val);
                            u32 k;
                            while ((k = s_slot[h]) != kEmpty) {
                                if (s_tup[k].val == t[j].val) {
                                    u64 bk = s_tup[k].key;
                                    if (at < a.capacity) {
                                        if (a.build_is_S) st_stream(a.out + at, t[j].key, bk);
                                        else st_stream(a.out + at, bk, t[j].key);
                                    }
                                    ++at;
                                }
                                h = (h + 1) & (kSlots - 1);
                            }
                        }
                    }
                    rr ^= 1;
                }
            }
        }
        if (MODE == kJoinCount) {
            u64 w = warp_sum64(my_count);
            if (lane == 0) s_red[warp] = w;
            __syncthreads();
            if (warp == 0) {
                u64 x = lane < (kJoinThreads / 32) ? s_red[lane] : 0;
                x = warp_sum64(x);
                if (lane == 0) a.item_cnt[idx] = x;
            }
        }
        __syncthreads();  // table and buffer `buf` are free: the next top-of-stage may overwrite them
        cur = nxt;
        nxt = nx2;
#pragma unroll
        for (int j = 0; j < kJoinItems; ++j) t[j] = tn[j];
        buf ^= 1;
        idx += G;
    }
}

}  // namespace rhj
