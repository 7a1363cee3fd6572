// rhj_join.cuh -- (4)+(5): per-partition build / probe / emit.
//
// Replaces JoinJob::run + Result::join_buckets (JobScheduler.cpp:186-192, Result.cpp:43-76: a
// bucket-chain index with `payload % prime` hashing, one pthread job per bucket) and the
// add_result / addAll page list (Result.cpp:21-35,78-84,111-121).
//
// Persistent CTAs of 384 threads, three per SM (72 KiB of shared memory each: measured, the third
// CTA hides the barriers and latencies of the other two: 2.03 -> 1.78 ms), pull work items
// (partition p, probe chunk c) from a global counter.  For each build chunk of <= 2560 tuples:
//   - one elected thread TMA-bulk-loads the chunk's tuples verbatim into shared memory
//     (cp.async.bulk + mbarrier); meanwhile every thread issues the coalesced 16-B loads of its
//     first 4 probe tuples and clears the slot table, so both global latencies overlap;
//   - build: every staged tuple claims a slot of the open-addressing table (u32 index into the
//     staged tuples, linear probing, shared-memory atomicCAS; 8192 slots => load factor <= 0.31,
//     0.25 on average -- measured: the warp pays for its longest probe, load 0.5 costs +40 %).
//     A claim that walks past an equal value flags the chunk as "has duplicate keys";
//   - probe: rounds of 1536 probe tuples.  Unique-key chunks stop at the first hit;
//     duplicate-key chunks count, then re-walk to write;
//   - emit: matches of a round are ranked with ballots + one shared atomic per warp; one thread
//     reserves the round's output range (FUSED: one global atomic per round; WRITE: running
//     offset from the count pass; COUNT: nothing is written) and lanes store 16-B pairs at
//     consecutive positions.
// A build partition larger than one chunk (duplicate-heavy keys that no radix bit can split) is
// processed chunk by chunk, re-reading the probe chunk per build chunk.
// A double-buffered, register-prefetching variant was measured and was not faster
// (profiles/r01_tuning_notes.md).
// Algorithmic bytes: 16 per input tuple read + 16 per result pair written.
#pragma once
#include "rhj_device.cuh"

namespace rhj {

#ifndef RHJ_JOIN_THREADS
#define RHJ_JOIN_THREADS 384
#endif
#ifndef RHJ_JOIN_CAP
#define RHJ_JOIN_CAP 2560
#endif
#ifndef RHJ_JOIN_SLOTS
#define RHJ_JOIN_SLOTS 8192
#endif
#ifndef RHJ_JOIN_TARGET
#define RHJ_JOIN_TARGET 2048
#endif
#ifndef RHJ_JOIN_MINBLOCKS
#define RHJ_JOIN_MINBLOCKS 3
#endif
constexpr int kJoinThreads = RHJ_JOIN_THREADS;
#ifndef RHJ_JOIN_ITEMS
#define RHJ_JOIN_ITEMS 4
#endif
#ifndef RHJ_JOIN_ITEMS_POS
#define RHJ_JOIN_ITEMS_POS 4
#endif
constexpr int kJoinItems = RHJ_JOIN_ITEMS;           // probe tuples per thread per round
constexpr int kRound = kJoinThreads * kJoinItems;    // probe tuples per round
constexpr u32 kBuildCap = RHJ_JOIN_CAP;              // build tuples per shared-memory table
typedef u32 slot_t;
constexpr u32 kSlots = RHJ_JOIN_SLOTS;               // open-addressing slots (u32 index)
constexpr slot_t kSlotEmpty = 0xFFFFFFFFu;
constexpr u32 kProbeChunk = 16384;                   // probe tuples per work item
constexpr u32 kTargetBuildPerPart = RHJ_JOIN_TARGET; // radix bits are chosen for this average
constexpr size_t kJoinSmemBytes = (size_t) kBuildCap * sizeof(Tup) + (size_t) kSlots * sizeof(slot_t);

struct Item {
    u32 part;
    u32 chunk;
};

enum JoinMode { kJoinCount = 0, kJoinWrite = 1, kJoinFused = 2 };

struct JoinArgs {
    const Tup *build;    // partitioned build relation (the smaller input)
    const Tup *probe;    // partitioned probe relation
    const u64 *offB;     // [nparts] where each build partition starts ...
    const u64 *offP;
    const u64 *endB;     // ... and ends: offB + 1 for packed layouts, the scatter cursors for fixed-capacity regions
    const u64 *endP;
    const Item *items;
    const u32 *nitems;
    u32 *work_counter;   // dynamic item scheduler
    u64 *item_cnt;       // COUNT: out; per-item match count
    const u64 *item_off; // WRITE: per-item output offset
    u64 *out_cursor;     // FUSED: global reservation cursor (ends as the total match count)
    Pair *out;
    u64 capacity;
    int build_is_S;      // output is always (rowidR, rowidS): Result.cpp:66-69
    u64 *holes;          // POS: output slots reserved by position that stayed without a match (they hold kHolePair)
};

// POSITIONAL emit (FUSED mode, template flag POS).  When a partition's build side is one chunk of unique keys, a probe tuple
// has at most one match, so the item reserves ONE output slot per probe tuple up front -- a single global atomic per item,
// issued at item fetch and hidden behind the build -- and tuple i writes its pair to slot base + i: no ballots, no ranking,
// no CTA barrier and no reservation latency inside the probe loop.  A tuple without a match leaves kHolePair in its slot and
// is counted in *holes; the host closes the holes afterwards (k_holes_*), which costs nothing on foreign-key style joins
// where every probe tuple matches.  Items with duplicate build keys or several build chunks take the ranked path.
#define RHJ_HOLE 0xFFFFFFFFFFFFFFFFull

__device__ __forceinline__ u32 slot_of(u64 v) { return hash32(v) & (kSlots - 1); }

template <int MODE, bool POS = false>
__global__ void __launch_bounds__(kJoinThreads, RHJ_JOIN_MINBLOCKS) k_join(JoinArgs a) {
    // probe tuples per thread and round (more than four were measured slower in either emitter: 5 -> 1.90 ms, 6 -> 1.83 ms
    // against 1.61 ms for the positional kernel; registers / spills)
    constexpr int ITEMS = POS ? RHJ_JOIN_ITEMS_POS : kJoinItems;
    constexpr int ROUND = kJoinThreads * ITEMS;
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    Tup *s_tup = reinterpret_cast<Tup *>(dyn_smem);
    slot_t *s_slot = reinterpret_cast<slot_t *>(dyn_smem + (size_t) kBuildCap * sizeof(Tup));
    __shared__ __align__(8) u64 s_bar;
    __shared__ u32 s_item;
    __shared__ u32 s_cnt[2];
    __shared__ u64 s_base[2];
    __shared__ u64 s_red[32];
    __shared__ u64 s_posbase;
    u64 my_miss = 0;  // POS: reserved slots of this thread's probe tuples that found no match

    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 lt_mask = lanemask_lt();
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        s_cnt[0] = 0;
        s_cnt[1] = 0;
    }
    __syncthreads();
    const u32 nitems = *a.nitems;
    u32 phase = 0, rr = 0;

    while (true) {
        if (tid == 0) s_item = atomicAdd(a.work_counter, 1u);
        __syncthreads();
        const u32 item = s_item;
        if (item >= nitems) break;
        const Item it = a.items[item];
        const u64 b0 = a.offB[it.part], b1 = a.endB[it.part];
        const u64 p0 = a.offP[it.part] + (u64) it.chunk * kProbeChunk;
        const u64 p1 = min(a.endP[it.part], p0 + (u64) kProbeChunk);
        u64 my_count = 0;                                   // COUNT
        u64 run_base = (MODE == kJoinWrite && tid == 0) ? a.item_off[item] : 0;  // WRITE (thread 0 only)
        // POS: one slot per probe tuple, reserved now; the value is first looked at behind the build
        const bool single = POS && MODE == kJoinFused && b1 - b0 <= kBuildCap;
        u64 posres = 0;
        if (single && tid == 0) posres = atomicAdd(a.out_cursor, p1 - p0);

        for (u64 bb = b0; bb < b1; bb += kBuildCap) {
            const u32 nb = (u32) min((u64) kBuildCap, b1 - bb);
            if (tid == 0) {
                mbar_expect_tx(&s_bar, nb * (u32) sizeof(Tup));
                bulk_g2s(s_tup, a.build + bb, nb * (u32) sizeof(Tup), &s_bar);
            }
            Tup t[ITEMS];
            {
#pragma unroll
                for (int j = 0; j < ITEMS; ++j) {
                    u64 idx = p0 + (u64) j * kJoinThreads + tid;
                    if (idx < p1) t[j] = ld_stream(a.probe + idx);
                }
            }
            for (u32 i = tid; i < kSlots * sizeof(slot_t) / 16; i += kJoinThreads)
                reinterpret_cast<uint4 *>(s_slot)[i] = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
            mbar_wait(&s_bar, phase);
            phase ^= 1;
            __syncthreads();
            // build
            int dup = 0;
            for (u32 i = tid; i < nb; i += kJoinThreads) {
                const u64 v = s_tup[i].val;
                u32 h = slot_of(v);
                while (true) {
                    u32 old = atomicCAS(&s_slot[h], kSlotEmpty, (slot_t) i);
                    if (old == kSlotEmpty) break;
                    if (s_tup[old].val == v) dup = 1;
                    h = (h + 1) & (kSlots - 1);
                }
            }
            if (single && tid == 0) s_posbase = posres;
            dup = __syncthreads_or(dup);

            if (single && !dup) {
                // positional probe + emit: slot = reserved base + position of the tuple in the item's probe range
                const u64 base = s_posbase;
                for (u64 q0 = p0; q0 < p1; q0 += ROUND) {
#pragma unroll
                    for (int j = 0; j < ITEMS; ++j) {
                        const u64 idx = q0 + (u64) j * kJoinThreads + tid;
                        if (idx < p1 && q0 != p0) t[j] = ld_stream(a.probe + idx);
                    }
#pragma unroll
                    for (int j = 0; j < ITEMS; ++j) {
                        const u64 idx = q0 + (u64) j * kJoinThreads + tid;
                        if (idx < p1) {
                            u32 h = slot_of(t[j].val);
                            u32 c, hit = kEmpty;
                            while ((c = s_slot[h]) != kSlotEmpty) {
                                if (s_tup[c].val == t[j].val) { hit = c; break; }
                                h = (h + 1) & (kSlots - 1);
                            }
                            const u64 at = base + (idx - p0);
                            if (at < a.capacity) {
                                if (hit != kEmpty) {
                                    const u64 bk = s_tup[hit].key;
                                    if (a.build_is_S) st_stream(a.out + at, t[j].key, bk);
                                    else st_stream(a.out + at, bk, t[j].key);
                                } else {
                                    st_stream(a.out + at, RHJ_HOLE, RHJ_HOLE);
                                }
                            }
                            my_miss += hit == kEmpty;
                        }
                    }
                }
                __syncthreads();  // everyone is done with this table before it is overwritten
                continue;
            }
            if (single) {
                // duplicate build keys after all: the reserved slots stay holes, the ranked path below reserves its own
                const u64 base = s_posbase;
                for (u64 i = tid; i < p1 - p0; i += kJoinThreads)
                    if (base + i < a.capacity) st_stream(a.out + base + i, RHJ_HOLE, RHJ_HOLE);
                if (tid == 0) my_miss += p1 - p0;
            }

            // probe
            for (u64 q0 = p0; q0 < p1; q0 += ROUND) {
                bool ok[ITEMS];
#pragma unroll
                for (int j = 0; j < ITEMS; ++j) {
                    u64 idx = q0 + (u64) j * kJoinThreads + tid;
                    ok[j] = idx < p1;
                    if (ok[j] && q0 != p0) t[j] = ld_stream(a.probe + idx);
                }
                if (!dup) {
                    // unique build keys: at most one match per probe tuple
                    u32 m[ITEMS], ball[ITEMS];
                    u32 wtotal = 0;
#pragma unroll
                    for (int j = 0; j < ITEMS; ++j) {
                        m[j] = kEmpty;
                        if (ok[j]) {
                            u32 h = slot_of(t[j].val);
                            u32 idx;
                            while ((idx = s_slot[h]) != kSlotEmpty) {
                                if (s_tup[idx].val == t[j].val) { m[j] = idx; break; }
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
                    for (int j = 0; j < ITEMS; ++j) {
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
                    u32 cnt[ITEMS];
                    u32 mine = 0;
#pragma unroll
                    for (int j = 0; j < ITEMS; ++j) {
                        cnt[j] = 0;
                        if (ok[j]) {
                            u32 h = slot_of(t[j].val);
                            u32 idx;
                            while ((idx = s_slot[h]) != kSlotEmpty) {
                                if (s_tup[idx].val == t[j].val) cnt[j]++;
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
                    for (int j = 0; j < ITEMS; ++j) {
                        if (cnt[j]) {
                            u32 h = slot_of(t[j].val);
                            u32 idx;
                            while ((idx = s_slot[h]) != kSlotEmpty) {
                                if (s_tup[idx].val == t[j].val) {
                                    u64 bk = s_tup[idx].key;
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
            __syncthreads();  // everyone is done with this table before it is overwritten
        }
        if (MODE == kJoinCount) {
            u64 w = warp_sum64(my_count);
            if (lane == 0) s_red[warp] = w;
            __syncthreads();
            if (warp == 0) {
                u64 x = lane < (kJoinThreads / 32) ? s_red[lane] : 0;
                x = warp_sum64(x);
                if (lane == 0) a.item_cnt[item] = x;
            }
        }
    }
    if (POS) {
        const u64 w = warp_sum64(my_miss);
        if (lane == 0 && w) atomicAdd(a.holes, w);
    }
}


}  // namespace rhj
