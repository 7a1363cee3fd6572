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
// The fused emitter's default, k_join_pos (bottom of this file), takes the common case -- one table load, unique build keys --
// with instruction-level-parallel loops and leaves everything else to k_join<FUSED>.
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

// ---- k_join_pos: the positional emitter as a kernel of its own ---------------------------------------------------------
// k_join<FUSED, POS> above handles one tuple at a time: every build tuple pays the dependent chain value read -> hash ->
// CAS, every probe tuple hash -> slot read -> value read -> key read, with nothing else of the same warp to issue meanwhile
// (ncu of the r02 build: 17.7 cycles between two issues of a warp, 0.9 eligible warps per scheduler, issue slots 51 % busy),
// and the kernel carries the state of the ranked and duplicate-key paths through its hot loops (56 registers, the hash is
// re-derived and the shared-memory window base re-read, S2R SR_CgaCtaId, in front of every access).
// This kernel takes ONLY the common case -- a build partition that fits one table and has no duplicate keys -- and hands
// every other item to k_join<FUSED> through the list left[*nleft] (a second, usually empty launch).  That leaves registers for
// instruction-level parallelism:
//   * build and probe do each step for ITEMS (3) tuples of the thread back to back (hashes, slot reads / CAS, value
//     reads, key reads: independent, so their latencies overlap); only a tuple whose first slot held another key walks on;
//   * pk = slot index << 16 | build index found there (0xFFFF: none) carries both in one register, so the walk needs no
//     second hash;
//   * shared memory is addressed from one opaque base register (ld.shared / atom.shared by 32-bit address);
//   * the probe tuples of round r + 1 are loaded while round r is probed (three tuples per thread and round: a 2048-tuple
//     partition is two rounds at 89 % lane use instead of 1 1/3 rounds of four at 67 %);
//   * offsets inside an item are 32-bit; the output capacity is checked once per item, the build side once per round.
static_assert(kBuildCap < 0xFFFFu && kSlots <= 0x10000u, "k_join_pos packs (slot index, build index) into 16 + 16 bits");

__device__ __forceinline__ u32 smem_base_opaque(const void *p) {
    u32 a = smem_u32(p), b;
    asm volatile("mov.u32 %0, %1;" : "=r"(b) : "r"(a));
    return b;
}
// `volatile` + "memory": these stay behind the barrier / mbarrier wait they follow; independent ones still issue back to back
__device__ __forceinline__ u32 lds32(u32 addr) {
    u32 v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ u64 lds64(u32 addr) {
    u64 v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ u32 cas32_shared(u32 addr, u32 cmp, u32 val) {
    u32 old;
    asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(addr), "r"(cmp), "r"(val) : "memory");
    return old;
}

// Four tuples per thread and round without the prefetch (the second tuple set would spill) were measured slower: 1.77 ms
// against 1.55 ms (profiles/r02_call95_*).
// V, what the per-instruction counts of the first build showed (profiles/r02_ncu/k_join_pos3_source_sass.csv):
//   bit 0: nvcc re-derives the 12-instruction hash wherever pk's slot half is used (once more per probe tuple, and at the
//          head of every walk) instead of keeping it in a register; an empty asm makes the value opaque, so it is kept;
//   bit 1: thread 0 claims the NEXT work item while the current one is probed and publishes it in front of the barrier
//          that ends the item, which then also starts the next one: one CTA barrier and one exposed global round trip
//          (the claiming atomic) less per item.
template <int V>
__device__ __forceinline__ u32 slot_hi16(u64 v) {
    u32 h = slot_of(v) << 16;
    if (V & 1) asm volatile("" : "+r"(h));
    return h;
}
template <int ITEMS, int V>
__global__ void __launch_bounds__(kJoinThreads, RHJ_JOIN_MINBLOCKS) k_join_pos(JoinArgs a, Item *left, u32 *nleft) {
    constexpr u32 T = kJoinThreads, ROUND = T * ITEMS;
    extern __shared__ __align__(128) unsigned char dyn_smem[];   // [kBuildCap] staged build tuples | [kSlots] u32 slots
    __shared__ __align__(8) u64 s_bar;
    __shared__ u32 s_item;
    __shared__ u64 s_posbase;

    const u32 tid = threadIdx.x;
    const u32 sb = smem_base_opaque(dyn_smem);                 // shared address of the staged tuples ...
    const u32 sl = sb + kBuildCap * (u32) sizeof(Tup);         // ... and of the slot table
    u32 miss = 0;      // reserved slots of this thread's probe tuples that found no match (far fewer than 2^32 per launch)
    u64 miss_items = 0;  // thread 0: slots of whole items handed to the ranked kernel
    if (tid == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    const u32 nitems = *a.nitems;
    u32 phase = 0;
    u32 nxt = 0;   // V & 2, thread 0: the next item, claimed while the current one is probed
    if (V & 2) {
        if (tid == 0) s_item = atomicAdd(a.work_counter, 1u);
        __syncthreads();
    }

    while (true) {
        if (!(V & 2)) {
            if (tid == 0) s_item = atomicAdd(a.work_counter, 1u);
            __syncthreads();
        }
        const u32 item = s_item;
        if (item >= nitems) break;
        const Item it = a.items[item];
        const u64 b0 = a.offB[it.part], b1 = a.endB[it.part];
        const u64 p0 = a.offP[it.part] + (u64) it.chunk * kProbeChunk;
        const u64 p1 = min(a.endP[it.part], p0 + (u64) kProbeChunk);
        if (b1 - b0 > kBuildCap) {   // several build chunks (duplicate-heavy keys no radix bit can split)
            if (V & 2) {
                __syncthreads();     // everybody has read s_item
                if (tid == 0) {
                    left[atomicAdd(nleft, 1u)] = it;
                    s_item = atomicAdd(a.work_counter, 1u);
                }
                __syncthreads();
                continue;
            }
            if (tid == 0) left[atomicAdd(nleft, 1u)] = it;
            __syncthreads();         // nobody is still reading s_item when thread 0 overwrites it
            continue;
        }
        const u32 nb = (u32) (b1 - b0), n = (u32) (p1 - p0);   // n <= kProbeChunk
        u64 posres = 0;
        if (tid == 0) {
            // one output slot per probe tuple, reserved now; the value is first looked at behind the build
            posres = atomicAdd(a.out_cursor, (u64) n);
            mbar_expect_tx(&s_bar, nb * (u32) sizeof(Tup));
            bulk_g2s(dyn_smem, a.build + b0, nb * (u32) sizeof(Tup), &s_bar);
        }
        const Tup *const inp = a.probe + p0;
        Tup t[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const u32 o = j * T + tid;
            if (o < n) t[j] = ld_stream(inp + o);
        }
        {
            uint4 *slots = reinterpret_cast<uint4 *>(dyn_smem + (size_t) kBuildCap * sizeof(Tup));
            for (u32 i = tid; i < kSlots * sizeof(slot_t) / 16; i += T) slots[i] = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
        }
        mbar_wait(&s_bar, phase);
        phase ^= 1;
        __syncthreads();

        // ---- build: ITEMS claims in flight per thread; only the losers of the first CAS walk on ----
        int dup = 0;
        for (u32 i0 = tid; i0 < nb; i0 += ROUND) {
            u64 v[ITEMS];
            u32 pk[ITEMS];
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const u32 i = i0 + k * T;
                v[k] = i < nb ? lds64(sb + i * 16 + 8) : 0;
            }
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) pk[k] = slot_hi16<V>(v[k]);
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const u32 i = i0 + k * T;
                const u32 old = i < nb ? cas32_shared(sl + (pk[k] >> 14), kSlotEmpty, i) : kSlotEmpty;
                pk[k] |= old & 0xFFFFu;
            }
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                u32 o = pk[k] & 0xFFFFu, hs = pk[k] >> 16;
                while (o != 0xFFFFu) {   // the slot was taken: by an equal key (then this partition is the ranked kernel's), or not
                    if (lds64(sb + o * 16 + 8) == v[k]) dup = 1;
                    hs = (hs + 1) & (kSlots - 1);
                    o = cas32_shared(sl + (hs << 2), kSlotEmpty, i0 + k * T) & 0xFFFFu;
                }
            }
        }
        if (tid == 0) {
            s_posbase = posres;
            // everybody read s_item in front of the barrier behind the table load: the claim's latency hides behind the probe
            if (V & 2) nxt = atomicAdd(a.work_counter, 1u);
        }
        dup = __syncthreads_or(dup);
        const u64 base = s_posbase;
        // one capacity check per item: when an item's slots do not all fit, the cursor ends beyond the capacity and the host
        // redoes the join with the ranked emitter, so nothing of such an item needs to be written
        const bool fits = base + n <= a.capacity;
        Pair *const outp = a.out + base;
        if (dup) {
            // duplicate build keys: the reserved slots stay holes, the ranked kernel reserves its own
            if (fits)
                for (u32 i = tid; i < n; i += T) st_stream(outp + i, RHJ_HOLE, RHJ_HOLE);
            if (tid == 0) {
                miss_items += n;
                left[atomicAdd(nleft, 1u)] = it;
                if (V & 2) s_item = nxt;
            }
            __syncthreads();
            continue;
        }

        // ---- probe + positional emit: tuple i of the item writes slot base + i ----
        // One round: load the NEXT round's tuples into tn (in flight while this round is probed), probe the tuples in tc.
        // Two tuple sets that swap roles from round to round (no register copies).
        auto round = [&](Tup (&tc)[ITEMS], Tup (&tn)[ITEMS], const u32 r0) {
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) {
                const u32 o = r0 + ROUND + j * T + tid;
                if (o < n) tn[j] = ld_stream(inp + o);
            }
            const u32 o0 = r0 + tid;
            u32 pk[ITEMS];
            u64 bw[ITEMS];
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) pk[j] = slot_hi16<V>(tc[j].val);
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) pk[j] |= (o0 + j * T < n ? lds32(sl + (pk[j] >> 14)) : kSlotEmpty) & 0xFFFFu;
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) {
                const u32 c = pk[j] & 0xFFFFu;
                bw[j] = c != 0xFFFFu ? lds64(sb + c * 16 + 8) : 0;
            }
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) {
                // the low half becomes the matching build tuple or 0xFFFF; a first slot holding another key is the rare case
                u32 c = pk[j] & 0xFFFFu;
                if (c != 0xFFFFu && bw[j] != tc[j].val) {
                    u32 hs = pk[j] >> 16;
                    do {
                        hs = (hs + 1) & (kSlots - 1);
                        c = lds32(sl + (hs << 2)) & 0xFFFFu;
                    } while (c != 0xFFFFu && lds64(sb + c * 16 + 8) != tc[j].val);
                    pk[j] = c;
                }
            }
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) {
                const u32 c = pk[j] & 0xFFFFu;
                bw[j] = c != 0xFFFFu ? lds64(sb + c * 16) : RHJ_HOLE;
            }
            Pair *const q = outp + o0;
            if (fits) {
                if (a.build_is_S) {
#pragma unroll
                    for (int j = 0; j < ITEMS; ++j)
                        if (o0 + j * T < n) st_stream(q + j * T, (pk[j] & 0xFFFFu) != 0xFFFFu ? tc[j].key : RHJ_HOLE, bw[j]);
                } else {
#pragma unroll
                    for (int j = 0; j < ITEMS; ++j)
                        if (o0 + j * T < n) st_stream(q + j * T, bw[j], (pk[j] & 0xFFFFu) != 0xFFFFu ? tc[j].key : RHJ_HOLE);
                }
            }
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) miss += (o0 + j * T < n) && (pk[j] & 0xFFFFu) == 0xFFFFu;
        };
        Tup t2[ITEMS];
        for (u32 r0 = 0; r0 < n; r0 += 2 * ROUND) {
            round(t, t2, r0);
            if (r0 + ROUND >= n) break;
            round(t2, t, r0 + ROUND);
        }
        if ((V & 2) && tid == 0) s_item = nxt;
        __syncthreads();  // everyone is done with this table before it is overwritten (V & 2: ... and sees the next item)
    }
    const u64 w = warp_sum64((u64) miss + miss_items);
    if ((tid & 31) == 0 && w) atomicAdd(a.holes, w);
}


}  // namespace rhj
