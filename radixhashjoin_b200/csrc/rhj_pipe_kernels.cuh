// rhj_pipe_kernels.cuh -- device side of the pipelined multi-GPU exchange (rhj_pipe_* in include/rhj.h).
//
// No reference equivalent (the reference is one process, join.cpp:42-50); SURVEY.md 8e.  Every rank
// cuts each local relation into C row chunks.  For every chunk
//   pass 1   k_scatter<kDigitShard, LIMIT> partitions the rows on (destination rank | sub-digit) WITHOUT a
//            histogram: every (destination, sub-digit) digit owns a fixed-capacity region of the chunk's
//            staging area (the region of the rank's own digits lies directly in its receive buffer);
//   ship     k_pipe_ship, a small persistent copy kernel next to the partitioning kernels: one thread per
//            CTA drives a shared-memory ring with TMA bulk copies (cp.async.bulk global -> shared ->
//            PEER global over NVLink / NVSwitch, SASS UBLKCP both ways), copies only the filled part of
//            every region into the same region of the destination's receive buffer, stores the region's
//            end there too, and the last CTA releases one flag per destination (st.release.sys);
//   pass 2   on the destination, k_pipe_arrive waits for the chunk's flags of all sources (ld.acquire.sys,
//            bounded spin) and turns the received region ends into a segment table; the segmented
//            k_scatter<kDigitHash, LIMIT> appends the chunk to the fixed-capacity final partitions.
// No histogram, no all-gather, no host synchronisation and no NCCL call inside a step; receive buffers,
// region ends and flags are double-buffered by step parity, which makes acknowledgements unnecessary
// (see DESIGN.md 6 for the argument).  An overflowing region anywhere is reported to every rank through
// the flags / status words so that all ranks redo the step through the exact (histogram) exchange.
#pragma once
#include "rhj_device.cuh"
#include "rhj_kernels.cuh"

namespace rhj {

constexpr int kPipeMaxChunks = 8;
constexpr u32 kPipeShipThreads = 64;
constexpr u32 kPipeMaxStages = 32;                // ring = stages x stage_bytes of dynamic shared memory, chosen at rhj_pipe_open:
                                                  // default 8 x 8 KiB = 64 KiB per CTA, which fits next to two k_scatter CTAs on an SM
constexpr u64 kPipeSpinNs = 4000000000ull;        // a wait gives up after 4 s and reports RHJ_PIPE_TIMEOUT

// status bits (low 8 bits of a flag / status word; the rest is the epoch)
constexpr u64 kPipeOvf = 1;      // a fixed-capacity region overflowed: redo the step through the exact exchange
constexpr u64 kPipeTimeout = 2;  // a flag did not arrive in time
constexpr u64 kPipeBad = 4;      // a received region end was out of range
constexpr u64 kPipeWide = 8;     // 12-byte wire format: a row id did not fit 32 bits
constexpr u32 kPipeShip12Threads = 128;   // the repacking copy kernel: 4 tuples (64 B -> 48 B) per thread and block of 512 tuples
constexpr u32 kPipe12Block = 512;         // tuples per ring stage of the repacking copy kernel (8 KiB in, 6 KiB out)
constexpr u32 kPipe12Stages = 4;          // 4 x (8 + 6) KiB = 56 KiB of shared memory per CTA

__device__ __forceinline__ void st_release_sys(u64 *p, u64 v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u64 ld_acquire_sys(const u64 *p) {
    u64 v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ u64 global_ns() {
    u64 t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void bulk_wait_read(int pending) {
    if (pending == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Pass-1 cursors of all chunks of both relations: digit g of a chunk starts at g * cap1 of the chunk's area.
struct PipeBeginArgs {
    u64 *cursor[2];   // [chunks][kMaxDigits]
    u64 cap1[2];
    u32 chunks, ndig;
};
__global__ void k_pipe_begin(PipeBeginArgs a) {
    const int rel = blockIdx.y;
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.chunks * a.ndig) a.cursor[rel][(u64) (i / a.ndig) * kMaxDigits + i % a.ndig] = (u64) (i % a.ndig) * a.cap1[rel];
}

// The copy kernel of one (relation, chunk).
struct PipeShipArgs {
    const Tup *stage;            // the chunk's staging area: region g at g * cap1
    const u64 *cursor;           // [ndig] pass-1 cursors of the chunk (region g holds cursor[g] - g * cap1 tuples)
    u64 cap1;
    u32 world, rank, ndig;
    int sub_bits;
    Tup *peer_recv[kMaxPeers];   // every rank's receive buffer of this (relation, parity)
    u64 *peer_end[kMaxPeers];    // every rank's region-end table of this (relation, parity)
    u64 *peer_flag[kMaxPeers];   // every rank's flag word of this (parity, relation, chunk), indexed [source]
    u64 region0;                 // index of region (chunk, source = rank, p1 = 0) inside a receive buffer
    u64 epoch;
    u32 *done;                   // CTA arrival counter (returns to 0)
    u32 *overflow;               // the local overflow flag (set by pass 1 or here)
    u32 stages, stage_bytes;     // shared-memory ring
};

__global__ void __launch_bounds__(kPipeShipThreads) k_pipe_ship(PipeShipArgs a) {
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ __align__(8) u64 s_full[kPipeMaxStages];
    __shared__ u32 s_last;
    const u32 tid = threadIdx.x;
    const u32 nd1 = 1u << a.sub_bits;
    if (tid == 0)
        for (u32 s = 0; s < a.stages; ++s) mbar_init(&s_full[s], 1);
    __syncthreads();

    // region ends, one thread per region of this CTA's share: region (dest, p1) of this chunk lands in the
    // destination's receive buffer at (region0 + p1) * cap1, the same capacity as here
    for (u32 g = blockIdx.x * blockDim.x + tid; g < a.ndig; g += gridDim.x * blockDim.x) {
        const u32 dest = g >> a.sub_bits, p1 = g & (nd1 - 1);
        u64 cnt = a.cursor[g] - (u64) g * a.cap1;
        if (cnt > a.cap1) {
            cnt = a.cap1;
            *a.overflow = 1;
        }
        a.peer_end[dest][a.region0 + p1] = (a.region0 + p1) * a.cap1 + cnt;
        __threadfence_system();
    }

    if (tid == 0) {
        // this CTA's regions: every gridDim.x-th REMOTE region
        const u32 nremote = (a.world - 1) << a.sub_bits;
        // block sequence of the CTA, produced on the fly by two cursors (load side, store side)
        struct Cur {
            u32 k;      // index into the CTA's region list
            u64 off;    // tuples of the current region already handled
            u64 cnt;    // tuples in the current region
            u32 g;      // digit of the current region
        };
        auto region_of = [&](u32 k, u32 &g, u64 &cnt) -> bool {
            const u32 r = blockIdx.x + k * gridDim.x;
            if (r >= nremote) return false;
            // consecutive regions go to consecutive destinations, starting behind the own rank: at any moment the
            // CTAs of a rank feed all its peers, and no two ranks start on the same destination (no incast)
            const u32 dest = (a.rank + 1 + r % (a.world - 1)) % a.world;
            g = (dest << a.sub_bits) | (r / (a.world - 1));
            cnt = min(a.cursor[g] - (u64) g * a.cap1, a.cap1);
            return true;
        };
        auto next_block = [&](Cur &c, u32 &g, u64 &off, u32 &n) -> bool {
            while (true) {
                if (c.off < c.cnt) {
                    g = c.g;
                    off = c.off;
                    n = (u32) min((u64) (a.stage_bytes / sizeof(Tup)), c.cnt - c.off);
                    c.off += n;
                    return true;
                }
                if (!region_of(c.k, c.g, c.cnt)) return false;
                c.k++;
                c.off = 0;
            }
        };
        Cur ld{0, 0, 0, 0}, stc{0, 0, 0, 0};
        u32 issued = 0, stored = 0;
        u32 g, n;
        u64 off;
        // prime the ring
        while (issued < a.stages - 1 && next_block(ld, g, off, n)) {
            const u32 s = issued % a.stages;
            mbar_expect_tx(&s_full[s], n * (u32) sizeof(Tup));
            bulk_g2s(ring + (size_t) s * a.stage_bytes, a.stage + (u64) g * a.cap1 + off, n * (u32) sizeof(Tup), &s_full[s]);
            ++issued;
        }
        while (next_block(stc, g, off, n)) {
            const u32 s = stored % a.stages;
            mbar_wait(&s_full[s], (stored / a.stages) & 1);
            const u32 dest = g >> a.sub_bits, p1 = g & (nd1 - 1);
            bulk_s2g(a.peer_recv[dest] + (a.region0 + p1) * a.cap1 + off, ring + (size_t) s * a.stage_bytes, n * (u32) sizeof(Tup));
            bulk_commit();
            ++stored;
            // refill the stage the PREVIOUS store read from (its read-out is done once at most one group is pending)
            u32 g2, n2;
            u64 off2;
            if (next_block(ld, g2, off2, n2)) {
                bulk_wait_read(1);
                const u32 s2 = issued % a.stages;
                mbar_expect_tx(&s_full[s2], n2 * (u32) sizeof(Tup));
                bulk_g2s(ring + (size_t) s2 * a.stage_bytes, a.stage + (u64) g2 * a.cap1 + off2, n2 * (u32) sizeof(Tup), &s_full[s2]);
                ++issued;
            }
        }
        bulk_wait_all();  // every bulk store of this CTA has been performed
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();
        s_last = atomicAdd(a.done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        const u64 v = (a.epoch << 8) | (*a.overflow ? kPipeOvf : 0);
        if (tid < a.world) st_release_sys(a.peer_flag[tid] + a.rank, v);
        if (tid == 0) *a.done = 0;
    }
}

// The copy kernel of one (relation, chunk) for the 12-BYTE WIRE FORMAT: same job as k_pipe_ship, but every block of 512
// tuples is repacked on its way through shared memory from 16-byte {row id, value} tuples to 12-byte {value, u32 row id}
// records (4 tuples = 64 B -> 48 B = three 16-byte words), so a quarter fewer bytes cross NVLink and a quarter fewer land
// in the destination's HBM.  Thread 0 drives the TMA bulk copies (global -> in-ring, out-ring -> peer global), all 128
// threads repack.  The rank's own regions take the same route (a local repacking copy), so that a receive buffer holds
// one format.  A row id that does not fit 32 bits sets kPipeWide.
__global__ void __launch_bounds__(kPipeShip12Threads) k_pipe_ship12(PipeShipArgs a) {
    extern __shared__ __align__(128) unsigned char ring[];
    unsigned char *in_ring = ring;
    unsigned char *out_ring = ring + (size_t) kPipe12Stages * kPipe12Block * 16;
    __shared__ __align__(8) u64 s_full[kPipe12Stages];
    __shared__ u32 s_g[kPipe12Stages], s_n[kPipe12Stages];
    __shared__ u64 s_off[kPipe12Stages];
    __shared__ u32 s_last, s_wide;
    const u32 tid = threadIdx.x;
    const u32 nd1 = 1u << a.sub_bits;
    if (tid == 0) {
        for (u32 s = 0; s < kPipe12Stages; ++s) mbar_init(&s_full[s], 1);
        s_wide = 0;
    }
    __syncthreads();
    for (u32 g = blockIdx.x * blockDim.x + tid; g < a.ndig; g += gridDim.x * blockDim.x) {
        const u32 dest = g >> a.sub_bits, p1 = g & (nd1 - 1);
        u64 cnt = a.cursor[g] - (u64) g * a.cap1;
        if (cnt > a.cap1) {
            cnt = a.cap1;
            *a.overflow = 1;
        }
        a.peer_end[dest][a.region0 + p1] = (a.region0 + p1) * a.cap1 + cnt;
        __threadfence_system();
    }
    // thread 0's view of the CTA's block sequence: every gridDim.x-th region (ALL destinations, the own rank included),
    // consecutive regions going to consecutive destinations starting behind the own rank
    struct Cur {
        u32 k;
        u64 off, cnt;
        u32 g;
    };
    Cur ld{0, 0, 0, 0};
    auto next_block = [&](u32 &g, u64 &off, u32 &n) -> bool {
        while (true) {
            if (ld.off < ld.cnt) {
                g = ld.g;
                off = ld.off;
                n = (u32) min((u64) kPipe12Block, ld.cnt - ld.off);
                ld.off += n;
                return true;
            }
            const u32 r = blockIdx.x + ld.k * gridDim.x;
            if (r >= a.ndig) return false;
            const u32 dest = (a.rank + 1 + r % a.world) % a.world;
            ld.g = (dest << a.sub_bits) | (r / a.world);
            ld.cnt = min(a.cursor[ld.g] - (u64) ld.g * a.cap1, a.cap1);
            ld.k++;
            ld.off = 0;
        }
    };
    auto issue = [&](u32 s) {  // thread 0: the next block into in-stage s, or the end-of-stream sentinel
        u32 g, n;
        u64 off;
        if (next_block(g, off, n)) {
            s_g[s] = g;
            s_n[s] = n;
            s_off[s] = off;
            const u32 bytes = ((n + 3) & ~3u) * 16;  // whole groups of 4 tuples; the tail reads a few tuples of slack
            mbar_expect_tx(&s_full[s], bytes);
            bulk_g2s(in_ring + (size_t) s * kPipe12Block * 16, a.stage + (u64) g * a.cap1 + off, bytes, &s_full[s]);
        } else {
            s_n[s] = 0;
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_full[s])) : "memory");
        }
    };
    if (tid == 0)
        for (u32 s = 0; s < kPipe12Stages; ++s) issue(s);
    u32 wide = 0;
    for (u32 j = 0;; ++j) {
        const u32 s = j % kPipe12Stages;
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPipe12Stages - 1) : "memory");  // out-stage s is free
        __syncthreads();
        mbar_wait(&s_full[s], (j / kPipe12Stages) & 1);
        const u32 n = s_n[s];
        if (n == 0) break;
        const u32 g = s_g[s];
        const u64 off = s_off[s];
        const uint4 *in = reinterpret_cast<const uint4 *>(in_ring + (size_t) s * kPipe12Block * 16);
        uint4 *out = reinterpret_cast<uint4 *>(out_ring + (size_t) s * kPipe12Block * 12);
        const u32 groups = (n + 3) / 4;
        if (tid < groups) {
            const uint4 t0 = in[4 * tid], t1 = in[4 * tid + 1], t2 = in[4 * tid + 2], t3 = in[4 * tid + 3];  // {key.lo, key.hi, val.lo, val.hi}
            const u32 left = n - 4 * tid;
            if (t0.y | (left > 1 ? t1.y : 0) | (left > 2 ? t2.y : 0) | (left > 3 ? t3.y : 0)) wide = 1;
            out[3 * tid] = make_uint4(t0.z, t0.w, t0.x, t1.z);
            out[3 * tid + 1] = make_uint4(t1.w, t1.x, t2.z, t2.w);
            out[3 * tid + 2] = make_uint4(t2.x, t3.z, t3.w, t3.x);
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            const u32 dest = g >> a.sub_bits, p1 = g & (nd1 - 1);
            unsigned char *dst = reinterpret_cast<unsigned char *>(a.peer_recv[dest]) + ((a.region0 + p1) * a.cap1 + off) * 12;
            bulk_s2g(dst, out, groups * 48);
            bulk_commit();
            issue(s);  // in-stage s has been consumed
        }
    }
    if (wide) s_wide = 1;
    if (tid == 0) bulk_wait_all();
    __syncthreads();
    if (tid == 0) {
        if (s_wide) atomicOr(a.overflow, 2u);
        __threadfence_system();
        s_last = atomicAdd(a.done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        const u32 o = *a.overflow;
        const u64 v = (a.epoch << 8) | (o & 1 ? kPipeOvf : 0) | (o & 2 ? kPipeWide : 0);
        if (tid < a.world) st_release_sys(a.peer_flag[tid] + a.rank, v);
        if (tid == 0) *a.done = 0;
    }
}

// Destination side of one (relation, chunk): wait for the flags of all sources, then build the segment table
// of what arrived.  Segment s = source * nd1 + p1 is region (chunk, source, p1) of the receive buffer.
struct PipeArriveArgs {
    const u64 *flag;       // [world] this rank's flags of (parity, relation, chunk)
    const u64 *region_end; // [nseg] region ends of the chunk, written by the sources
    u64 region0;           // index of the chunk's first region
    u64 cap1;
    u32 world, nseg;
    u64 epoch;
    u64 *seg_off;          // [nseg + 1]
    u64 *seg_end;          // [nseg]
    u32 *seg_tile0;        // [nseg + 1]
    TileDesc *tiles;       // [ntiles] one descriptor per pass-2 tile (what k_tile_table builds on the other paths)
    u32 ntiles;            // host-side bound; entries behind the last tile get len 0
    u64 *status;           // local status word (kPipe* bits are OR-ed in)
};
__global__ void __launch_bounds__(1024) k_pipe_arrive(PipeArriveArgs a) {
    __shared__ u64 s_w[33];
    __shared__ u32 s_bad;
    const u32 tid = threadIdx.x;
    if (tid == 0) s_bad = 0;
    __syncthreads();
    if (tid < a.world) {
        const u64 t0 = global_ns();
        u64 v;
        while (((v = ld_acquire_sys(a.flag + tid)) >> 8) < a.epoch) {
            if (global_ns() - t0 > kPipeSpinNs) {
                atomicOr((unsigned long long *) a.status, kPipeTimeout);
                s_bad = 1;
                break;
            }
            __nanosleep(200);
        }
        if (v & 0xff) atomicOr((unsigned long long *) a.status, v & 0xff);
    }
    __syncthreads();
    u64 beg = 0, cnt = 0;
    if (tid < a.nseg) {
        beg = (a.region0 + tid) * a.cap1;
        const u64 e = s_bad ? beg : a.region_end[tid];
        if (e < beg || e > beg + a.cap1) atomicOr((unsigned long long *) a.status, kPipeBad);
        else cnt = e - beg;
    }
    u64 ttotal;
    const u64 tl = (cnt + kTile - 1) / kTile;
    const u64 tex = block_excl_scan64(tl, s_w, &ttotal);
    if (tid < a.nseg) {
        a.seg_off[tid] = beg;
        a.seg_end[tid] = beg + cnt;
        a.seg_tile0[tid] = (u32) tex;
        if (tid == a.nseg - 1) {
            a.seg_off[a.nseg] = beg + cnt;
            a.seg_tile0[a.nseg] = (u32) ttotal;
        }
        for (u32 t = 0; t < (u32) tl; ++t) {
            const u64 tb = beg + (u64) t * kTile;
            a.tiles[tex + t] = TileDesc{tb, (u32) min((u64) kTile, beg + cnt - tb), tid};
        }
    }
    for (u32 t = (u32) ttotal + tid; t < a.ntiles; t += blockDim.x) a.tiles[t] = TileDesc{0, 0, 0};
}

// After the last pass 2: tell every rank whether anything overflowed here (sender or receiver side).
struct PipePostArgs {
    u64 *peer_status[kMaxPeers];  // every rank's status table of this parity, indexed [source]
    const u32 *overflow;
    const u64 *status;
    u32 world, rank;
    u64 epoch;
};
__global__ void k_pipe_post(PipePostArgs a) {
    const u32 tid = threadIdx.x;
    if (tid < a.world) {
        const u32 o = *a.overflow;
        const u64 bits = (o & 1 ? kPipeOvf : 0) | (o & 2 ? kPipeWide : 0) | (*a.status & 0xff);
        __threadfence_system();
        st_release_sys(a.peer_status[tid] + a.rank, (a.epoch << 8) | bits);
    }
}
// After the join: collect every rank's verdict, so that all ranks take the same decision.
struct PipeCollectArgs {
    const u64 *status_in;  // [world]
    u64 *status;           // local status word
    u32 world;
    u64 epoch;
};
__global__ void k_pipe_collect(PipeCollectArgs a) {
    const u32 tid = threadIdx.x;
    if (tid < a.world) {
        const u64 t0 = global_ns();
        u64 v;
        while (((v = ld_acquire_sys(a.status_in + tid)) >> 8) < a.epoch) {
            if (global_ns() - t0 > kPipeSpinNs) {
                v = kPipeTimeout;
                break;
            }
            __nanosleep(200);
        }
        if (v & 0xff) atomicOr((unsigned long long *) a.status, v & 0xff);
    }
}

}  // namespace rhj
