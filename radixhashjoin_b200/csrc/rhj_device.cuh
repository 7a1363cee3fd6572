// rhj_device.cuh -- device-side primitives shared by the sm_100a kernels.
//
// Everything here is integer / byte work bounded by HBM bandwidth; there is no tensor-core
// work on this path.  Blackwell features used: TMA bulk copies (cp.async.bulk, SASS UBLKCP)
// with mbarrier completion for global->shared staging of build partitions and shared->global
// write-out of partition runs.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rhj {

typedef unsigned long long u64;
typedef unsigned int u32;

// == rhj_tuple == reference `tuple` (structs.h:33-36): key = row id, val = join value
struct __align__(16) Tup {
    u64 key;
    u64 val;
};
// == rhj_pair == reference `key_tuple` (Result.h:9-12)
struct __align__(16) Pair {
    u64 r;
    u64 s;
};

constexpr u32 kEmpty = 0xFFFFFFFFu;

// Mixed hash of the join value (two multiply-xorshift rounds).  hash32 = low word: partition id =
// its TOP bits, shared-memory table slot = its LOW bits.  hash_hi32 = high word: destination rank of
// the multi-GPU shuffle (independent bits, so a rank's local partitions stay balanced).  The
// reference buckets on the raw low byte (JobScheduler.cpp:151); any function of the value gives
// the same join result, and mixed bits keep the low-entropy contest columns balanced.
// (A one-multiply Fibonacci hash was measured: no kernel got faster, the join got 4 % slower --
// profiles/r01_tuning_notes.md.)
__device__ __forceinline__ u64 hash64(u64 v) {
    v ^= v >> 32;
    v *= 0xd6e8feb86659fd93ULL;
    v ^= v >> 32;
    v *= 0xd6e8feb86659fd93ULL;
    v ^= v >> 32;
    return v;
}
__device__ __forceinline__ u32 hash32(u64 v) { return (u32) hash64(v); }
__device__ __forceinline__ u32 hash_hi32(u64 v) { return (u32) (hash64(v) >> 32); }

// splitmix64 finalizer -- digest only (same function as orc_mix64)
__device__ __forceinline__ u64 mix64(u64 x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    return x;
}

// ---- streaming 16-byte global accesses (read-once inputs, write-once outputs) ---------------
__device__ __forceinline__ Tup ld_stream(const Tup *p) {
    Tup t;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(t.key), "=l"(t.val) : "l"(p));
    return t;
}
__device__ __forceinline__ void st_stream(Tup *p, const Tup &t) {
    asm volatile("st.global.L1::no_allocate.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(t.key), "l"(t.val) : "memory");
}
__device__ __forceinline__ void st_stream(Pair *p, u64 r, u64 s) {
    asm volatile("st.global.L1::no_allocate.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(r), "l"(s) : "memory");
}

// ---- mbarrier + TMA bulk copy (cp.async.bulk) -------------------------------------------------
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy, completion signalled on `bar` (bytes % 16 == 0, both 16-B aligned)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, u32 bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy (bulk-group completion)
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, u32 bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (before bulk_s2g)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- warp / block helpers ---------------------------------------------------------------------
__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ u32 lanemask_lt() {
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ u32 warp_incl_scan(u32 v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane_id() >= (u32) o) v += t;
    }
    return v;
}
__device__ __forceinline__ u64 warp_incl_scan64(u64 v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u64 t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane_id() >= (u32) o) v += t;
    }
    return v;
}
__device__ __forceinline__ u64 warp_sum64(u64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ u64 warp_xor64(u64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v ^= __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace rhj
