// rhj_query_kernels.cuh -- kernels of the join's neighbours on the query path (included by
// rhj_query.cu only): filters (Query.cpp:94-146), row-id gathers (structs.cpp:217-243), the
// projection checksum (Query.cpp:66-74), the multiset digest of the parity checks, and
// update_intermediate (intermediate.cpp:52-183) as join + gather.
#pragma once
#include "rhj_device.cuh"

namespace rhj {

// ---- digest of a pair list (parity checks at sizes where sorting is too slow) ------------------
__global__ void __launch_bounds__(256) k_pairs_digest(const Pair *p, u64 n, u64 *sum, u64 *xr) {
    u64 s = 0, x = 0;
    for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64) gridDim.x * blockDim.x) {
        Pair q = p[i];
        u64 h = mix64(q.r * 0x100000001b3ULL + q.s);
        s += h;
        x ^= h;
    }
    s = warp_sum64(s);
    x = warp_xor64(x);
    if (lane_id() == 0) {
        atomicAdd(sum, s);
        atomicXor(xr, x);
    }
}

// ---- filters and gathers (Query.cpp:94-146, structs.cpp:217-226, Query.cpp:66-74) --------------
constexpr int kFiltThreads = 256;
constexpr int kFiltItems = 8;
constexpr int kFiltTile = kFiltThreads * kFiltItems;

__device__ __forceinline__ bool pred(u64 v, int op, u64 c) {
    return op == '>' ? v > c : op == '<' ? v < c : v == c;
}

// pass 1: survivors per tile
__global__ void __launch_bounds__(kFiltThreads) k_filter_count(const u64 *col, const u64 *rowids, u64 n, int op, u64 c,
                                                               u32 *tile_cnt) {
    __shared__ u32 s_c;
    if (threadIdx.x == 0) s_c = 0;
    __syncthreads();
    u64 base = (u64) blockIdx.x * kFiltTile;
    u32 mine = 0;
#pragma unroll
    for (int j = 0; j < kFiltItems; ++j) {
        u64 i = base + (u64) threadIdx.x * kFiltItems + j;
        if (i < n) {
            u64 row = rowids ? rowids[i] : i;
            mine += pred(col[row], op, c);
        }
    }
    u32 w = (u32) warp_sum64(mine);
    if (lane_id() == 0 && w) atomicAdd(&s_c, w);
    __syncthreads();
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = s_c;
}

__global__ void __launch_bounds__(1024) k_scan_tiles(const u32 *cnt, u32 n, u64 *off, u64 *total) {
    __shared__ u64 s_w[32];
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

// pass 2: order-preserving compaction (thread t owns 8 consecutive rows -> ascending output)
__global__ void __launch_bounds__(kFiltThreads) k_filter_write(const u64 *col, const u64 *rowids, u64 n, int op, u64 c,
                                                               const u64 *tile_off, u64 *out) {
    __shared__ u32 s_w[kFiltThreads / 32];
    u64 base = (u64) blockIdx.x * kFiltTile;
    u64 row[kFiltItems];
    bool keep[kFiltItems];
    u32 mine = 0;
#pragma unroll
    for (int j = 0; j < kFiltItems; ++j) {
        u64 i = base + (u64) threadIdx.x * kFiltItems + j;
        keep[j] = false;
        if (i < n) {
            row[j] = rowids ? rowids[i] : i;
            keep[j] = pred(col[row[j]], op, c);
        }
        mine += keep[j];
    }
    u32 incl = warp_incl_scan(mine);
    if (lane_id() == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    u32 before = 0;
    for (u32 w = 0; w < (threadIdx.x >> 5); ++w) before += s_w[w];
    u64 at = tile_off[blockIdx.x] + before + (incl - mine);
#pragma unroll
    for (int j = 0; j < kFiltItems; ++j)
        if (keep[j]) out[at++] = row[j];
}

__global__ void __launch_bounds__(256) k_gather_tuples(const u64 *col, const u64 *rowids, u64 n, Tup *out) {
    for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64) gridDim.x * blockDim.x) {
        u64 row = rowids[i];
        Tup t;
        t.key = row;
        t.val = col[row];
        out[i] = t;
    }
}

__global__ void __launch_bounds__(256) k_gather_sum(const u64 *col, const u64 *rowids, u64 n, u64 *sum) {
    u64 s = 0;
    for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64) gridDim.x * blockDim.x)
        s += col[rowids[i]];
    s = warp_sum64(s);
    if (lane_id() == 0 && s) atomicAdd(sum, s);
}

// ---- update_intermediate as join + gather (intermediate.cpp:52-183) ---------------------------------
// The reference scans the whole intermediate once per result pair (O(pairs x rows), 99 % of the
// small.work wall time).  Case 2 is an N:M equi-join between the result pairs and the existing
// row-id column, case 3 a semi-join on the (row id, row id) pair: both are run with the join
// kernels above on index-keyed relations, followed by coalesced gathers.

// A[e] = {e, col[e]}: the existing row-id column as a relation keyed by its row index
__global__ void __launch_bounds__(256) k_index_tuples(const u64 *col, u64 n, Tup *out) {
    for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64) gridDim.x * blockDim.x) {
        Tup t;
        t.key = i;
        t.val = col[i];
        out[i] = t;
    }
}
// B[p] = {p, pairs[p].keyS or .keyR}: one side of the join result keyed by the pair index
__global__ void __launch_bounds__(256) k_pair_side_tuples(const Pair *pairs, u64 n, int take_s, Tup *out) {
    for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64) gridDim.x * blockDim.x) {
        Pair q = pairs[i];
        Tup t;
        t.key = i;
        t.val = take_s ? q.s : q.r;
        out[i] = t;
    }
}
// composite keys (a << 32 | b) for case 3; *overflow is set if a row id does not fit 32 bits
__global__ void __launch_bounds__(256) k_index_tuples2(const u64 *c1, const u64 *c2, u64 n, Tup *out, u32 *overflow) {
    for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64) gridDim.x * blockDim.x) {
        u64 a = c1[i], b = c2[i];
        if ((a | b) >> 32) *overflow = 1;
        Tup t;
        t.key = i;
        t.val = (a << 32) | (b & 0xffffffffull);
        out[i] = t;
    }
}
__global__ void __launch_bounds__(256) k_pair_both_tuples(const Pair *pairs, u64 n, Tup *out, u32 *overflow) {
    for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64) gridDim.x * blockDim.x) {
        Pair q = pairs[i];
        if ((q.r | q.s) >> 32) *overflow = 1;
        Tup t;
        t.key = i;
        t.val = (q.r << 32) | (q.s & 0xffffffffull);
        out[i] = t;
    }
}
// out[i] = col[ep[i].r]: carry a column of the old intermediate to the new rows
__global__ void __launch_bounds__(256) k_gather_by_elem(const u64 *col, const Pair *ep, u64 m, u64 *out) {
    for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (u64) gridDim.x * blockDim.x)
        out[i] = col[ep[i].r];
}
// out[i] = the other side of the matched result pair: the new binding's row id
__global__ void __launch_bounds__(256) k_gather_pair_value(const Pair *pairs, const Pair *ep, u64 m, int take_r, u64 *out) {
    for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (u64) gridDim.x * blockDim.x) {
        Pair q = pairs[ep[i].s];
        out[i] = take_r ? q.r : q.s;
    }
}
// case 3 fallback for row ids >= 2^32: candidates matched on the first id, verified on the second
__global__ void __launch_bounds__(kFiltThreads) k_ep_verify_count(const Pair *ep, u64 m, const u64 *c2, const Pair *pairs,
                                                                  u32 *tile_cnt) {
    __shared__ u32 s_c;
    if (threadIdx.x == 0) s_c = 0;
    __syncthreads();
    u64 base = (u64) blockIdx.x * kFiltTile;
    u32 mine = 0;
#pragma unroll
    for (int j = 0; j < kFiltItems; ++j) {
        u64 i = base + (u64) threadIdx.x * kFiltItems + j;
        if (i < m) mine += c2[ep[i].r] == pairs[ep[i].s].s;
    }
    u32 w = (u32) warp_sum64(mine);
    if (lane_id() == 0 && w) atomicAdd(&s_c, w);
    __syncthreads();
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = s_c;
}
__global__ void __launch_bounds__(kFiltThreads) k_ep_verify_write(const Pair *ep, u64 m, const u64 *c2, const Pair *pairs,
                                                                  const u64 *tile_off, Pair *out) {
    __shared__ u32 s_w[kFiltThreads / 32];
    u64 base = (u64) blockIdx.x * kFiltTile;
    Pair q[kFiltItems];
    bool keep[kFiltItems];
    u32 mine = 0;
#pragma unroll
    for (int j = 0; j < kFiltItems; ++j) {
        u64 i = base + (u64) threadIdx.x * kFiltItems + j;
        keep[j] = false;
        if (i < m) {
            q[j] = ep[i];
            keep[j] = c2[q[j].r] == pairs[q[j].s].s;
        }
        mine += keep[j];
    }
    u32 incl = warp_incl_scan(mine);
    if (lane_id() == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    u32 before = 0;
    for (u32 w = 0; w < (threadIdx.x >> 5); ++w) before += s_w[w];
    u64 at = tile_off[blockIdx.x] + before + (incl - mine);
#pragma unroll
    for (int j = 0; j < kFiltItems; ++j)
        if (keep[j]) out[at++] = q[j];
}

}  // namespace rhj
