// rhj_exec_kernels.cuh -- kernels of the device-resident query executor (rhj_exec.cu): everything around the join
// that the reference does with unordered_sets and vectors on the host (Query.cpp:81-158 run_filters, structs.cpp:217-243
// create_relation / foo incl. the row-id de-duplication, intermediate.cpp:11-183 parse_table / update_intermediate,
// Query.cpp:66-74 column_proj), on row-id lists and intermediate columns that never leave HBM.
#pragma once
#include "rhj_device.cuh"
#include "rhj_query_kernels.cuh"

namespace rhj {

// One order-preserving selection primitive (count per tile -> k_scan_tiles -> write) for the three row predicates of the
// query path.  Element i of the input is
//   kSelConst   row = idx ? idx[i] : i;   kept when colA[row] <op> c          -> emits row        (run_filters)
//   kSelSameRow row = idx ? idx[i] : i;   kept when colA[row] == colB[row]     -> emits row        (parse_table, first branch)
//   kSelTwoCols kept when colA[idx[i]] == colB[idxB[i]]                        -> emits i          (update_intermediate case 3,
//                                                                                 parse_table second branch: a row filter)
enum SelKind { kSelConst = 0, kSelSameRow = 1, kSelTwoCols = 2 };
struct SelArgs {
    int kind;
    const u64 *colA, *colB;
    const u64 *idx, *idxB;
    u64 n;
    int op;
    u64 c;
};
__device__ __forceinline__ bool sel_keep(const SelArgs &a, u64 i, u64 &emit) {
    if (a.kind == kSelTwoCols) {
        emit = i;
        return a.colA[a.idx[i]] == a.colB[a.idxB[i]];
    }
    const u64 row = a.idx ? a.idx[i] : i;
    emit = row;
    if (a.kind == kSelSameRow) return a.colA[row] == a.colB[row];
    return pred(a.colA[row], a.op, a.c);
}
__global__ void __launch_bounds__(kFiltThreads) k_select_count(SelArgs a, u32 *tile_cnt) {
    __shared__ u32 s_c;
    if (threadIdx.x == 0) s_c = 0;
    __syncthreads();
    const u64 base = (u64) blockIdx.x * kFiltTile;
    u32 mine = 0;
#pragma unroll
    for (int j = 0; j < kFiltItems; ++j) {
        const u64 i = base + (u64) threadIdx.x * kFiltItems + j;
        u64 e;
        if (i < a.n) mine += sel_keep(a, i, e);
    }
    const u32 w = (u32) warp_sum64(mine);
    if (lane_id() == 0 && w) atomicAdd(&s_c, w);
    __syncthreads();
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = s_c;
}
__global__ void __launch_bounds__(kFiltThreads) k_select_write(SelArgs a, const u64 *tile_off, u64 *out) {
    __shared__ u32 s_w[kFiltThreads / 32];
    const u64 base = (u64) blockIdx.x * kFiltTile;
    u64 emit[kFiltItems];
    bool keep[kFiltItems];
    u32 mine = 0;
#pragma unroll
    for (int j = 0; j < kFiltItems; ++j) {
        const u64 i = base + (u64) threadIdx.x * kFiltItems + j;
        keep[j] = i < a.n && sel_keep(a, i, emit[j]);
        mine += keep[j];
    }
    const u32 incl = warp_incl_scan(mine);
    if (lane_id() == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    u32 before = 0;
    for (u32 w = 0; w < (threadIdx.x >> 5); ++w) before += s_w[w];
    u64 at = tile_off[blockIdx.x] + before + (incl - mine);
#pragma unroll
    for (int j = 0; j < kFiltItems; ++j)
        if (keep[j]) out[at++] = emit[j];
}

// create_relation / foo (structs.cpp:217-243): out[i] = {key, col[row]} with row = rows ? rows[i] : i and
//   key = row            for a binding that was not joined yet (the reference's tuple: key = row id), or
//   key = i              (by_index) for a binding of the intermediate: the relation is keyed by the INTERMEDIATE ROW, one
//                        tuple per row and no de-duplication, so that the join result indexes the old intermediate directly.
__global__ void __launch_bounds__(256) k_make_tuples(const u64 *col, const u64 *rows, u64 n, int by_index, Tup *out) {
    for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64) gridDim.x * blockDim.x) {
        const u64 row = rows ? rows[i] : i;
        Tup t;
        t.key = by_index ? i : row;
        t.val = col[row];
        out[i] = t;
    }
}
// out[i] = src[take_s ? pairs[i].s : pairs[i].r]  (src == null: the pair component itself): carries a column of the old
// intermediate to the rows of the new one / unzips the pairs (update_intermediate cases 1 and 2)
__global__ void __launch_bounds__(256) k_pairs_gather(const Pair *pairs, u64 n, int take_s, const u64 *src, u64 *out) {
    for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64) gridDim.x * blockDim.x) {
        const Pair q = pairs[i];
        const u64 k = take_s ? q.s : q.r;
        out[i] = src ? src[k] : k;
    }
}
// out[i] = src[idx[i]]: compaction of an intermediate column by the list of surviving rows
__global__ void __launch_bounds__(256) k_gather_u64(const u64 *src, const u64 *idx, u64 n, u64 *out) {
    for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64) gridDim.x * blockDim.x) out[i] = src[idx[i]];
}

// Row-id de-duplication (the unordered_set of create_relation, structs.cpp:238-241): row ids are < the relation's row
// count, so a bitmap of that many bits marks the present ones (one atomicOr per row id) and an ordered selection over the
// bitmap words emits them -- O(n + rows / 32), ascending output.
__global__ void __launch_bounds__(256) k_bitmap_mark(const u64 *rowids, u64 n, u64 nbits, u32 *bitmap, u32 *err) {
    for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64) gridDim.x * blockDim.x) {
        const u64 r = rowids[i];
        if (r >= nbits) {
            *err = 1;
            continue;
        }
        atomicOr(&bitmap[r >> 5], 1u << (r & 31));
    }
}
__global__ void __launch_bounds__(kFiltThreads) k_bitmap_count(const u32 *bitmap, u64 nwords, u32 *tile_cnt) {
    __shared__ u32 s_c;
    if (threadIdx.x == 0) s_c = 0;
    __syncthreads();
    const u64 base = (u64) blockIdx.x * kFiltTile;
    u32 mine = 0;
#pragma unroll
    for (int j = 0; j < kFiltItems; ++j) {
        const u64 i = base + (u64) threadIdx.x * kFiltItems + j;
        if (i < nwords) mine += __popc(bitmap[i]);
    }
    const u32 w = (u32) warp_sum64(mine);
    if (lane_id() == 0 && w) atomicAdd(&s_c, w);
    __syncthreads();
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = s_c;
}
__global__ void __launch_bounds__(kFiltThreads) k_bitmap_write(const u32 *bitmap, u64 nwords, const u64 *tile_off, u64 *out) {
    __shared__ u32 s_w[kFiltThreads / 32];
    const u64 base = (u64) blockIdx.x * kFiltTile;
    u32 word[kFiltItems];
    u32 mine = 0;
#pragma unroll
    for (int j = 0; j < kFiltItems; ++j) {
        const u64 i = base + (u64) threadIdx.x * kFiltItems + j;
        word[j] = i < nwords ? bitmap[i] : 0u;
        mine += __popc(word[j]);
    }
    const u32 incl = warp_incl_scan(mine);
    if (lane_id() == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    u32 before = 0;
    for (u32 w = 0; w < (threadIdx.x >> 5); ++w) before += s_w[w];
    u64 at = tile_off[blockIdx.x] + before + (incl - mine);
#pragma unroll
    for (int j = 0; j < kFiltItems; ++j) {
        const u64 first = (base + (u64) threadIdx.x * kFiltItems + j) * 32;
        u32 m = word[j];
        while (m) {
            const int b = __ffs(m) - 1;
            out[at++] = first + b;
            m &= m - 1;
        }
    }
}

}  // namespace rhj
