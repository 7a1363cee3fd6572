// rhj_query.cu -- the join's neighbours on the query path, behind the same C ABI (include/rhj.h):
// filters (Query.cpp:94-146), row-id gathers (structs.cpp:217-243), the projection checksum
// (Query.cpp:66-74), the multiset digest used by the parity checks, and update_intermediate
// (intermediate.cpp:52-183) re-expressed as join + gather.  No CPU fallback.
#include "rhj_ctx.cuh"
#include "rhj_query_kernels.cuh"

extern "C" {

// ---- filters / gathers ------------------------------------------------------------------------------

int rhj_filter_u64_device(rhj_ctx *ctx, const uint64_t *d_col, const uint64_t *d_rowids_in, uint64_t n_in, int op,
                          uint64_t constant, uint64_t *d_rowids_out, uint64_t *count, void *stream) {
    if (!ctx || !count || (op != '>' && op != '<' && op != '=')) return RHJ_ERR_ARG;
    *count = 0;
    if (n_in == 0) return RHJ_OK;
    if (!d_col || !d_rowids_out) return fail(ctx, RHJ_ERR_ARG, "null pointer");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    u64 ntile64 = (n_in + kFiltTile - 1) / kFiltTile;
    if (ntile64 > 0x7fffffffull) return fail(ctx, RHJ_ERR_ARG, "column too large");
    u32 ntile = (u32) ntile64;
    int rc;
    if ((rc = ensure(ctx, ctx->filt_cnt, (size_t) ntile * 4))) return rc;
    if ((rc = ensure(ctx, ctx->filt_off, (size_t) ntile * 8 + 8))) return rc;
    u64 *total = (u64 *) ctx->filt_off.p + ntile;
    k_filter_count<<<ntile, kFiltThreads, 0, st>>>((const u64 *) d_col, (const u64 *) d_rowids_in, n_in, op, constant,
                                                   (u32 *) ctx->filt_cnt.p);
    k_scan_tiles<<<1, 1024, 0, st>>>((const u32 *) ctx->filt_cnt.p, ntile, (u64 *) ctx->filt_off.p, total);
    // in-place compaction is safe tile by tile only if no tile writes ahead of an unread tile:
    // output index <= input index always holds, but tiles run concurrently -> stage when aliased.
    u64 *dst = (u64 *) d_rowids_out;
    DevBuf &tmp = ctx->filt_tmp;
    bool aliased = d_rowids_in && d_rowids_out == d_rowids_in;
    if (aliased) {
        if ((rc = ensure(ctx, tmp, n_in * 8))) return rc;
        dst = (u64 *) tmp.p;
    }
    k_filter_write<<<ntile, kFiltThreads, 0, st>>>((const u64 *) d_col, (const u64 *) d_rowids_in, n_in, op, constant,
                                                   (const u64 *) ctx->filt_off.p, dst);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_scalars, total, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *count = ctx->h_scalars[0];
    if (aliased && *count) {
        CK(cudaMemcpyAsync(d_rowids_out, dst, *count * 8, cudaMemcpyDeviceToDevice, st));
        CK(cudaStreamSynchronize(st));
    }
    return RHJ_OK;
}

int rhj_gather_tuples_device(rhj_ctx *ctx, const uint64_t *d_col, const uint64_t *d_rowids, uint64_t n,
                             rhj_tuple *d_out, void *stream) {
    if (!ctx) return RHJ_ERR_ARG;
    if (n == 0) return RHJ_OK;
    if (!d_col || !d_rowids || !d_out) return fail(ctx, RHJ_ERR_ARG, "null pointer");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    u32 grid = (u32) std::min<u64>((n + 255) / 256, (u64) ctx->num_sms * 16);
    k_gather_tuples<<<grid, 256, 0, st>>>((const u64 *) d_col, (const u64 *) d_rowids, n, (Tup *) d_out);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    return RHJ_OK;
}

int rhj_gather_sum_u64_device(rhj_ctx *ctx, const uint64_t *d_col, const uint64_t *d_rowids, uint64_t n,
                              uint64_t *sum, void *stream) {
    if (!ctx || !sum) return RHJ_ERR_ARG;
    *sum = 0;
    if (n == 0) return RHJ_OK;
    if (!d_col || !d_rowids) return fail(ctx, RHJ_ERR_ARG, "null pointer");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    int rc;
    if ((rc = ensure(ctx, ctx->filt_off, 16))) return rc;
    u64 *acc = (u64 *) ctx->filt_off.p;
    CK(cudaMemsetAsync(acc, 0, 8, st));
    u32 grid = (u32) std::min<u64>((n + 255) / 256, (u64) ctx->num_sms * 16);
    k_gather_sum<<<grid, 256, 0, st>>>((const u64 *) d_col, (const u64 *) d_rowids, n, acc);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_scalars, acc, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *sum = ctx->h_scalars[0];
    return RHJ_OK;
}

int rhj_pairs_digest_device(rhj_ctx *ctx, const rhj_pair *d_pairs, uint64_t n, uint64_t *sum, uint64_t *xr,
                            void *stream) {
    if (!ctx || !sum || !xr) return RHJ_ERR_ARG;
    *sum = 0;
    *xr = 0;
    if (n == 0) return RHJ_OK;
    if (!d_pairs) return fail(ctx, RHJ_ERR_ARG, "null pointer");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    int rc;
    if ((rc = ensure(ctx, ctx->filt_off, 16))) return rc;
    u64 *acc = (u64 *) ctx->filt_off.p;
    CK(cudaMemsetAsync(acc, 0, 16, st));
    u32 grid = (u32) std::min<u64>((n + 255) / 256, (u64) ctx->num_sms * 16);
    k_pairs_digest<<<grid, 256, 0, st>>>((const Pair *) d_pairs, n, acc, acc + 1);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_scalars, acc, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *sum = ctx->h_scalars[0];
    *xr = ctx->h_scalars[1];
    return RHJ_OK;
}


// ---- update_intermediate (intermediate.cpp:146-183) ---------------------------------------------------

namespace {

u32 grid_for(const rhj_ctx *ctx, u64 n) { return (u32) std::min<u64>((n + 255) / 256, (u64) ctx->num_sms * 16); }

// Joins A (index-keyed old rows) with B (index-keyed result pairs) on the device and leaves the
// matched (row index e, pair index p) list in ctx->iu_ep; *m = number of new rows.
int iu_match(rhj_ctx *ctx, cudaStream_t st, u64 n_rows, u64 n_pairs, u64 *m) {
    uint64_t cnt = 0;
    int rc = rhj_join_count_device(ctx, (const rhj_tuple *) ctx->iu_A.p, n_rows, (const rhj_tuple *) ctx->iu_B.p, n_pairs,
                                   &cnt, st);
    *m = cnt;
    if (rc) return rc;
    if (*m == 0) return RHJ_OK;
    if ((rc = ensure(ctx, ctx->iu_ep, *m * sizeof(Pair)))) return rc;
    return rhj_join_write_device(ctx, (rhj_pair *) ctx->iu_ep.p, *m, st);
}

int iu_pinned(rhj_ctx *ctx, size_t bytes) { return ensure_pinned(ctx, &ctx->h_iu, &ctx->h_iu_cap, bytes); }

// Carries every live column of the old intermediate to the m new rows (device gather through
// ctx->iu_ep) and copies it into the pinned result block; column c lands at h_iu + c*m.
int iu_gather_columns(rhj_ctx *ctx, cudaStream_t st, const uint64_t *const *cols, uint32_t n_cols, u64 n_rows, u64 m,
                      uint64_t **out_cols) {
    int rc;
    if ((rc = ensure(ctx, ctx->iu_col, n_rows * 8))) return rc;
    if ((rc = ensure(ctx, ctx->iu_out, m * 8))) return rc;
    u64 *h = (u64 *) ctx->h_iu;
    for (uint32_t c = 0; c < n_cols; ++c) {
        CK(cudaMemcpyAsync(ctx->iu_col.p, cols[c], n_rows * 8, cudaMemcpyHostToDevice, st));
        k_gather_by_elem<<<grid_for(ctx, m), 256, 0, st>>>((const u64 *) ctx->iu_col.p, (const Pair *) ctx->iu_ep.p, m,
                                                           (u64 *) ctx->iu_out.p);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(h + (size_t) c * m, ctx->iu_out.p, m * 8, cudaMemcpyDeviceToHost, st));
        out_cols[c] = (uint64_t *) (h + (size_t) c * m);
    }
    return RHJ_OK;
}

}  // namespace

// Case 2 of update_intermediate (intermediate.cpp:52-66,108-125,162-170): one binding of the join is
// already in the intermediate (row-id column match_col[n_rows]); every result pair whose row id on
// that side equals match_col[e] produces a new row = old row e + the pair's other row id.
// match_on_S = 1 when the already-joined binding is the join's S side (match keyS, append keyR).
// out_cols[0..n_cols-1] = the carried columns, out_cols[n_cols] = the new binding's column, all
// *out_rows long, in context-owned pinned memory valid until the next call on this context.
int rhj_intermediate_expand_host(rhj_ctx *ctx, const uint64_t *match_col, uint64_t n_rows, const rhj_pair *pairs,
                                 uint64_t n_pairs, int match_on_S, const uint64_t *const *cols, uint32_t n_cols,
                                 uint64_t **out_cols, uint64_t *out_rows) {
    if (!ctx || !out_cols || !out_rows) return RHJ_ERR_ARG;
    *out_rows = 0;
    for (uint32_t c = 0; c <= n_cols; ++c) out_cols[c] = nullptr;
    if (n_rows == 0 || n_pairs == 0) return RHJ_OK;
    if (!match_col || !pairs || (n_cols && !cols)) return fail(ctx, RHJ_ERR_ARG, "null pointer");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int rc;
    if ((rc = ensure(ctx, ctx->iu_col, n_rows * 8))) return rc;
    if ((rc = ensure(ctx, ctx->iu_pairs, n_pairs * sizeof(Pair)))) return rc;
    if ((rc = ensure(ctx, ctx->iu_A, n_rows * sizeof(Tup)))) return rc;
    if ((rc = ensure(ctx, ctx->iu_B, n_pairs * sizeof(Tup)))) return rc;
    CK(cudaMemcpyAsync(ctx->iu_col.p, match_col, n_rows * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->iu_pairs.p, pairs, n_pairs * sizeof(Pair), cudaMemcpyHostToDevice, st));
    k_index_tuples<<<grid_for(ctx, n_rows), 256, 0, st>>>((const u64 *) ctx->iu_col.p, n_rows, (Tup *) ctx->iu_A.p);
    k_pair_side_tuples<<<grid_for(ctx, n_pairs), 256, 0, st>>>((const Pair *) ctx->iu_pairs.p, n_pairs, match_on_S ? 1 : 0,
                                                               (Tup *) ctx->iu_B.p);
    CK(cudaGetLastError());
    u64 m = 0;
    if ((rc = iu_match(ctx, st, n_rows, n_pairs, &m))) return rc;
    if (m == 0) return RHJ_OK;
    if ((rc = iu_pinned(ctx, (size_t) (n_cols + 1) * m * 8))) return rc;
    if ((rc = iu_gather_columns(ctx, st, cols, n_cols, n_rows, m, out_cols))) return rc;
    k_gather_pair_value<<<grid_for(ctx, m), 256, 0, st>>>((const Pair *) ctx->iu_pairs.p, (const Pair *) ctx->iu_ep.p, m,
                                                          match_on_S ? 1 : 0, (u64 *) ctx->iu_out.p);
    CK(cudaGetLastError());
    u64 *h = (u64 *) ctx->h_iu + (size_t) n_cols * m;
    CK(cudaMemcpyAsync(h, ctx->iu_out.p, m * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    out_cols[n_cols] = (uint64_t *) h;
    *out_rows = m;
    return RHJ_OK;
}

// Case 3 of update_intermediate (intermediate.cpp:72-87,130-138,171-180): both bindings are already
// in the intermediate; a row e survives once per result pair equal to (col1[e], col2[e]).
// out_cols[0..n_cols-1] = the surviving rows of every carried column.
int rhj_intermediate_filter_host(rhj_ctx *ctx, const uint64_t *col1, const uint64_t *col2, uint64_t n_rows,
                                 const rhj_pair *pairs, uint64_t n_pairs, const uint64_t *const *cols, uint32_t n_cols,
                                 uint64_t **out_cols, uint64_t *out_rows) {
    if (!ctx || !out_cols || !out_rows) return RHJ_ERR_ARG;
    *out_rows = 0;
    for (uint32_t c = 0; c < n_cols; ++c) out_cols[c] = nullptr;
    if (n_rows == 0 || n_pairs == 0) return RHJ_OK;
    if (!col1 || !col2 || !pairs || (n_cols && !cols)) return fail(ctx, RHJ_ERR_ARG, "null pointer");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int rc;
    if ((rc = ensure(ctx, ctx->iu_col, 2 * n_rows * 8))) return rc;
    if ((rc = ensure(ctx, ctx->iu_pairs, n_pairs * sizeof(Pair)))) return rc;
    if ((rc = ensure(ctx, ctx->iu_A, n_rows * sizeof(Tup)))) return rc;
    if ((rc = ensure(ctx, ctx->iu_B, n_pairs * sizeof(Tup)))) return rc;
    if ((rc = ensure(ctx, ctx->filt_off, 16))) return rc;
    u64 *d_c1 = (u64 *) ctx->iu_col.p, *d_c2 = d_c1 + n_rows;
    u32 *d_ovf = (u32 *) ctx->filt_off.p;
    CK(cudaMemcpyAsync(d_c1, col1, n_rows * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_c2, col2, n_rows * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->iu_pairs.p, pairs, n_pairs * sizeof(Pair), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(d_ovf, 0, 4, st));
    k_index_tuples2<<<grid_for(ctx, n_rows), 256, 0, st>>>(d_c1, d_c2, n_rows, (Tup *) ctx->iu_A.p, d_ovf);
    k_pair_both_tuples<<<grid_for(ctx, n_pairs), 256, 0, st>>>((const Pair *) ctx->iu_pairs.p, n_pairs, (Tup *) ctx->iu_B.p,
                                                               d_ovf);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_scalars, d_ovf, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const bool wide = (*(u32 *) ctx->h_scalars) != 0;
    u64 m = 0;
    if (!wide) {
        if ((rc = iu_match(ctx, st, n_rows, n_pairs, &m))) return rc;
    } else {
        // row ids do not fit the 32+32-bit composite: match on the first id, verify the second
        k_index_tuples<<<grid_for(ctx, n_rows), 256, 0, st>>>(d_c1, n_rows, (Tup *) ctx->iu_A.p);
        k_pair_side_tuples<<<grid_for(ctx, n_pairs), 256, 0, st>>>((const Pair *) ctx->iu_pairs.p, n_pairs, 0,
                                                                   (Tup *) ctx->iu_B.p);
        CK(cudaGetLastError());
        u64 cand = 0;
        if ((rc = iu_match(ctx, st, n_rows, n_pairs, &cand))) return rc;
        if (cand) {
            u64 ntile64 = (cand + kFiltTile - 1) / kFiltTile;
            if (ntile64 > 0x7fffffffull) return fail(ctx, RHJ_ERR_ARG, "candidate list too large");
            u32 ntile = (u32) ntile64;
            if ((rc = ensure(ctx, ctx->filt_cnt, (size_t) ntile * 4))) return rc;
            if ((rc = ensure(ctx, ctx->filt_off, (size_t) ntile * 8 + 8))) return rc;
            if ((rc = ensure(ctx, ctx->filt_tmp, cand * sizeof(Pair)))) return rc;
            u64 *total = (u64 *) ctx->filt_off.p + ntile;
            k_ep_verify_count<<<ntile, kFiltThreads, 0, st>>>((const Pair *) ctx->iu_ep.p, cand, d_c2,
                                                              (const Pair *) ctx->iu_pairs.p, (u32 *) ctx->filt_cnt.p);
            k_scan_tiles<<<1, 1024, 0, st>>>((const u32 *) ctx->filt_cnt.p, ntile, (u64 *) ctx->filt_off.p, total);
            k_ep_verify_write<<<ntile, kFiltThreads, 0, st>>>((const Pair *) ctx->iu_ep.p, cand, d_c2,
                                                              (const Pair *) ctx->iu_pairs.p, (const u64 *) ctx->filt_off.p,
                                                              (Pair *) ctx->filt_tmp.p);
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(ctx->h_scalars, total, 8, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            m = ctx->h_scalars[0];
            if (m) CK(cudaMemcpyAsync(ctx->iu_ep.p, ctx->filt_tmp.p, m * sizeof(Pair), cudaMemcpyDeviceToDevice, st));
        }
    }
    if (m == 0) return RHJ_OK;
    if ((rc = iu_pinned(ctx, (size_t) n_cols * m * 8))) return rc;
    if ((rc = iu_gather_columns(ctx, st, cols, n_cols, n_rows, m, out_cols))) return rc;
    CK(cudaStreamSynchronize(st));
    *out_rows = m;
    return RHJ_OK;
}

}  // extern "C"

#include "rhj_exec.cuh"
