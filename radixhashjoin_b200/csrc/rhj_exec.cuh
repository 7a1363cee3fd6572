// rhj_exec.cuh -- device-resident query path, behind the C ABI of include/rhj.h; part of the
// rhj_query.cu translation unit (shares its kernels).
//
// The reference runs a query as  Query::execute -> run_filters -> run_joins -> column_proj  (Query.cpp:81-211) with
// unordered_sets of row ids, `new tuple[]` relations, malloc'd Result pages and vector<vector<u64>> intermediates, all on
// the host, so a three-join query built on the CUDA join alone crosses PCIe six times.  Here
//   - a relList column is uploaded ONCE (rhj_column_device: a process-wide cache keyed by the host pointer, the columns
//     are read-only mmap'ed files, structs.cpp:18-60) and stays in HBM;
//   - filters compact row-id lists on the device (k_select_*), create_relation gathers {row id, value} tuples on the
//     device (k_make_tuples), the join result never becomes a page list, update_intermediate's three cases are an
//     unzip, ONE join keyed by the intermediate row (which replaces de-duplication + join + the O(pairs x rows) expansion
//     of intermediate.cpp:52-125) and a row filter, and column_proj is a gather-sum;
//   - the only bytes that cross PCIe per query are a handful of counts and the projection sums.
// rhj_unique_rowids_device is the reference's row-id de-duplication (structs.cpp:238-241) as a stand-alone kernel for
// callers that keep the reference's formulation.  No CPU fallback.
#include <mutex>
#include <unordered_map>
#include <vector>

#pragma once
#include "rhj_ctx.cuh"
#include "rhj_exec_kernels.cuh"

namespace {

u32 grid_of(const rhj_ctx *ctx, u64 n) { return (u32) std::max<u64>(1, std::min<u64>((n + 255) / 256, (u64) ctx->num_sms * 16)); }

// ---- resident columns ---------------------------------------------------------------------------------------------
struct ColKey {
    const void *host;
    int device;
    bool operator==(const ColKey &o) const { return host == o.host && device == o.device; }
};
struct ColKeyHash {
    size_t operator()(const ColKey &k) const { return std::hash<const void *>()(k.host) ^ ((size_t) k.device * 0x9E3779B97F4A7C15ull); }
};
struct ColVal {
    u64 *dev;
    u64 n;
    u64 probe[3];  // first / middle / last value at upload time: a recycled host address with other contents is re-uploaded
};
inline void col_probe(const uint64_t *h, u64 n, u64 out[3]) {
    out[0] = h[0];
    out[1] = h[n / 2];
    out[2] = h[n - 1];
}
std::mutex g_col_mu;
std::unordered_map<ColKey, ColVal, ColKeyHash> g_cols;

// stream-ordered scratch of one query: everything is freed when the executor returns
struct Arena {
    rhj_ctx *ctx;
    cudaStream_t st;
    std::vector<void *> blocks;
    int alloc(void **p, size_t bytes) {
        *p = nullptr;
        cudaError_t e = cudaMallocAsync(p, std::max<size_t>(bytes, 16), st);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, RHJ_ERR_NOMEM, "cudaMallocAsync (query scratch)", e);
        }
        blocks.push_back(*p);
        return RHJ_OK;
    }
    void release(void *p) {
        for (auto &b : blocks)
            if (b == p) {
                cudaFreeAsync(p, st);
                b = nullptr;
                return;
            }
    }
    ~Arena() {
        for (void *b : blocks)
            if (b) cudaFreeAsync(b, st);
    }
};

// order-preserving selection: out[0..*count) = the emitted values of the kept elements.  One host sync (the count).
int select_rows(rhj_ctx *ctx, cudaStream_t st, const SelArgs &a, u64 *d_out, u64 *count, rhj_query_stats *qs) {
    *count = 0;
    if (a.n == 0) return RHJ_OK;
    const u64 ntile64 = (a.n + kFiltTile - 1) / kFiltTile;
    if (ntile64 > 0x7fffffffull) return fail(ctx, RHJ_ERR_ARG, "row list too large");
    const u32 ntile = (u32) ntile64;
    int rc;
    if ((rc = ensure(ctx, ctx->filt_cnt, (size_t) ntile * 4))) return rc;
    if ((rc = ensure(ctx, ctx->filt_off, (size_t) ntile * 8 + 8))) return rc;
    u64 *total = (u64 *) ctx->filt_off.p + ntile;
    k_select_count<<<ntile, kFiltThreads, 0, st>>>(a, (u32 *) ctx->filt_cnt.p);
    k_scan_tiles<<<1, 1024, 0, st>>>((const u32 *) ctx->filt_cnt.p, ntile, (u64 *) ctx->filt_off.p, total);
    k_select_write<<<ntile, kFiltThreads, 0, st>>>(a, (const u64 *) ctx->filt_off.p, d_out);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_scalars, total, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *count = ctx->h_scalars[0];
    if (qs) {
        qs->kernel_launches += 3;
        qs->d2h_bytes += 8;
    }
    return RHJ_OK;
}

}  // namespace

extern "C" {

// relList::relList keeps every column as a read-only host array for the life of the process (structs.cpp:18-60); the
// first query that touches a column uploads it, every later one -- on any context of the process -- gets the same
// device copy.  *d_col is valid until rhj_column_cache_clear().
int rhj_column_device(rhj_ctx *ctx, const uint64_t *host_col, uint64_t n, const uint64_t **d_col, uint64_t *uploaded_bytes) {
    if (!ctx || !d_col || (n && !host_col)) return RHJ_ERR_ARG;
    if (uploaded_bytes) *uploaded_bytes = 0;
    *d_col = nullptr;
    if (n == 0) return RHJ_OK;
    CK(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> lock(g_col_mu);
    const ColKey key{host_col, ctx->device};
    auto it = g_cols.find(key);
    u64 pr[3];
    col_probe(host_col, n, pr);
    if (it != g_cols.end() && it->second.n == n && it->second.probe[0] == pr[0] && it->second.probe[1] == pr[1] &&
        it->second.probe[2] == pr[2]) {
        *d_col = (const uint64_t *) it->second.dev;
        return RHJ_OK;
    }
    u64 *dev = nullptr;
    cudaError_t e = cudaMalloc(&dev, n * 8);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, RHJ_ERR_NOMEM, "cudaMalloc (resident column)", e);
    }
    // pageable source (an mmap'ed file).  A plain cudaMemcpy may return once the data is STAGED, with the DMA still in
    // flight on the null stream, which the contexts' non-blocking streams do not wait for: copy on this context's stream
    // and wait for it before any context can see the column.
    e = cudaMemcpyAsync(dev, host_col, n * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        cudaFree(dev);
        return fail(ctx, RHJ_ERR_CUDA, "cudaMemcpy (resident column)", e);
    }
    if (it != g_cols.end()) {
        cudaFree(it->second.dev);
        it->second = ColVal{dev, n, {pr[0], pr[1], pr[2]}};
    } else {
        g_cols.emplace(key, ColVal{dev, n, {pr[0], pr[1], pr[2]}});
    }
    if (uploaded_bytes) *uploaded_bytes = n * 8;
    *d_col = (const uint64_t *) dev;
    return RHJ_OK;
}

int rhj_column_cache_clear(void) {
    std::lock_guard<std::mutex> lock(g_col_mu);
    for (auto &kv : g_cols) {
        cudaSetDevice(kv.first.device);
        cudaFree(kv.second.dev);
    }
    g_cols.clear();
    return RHJ_OK;
}

// create_relation's row-id de-duplication (structs.cpp:238-241: the intermediate column goes through an unordered_set):
// d_out[0..*count) = the distinct values of d_rowids[n], ascending.  Every row id must be < n_rows (the relation's row
// count); d_out needs room for min(n, n_rows) values.
int rhj_unique_rowids_device(rhj_ctx *ctx, const uint64_t *d_rowids, uint64_t n, uint64_t n_rows, uint64_t *d_out,
                             uint64_t *count, void *stream) {
    if (!ctx || !count) return RHJ_ERR_ARG;
    *count = 0;
    if (n == 0) return RHJ_OK;
    if (!d_rowids || !d_out || n_rows == 0) return fail(ctx, RHJ_ERR_ARG, "null pointer / empty relation");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    const u64 nwords = (n_rows + 31) / 32;
    const u64 ntile64 = (nwords + kFiltTile - 1) / kFiltTile;
    if (ntile64 > 0x7fffffffull) return fail(ctx, RHJ_ERR_ARG, "relation too large");
    const u32 ntile = (u32) ntile64;
    int rc;
    if ((rc = ensure(ctx, ctx->filt_tmp, nwords * 4 + 16))) return rc;
    if ((rc = ensure(ctx, ctx->filt_cnt, (size_t) ntile * 4))) return rc;
    if ((rc = ensure(ctx, ctx->filt_off, (size_t) ntile * 8 + 16))) return rc;
    u32 *bitmap = (u32 *) ctx->filt_tmp.p;
    u32 *err = bitmap + nwords;
    u64 *total = (u64 *) ctx->filt_off.p + ntile;
    CK(cudaMemsetAsync(bitmap, 0, nwords * 4 + 4, st));
    k_bitmap_mark<<<grid_of(ctx, n), 256, 0, st>>>((const u64 *) d_rowids, n, n_rows, bitmap, err);
    k_bitmap_count<<<ntile, kFiltThreads, 0, st>>>(bitmap, nwords, (u32 *) ctx->filt_cnt.p);
    k_scan_tiles<<<1, 1024, 0, st>>>((const u32 *) ctx->filt_cnt.p, ntile, (u64 *) ctx->filt_off.p, total);
    k_bitmap_write<<<ntile, kFiltThreads, 0, st>>>(bitmap, nwords, (const u64 *) ctx->filt_off.p, (u64 *) d_out);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_scalars, total, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ctx->h_scalars + 1, err, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if ((u32) ctx->h_scalars[1]) return fail(ctx, RHJ_ERR_ARG, "rhj_unique_rowids_device: a row id is >= n_rows");
    *count = ctx->h_scalars[0];
    return RHJ_OK;
}

// Query::execute (Query.cpp:204-211) on the device.  sums[n_projs] receive the projection checksums (Query.cpp:66-74);
// *empty = 1 when a filter or a join left nothing (the reference prints NULL for every projection then).
int rhj_query_execute(rhj_ctx *ctx, const rhj_query_desc *q, uint64_t *sums, int *empty, rhj_query_stats *stats) {
    if (!ctx || !q || !empty || (q->n_projs && !sums)) return RHJ_ERR_ARG;
    if (q->n_bindings == 0 || q->n_bindings > RHJ_MAX_BINDINGS || !q->bindings) return fail(ctx, RHJ_ERR_ARG, "bad binding list");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    rhj_query_stats local{};
    rhj_query_stats *qs = stats ? stats : &local;
    *qs = rhj_query_stats{};
    *empty = 0;
    for (u32 i = 0; i < q->n_projs; ++i) sums[i] = 0;
    Arena ar{ctx, st, {}};
    int rc;
    const u32 nb = q->n_bindings;

    auto column = [&](u32 b, u32 c, const u64 **d) -> int {
        const rhj_q_relation &r = q->bindings[b];
        if (c >= r.num_columns) return fail(ctx, RHJ_ERR_ARG, "column index out of range");
        uint64_t up = 0;
        int rc2 = rhj_column_device(ctx, r.columns[c], r.num_tuples, (const uint64_t **) d, &up);
        qs->h2d_bytes += up;
        return rc2;
    };

    // ---- run_filters (Query.cpp:81-158): per binding, the surviving row ids; null list = all rows ----
    struct Rows {
        u64 *list = nullptr;
        u64 n = 0;
    };
    Rows filtered[RHJ_MAX_BINDINGS];
    for (u32 b = 0; b < nb; ++b) filtered[b].n = q->bindings[b].num_tuples;
    for (u32 f = 0; f < q->n_filters; ++f) {
        const rhj_q_filter &fl = q->filters[f];
        if (fl.binding >= nb) return fail(ctx, RHJ_ERR_ARG, "filter binding out of range");
        if (fl.op != '>' && fl.op != '<' && fl.op != '=') return fail(ctx, RHJ_ERR_ARG, "unknown filter operator");
        Rows &fr = filtered[fl.binding];
        const u64 *col;
        if ((rc = column(fl.binding, fl.column, &col))) return rc;
        u64 *out;
        if ((rc = ar.alloc((void **) &out, fr.n * 8))) return rc;
        SelArgs sa{kSelConst, col, nullptr, fr.list, nullptr, fr.n, fl.op, fl.constant};
        u64 kept = 0;
        if ((rc = select_rows(ctx, st, sa, out, &kept, qs))) return rc;
        if (fr.list) ar.release(fr.list);
        fr.list = out;
        fr.n = kept;
        if (kept == 0) {  // Query.cpp:104-106,124-126,140-142
            *empty = 1;
            return RHJ_OK;
        }
    }

    // ---- run_joins (Query.cpp:161-201): the intermediate is one device column of row ids per joined binding ----
    u64 *inter[RHJ_MAX_BINDINGS] = {};
    u64 rows = 0;
    auto in_inter = [&](u32 b) { return inter[b] != nullptr; };
    auto any_inter = [&]() {
        for (u32 b = 0; b < nb; ++b)
            if (inter[b]) return true;
        return false;
    };
    // keeps the rows listed in d_keep[kept] (row indices) of every intermediate column
    auto compact_inter = [&](const u64 *d_keep, u64 kept) -> int {
        for (u32 b = 0; b < nb; ++b) {
            if (!inter[b]) continue;
            u64 *nu;
            int rc2;
            if ((rc2 = ar.alloc((void **) &nu, kept * 8))) return rc2;
            if (kept) k_gather_u64<<<grid_of(ctx, kept), 256, 0, st>>>(inter[b], d_keep, kept, nu);
            qs->kernel_launches++;
            ar.release(inter[b]);
            inter[b] = nu;
        }
        rows = kept;
        return RHJ_OK;
    };

    // ---- join order.  The reference runs the joins as written (README.md:63-64 lists reordering as
    // future work, the statistics of structs.cpp:40-60 / Query.cpp:88-154 are collected and never used).  With
    // query->reorder_joins the executor goes cheapest-first over the join GRAPH: start with the join whose two filtered
    // inputs are smallest, then always take a join that touches the bindings joined so far -- one that closes a cycle
    // first (it is a row filter), else the one that brings in the smallest new binding.  Any order of a connected join
    // graph yields the same multiset of result rows, so the checksums are unchanged; queries with a same-binding
    // predicate or a disconnected graph keep the written order.
    u32 order[64];
    const u32 nj = q->n_joins;
    if (nj > 64) return fail(ctx, RHJ_ERR_ARG, "more than 64 joins");
    for (u32 j = 0; j < nj; ++j) order[j] = j;
    if (q->reorder_joins && nj > 1) {
        bool plain = true;
        for (u32 j = 0; j < nj; ++j) {
            const rhj_q_join &x = q->joins[j];
            if (x.binding1 >= nb || x.binding2 >= nb) return fail(ctx, RHJ_ERR_ARG, "join binding out of range");
            if (x.binding1 == x.binding2) plain = false;
        }
        if (plain) {
            bool used[64] = {}, joined[RHJ_MAX_BINDINGS] = {};
            u32 planned[64], np = 0;
            for (; np < nj; ++np) {
                u32 best = nj;
                u64 best_cost = ~0ull;
                for (u32 j = 0; j < nj; ++j) {
                    if (used[j]) continue;
                    const rhj_q_join &x = q->joins[j];
                    const bool in1 = joined[x.binding1], in2 = joined[x.binding2];
                    u64 cost;
                    if (np == 0) cost = filtered[x.binding1].n + filtered[x.binding2].n;
                    else if (in1 && in2) cost = 0;
                    else if (in1 || in2) cost = 1 + filtered[in1 ? x.binding2 : x.binding1].n;
                    else continue;  // not connected to what has been joined so far
                    if (cost < best_cost) {
                        best_cost = cost;
                        best = j;
                    }
                }
                if (best == nj) break;  // disconnected graph: keep the written order
                used[best] = true;
                planned[np] = best;
                joined[q->joins[best].binding1] = joined[q->joins[best].binding2] = true;
            }
            if (np == nj) {
                for (u32 j = 0; j < nj; ++j) order[j] = planned[j];
                for (u32 j = 0; j < nj; ++j)
                    if (order[j] != j) qs->joins_reordered = 1;
            }
        }
    }

    for (u32 jo = 0; jo < q->n_joins && !*empty; ++jo) {
        const rhj_q_join &jn = q->joins[order[jo]];
        if (jn.binding1 >= nb || jn.binding2 >= nb) return fail(ctx, RHJ_ERR_ARG, "join binding out of range");
        const u64 *c1, *c2;
        if ((rc = column(jn.binding1, jn.column1, &c1))) return rc;
        if ((rc = column(jn.binding2, jn.column2, &c2))) return rc;
        const u32 b1 = jn.binding1, b2 = jn.binding2;
        if (b1 == b2) {
            // same-binding predicate t.a = t.b: parse_table (intermediate.cpp:11-44)
            if (!in_inter(b1)) {
                if (any_inter()) return fail(ctx, RHJ_ERR_ARG, "same-binding predicate on an unjoined binding next to joined ones: undefined in the reference");
                // first branch (16-25): the filtered rows with equal columns BECOME the binding's intermediate column
                Rows &fr = filtered[b1];
                u64 *out;
                if ((rc = ar.alloc((void **) &out, fr.n * 8))) return rc;
                SelArgs sa{kSelSameRow, c1, c2, fr.list, nullptr, fr.n, 0, 0};
                u64 kept = 0;
                if ((rc = select_rows(ctx, st, sa, out, &kept, qs))) return rc;
                if (kept) {  // (an empty result leaves the binding unjoined, exactly as the reference's empty vector does)
                    inter[b1] = out;
                    rows = kept;
                } else {
                    ar.release(out);
                }
            } else {
                // second branch: a row filter of the intermediate (the reference's code dereferences end() here)
                u64 *keep;
                if ((rc = ar.alloc((void **) &keep, rows * 8))) return rc;
                SelArgs sa{kSelTwoCols, c1, c2, inter[b1], inter[b1], rows, 0, 0};
                u64 kept = 0;
                if ((rc = select_rows(ctx, st, sa, keep, &kept, qs))) return rc;
                if ((rc = compact_inter(keep, kept))) return rc;
                ar.release(keep);
            }
            continue;
        }
        const bool has1 = in_inter(b1), has2 = in_inter(b2);
        if (has1 && has2) {
            // update_intermediate case 3 (intermediate.cpp:72-87,171-180).  The reference joins the two de-duplicated row-id
            // columns and keeps a row once per result pair equal to its (row id, row id): the pair is in the result exactly
            // when the two VALUES are equal, so the whole step is a row filter.
            u64 *keep;
            if ((rc = ar.alloc((void **) &keep, rows * 8))) return rc;
            SelArgs sa{kSelTwoCols, c1, c2, inter[b1], inter[b2], rows, 0, 0};
            u64 kept = 0;
            if ((rc = select_rows(ctx, st, sa, keep, &kept, qs))) return rc;
            if ((rc = compact_inter(keep, kept))) return rc;
            ar.release(keep);
            if (kept == 0) *empty = 1;  // results.isEmpty() -> filtered_out (Query.cpp:187-190)
            continue;
        }
        if (!has1 && !has2 && any_inter())
            return fail(ctx, RHJ_ERR_ARG, "join between two unjoined bindings next to joined ones (a cross product): the reference's intermediate is inconsistent there");
        // relR / relS: create_relation (structs.cpp:228-243).  A binding of the intermediate is keyed by the intermediate ROW
        // (one tuple per row, no de-duplication): the pairs then index the old intermediate directly, which turns
        // de-duplication + join + the per-pair rescan of change_intermediate (intermediate.cpp:52-66) into one join.
        const u64 nR = has1 ? rows : filtered[b1].n, nS = has2 ? rows : filtered[b2].n;
        Tup *dR, *dS;
        if ((rc = ar.alloc((void **) &dR, nR * sizeof(Tup)))) return rc;
        if ((rc = ar.alloc((void **) &dS, nS * sizeof(Tup)))) return rc;
        k_make_tuples<<<grid_of(ctx, nR), 256, 0, st>>>(c1, has1 ? inter[b1] : filtered[b1].list, nR, has1 ? 1 : 0, dR);
        k_make_tuples<<<grid_of(ctx, nS), 256, 0, st>>>(c2, has2 ? inter[b2] : filtered[b2].list, nS, has2 ? 1 : 0, dS);
        CK(cudaGetLastError());
        qs->kernel_launches += 2;
        uint64_t m = 0;
        if ((rc = rhj_join_count_device(ctx, (const rhj_tuple *) dR, nR, (const rhj_tuple *) dS, nS, &m, st))) return rc;
        qs->kernel_launches += ctx->info.kernel_launches;
        qs->d2h_bytes += kScCount * 8;
        qs->joins++;
        qs->join_input_tuples += nR + nS;
        qs->join_output_pairs += m;
        if (m == 0) {
            *empty = 1;
            break;
        }
        Pair *pairs;
        if ((rc = ar.alloc((void **) &pairs, m * sizeof(Pair)))) return rc;
        if ((rc = rhj_join_write_device(ctx, (rhj_pair *) pairs, m, st))) return rc;
        qs->kernel_launches++;
        ar.release(dR);
        ar.release(dS);
        // update_intermediate cases 1 and 2 (intermediate.cpp:153-170): every pair is one row of the new intermediate
        u64 *nu[RHJ_MAX_BINDINGS] = {};
        for (u32 b = 0; b < nb; ++b) {
            const bool fresh = (b == b1 && !has1) || (b == b2 && !has2);
            if (!inter[b] && !fresh) continue;
            if ((rc = ar.alloc((void **) &nu[b], m * 8))) return rc;
            if (fresh) {  // the new binding's row id is the pair's own component
                k_pairs_gather<<<grid_of(ctx, m), 256, 0, st>>>(pairs, m, b == b2 ? 1 : 0, nullptr, nu[b]);
            } else {      // carried column: the pair's component on the joined side is the old row index
                k_pairs_gather<<<grid_of(ctx, m), 256, 0, st>>>(pairs, m, has2 ? 1 : 0, inter[b], nu[b]);
            }
            qs->kernel_launches++;
        }
        CK(cudaGetLastError());
        for (u32 b = 0; b < nb; ++b) {
            if (inter[b]) ar.release(inter[b]);
            inter[b] = nu[b];
        }
        ar.release(pairs);
        rows = m;
    }

    // ---- column_proj (Query.cpp:66-74, 198-200) ----
    if (!*empty && q->n_projs) {
        u64 *acc;
        if ((rc = ar.alloc((void **) &acc, (size_t) q->n_projs * 8))) return rc;
        CK(cudaMemsetAsync(acc, 0, (size_t) q->n_projs * 8, st));
        for (u32 p = 0; p < q->n_projs; ++p) {
            const rhj_q_proj &pr = q->projs[p];
            if (pr.binding >= nb) return fail(ctx, RHJ_ERR_ARG, "projection binding out of range");
            if (!inter[pr.binding] || rows == 0) continue;  // a binding that never joined sums an empty vector: 0
            const u64 *col;
            if ((rc = column(pr.binding, pr.column, &col))) return rc;
            k_gather_sum<<<grid_of(ctx, rows), 256, 0, st>>>(col, inter[pr.binding], rows, acc + p);
            qs->kernel_launches++;
        }
        CK(cudaGetLastError());
        std::vector<u64> h(q->n_projs);
        CK(cudaMemcpyAsync(h.data(), acc, (size_t) q->n_projs * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        qs->d2h_bytes += (u64) q->n_projs * 8;
        for (u32 p = 0; p < q->n_projs; ++p) sums[p] = h[p];
    }
    qs->result_rows = *empty ? 0 : rows;
    return RHJ_OK;
}

}  // extern "C"
