/*
 * rhj.h -- C ABI of the B200-native radix hash join (librhj.so).
 *
 * Drop-in boundary for the reference's join path.  The reference has no FFI:
 * its seam is one C++ member function and one POD page format,
 *     void Result::multiRadixHashJoin(JobScheduler&, relation&, relation&)   (Result.h:30, Result.cpp:90-124)
 * called from Query::run_joins (Query.cpp:185-186) and consumed by
 * update_intermediate (intermediate.cpp:146-183).  Every entry point below
 * cites the reference code it replaces.  Plain pointers and sizes only; no
 * C++ or torch types cross this boundary.
 *
 * Conventions
 *   - every function returns an rhj_status (0 = RHJ_OK); rhj_last_error(ctx)
 *     gives the text of the last failure on that context.  There is NO CPU
 *     fallback: without a usable sm_100 device rhj_create fails.
 *   - `*_device` entry points take DEVICE pointers and a cudaStream_t passed as
 *     void* (NULL = CUDA's default stream, as in the runtime API).  They enqueue
 *     on that stream -- after whatever the caller enqueued there before -- and,
 *     unless stated otherwise, synchronise it before returning so that host
 *     out-parameters are valid.
 *   - `*_host` entry points take HOST pointers, do their own H2D/D2H.
 *   - a context is NOT thread-safe; the reference calls the join from up to
 *     NUM_OF_THREADS=8 query threads (MainScheduler.cpp:6-14), so the host side
 *     keeps one context per query thread (contexts are independent).
 *   - rhj_tuple.key is the ROW ID and rhj_tuple.payload the JOIN VALUE, exactly
 *     as in the reference (structs.h:33-36, structs.cpp:222-223).
 */
#ifndef RHJ_H
#define RHJ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RHJ_VERSION "0.1.0"

/* layout-identical to `tuple` (structs.h:33-36) */
typedef struct rhj_tuple { uint64_t key; uint64_t payload; } rhj_tuple;
/* layout-identical to `key_tuple` (Result.h:9-12): always R's row id first
 * (Result.cpp:66-69, JobScheduler.cpp:187-190) */
typedef struct rhj_pair { uint64_t keyR; uint64_t keyS; } rhj_pair;

typedef struct rhj_ctx rhj_ctx;

typedef enum rhj_status {
    RHJ_OK = 0,
    RHJ_ERR_CUDA = 1,        /* a CUDA runtime call or kernel failed            */
    RHJ_ERR_ARG = 2,         /* bad argument                                     */
    RHJ_ERR_NOMEM = 3,       /* device or pinned-host allocation failed          */
    RHJ_ERR_CAPACITY = 4,    /* output buffer too small; *count holds the need   */
    RHJ_ERR_STATE = 5,       /* call sequence error (e.g. write before count)    */
    RHJ_ERR_NO_DEVICE = 6    /* no CUDA device / not sm_100                       */
} rhj_status;

/* digit function of the stand-alone partition entry points */
#define RHJ_DIGIT_RAW  0     /* (payload >> shift) & (2^bits-1): the reference's payload & 0xFF when shift=0,bits=8
                                (JobScheduler.cpp:151,171) */
#define RHJ_DIGIT_HASH 1     /* same bits taken from the 32-bit mixed hash the join itself partitions on */

/* emitter selection for rhj_join_device */
#define RHJ_EMIT_FUSED 0           /* one probe pass, block-level atomic reservation in the caller's buffer   */
#define RHJ_EMIT_COUNT_THEN_WRITE 1 /* two probe passes: count -> prefix sum -> write at fixed offsets        */

/* filter operators = the reference's `op` characters (Query.cpp:94,114,133) */
#define RHJ_OP_GT '>'
#define RHJ_OP_LT '<'
#define RHJ_OP_EQ '='

/* ---- context ------------------------------------------------------------------------------ */

/* Creates a context on CUDA device `device` (one per query thread; replaces the per-thread
 * JobScheduler of MainScheduler.cpp:6-14 on this path). */
int rhj_create(int device, rhj_ctx **ctx);
int rhj_destroy(rhj_ctx *ctx);
const char *rhj_last_error(const rhj_ctx *ctx);
const char *rhj_version(void);
/* Pre-sizes the device workspace for joins of up to nR x nS tuples (optional; the workspace
 * grows on demand, and growing is a cudaMalloc that benchmarks want outside the timed region). */
int rhj_reserve(rhj_ctx *ctx, uint64_t nR, uint64_t nS);
/* Bytes of device workspace currently held. */
uint64_t rhj_workspace_bytes(const rhj_ctx *ctx);

/* ---- the join: Result::multiRadixHashJoin (Result.cpp:90-124) -------------------------------- */

/* Device-resident join.  dR[nR], dS[nS] are read-only device arrays; d_out receives up to
 * `capacity` pairs, *count the number of matching pairs (valid even on RHJ_ERR_CAPACITY, in
 * which case only the first `capacity` reservations were written).  Output is the full multiset
 * of (rowidR,rowidS) with equal payloads; order unspecified, as in the reference.
 * `emit` is RHJ_EMIT_FUSED or RHJ_EMIT_COUNT_THEN_WRITE. */
int rhj_join_device(rhj_ctx *ctx, const rhj_tuple *dR, uint64_t nR, const rhj_tuple *dS, uint64_t nS,
                    rhj_pair *d_out, uint64_t capacity, uint64_t *count, int emit, void *stream);

/* Two-call form of the count-then-write emitter (replaces add_result/addAll paging,
 * Result.cpp:21-35,78-84,111-121): the first call partitions both relations and counts, the
 * second writes exactly *count pairs into a buffer the caller sized from it. */
int rhj_join_count_device(rhj_ctx *ctx, const rhj_tuple *dR, uint64_t nR, const rhj_tuple *dS, uint64_t nS,
                          uint64_t *count, void *stream);
int rhj_join_write_device(rhj_ctx *ctx, rhj_pair *d_out, uint64_t capacity, void *stream);

/* Host-resident join: what Result::multiRadixHashJoin does for the query threads.  R and S are
 * host arrays (the caller's relation::tuples, structs.h:38-40).  *out is a pinned host array of
 * *count pairs owned by the context and valid until the next call on it (NULL when *count == 0,
 * matching head == nullptr, Result.cpp:16-18). */
int rhj_join_host(rhj_ctx *ctx, const rhj_tuple *R, uint64_t nR, const rhj_tuple *S, uint64_t nS,
                  const rhj_pair **out, uint64_t *count);

/* Materialises pairs[count] as the reference's page list (Result.cpp:21-35): malloc'd 128 KiB
 * pages [next*][8191 x key_tuple], newest page first; *head_size = pairs in the head page
 * (Result.size), every other page is full.  The pages are free()'d by ~Result (Result.cpp:127-133).
 * Returns the head page (NULL when count == 0; NULL with *head_size == 0 when a page could not be allocated --
 * nothing is leaked in that case). */
void *rhj_pairs_to_pages(const rhj_pair *pairs, uint64_t count, uint64_t *head_size);

/* ---- the steps, individually callable (parity tests compare each with the reference) -------- */

/* HistogramJob::run + global sum (JobScheduler.cpp:149-155, structs.cpp:168-173):
 * d_hist[2^bits] (u64) = number of tuples per digit. */
int rhj_histogram_device(rhj_ctx *ctx, const rhj_tuple *d_in, uint64_t n, int bits, int shift, int digit_kind,
                         uint64_t *d_hist, void *stream);

/* hash_relation (structs.cpp:144-204): d_out[n] = d_in[n] grouped by digit (not stable -- the
 * reference's consumers never depend on the order inside a bucket), d_offsets[2^bits+1] (u64) =
 * exclusive prefix sum of the histogram. */
int rhj_partition_device(rhj_ctx *ctx, const rhj_tuple *d_in, uint64_t n, int bits, int shift, int digit_kind,
                         rhj_tuple *d_out, uint64_t *d_offsets, void *stream);

/* ---- neighbours of the join on the query path ------------------------------------------------ */

/* Query::run_filters predicate scans (Query.cpp:94-146).  d_rowids_in == NULL means "all rows
 * 0..n_in-1" of the column (the initial set fill, Query.cpp:85-87); otherwise only those rows
 * are tested.  Survivors keep their input order.  d_rowids_out may alias d_rowids_in. */
int rhj_filter_u64_device(rhj_ctx *ctx, const uint64_t *d_col, const uint64_t *d_rowids_in, uint64_t n_in,
                          int op, uint64_t constant, uint64_t *d_rowids_out, uint64_t *count, void *stream);

/* relation::foo (structs.cpp:217-226): d_out[i] = {rowids[i], col[rowids[i]]}. */
int rhj_gather_tuples_device(rhj_ctx *ctx, const uint64_t *d_col, const uint64_t *d_rowids, uint64_t n,
                             rhj_tuple *d_out, void *stream);

/* column_proj (Query.cpp:66-74): *sum = sum of col[rowids[i]] mod 2^64. */
int rhj_gather_sum_u64_device(rhj_ctx *ctx, const uint64_t *d_col, const uint64_t *d_rowids, uint64_t n,
                              uint64_t *sum, void *stream);

/* Order-independent digest of a device pair list: *sum / *xr = sum / xor over pairs of
 * mix64(keyR * 0x100000001b3 + keyS)  (same function as orc_pairs_digest). */
int rhj_pairs_digest_device(rhj_ctx *ctx, const rhj_pair *d_pairs, uint64_t n, uint64_t *sum, uint64_t *xr,
                            void *stream);

/* update_intermediate, case 2 (intermediate.cpp:52-66,108-125,162-170): one binding of the join is
 * already in the intermediate (its row-id column is match_col[n_rows]); every result pair whose row
 * id on that side equals match_col[e] yields a new row = old row e + the pair's other row id.  The
 * reference scans all rows once per pair (99 % of the small.work wall time); here it is one join
 * (index-keyed relations) + gathers.  match_on_S = 1 when the already-joined binding is the join's
 * S side.  cols[n_cols] are the live columns of the old intermediate (n_rows each, HOST memory).
 * out_cols[0..n_cols-1] = carried columns, out_cols[n_cols] = the new binding's column, each
 * *out_rows long, in context-owned pinned HOST memory valid until the next call on the context.
 * Row order differs from the reference; the multiset of rows is identical. */
int rhj_intermediate_expand_host(rhj_ctx *ctx, const uint64_t *match_col, uint64_t n_rows, const rhj_pair *pairs,
                                 uint64_t n_pairs, int match_on_S, const uint64_t *const *cols, uint32_t n_cols,
                                 uint64_t **out_cols, uint64_t *out_rows);

/* update_intermediate, case 3 (intermediate.cpp:72-87,130-138,171-180): both bindings are already in
 * the intermediate; row e survives once per result pair equal to (col1[e], col2[e]).
 * out_cols[0..n_cols-1] = the surviving rows of every carried column. */
int rhj_intermediate_filter_host(rhj_ctx *ctx, const uint64_t *col1, const uint64_t *col2, uint64_t n_rows,
                                 const rhj_pair *pairs, uint64_t n_pairs, const uint64_t *const *cols, uint32_t n_cols,
                                 uint64_t **out_cols, uint64_t *out_rows);

/* ---- the query path around the join, device resident ------------------------------ */

/* relList columns (structs.cpp:18-60: read-only host arrays for the life of the process) are uploaded ONCE: the first
 * call with a host column copies it to the device, every later call -- from any context of the process -- returns the
 * same device copy.  *uploaded_bytes (optional) = bytes this call moved over PCIe (0 on a hit).  The cache is keyed by the
 * host address (plus length and a first / middle / last fingerprint, so a recycled address with other contents is uploaded
 * again); a column must not be modified or freed while a query that uses it is running. */
int rhj_column_device(rhj_ctx *ctx, const uint64_t *host_col, uint64_t n, const uint64_t **d_col, uint64_t *uploaded_bytes);
int rhj_column_cache_clear(void);

/* create_relation's row-id de-duplication (structs.cpp:238-241): d_out[0..*count) = the distinct values of d_rowids[n],
 * ascending (the reference's unordered_set order is unspecified).  Row ids must be < n_rows, the relation's row count;
 * d_out needs room for min(n, n_rows) values. */
int rhj_unique_rowids_device(rhj_ctx *ctx, const uint64_t *d_rowids, uint64_t n, uint64_t n_rows, uint64_t *d_out,
                             uint64_t *count, void *stream);

/* Query::execute (Query.h:50, Query.cpp:204-211: run_filters -> run_joins -> column_proj) with row-id lists, relations,
 * join results and the intermediate resident in HBM from the first filter to the last checksum.  A binding refers to a
 * relList (its columns are HOST pointers, uploaded once through rhj_column_device); filters / joins / projections are
 * the reference's filter_info / join_info / proj_info (Query.h:8-33) with `table` = binding index.
 * sums[n_projs] = the projection checksums (Query.cpp:66-74); *empty = 1 when a filter or join left no row (the reference
 * prints NULL for every projection, Query.cpp:226-235).  Rows of the intermediate come out in a different order than the
 * reference's; the multiset is the same and only sums observe it. */
#define RHJ_MAX_BINDINGS 16
typedef struct rhj_q_relation { const uint64_t *const *columns; uint64_t num_tuples, num_columns; } rhj_q_relation;
typedef struct rhj_q_filter { uint32_t binding, column; int32_t op; uint32_t reserved0; uint64_t constant; } rhj_q_filter;
typedef struct rhj_q_join { uint32_t binding1, column1, binding2, column2; } rhj_q_join;
typedef struct rhj_q_proj { uint32_t binding, column; } rhj_q_proj;
typedef struct rhj_query_desc {
    uint32_t n_bindings, n_filters, n_joins, n_projs;
    uint32_t reorder_joins;       /* 0: run the joins as written, like the reference (README.md:63-64); 1: cheapest-first over
                                     the join graph using the filtered row counts (same result rows, same checksums) */
    uint32_t reserved0;
    const rhj_q_relation *bindings;
    const rhj_q_filter *filters;
    const rhj_q_join *joins;
    const rhj_q_proj *projs;
} rhj_query_desc;
typedef struct rhj_query_stats {
    uint64_t h2d_bytes;           /* column uploads of this query (0 once the columns are resident)          */
    uint64_t d2h_bytes;           /* counts + checksums read back                                           */
    uint64_t kernel_launches, joins, join_input_tuples, join_output_pairs, result_rows;
    uint64_t joins_reordered;     /* 1 when reorder_joins changed the order                                  */
} rhj_query_stats;
int rhj_query_execute(rhj_ctx *ctx, const rhj_query_desc *query, uint64_t *sums, int *empty, rhj_query_stats *stats);

/* ---- multi-GPU (no reference equivalent) --------------------------------------- */

/* Groups d_in[n] by destination rank = top log2(world) bits of the join hash (world a power of
 * two <= 256).  d_out[n] is grouped by rank, h_counts[world] (host) the per-rank tuple counts;
 * the caller exchanges the groups (NCCL all-to-all) and joins what it receives with
 * rhj_join_device -- tuples with equal payloads always land on the same rank. */
int rhj_shuffle_partition_device(rhj_ctx *ctx, const rhj_tuple *d_in, uint64_t n, int world,
                                 rhj_tuple *d_out, uint64_t *h_counts, void *stream);

/* Radix plan shared by the sharded joins: destination rank = top rank_bits of the HIGH hash word, local partition = top
 * bits_total bits of the low word, bits_pass1 of them taken in the pass that also separates the destinations. */
typedef struct rhj_shard_plan {
    uint32_t world, rank_bits;                       /* ranks (power of two <= 16), log2(world)        */
    uint32_t bits_total, bits_pass1, bits_pass2;     /* local radix bits: total, fused pass, second pass */
    uint32_t build_is_S;
} rhj_shard_plan;
int rhj_shard_plan_make(uint64_t nR_global, uint64_t nS_global, int world, rhj_shard_plan *plan);

/* Exact (histogram) exchange: the fallback of the pipelined exchange below for skewed / duplicate-heavy inputs.  Pass 1
 * partitions each local shard on (destination rank | sub-digit) into a LOCAL staging buffer, so the data for one
 * destination is one contiguous chunk that is already pass-1 partitioned; the caller all-gathers the histograms, ships the
 * chunks with the copy engines (peer cudaMemcpyAsync over NVLink, no SM time) while the SMs partition the other relation;
 * pass 2 consumes the received chunks as (source, partition) pieces.  Per relation (rel 0 = R, 1 = S), on every rank:
 *   rhj_shardx_begin -> rhj_shardx_pass1_device(rel) -> [all-gather d_hist] -> rhj_shardx_layout_device(rel)
 *   -> [barrier; peer copies of send_cnt[d] tuples from stage+send_off[d] to dest d's buffer+dst_off[d];
 *      barrier] -> rhj_shardx_pass2_device(rel) ... -> rhj_shardx_join_device. */
int rhj_shardx_begin(rhj_ctx *ctx, const rhj_shard_plan *plan, void *stream);
int rhj_shardx_pass1_device(rhj_ctx *ctx, const rhj_shard_plan *plan, int rel, const rhj_tuple *d_in, uint64_t n,
                            rhj_tuple *d_stage, uint64_t *d_hist, void *stream);
int rhj_shardx_layout_device(rhj_ctx *ctx, const rhj_shard_plan *plan, int rank, int rel, const uint64_t *d_all_hist,
                             uint64_t *send_off, uint64_t *send_cnt, uint64_t *dst_off, uint64_t *recv_total,
                             uint64_t *recv_max /* optional: largest recv_total of any rank, equal on all ranks */,
                             void *stream);
int rhj_shardx_pass2_device(rhj_ctx *ctx, const rhj_shard_plan *plan, int rel, const rhj_tuple *d_recv, uint64_t n_recv,
                            void *stream);
int rhj_shardx_join_device(rhj_ctx *ctx, const rhj_shard_plan *plan, rhj_pair *d_out, uint64_t capacity, uint64_t *count,
                           void *stream);

/* Pipelined exchange (the default of bench.py at N > 1): histogram-free, chunked, no host synchronisation and
 * no collective call inside a step.  Every rank cuts each local relation into `chunks` row chunks; pass 1 of a
 * chunk scatters on (destination rank | sub-digit) into fixed-capacity regions; a hand-written copy kernel
 * (TMA bulk copies through a shared-memory ring, peer stores over NVLink / NVSwitch) ships the filled part of
 * every region into the same region of the destination's receive buffer while the SMs partition the next chunk;
 * the destination appends each chunk, as soon as its flags have arrived, to fixed-capacity final partitions;
 * one build/probe/emit pass follows.  Receive buffers, region ends and flags live in one SYMMETRIC block per
 * rank (rhj_pipe_sym_bytes bytes, zero-filled once, mapped into every peer: CUDA IPC / symmetric memory) and
 * are double-buffered by step parity.  Per step, on every rank, with the same epoch = 1, 2, 3, ...:
 *   rhj_pipe_begin -> for rel, chunk: rhj_pipe_pass1_device (compute stream), rhj_pipe_ship_device (copy stream,
 *   behind an event) -> for rel, chunk: rhj_pipe_pass2_device -> rhj_pipe_post_device -> rhj_pipe_join_device.
 * A region that overflows anywhere (skewed or duplicate-heavy data) reaches every rank as RHJ_PIPE_OVERFLOW in
 * *status: all ranks then redo the step through the exact exchange (rhj_shardx_*). */
#define RHJ_MAX_PEERS 16
#define RHJ_PIPE_OVERFLOW 1u   /* a fixed-capacity region overflowed somewhere: redo through rhj_shardx_*   */
#define RHJ_PIPE_TIMEOUT 2u    /* a peer's flag did not arrive within 4 s                                   */
#define RHJ_PIPE_BAD 4u        /* a received region end was out of range                                    */
#define RHJ_PIPE_WIDE 8u       /* 12-byte wire format: a row id did not fit 32 bits (use wire_bytes = 16)   */
typedef struct rhj_pipe_cfg {
    uint32_t world, rank;
    uint32_t chunks;                       /* row chunks per relation, 1..8                                  */
    uint32_t ship_ctas;                    /* CTAs of the copy kernel (0 = default 48)                       */
    uint32_t wire_bytes;                   /* bytes per tuple on the wire: 16 (0 = default), or 12 = {u64 value, u32 row
                                              id} records repacked by the copy kernel -- the caller promises that row ids
                                              fit 32 bits, a wider one is reported as RHJ_PIPE_WIDE                      */
    uint32_t reserved0;
    uint64_t nR_local_max, nS_local_max;   /* rows per rank (upper bound over ranks) of R and S              */
    void *sym[RHJ_MAX_PEERS];              /* base of every rank's symmetric block, valid in this process    */
} rhj_pipe_cfg;
uint64_t rhj_pipe_sym_bytes(const rhj_shard_plan *plan, const rhj_pipe_cfg *cfg);
int rhj_pipe_open(rhj_ctx *ctx, const rhj_shard_plan *plan, const rhj_pipe_cfg *cfg);
int rhj_pipe_begin(rhj_ctx *ctx, uint64_t epoch, void *stream);
int rhj_pipe_pass1_device(rhj_ctx *ctx, int rel, int chunk, const rhj_tuple *d_rows, uint64_t n, void *stream);
int rhj_pipe_ship_device(rhj_ctx *ctx, int rel, int chunk, void *stream);
int rhj_pipe_pass2_device(rhj_ctx *ctx, int rel, int chunk, void *stream);
int rhj_pipe_post_device(rhj_ctx *ctx, void *stream);
int rhj_pipe_join_device(rhj_ctx *ctx, rhj_pair *d_out, uint64_t capacity, uint64_t *count, uint32_t *status, void *stream);

/* ---- introspection for benchmarks ------------------------------------------------------------ */

typedef struct rhj_plan_info {
    uint32_t bits_total, bits_pass1, bits_pass2; /* radix bits: total and per pass (0 = pass skipped) */
    uint32_t build_is_S;                         /* 1 if S (the smaller side) is the build side       */
    uint32_t n_partitions;                       /* 2^bits_total                                      */
    uint32_t n_items;                            /* probe work items of the last join                 */
    uint32_t kernel_launches;                    /* kernels launched by the last join call            */
    uint32_t optimistic_pass1;                   /* bit 0 / 1: pass 1 of the build / probe relation ran without a
                                                    histogram (fixed-capacity regions); bit 2: so did pass 2
                                                    of both (fixed-capacity final partitions)              */
} rhj_plan_info;
int rhj_last_plan(const rhj_ctx *ctx, rhj_plan_info *info);

/* Per-phase device times of the last join call, measured with CUDA events recorded on the
 * launching stream between the kernels (off by default; costs one event record per phase).
 * ms[RHJ_NUM_PHASES], indexed by the RHJ_PHASE_* constants; phases that did not run are 0. */
#define RHJ_PHASE_HIST1 0      /* (1) histogram, pass 1                 */
#define RHJ_PHASE_SCAN1 1      /* (2) prefix sum, pass 1                */
#define RHJ_PHASE_SCATTER1 2   /* (3) partition scatter, pass 1         */
#define RHJ_PHASE_HIST2 3      /* (1) histogram, pass 2                 */
#define RHJ_PHASE_SCAN2 4      /* (2) prefix sum, pass 2                */
#define RHJ_PHASE_SCATTER2 5   /* (3) partition scatter, pass 2         */
#define RHJ_PHASE_PLAN 6       /* work-item planning                    */
#define RHJ_PHASE_JOIN 7       /* (4)+(5) build/probe + fused emit, or the count pass */
#define RHJ_PHASE_SCAN_ITEMS 8 /* (5) prefix sum of per-item counts     */
#define RHJ_PHASE_JOIN_WRITE 9 /* (4)+(5) write pass                    */
#define RHJ_NUM_PHASES 10
int rhj_set_profiling(rhj_ctx *ctx, int on);
int rhj_last_phase_ms(rhj_ctx *ctx, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* RHJ_H */
