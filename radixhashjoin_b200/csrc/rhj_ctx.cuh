// rhj_ctx.cuh -- the context object behind `rhj_ctx*` and the helpers shared by the translation
// units of librhj.so (rhj_api.cu: the join; rhj_query.cu: the join's neighbours on the query path).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/rhj.h"
#include "rhj_device.cuh"

using namespace rhj;

static_assert(sizeof(rhj_tuple) == sizeof(Tup), "tuple layout");
static_assert(sizeof(rhj_pair) == sizeof(Pair), "pair layout");

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

// indices into the zeroed scalar block (u64 units)
enum Scalar {
    kScWork0 = 0,   // work counter of the first join kernel
    kScWork1 = 1,   // work counter of the second join kernel (write pass)
    kScCursor = 2,  // FUSED output cursor / final count
    kScNItems = 3,
    kScTotal = 4,   // COUNT_THEN_WRITE total
    kScErr = 5,
    kScDigSum = 6,
    kScDigXor = 7,
    kScFilt = 8,
    kScOverflow = 9,  // optimistic pass 1: a partition outgrew its fixed-capacity region
    kScHoles = 10,       // positional emit: reserved output slots left without a match
    kScPipeStatus = 11,  // pipelined exchange: RHJ_PIPE_* bits of this step, own and received
    kScLeft = 12,        // k_join_pos: items left to the ranked kernel
    kScCount = 16
};

struct rhj_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    std::string err;
    bool hist_agg = false;
    bool optimistic = true;   // skip the pass-1 histogram when a sampled histogram says it is safe (RHJ_NO_OPT=1 disables)
    bool force_optimistic = false;  // RHJ_FORCE_OPT=1 (tests): take the optimistic path even when the sample says skewed
    bool optimistic2 = true;  // also skip the pass-2 histogram (fixed-capacity final partitions; RHJ_NO_OPT2=1 disables)
    int opt2_skip = 0;        // joins left to run without the pass-2 shortcut after it overflowed (duplicate-heavy data)
    // The sample only PREDICTS whether the histogram-free layouts will fit; an overflow is detected and redone exactly
    // either way.  After two sampled joins of the same shape agreed and succeeded, the next joins of that shape reuse the
    // decision without sampling (one kernel, one D2H and one host synchronisation less per join); an overflow, another shape
    // or 64 joins bring the sample back.
    struct {
        u64 nB = 0, nP = 0;
        bool opt[2] = {false, false}, poisson[2] = {false, false};
        int streak = 0, age = 0;
    } trust, pending;
    bool trust_sample = true;  // RHJ_NO_TRUST=1 samples every join
    // positional emit (k_join<FUSED, POS>): one output slot per probe tuple, holes closed afterwards.  Wins when (almost) every
    // probe tuple matches (foreign-key style joins); a join that left more than 1/64 holes switches it off for the next 16.
    bool positional = true;    // RHJ_NO_POS=1 disables
    int pos_skip = 0;
    int join_pos_v = 3;        // variant of k_join_pos (RHJ_JOIN_POS_V): 3 (default) = hash kept in a register + next item claimed early; 1, 0 = without
    bool join_lean = true;     // positional emitter = k_join_pos + leftover launch (RHJ_JOIN_LEAN=0: the r02 kernel k_join<FUSED, POS>)
    DevBuf sample;            // sampled pass-1 histogram
    int scatter_mode = 0;     // 0 staged per-thread stores, 1 TMA bulk stores (RHJ_SCATTER_MODE)
    int shard_scatter_mode = 1;  // same choice for pass 1 of the exact sharded exchange (local staging): bulk stores were
                                 // measured faster for its 1024-digit runs (RHJ_SHARD_SCATTER_MODE)

    DevBuf bufA, bufB;        // pass-1 / pass-2 partitioned tuples (build side first, then probe side)
    DevBuf tiles;             // TileDesc tables of the second pass (both relations)
    DevBuf bufB2;             // sharded join: pass-2 output of relation S (the relations arrive separately)
    DevBuf shard_meta;        // sharded join: local offsets, piece tables, ship matrix
    DevBuf zero;              // hist1 | hist2 | scalars   (memset to 0 per call)
    DevBuf meta;              // offsets, cursors, tile tables
    DevBuf items, item_cnt, item_off;
    DevBuf items_left;        // work items k_join_pos hands to the ranked kernel
    DevBuf filt_cnt, filt_off, filt_tmp;
    DevBuf inR, inS, outP;    // device staging of the host entry point
    DevBuf pin[2], pout[2], pA, pB;      // pipelined host join: probe chunk in (x2), result out (x2), chunk partitions
    cudaStream_t s_in = nullptr, s_out = nullptr;    // pipelined host join: H2D and D2H streams
    cudaEvent_t ev_in[2] = {}, ev_cmp[2] = {}, ev_out[2] = {};
    uint64_t host_chunk = (uint64_t) 1 << 24;        // probe tuples per pipelined chunk (RHJ_HOST_CHUNK)
    DevBuf iu_col, iu_pairs, iu_A, iu_B, iu_ep, iu_out;  // update_intermediate staging (rhj_query.cu)
    void *h_iu = nullptr;     // pinned host result columns of the intermediate update
    size_t h_iu_cap = 0;
    // staged upload of PAGEABLE host inputs (the reference's relation::tuples are `new tuple[]`): a pinned ring the host
    // fills with a few memcpy threads while the copy engine drains it (rhj_api.cu: upload_host)
    void *stage_pin = nullptr;
    static constexpr size_t kStageSlot = (size_t) 32 << 20;
    static constexpr int kStageSlots = 4;
    cudaEvent_t stage_ev[kStageSlots] = {};
    int stage_next = 0;
    unsigned stage_threads = 8;   // memcpy threads per staged slice (RHJ_STAGE_THREADS), at most half the host's CPUs
    void *h_out = nullptr;    // pinned host result of rhj_join_host
    size_t h_out_cap = 0;
    u64 *h_scalars = nullptr; // pinned, kScCount u64

    // state left by the partition + plan phase for the emit phase
    struct {
        bool valid = false;
        bool counted = false;
        const Tup *build = nullptr, *probe = nullptr;
        const u64 *offB = nullptr, *offP = nullptr;   // partition starts
        const u64 *endB = nullptr, *endP = nullptr;   // partition ends (off + 1 unless the layout has fixed-capacity regions)
        u32 nparts = 0;
        u32 item_cap = 0;
        int build_is_S = 0;
        u64 count = 0;
    } cur;
    rhj_plan_info info{};
    u64 shard_n[3] = {0, 0, 0};            // sharded join: tuples received per slot
    const Tup *shard_recv[3] = {nullptr, nullptr, nullptr};
    bool shard_poisson[3] = {false, false, false};  // the received pass-1 partition sizes look like hashed distinct keys
    u64 shard_cap[3] = {0, 0, 0};                   // > 0: the slot's final partitions lie in fixed-capacity regions of this size
    u64 shard_count = 0;                            // pairs emitted by the joins of this sharded step so far
    bool shard_optimistic2 = true;                  // RHJ_NO_SHARD_OPT2=1 disables the histogram-free second pass
    u32 shard_opt2_world = 2;                       // ... which is on by default up to this many ranks (RHJ_SHARD_OPT2_WORLD)

    // pipelined exchange (rhj_pipe_*): wiring + layout of the symmetric blocks, fixed at rhj_pipe_open
    struct PipeState {
        bool open = false;
        rhj_shard_plan plan{};
        u32 world = 1, rank = 0, chunks = 1, ship_ctas = 48, stages = 8, stage_bytes = 8192, wire = 16;
        u64 nmax[2] = {0, 0};          // rows per rank (upper bound) of R / S
        u64 chunk_rows[2] = {0, 0};    // rows per chunk
        u64 cap1[2] = {0, 0};          // capacity of one (chunk, destination, sub-digit) region, tuples
        u64 cap2[2] = {0, 0};          // capacity of one final partition, tuples
        void *sym[16] = {};            // every rank's symmetric block (sym[rank] = the own one)
        u64 off_recv[2][2] = {}, off_end[2][2] = {}, off_flag = 0, off_status = 0, sym_bytes = 0;  // byte offsets
        DevBuf stage[2];               // local staging of R / S: chunks * ndig * cap1 tuples + one dump tile
        DevBuf cursors;                // [2][chunks * ndig] pass-1 cursors
        DevBuf segs;                   // [2][chunks] segment tables of what arrived
        DevBuf done;                   // arrival counter of the copy kernel (lives outside the per-step zero block)
        u64 epoch = 0;
        bool ship_smem_set = false;
    } pipe;

    // optional per-phase timing (rhj_set_profiling)
    bool profiling = false;
    static constexpr int kMaxMarks = 24;
    cudaEvent_t ev[kMaxMarks] = {};
    int mark_phase[kMaxMarks] = {};
    int nmarks = 0;
};

template <typename F>
inline void for_each_buf(rhj_ctx *c, F f) {
    DevBuf *bufs[] = {&c->bufA, &c->bufB, &c->tiles, &c->sample, &c->bufB2, &c->shard_meta, &c->zero, &c->meta, &c->items, &c->items_left, &c->item_cnt, &c->item_off, &c->filt_cnt,
                      &c->filt_off, &c->filt_tmp, &c->inR, &c->inS, &c->outP, &c->pin[0], &c->pin[1], &c->pout[0], &c->pout[1],
                      &c->pA, &c->pB, &c->iu_col, &c->iu_pairs, &c->iu_A,
                      &c->iu_B, &c->iu_ep, &c->iu_out, &c->pipe.stage[0], &c->pipe.stage[1], &c->pipe.cursors, &c->pipe.segs,
                      &c->pipe.done};
    for (DevBuf *b : bufs) f(*b);
}

inline int fail(rhj_ctx *c, int code, const char *what, cudaError_t e = cudaSuccess) {
    if (c) {
        c->err = what;
        if (e != cudaSuccess) {
            c->err += ": ";
            c->err += cudaGetErrorString(e);
        }
    }
    return code;
}

#define CK(call)                                                           \
    do {                                                                   \
        cudaError_t e_ = (call);                                           \
        if (e_ != cudaSuccess) return fail(ctx, RHJ_ERR_CUDA, #call, e_);  \
    } while (0)

// Grows a device buffer geometrically (never shrinks): the contest workload calls the join ~100
// times per thread with varying sizes, and every cudaFree/cudaMalloc pair is a device-wide sync.
inline int ensure(rhj_ctx *ctx, DevBuf &b, size_t bytes) {
    if (bytes <= b.cap) return RHJ_OK;
    CK(cudaSetDevice(ctx->device));
    if (b.p) CK(cudaFree(b.p));
    b.p = nullptr;
    size_t want = std::max(bytes, std::min(2 * b.cap, b.cap + ((size_t) 1 << 30)));
    want = std::max(want, (size_t) 1 << 16);
    want = (want + 255) & ~(size_t) 255;
    b.cap = 0;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess && want > bytes) {  // the slack did not fit: retry with the exact size
        cudaGetLastError();
        want = (bytes + 255) & ~(size_t) 255;
        e = cudaMalloc(&b.p, want);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, RHJ_ERR_NOMEM, "cudaMalloc workspace", e);
    }
    b.cap = want;
    return RHJ_OK;
}

// Same policy for the context-owned pinned host result blocks (pinning is slow: ~0.3 s / GiB).
inline int ensure_pinned(rhj_ctx *ctx, void **p, size_t *cap, size_t bytes) {
    if (bytes <= *cap) return RHJ_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    size_t want = std::max(bytes, std::min(2 * *cap, *cap + ((size_t) 1 << 30)));
    want = std::max(want, (size_t) 1 << 20);
    *cap = 0;
    cudaError_t e = cudaHostAlloc(p, want, cudaHostAllocDefault);
    if (e != cudaSuccess && want > bytes) {
        cudaGetLastError();
        want = bytes;
        e = cudaHostAlloc(p, want, cudaHostAllocDefault);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, RHJ_ERR_NOMEM, "cudaHostAlloc result", e);
    }
    *cap = want;
    return RHJ_OK;
}

inline void mark(rhj_ctx *ctx, cudaStream_t st, int phase) {
    if (!ctx->profiling || ctx->nmarks >= rhj_ctx::kMaxMarks) return;
    if (!ctx->ev[ctx->nmarks]) cudaEventCreate(&ctx->ev[ctx->nmarks]);
    cudaEventRecord(ctx->ev[ctx->nmarks], st);
    ctx->mark_phase[ctx->nmarks++] = phase;
}


// `stream` is the caller's cudaStream_t; NULL is CUDA's (legacy) default stream, exactly as in the
// CUDA runtime API -- the caller's preceding work on that stream is what our kernels must follow.
inline cudaStream_t pick(rhj_ctx *, void *stream) { return (cudaStream_t) stream; }


