// thread_ctx.h -- one CUDA join context per query thread, shared by the drop-in translation units.
//
// The reference enters the join path concurrently from up to NUM_OF_THREADS query threads
// (join.cpp:42-48, MainScheduler.cpp:6-14,23-26), each with its own JobScheduler; an rhj_ctx plays
// that role here (own stream, workspace, pinned result buffers) and is not thread-safe, so every
// query thread lazily creates its own.  RHJ_DEVICE=<n> selects the GPU (default 0).
#ifndef RHJ_HOST_THREAD_CTX_H
#define RHJ_HOST_THREAD_CTX_H

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "rhj.h"

namespace rhj_host {

// Optional host-side accounting: RHJ_HOST_TIMING=1 prints, at exit, the thread-time spent inside
// each drop-in entry point (summed over the query threads).
struct Timing {
    std::atomic<long long> ns[4];
    std::atomic<long long> calls[4];
    const char *names[4] = {"multiRadixHashJoin | Query::execute", "update_intermediate/unzip", "update_intermediate/expand",
                            "update_intermediate/filter"};
    std::chrono::steady_clock::time_point start = std::chrono::steady_clock::now();
    std::atomic<long long> first_ns, last_ns, ctx_ns, ctx_done_ns;
    Timing() {
        for (int i = 0; i < 4; i++) { ns[i] = 0; calls[i] = 0; }
        first_ns = -1; last_ns = 0; ctx_ns = 0; ctx_done_ns = 0;
    }
    long long now_ns() const {
        return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - start).count();
    }
    ~Timing() {
        if (!getenv("RHJ_HOST_TIMING")) return;
        fprintf(stderr, "[rhj host timing] first GPU call at +%.1f ms, last one ended at +%.1f ms, exit at +%.1f ms; "
                        "context creation thread-time %.1f ms\n", first_ns.load() / 1e6, last_ns.load() / 1e6, now_ns() / 1e6,
                ctx_ns.load() / 1e6);
        // what a long-running process would see per workload: the span from the moment the last query thread had its CUDA
        // context to the end of the last call
        fprintf(stderr, "[rhj host timing] query phase after the last context was created: %.1f ms\n",
                (last_ns.load() - ctx_done_ns.load()) / 1e6);
        for (int i = 0; i < 4; i++)
            fprintf(stderr, "[rhj host timing] %-28s calls %6lld  thread-time %9.3f ms\n", names[i], calls[i].load(),
                    ns[i].load() / 1e6);
    }
};
inline Timing &timing() {
    static Timing t;
    return t;
}
struct Scope {
    int slot;
    std::chrono::steady_clock::time_point t0;
    explicit Scope(int s) : slot(s), t0(std::chrono::steady_clock::now()) {
        long long expect = -1;
        timing().first_ns.compare_exchange_strong(expect, timing().now_ns());
    }
    ~Scope() {
        timing().last_ns = timing().now_ns();
        timing().ns[slot] += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
        timing().calls[slot]++;
    }
};

// PCIe accounting of the device-resident query path (Query_execute.cpp): RHJ_HOST_TIMING=1 prints the totals at exit.
struct Pcie {
    std::atomic<unsigned long long> h2d, d2h, queries, launches, joins;
    Pcie() : h2d(0), d2h(0), queries(0), launches(0), joins(0) {}
    void add(const rhj_query_stats &s) {
        h2d += s.h2d_bytes;
        d2h += s.d2h_bytes;
        queries++;
        launches += s.kernel_launches;
        joins += s.joins;
    }
    ~Pcie() {
        if (!getenv("RHJ_HOST_TIMING") || !queries.load()) return;
        fprintf(stderr, "[rhj host timing] %llu queries on the device: %llu joins, %llu kernel launches, H2D %llu bytes (columns, once), "
                        "D2H %llu bytes (%.0f per query)\n", queries.load(), joins.load(), launches.load(), h2d.load(), d2h.load(),
                (double) d2h.load() / (double) queries.load());
    }
};
inline Pcie &pcie() {
    static Pcie p;
    return p;
}

struct ThreadCtx {
    rhj_ctx *ctx = nullptr;
    ThreadCtx() {
        const char *d = getenv("RHJ_DEVICE");
        auto t0 = std::chrono::steady_clock::now();
        int rc = rhj_create(d ? atoi(d) : 0, &ctx);
        timing().ctx_ns += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
        long long done = timing().now_ns(), seen = timing().ctx_done_ns.load();
        while (seen < done && !timing().ctx_done_ns.compare_exchange_weak(seen, done)) {}
        if (rc != RHJ_OK) {
            fprintf(stderr, "rhj_create failed (status %d): the CUDA join needs an sm_100 GPU; there is no CPU path\n", rc);
            exit(EXIT_FAILURE);
        }
    }
    ~ThreadCtx() { rhj_destroy(ctx); }
};

inline rhj_ctx *thread_ctx() {
    static thread_local ThreadCtx t;
    return t.ctx;
}

inline void check(int rc, const char *what) {
    if (rc != RHJ_OK) {
        fprintf(stderr, "%s failed (status %d): %s\n", what, rc, rhj_last_error(thread_ctx()));
        exit(EXIT_FAILURE);
    }
}

}  // namespace rhj_host
#endif
