// ref_harness.cpp -- C entry point around the UNMODIFIED reference join.
//
// TEST INFRASTRUCTURE ONLY.  This file is ours; it is compiled together with
// the reference's own Result.cpp / structs.cpp / JobScheduler.cpp / auxFun.cpp
// where they lie under /root/reference (see oracle/Makefile) into
// oracle/_ref/libref_rhj.so.  No reference source is copied into the repo.
//
// It fills two `relation`s (structs.h:38-49), runs
// Result::multiRadixHashJoin (Result.cpp:90-124) on a process-wide JobScheduler with
// NUM_OF_THREADS workers (JobScheduler.h:11) and flattens the page list in the
// order a consumer walks it (intermediate.cpp:151-160).
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <chrono>

#include "structs.h"
#include "Result.h"
#include "JobScheduler.h"

extern "C" {

int ref_num_threads(void) { return NUM_OF_THREADS; }

// R and S are arrays of {key(rowid), payload(value)} (structs.h:33-36).
// On return *out is a malloc'd array of {keyR,keyS} (nullptr when empty) and
// *seconds the wall time of multiRadixHashJoin alone.
int ref_multi_radix_hash_join(const uint64_t *R, uint64_t nR, const uint64_t *S, uint64_t nS,
                              uint64_t **out, uint64_t *count, double *seconds, int want_pairs) {
    relation relR, relS;                       // ~relation does delete[] tuples
    relR.num_tuples = nR;
    relR.tuples = new tuple[nR ? nR : 1];
    relS.num_tuples = nS;
    relS.tuples = new tuple[nS ? nS : 1];
    memcpy(relR.tuples, R, nR * sizeof(tuple));
    memcpy(relS.tuples, S, nS * sizeof(tuple));

    // ONE scheduler for the life of the process, as in the reference program (a query thread keeps its JobScheduler across
    // all its joins, MainScheduler.cpp:6-14) -- and never stopped: JobScheduler::stop() sets `done` and broadcasts WITHOUT
    // holding queueLock (JobScheduler.cpp:139-145), so a worker that has just tested the flag and not yet reached
    // pthread_cond_wait (JobScheduler.cpp:29-31) misses the wake-up and stop() joins it forever.  A harness that created
    // and stopped a scheduler per call hit that window right behind tiny joins (observed: one hung run of the CPU suite).
    // The idle workers end with the process.
    static JobScheduler *const jsp = [] {
        JobScheduler *p = new JobScheduler;
        p->init(NUM_OF_THREADS);
        return p;
    }();
    JobScheduler &js = *jsp;
    uint64_t n = 0;
    uint64_t *flat = nullptr;
    {
        Result res;
        auto t0 = std::chrono::steady_clock::now();
        res.multiRadixHashJoin(js, relR, relS);
        auto t1 = std::chrono::steady_clock::now();
        if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();

        for (bucket_info *pg = res.head; pg != nullptr; pg = pg->next)
            n += (pg == res.head) ? res.size : res.capacity;
        if (n && want_pairs) {
            flat = (uint64_t *) malloc(n * 2 * sizeof(uint64_t));
            uint64_t at = 0;
            for (bucket_info *pg = res.head; pg != nullptr; pg = pg->next) {
                size_t cnt = (pg == res.head) ? res.size : res.capacity;
                memcpy(flat + 2 * at, &pg[1], cnt * sizeof(key_tuple));
                at += cnt;
            }
        }
    }
    *out = flat;
    *count = n;
    return 0;
}

void ref_free(void *p) { free(p); }

}  // extern "C"
