// ref_trace_hook.cpp -- TEST INFRASTRUCTURE ONLY (see ref_trace_impl.cpp).
//
// Defines Result::multiRadixHashJoin (Result.h:30) for the traced build of the
// reference program: forwards to the reference implementation and appends one
// line per join to $RHJ_TRACE_FILE:
//   nR nS count in_digest out_sum out_xor
// in_digest  = order-independent digest of both input relations,
// out_sum/xor = order-independent digest of the result pairs (orc_pairs_digest).
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "structs.h"
#include "Result.h"
#include "JobScheduler.h"

extern "C" void rhj_trace_call_reference(void *res, void *js, void *relR, void *relS);

static uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    return x;
}

static uint64_t relation_digest(const relation &r, uint64_t salt) {
    uint64_t s = 0;
    for (uint64_t i = 0; i < r.num_tuples; i++)
        s += mix64(mix64(r.tuples[i].key + salt) ^ r.tuples[i].payload);
    return s;
}

void Result::multiRadixHashJoin(JobScheduler &js, relation &relR, relation &relS) {
    static std::mutex mu;
    uint64_t din = relation_digest(relR, 1) * 31 + relation_digest(relS, 2);
    rhj_trace_call_reference(this, &js, &relR, &relS);
    uint64_t n = 0, sum = 0, xr = 0;
    for (bucket_info *pg = head; pg != nullptr; pg = pg->next) {
        size_t cnt = (pg == head) ? size : capacity;
        auto kt = (key_tuple *) &pg[1];
        for (size_t i = 0; i < cnt; i++) {
            uint64_t h = mix64(kt[i].keyR * 0x100000001b3ULL + kt[i].keyS);
            sum += h;
            xr ^= h;
        }
        n += cnt;
    }
    const char *path = getenv("RHJ_TRACE_FILE");
    if (path) {
        std::lock_guard<std::mutex> g(mu);
        FILE *f = fopen(path, "a");
        if (f) {
            fprintf(f, "%llu %llu %llu %llu %llu %llu\n", (unsigned long long) relR.num_tuples,
                    (unsigned long long) relS.num_tuples, (unsigned long long) n, (unsigned long long) din,
                    (unsigned long long) sum, (unsigned long long) xr);
            fclose(f);
        }
    }
}
