// Result.cpp -- drop-in replacement for the reference's Result.cpp (the file that holds
// Result::multiRadixHashJoin, Result.cpp:90-124).  Compile THIS file instead of the reference's
// Result.cpp, together with the reference's other, unmodified sources and headers, and link
// librhj.so: Query.cpp:185-186 and intermediate.cpp:146-183 then run unchanged on top of the
// CUDA join.  See INTEGRATION.md and the Makefile next to this file.
//
// Kept from the reference's contract (Result.h:19-38):
//   - Result() starts empty: head == nullptr, size == capacity == 8191           (Result.cpp:10-14)
//   - after the join `head` is the newest of a list of malloc'd 128 KiB pages
//     [bucket_info *next][8191 x key_tuple], only the head page partial (`size`)  (Result.cpp:21-35)
//   - pairs are (rowidR, rowidS) whichever side was the build side               (Result.cpp:66-69)
//   - ~Result() free()s page by page                                               (Result.cpp:127-133)
// Replaced: hash_relation x2, 256 JoinJobs, the barrier and the serial addAll merge
// (Result.cpp:93-121) by one rhj_join_host call; the JobScheduler argument is not used.
// No CPU fallback: if the GPU library fails the process aborts with the library's message,
// matching the reference's assert/exit error style (structs.cpp:19, JobScheduler.cpp:24-27).
#include <cstdio>
#include <cstdlib>

#include "Result.h"      // the reference's header (include path points at the reference tree)
#include "rhj.h"
#include "thread_ctx.h"

#define BUCKET_SIZE (128 * 1024)   // Result.cpp:7

static_assert(sizeof(tuple) == sizeof(rhj_tuple), "tuple must stay {u64 key; u64 payload}");
static_assert(sizeof(key_tuple) == sizeof(rhj_pair), "key_tuple must stay {u64 keyR; u64 keyS}");

using rhj_host::thread_ctx;

Result::Result() {
    capacity = (BUCKET_SIZE - sizeof(bucket_info)) / sizeof(tuple);
    size = capacity;
    head = nullptr;
}

bool Result::isEmpty() { return head == nullptr; }

// Kept for callers outside the join (none in the reference today): same page discipline.
void Result::add_result(uint64_t key1, uint64_t key2) {
    if (size == capacity) {
        auto page = (bucket_info *) malloc(BUCKET_SIZE);
        page->next = head;
        head = page;
        size = 0;
    }
    auto slots = (key_tuple *) &head[1];
    slots[size].keyR = key1;
    slots[size].keyS = key2;
    size++;
}

void Result::addAll(bucket_info *node, size_t n) {
    auto kt = (key_tuple *) &node[1];
    for (size_t i = 0; i < n; i++) add_result(kt[i].keyR, kt[i].keyS);
}

// JoinJob::run (JobScheduler.cpp:186-192) still references this symbol; nothing schedules a
// JoinJob any more, and there is deliberately no CPU implementation behind it.
void Result::join_buckets(relation_info *, relation_info *, size_t, size_t, size_t, size_t, bool) {
    fprintf(stderr, "Result::join_buckets: the CPU bucket join is not part of the CUDA build\n");
    abort();
}

void Result::multiRadixHashJoin(JobScheduler &, relation &relR, relation &relS) {
    rhj_host::Scope timer(0);
    rhj_ctx *ctx = thread_ctx();
    const rhj_pair *pairs = nullptr;
    uint64_t count = 0;
    int rc = rhj_join_host(ctx, (const rhj_tuple *) relR.tuples, relR.num_tuples, (const rhj_tuple *) relS.tuples,
                           relS.num_tuples, &pairs, &count);
    rhj_host::check(rc, "rhj_join_host");
    uint64_t head_size = capacity;
    head = (bucket_info *) rhj_pairs_to_pages(pairs, count, &head_size);
    if (count && head == nullptr) {  // the page list could not be allocated (the reference's add_result would have crashed)
        fprintf(stderr, "Result::multiRadixHashJoin: out of host memory for %llu result pairs\n", (unsigned long long) count);
        exit(EXIT_FAILURE);
    }
    size = head_size;
}

Result::~Result() {
    while (head != nullptr) {
        bucket_info *page = head;
        head = head->next;
        free(page);
    }
}
