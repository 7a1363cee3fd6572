/*
 * rhj_oracle.h -- CPU restatement of the reference radix hash join.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may link or load it, and only as the checker.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here
 * against (a) the reference's own code compiled from /root/reference into
 * oracle/_ref/libref_rhj.so (exact page-walk order, not only the multiset),
 * (b) the committed fixtures under tests/golden/ that were generated from
 * that library, and (c) small/small.result through the query-level oracle.
 *
 * All "ref:" citations are file:line in the reference tree.
 */
#ifndef RHJ_ORACLE_H
#define RHJ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ref: structs.h:33-36 -- `key` is the row id, `payload` is the join value. */
typedef struct { uint64_t key; uint64_t payload; } orc_tuple;
/* ref: Result.h:9-12 */
typedef struct { uint64_t keyR; uint64_t keyS; } orc_pair;

#define ORC_HASH_LSB 8                       /* ref: Result.cpp:5  */
#define ORC_PAGE_BYTES (128 * 1024)          /* ref: Result.cpp:7  */
#define ORC_PAGE_CAPACITY ((ORC_PAGE_BYTES - 8) / 16) /* ref: Result.cpp:11 -> 8191 */

/* ref: auxFun.cpp:4-22 */
size_t orc_next_prime(size_t x);

/* ref: structs.cpp:144-204 (+ JobScheduler.cpp:149-177): stable partition of
 * `in[n]` on payload & (fanout-1).  `out[n]` receives the partitioned tuples,
 * `hist[fanout]` the per-bucket counts. */
void orc_hash_relation(const orc_tuple *in, uint64_t n, size_t fanout,
                       orc_tuple *out, size_t *hist);

/* ref: Result.cpp:90-124.  Returns 0 and a malloc'd array of `*count` pairs in
 * the order a consumer sees when it walks the reference's page list from
 * `head` (intermediate.cpp:151-160).  `*out` is NULL when `*count` is 0
 * (ref: Result.cpp:16-18, head == nullptr).  Caller frees with orc_free. */
int orc_multi_radix_hash_join(const orc_tuple *R, uint64_t nR,
                              const orc_tuple *S, uint64_t nS,
                              orc_pair **out, uint64_t *count);

void orc_free(void *p);

/* Order-independent multiset digest of a pair list: (count, sum of
 * mix64(keyR * 0x100000001b3 + keyS) mod 2^64, xor of the same).  Used by the
 * full-size property tests. */
uint64_t orc_mix64(uint64_t x);
void orc_pairs_digest(const orc_pair *p, uint64_t n, uint64_t *sum, uint64_t *xr);

/* ---- neighbours of the join on the query path ---- */

/* ref: Query.cpp:94-146: rows j of col[n] that satisfy `col[j] OP c`
 * (OP in '>' '<' '='), ascending row id.  Returns the survivor count. */
uint64_t orc_filter(const uint64_t *col, uint64_t n, int op, uint64_t c, uint64_t *rowids_out);

/* ref: structs.cpp:217-226: tuples[i] = {rowids[i], col[rowids[i]]}. */
void orc_gather_tuples(const uint64_t *col, const uint64_t *rowids, uint64_t n, orc_tuple *out);

/* ref: Query.cpp:66-74: sum of col[rowids[i]] mod 2^64. */
uint64_t orc_column_sum(const uint64_t *col, const uint64_t *rowids, uint64_t n);

#ifdef __cplusplus
}
#endif
#endif
