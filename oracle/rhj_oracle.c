/*
 * rhj_oracle.c -- plain-C restatement of the reference's radix hash join path.
 *
 * TEST INFRASTRUCTURE ONLY (see rhj_oracle.h).  Single-threaded on purpose:
 * the reference's thread fan-out does not change its result, only its speed
 *   - the 8-range split + per-range index lists + (bucket, range, element)
 *     merge of structs.cpp:146-194 is exactly one stable partition;
 *   - the 256 JoinJobs write to private sub-results that are concatenated in
 *     bucket order (Result.cpp:100-121).
 * What IS restated faithfully is everything that decides the output and its
 * order: payload & 0xFF bucketing, smaller-side build with payload % prime
 * chains walked newest-first, R-first pair order, and the 8191-pair page list
 * whose head is the newest page.
 */
#include "rhj_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ref: auxFun.cpp:4-22 -- smallest prime > x, with the reference's special
 * cases (x < 2 -> 2, x == 3 -> 5; x == 2 falls through to 3). */
size_t orc_next_prime(size_t x) {
    if (x < 2) return 2;
    if (x == 3) return 5;
    size_t c = (x & 1) ? x + 2 : x + 1;
    for (;; c += 2) {
        if (c % 3 == 0) continue;
        int prime = 1;
        for (size_t d = 5; d * d <= c; d += 6) {
            if (c % d == 0 || c % (d + 2) == 0) { prime = 0; break; }
        }
        if (prime) return c;
    }
}

/* ref: structs.cpp:144-204; JobScheduler.cpp:149-155 (histogram),
 * 162-177 (prefix sum + index scatter); structs.cpp:183-194 (merge). */
void orc_hash_relation(const orc_tuple *in, uint64_t n, size_t fanout,
                       orc_tuple *out, size_t *hist) {
    const size_t mask = fanout - 1;
    memset(hist, 0, fanout * sizeof(size_t));
    for (uint64_t i = 0; i < n; i++) hist[in[i].payload & mask]++;
    size_t *cursor = (size_t *)malloc(fanout * sizeof(size_t));
    size_t run = 0;
    for (size_t b = 0; b < fanout; b++) { cursor[b] = run; run += hist[b]; }
    for (uint64_t i = 0; i < n; i++) out[cursor[in[i].payload & mask]++] = in[i];
    free(cursor);
}

/* growable pair list */
typedef struct { orc_pair *p; uint64_t n, cap; } pairvec;

static void pv_push(pairvec *v, uint64_t r, uint64_t s) {
    if (v->n == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 1024;
        v->p = (orc_pair *)realloc(v->p, v->cap * sizeof(orc_pair));
    }
    v->p[v->n].keyR = r;
    v->p[v->n].keyS = s;
    v->n++;
}

/* The order in which a consumer meets pairs that were add_result()'ed in the
 * order seq[0..n): pages are consecutive runs of 8191 appends, the newest
 * (possibly partial) page is the list head, older pages follow, and every page
 * is read front to back.  ref: Result.cpp:21-35 (append), 78-84 / 111-121 and
 * intermediate.cpp:151-160 (walk). */
static void page_walk(const orc_pair *seq, uint64_t n, pairvec *dst) {
    if (n == 0) return;
    const uint64_t cap = ORC_PAGE_CAPACITY;
    uint64_t pages = (n + cap - 1) / cap;
    for (uint64_t pg = pages; pg-- > 0;) {
        uint64_t lo = pg * cap;
        uint64_t hi = lo + cap < n ? lo + cap : n;
        for (uint64_t i = lo; i < hi; i++) pv_push(dst, seq[i].keyR, seq[i].keyS);
    }
}

/* ref: Result.cpp:43-76 with the build-side choice of JobScheduler.cpp:186-192.
 * `small`/`big` are the bucket slices; r_is_big tells which one came from R. */
static void join_bucket(const orc_tuple *small, size_t n_small,
                        const orc_tuple *big, size_t n_big,
                        int r_is_big, pairvec *seq) {
    size_t prime = orc_next_prime(n_small);
    int64_t *first = (int64_t *)malloc(prime * sizeof(int64_t));
    int64_t *older = (int64_t *)malloc(n_small * sizeof(int64_t));
    for (size_t i = 0; i < prime; i++) first[i] = -1;
    for (size_t i = 0; i < n_small; i++) {
        size_t h = small[i].payload % prime;
        older[i] = first[h];
        first[h] = (int64_t)i;
    }
    for (size_t j = 0; j < n_big; j++) {
        uint64_t v = big[j].payload;
        for (int64_t k = first[v % prime]; k != -1; k = older[k]) {
            if (small[k].payload != v) continue;
            if (r_is_big) pv_push(seq, big[j].key, small[k].key);
            else          pv_push(seq, small[k].key, big[j].key);
        }
    }
    free(first);
    free(older);
}

int orc_multi_radix_hash_join(const orc_tuple *R, uint64_t nR,
                              const orc_tuple *S, uint64_t nS,
                              orc_pair **out, uint64_t *count) {
    const size_t fanout = (size_t)1 << ORC_HASH_LSB;
    orc_tuple *Rp = (orc_tuple *)malloc((nR ? nR : 1) * sizeof(orc_tuple));
    orc_tuple *Sp = (orc_tuple *)malloc((nS ? nS : 1) * sizeof(orc_tuple));
    size_t *hR = (size_t *)malloc(fanout * sizeof(size_t));
    size_t *hS = (size_t *)malloc(fanout * sizeof(size_t));
    if (!Rp || !Sp || !hR || !hS) return -1;
    orc_hash_relation(R, nR, fanout, Rp, hR);   /* ref: Result.cpp:95 */
    orc_hash_relation(S, nS, fanout, Sp, hS);   /* ref: Result.cpp:96 */

    pairvec merged = {0, 0, 0};   /* the appends into *this, Result.cpp:111-121 */
    pairvec seq = {0, 0, 0};      /* the appends into res[i] */
    size_t begR = 0, begS = 0;
    for (size_t b = 0; b < fanout; b++) {
        if (hR[b] != 0 && hS[b] != 0) {         /* ref: Result.cpp:101 */
            seq.n = 0;
            if (hR[b] >= hS[b])                  /* ref: JobScheduler.cpp:187-188 */
                join_bucket(Sp + begS, hS[b], Rp + begR, hR[b], 1, &seq);
            else                                 /* ref: JobScheduler.cpp:189-190 */
                join_bucket(Rp + begR, hR[b], Sp + begS, hS[b], 0, &seq);
            page_walk(seq.p, seq.n, &merged);
        }
        begR += hR[b];
        begS += hS[b];
    }
    free(seq.p);
    free(Rp); free(Sp); free(hR); free(hS);

    pairvec final = {0, 0, 0};
    page_walk(merged.p, merged.n, &final);
    free(merged.p);
    *count = final.n;
    *out = final.n ? final.p : NULL;
    if (!final.n) free(final.p);
    return 0;
}

void orc_free(void *p) { free(p); }

/* splitmix64 finalizer (a bijection on u64); SURVEY.md section 8d. */
uint64_t orc_mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    return x;
}

void orc_pairs_digest(const orc_pair *p, uint64_t n, uint64_t *sum, uint64_t *xr) {
    uint64_t s = 0, x = 0;
    for (uint64_t i = 0; i < n; i++) {
        uint64_t h = orc_mix64(p[i].keyR * 0x100000001b3ULL + p[i].keyS);
        s += h;
        x ^= h;
    }
    *sum = s;
    *xr = x;
}

/* ref: Query.cpp:94-146.  The reference erases failing rows from a set of all
 * row ids; what survives is {j : col[j] OP c}.  '>' keeps value > c (erase on
 * value <= c, Query.cpp:100), '<' keeps value < c (120), '=' keeps == (136). */
uint64_t orc_filter(const uint64_t *col, uint64_t n, int op, uint64_t c, uint64_t *rowids_out) {
    uint64_t k = 0;
    for (uint64_t j = 0; j < n; j++) {
        uint64_t v = col[j];
        int keep = (op == '>') ? (v > c) : (op == '<') ? (v < c) : (v == c);
        if (keep) rowids_out[k++] = j;
    }
    return k;
}

/* ref: structs.cpp:217-226 */
void orc_gather_tuples(const uint64_t *col, const uint64_t *rowids, uint64_t n, orc_tuple *out) {
    for (uint64_t i = 0; i < n; i++) {
        out[i].key = rowids[i];
        out[i].payload = col[rowids[i]];
    }
}

/* ref: Query.cpp:66-74 */
uint64_t orc_column_sum(const uint64_t *col, const uint64_t *rowids, uint64_t n) {
    uint64_t s = 0;
    for (uint64_t i = 0; i < n; i++) s += col[rowids[i]];
    return s;
}
