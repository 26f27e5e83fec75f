/* CPU restatement of the reference's scan + top-k.  TEST INFRASTRUCTURE ONLY.
 *
 * Never linked into or called by the product (vlite_b200/); used by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg as the checker.
 * Parity: pinned against golden vectors produced by the reference's own code
 * (tests/golden/make_golden.py); see tests/test_oracle.py.
 *
 * Scores (paths relative to /root/reference):
 *   metric 0, HAMMING : popcount(q XOR e)  == np.sum(unpackbits(q) != unpackbits(e))
 *                       -- vlite/model.py:71-74 (hamming_distance), vlite/index.py:30
 *   metric 1, XORSUM  : sum_j (q_j XOR e_j) on byte VALUES
 *                       -- vlite/model.py:82 (the live torch.sum(q.int() ^ E.int(), dim=1))
 * Top-k: k smallest, ties by ascending row index (canonical order; the
 * reference's torch.topk at vlite/model.py:85 leaves tie order unspecified),
 * k clamped to the number of valid rows like vlite/model.py:84.
 *
 * Build: gcc -O3 -march=native -fopenmp -shared -fPIC oracle_scan.c -o liboracle_scan.so
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SENT_SCORE 0xFFFFFFFFu
#define SENT_ROW   0xFFFFFFFFFFFFFFFFull

static inline uint32_t score_hamming(const uint8_t* a, const uint8_t* b, int w) {
    uint32_t s = 0;
    int j = 0;
    for (; j + 8 <= w; j += 8) {
        uint64_t x, y;
        memcpy(&x, a + j, 8);
        memcpy(&y, b + j, 8);
        s += (uint32_t)__builtin_popcountll(x ^ y);
    }
    for (; j < w; ++j) s += (uint32_t)__builtin_popcount((unsigned)(a[j] ^ b[j]));
    return s;
}

static inline uint32_t score_xorsum(const uint8_t* a, const uint8_t* b, int w) {
    uint32_t s = 0;
    for (int j = 0; j < w; ++j) s += (uint32_t)(a[j] ^ b[j]);
    return s;
}

/* Full score vector of one query: out[n]. */
void oracle_scores(const uint8_t* corpus, uint64_t n, int w, const uint8_t* q, int metric,
                   uint32_t* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i)
        out[i] = metric == 0 ? score_hamming(q, corpus + (uint64_t)i * w, w)
                             : score_xorsum(q, corpus + (uint64_t)i * w, w);
}

/* Insert (s, r) into a list sorted ascending by (score, row); rows arrive in
 * ascending order so an equal score goes AFTER existing equals. */
static inline void insert_sorted(uint32_t* ls, uint64_t* lr, int k, int* fill, uint32_t s, uint64_t r) {
    int n = *fill;
    if (n == k) {
        if (s >= ls[k - 1]) return;
        n = k - 1;
    }
    int p = n;
    while (p > 0 && ls[p - 1] > s) { ls[p] = ls[p - 1]; lr[p] = lr[p - 1]; --p; }
    ls[p] = s;
    lr[p] = r;
    if (*fill < k) (*fill)++;
}

/* Batched search.  mask: optional bit per row (bit row&31 of word row>>5), 1 = row
 * is eligible.  Outputs out_score[nq*k], out_row[nq*k] padded with sentinels;
 * rows are reported as base_row + local row.  Returns 0. */
int oracle_search(const uint8_t* corpus, uint64_t n, int w, const uint8_t* queries, int nq, int k,
                  int metric, const uint32_t* mask, uint64_t base_row,
                  uint32_t* out_score, uint64_t* out_row) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int qi = 0; qi < nq; ++qi) {
        const uint8_t* q = queries + (size_t)qi * w;
        uint32_t* ls = out_score + (size_t)qi * k;
        uint64_t* lr = out_row + (size_t)qi * k;
        int fill = 0;
        for (uint64_t i = 0; i < n; ++i) {
            if (mask && !((mask[i >> 5] >> (i & 31)) & 1u)) continue;
            const uint8_t* e = corpus + i * (uint64_t)w;
            uint32_t s = metric == 0 ? score_hamming(q, e, w) : score_xorsum(q, e, w);
            insert_sorted(ls, lr, k, &fill, s, base_row + i);
        }
        for (int j = fill; j < k; ++j) { ls[j] = SENT_SCORE; lr[j] = SENT_ROW; }
    }
    return 0;
}

/* Same search with the row loop parallelised (for nq smaller than the core
 * count): per-thread lists over contiguous row ranges, merged in range order. */
int oracle_search_rowpar(const uint8_t* corpus, uint64_t n, int w, const uint8_t* queries, int nq,
                         int k, int metric, const uint32_t* mask, uint64_t base_row,
                         uint32_t* out_score, uint64_t* out_row) {
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    uint32_t* ts = (uint32_t*)malloc((size_t)nt * k * sizeof(uint32_t));
    uint64_t* tr = (uint64_t*)malloc((size_t)nt * k * sizeof(uint64_t));
    int* tf = (int*)malloc((size_t)nt * sizeof(int));
    for (int qi = 0; qi < nq; ++qi) {
        const uint8_t* q = queries + (size_t)qi * w;
#pragma omp parallel num_threads(nt)
        {
            int t = 0;
#ifdef _OPENMP
            t = omp_get_thread_num();
#endif
            uint64_t lo = n * (uint64_t)t / nt, hi = n * (uint64_t)(t + 1) / nt;
            int fill = 0;
            for (uint64_t i = lo; i < hi; ++i) {
                if (mask && !((mask[i >> 5] >> (i & 31)) & 1u)) continue;
                const uint8_t* e = corpus + i * (uint64_t)w;
                uint32_t s = metric == 0 ? score_hamming(q, e, w) : score_xorsum(q, e, w);
                insert_sorted(ts + (size_t)t * k, tr + (size_t)t * k, k, &fill, s, base_row + i);
            }
            tf[t] = fill;
        }
        uint32_t* ls = out_score + (size_t)qi * k;
        uint64_t* lr = out_row + (size_t)qi * k;
        int fill = 0;
        for (int t = 0; t < nt; ++t)      /* ranges ascend with t => rows still ascend */
            for (int j = 0; j < tf[t]; ++j)
                insert_sorted(ls, lr, k, &fill, ts[(size_t)t * k + j], tr[(size_t)t * k + j]);
        for (int j = fill; j < k; ++j) { ls[j] = SENT_SCORE; lr[j] = SENT_ROW; }
    }
    free(ts); free(tr); free(tf);
    return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
