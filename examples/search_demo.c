/* Plain-C use of libvlite_b200.so (no Python, no torch): build a synthetic shard on the GPU, search
 * it for a row that is known to be there, print the top 5.
 *
 *   gcc -std=c11 -Iinclude examples/search_demo.c -Lvlite_b200/lib -lvlite_b200 \
 *       -Wl,-rpath,$PWD/vlite_b200/lib -o /tmp/search_demo && /tmp/search_demo [rows]
 */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "vlite_b200.h"

#define CHECK(call)                                                              \
    do {                                                                         \
        int rc_ = (call);                                                        \
        if (rc_ != VL_OK) {                                                      \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, vl_last_error()); \
            return 1;                                                            \
        }                                                                        \
    } while (0)

int main(int argc, char** argv) {
    const uint64_t n = argc > 1 ? strtoull(argv[1], NULL, 10) : 1000000ull;
    const uint64_t needle = n / 3;
    enum { W = 128, K = 5 };
    if (vl_device_count() < 1) {
        fprintf(stderr, "no sm_100 device: %s has no CPU fallback\n", vl_version());
        return 2;
    }
    vl_corpus* c = NULL;
    CHECK(vl_corpus_create(0, W, n, &c));
    CHECK(vl_corpus_fill_synthetic(c, /*seed=*/7, /*first=*/0, n, /*period=*/0));

    uint8_t query[W];
    CHECK(vl_corpus_read(c, needle, 1, query)); /* the needle row itself ... */
    query[0] ^= 0x81;                           /* ... with two bits flipped */

    uint32_t scores[K];
    uint64_t rows[K];
    CHECK(vl_search(c, query, 1, K, VL_METRIC_HAMMING, NULL, scores, rows));
    printf("%s: %" PRIu64 " rows x %d B, top %d of 1 query (HAMMING)\n", vl_version(), n, W, K);
    for (int i = 0; i < K; ++i) printf("  #%d row %" PRIu64 " distance %u\n", i, rows[i], scores[i]);

    CHECK(vl_corpus_tombstone(c, &needle, 1)); /* delete = one bit in the live mask */
    CHECK(vl_search(c, query, 1, K, VL_METRIC_HAMMING, NULL, scores, rows));
    printf("after deleting row %" PRIu64 ": best is row %" PRIu64 " at distance %u\n", needle, rows[0], scores[0]);
    const int ok = scores[0] > 2;
    CHECK(vl_corpus_destroy(c));
    return ok ? 0 : 3;
}
