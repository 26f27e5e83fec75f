// Large k (k > 256; config 3's "binary top-1000 then rescore"): one sampled pass + one full pass.
//
//  1. sample pass: the ordinary scan over every m-th tile with k_s = 128 gives, per query, the
//     128-th best score v_q among the sampled rows (m = 3k/128, so about 3k corpus rows are
//     expected to score <= v_q).
//  2. collect pass: the same scan kernel in collect mode appends EVERY eligible row with
//     score <= v_q to a per-query buffer (one warp-aggregated atomicAdd per hit group; hits are rare).
//  3. select: one CTA per query sorts its buffer by (score, row) and writes the first k.
// The set collected in (2) is deterministic, so the output is too.  If a buffer overflowed (massive
// ties at v_q) or came up short, the caller falls back to the exact multi-pass cursor path.
#pragma once
#include "common.cuh"

namespace vl {

constexpr int kSelectThreads = 512;
constexpr int kLargeKSample = 32;   // sample list length cap: the register-list fast path of the scan
constexpr int kLargeKThrRank = 64;  // tensor-core sampled pass: the threshold is the 64th best of the merged 16-entry lists

__global__ void extract_thr_kernel(const uint32_t* __restrict__ sample_scores, const uint64_t* __restrict__ sample_rows,
                                   int n_queries, int ks, uint32_t* __restrict__ thr) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_queries) return;
    const size_t last = (size_t)q * ks + (ks - 1);
    thr[q] = (sample_rows[last] == kSentinelRow) ? kSentinelScore : sample_scores[last];
}

// flags: bit 0 = some buffer overflowed, bit 1 = some buffer holds fewer than k rows although its
// threshold was finite.  A non-zero *flags opens the gate of the exact multi-pass pass that the host
// has already enqueued behind this kernel (ScanParams::gate): no host round trip decides it.
__global__ void __launch_bounds__(kSelectThreads)
collect_select_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ count,
                      const uint32_t* __restrict__ thr, uint32_t cap, int k, uint64_t base_row,
                      uint32_t* __restrict__ out_scores, uint64_t* __restrict__ out_rows, int* __restrict__ flags) {
    extern __shared__ __align__(1024) uint8_t smem[];
    unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(smem);
    const int q = blockIdx.x;
    const uint32_t n = count[q];
    if (n > cap) {
        if (threadIdx.x == 0) atomicOr(flags, 1);
        return;
    }
    if (n < (uint32_t)k && thr[q] != kSentinelScore) {
        if (threadIdx.x == 0) atomicOr(flags, 2);
        return;
    }
    uint32_t n_sort = 2;
    while (n_sort < n) n_sort <<= 1;
    for (uint32_t i = threadIdx.x; i < n_sort; i += kSelectThreads)
        s_keys[i] = (i < n) ? keys[(size_t)q * cap + i] : ~0ull;
    __syncthreads();
    for (uint32_t size = 2; size <= n_sort; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = threadIdx.x; i < n_sort / 2; i += kSelectThreads) {
                const uint32_t lo = 2 * i - (i & (stride - 1));
                const uint32_t hi = lo + stride;
                const bool asc = ((lo & size) == 0);
                const unsigned long long a = s_keys[lo], b = s_keys[hi];
                if ((a > b) == asc) {
                    s_keys[lo] = b;
                    s_keys[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < k; i += kSelectThreads) {
        const bool ok = (uint32_t)i < n;
        const unsigned long long key = ok ? s_keys[i] : ~0ull;
        out_scores[(size_t)q * k + i] = ok ? (uint32_t)(key >> 32) : kSentinelScore;
        out_rows[(size_t)q * k + i] = ok ? base_row + (uint32_t)(key & 0xFFFFFFFFull) : kSentinelRow;
    }
}

}  // namespace vl
