// Kernel 3: rescoring of the binary scan's candidates against int8 document embeddings.
//
// North-star piece (3).  The reference has no rescore (nearest code: the unused cos_sim,
// vlite/utils.py:205-208); expected values are defined by oracle/search.py::rescore
// ("parity unpinned").  Per query the C candidate rows are scattered 1 KB gathers, so the
// kernel is gather-bandwidth bound, not a GEMM: one CTA per query, the query staged in
// shared memory, one warp per candidate row with 128-bit coalesced loads, DP4A (int8 query,
// exact int32) or DFMA (fp32 query: int8 x fp32 products are exact in fp64 and the 1024-term sum
// is accumulated in fp64, so the only rounding is the final one to float32 -- 6e-8 relative, for
// cancelling sums too; north_star asks for 1e-5), then an in-CTA bitonic sort by (score descending,
// row ascending) and the first k written out.  The fp64 pipe (64 DFMA/clk/SM on B200) stays under
// the gather time.
#pragma once
#include <math_constants.h>

#include "common.cuh"

namespace vl {

constexpr int kRescoreThreads = 256;
constexpr int kRescoreMaxCand = 2048;

struct RescoreParams {
    const int8_t* docs;        // n_docs x dims, LOCAL row index
    const void* queries;       // Q x dims, float32 or int8
    const uint64_t* cand;      // Q x n_cand GLOBAL rows, sentinel padded
    float* out_scores;         // Q x k
    uint64_t* out_rows;        // Q x k
    uint64_t base_row;
    uint64_t n_docs;
    int dims, n_cand, n_sort, k;  // n_sort = next pow2 >= n_cand
};

// QMODE 0: fp32 queries, 1: int8 queries, 2: gather probe (measurement only: the same loads, sort and
// stores with the arithmetic reduced to one XOR per word -- the floor any rescore of these candidates
// has to pay for moving Q x C x D bytes out of HBM)
template <int QMODE>
__global__ void __launch_bounds__(kRescoreThreads) rescore_kernel(const RescoreParams p) {
    constexpr bool QI8 = QMODE != 0;
    constexpr bool PROBE = QMODE == 2;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int q = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dims = p.dims;
    float* s_score = reinterpret_cast<float*>(smem);                       // n_sort
    uint64_t* s_row = reinterpret_cast<uint64_t*>(s_score + p.n_sort);     // n_sort (8B aligned: n_sort even)
    uint8_t* s_q = reinterpret_cast<uint8_t*>(s_row + p.n_sort);           // dims * (1 | 8): int8 query, or the fp32 query widened to fp64

    const size_t qbytes = (size_t)dims * (QI8 ? 1 : 4);
    {
        const uint4* src = reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(p.queries) + (size_t)q * qbytes);
        uint4* dst = reinterpret_cast<uint4*>(s_q);
        if constexpr (QI8) {
            for (int i = threadIdx.x; i < (int)(qbytes / 16); i += kRescoreThreads) dst[i] = src[i];
        } else {
            // fp32 query, widened to fp64 once: the 8 double2 that face one 16-byte doc chunk are stored
            // chunk-minor ([e][chunk]) so that the 32 lanes of a warp read consecutive 16 B (no bank conflicts)
            const int n_chunks = dims / 16;
            const float2* srcf = reinterpret_cast<const float2*>(src);
            double2* dstd = reinterpret_cast<double2*>(s_q);
            for (int i = threadIdx.x; i < dims / 2; i += kRescoreThreads) {
                const float2 f = srcf[i];
                dstd[(i & 7) * n_chunks + (i >> 3)] = make_double2((double)f.x, (double)f.y);
            }
            (void)dst;
        }
    }
    for (int i = threadIdx.x; i < p.n_sort; i += kRescoreThreads) {
        s_score[i] = -CUDART_INF_F;
        s_row[i] = kSentinelRow;
    }
    __syncthreads();

    const int chunks = dims / 16;  // 16 int8 per 128-bit load
    // one 1 KB gather per candidate is latency-bound: every warp keeps U rows (2 x 128-bit loads per
    // lane each) in flight before it starts the arithmetic of the first
    constexpr int U = 4;
    auto accumulate = [&](const uint4& v, int ch, double& facc, int& iacc) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        if constexpr (PROBE) {
            iacc ^= (int)(w[0] ^ w[1] ^ w[2] ^ w[3]) + ch;
        } else if constexpr (QI8) {
            const uint4 qv = reinterpret_cast<const uint4*>(s_q)[ch];
            iacc = __dp4a((int)w[0], (int)qv.x, iacc);
            iacc = __dp4a((int)w[1], (int)qv.y, iacc);
            iacc = __dp4a((int)w[2], (int)qv.z, iacc);
            iacc = __dp4a((int)w[3], (int)qv.w, iacc);
        } else {
            const double2* qd = reinterpret_cast<const double2*>(s_q) + ch;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // int8 -> double without the conversion unit: byte b of (x ^ 0x80808080) is d + 128 in
                // [0, 255]; PRMT drops it under the exponent of 2^52, so the double is 2^52 + 128 + d exactly
                const uint32_t u = w[j] ^ 0x80808080u;
                const double2 qa = qd[(2 * j) * chunks], qb = qd[(2 * j + 1) * chunks];
                constexpr double kMagic = 4503599627370624.0;   // 2^52 + 128
                facc = fma(__hiloint2double(0x43300000, (int)__byte_perm(u, 0u, 0x4440)) - kMagic, qa.x, facc);
                facc = fma(__hiloint2double(0x43300000, (int)__byte_perm(u, 0u, 0x4441)) - kMagic, qa.y, facc);
                facc = fma(__hiloint2double(0x43300000, (int)__byte_perm(u, 0u, 0x4442)) - kMagic, qb.x, facc);
                facc = fma(__hiloint2double(0x43300000, (int)__byte_perm(u, 0u, 0x4443)) - kMagic, qb.y, facc);
            }
        }
    };
    for (int c0 = warp * U; c0 < p.n_cand; c0 += (kRescoreThreads / 32) * U) {
        uint64_t grow[U];
        const uint4* d[U];
        uint4 v[U][2];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int c = c0 + u;
            grow[u] = c < p.n_cand ? p.cand[(size_t)q * p.n_cand + c] : kSentinelRow;
            const uint64_t lrow = grow[u] - p.base_row;
            if (grow[u] != kSentinelRow && lrow >= p.n_docs) grow[u] = kSentinelRow;  // not on this shard
            d[u] = reinterpret_cast<const uint4*>(p.docs + (grow[u] == kSentinelRow ? 0 : lrow) * (uint64_t)dims);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int ch = lane + 32 * h;
                v[u][h] = (grow[u] != kSentinelRow && ch < chunks) ? ldg128_stream(d[u] + ch) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (grow[u] == kSentinelRow) continue;   // warp-uniform
            double facc = 0.0;
            int iacc = 0;
#pragma unroll
            for (int h = 0; h < 2; ++h)
                if (lane + 32 * h < chunks) accumulate(v[u][h], lane + 32 * h, facc, iacc);
            for (int ch = lane + 64; ch < chunks; ch += 32) accumulate(ldg128_stream(d[u] + ch), ch, facc, iacc);
            if constexpr (QI8) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) iacc += __shfl_xor_sync(kFullMask, iacc, off);
                facc = (double)iacc;
            } else {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) facc += __shfl_xor_sync(kFullMask, facc, off);
            }
            VL_DCHECK(c0 + u < p.n_sort && grow[u] - p.base_row < p.n_docs, 401);
            if (lane == 0) {
                s_score[c0 + u] = (float)facc;
                s_row[c0 + u] = grow[u];
            }
        }
    }
    __syncthreads();

    // bitonic sort, order: score descending, then row ascending (sentinels sink to the end)
    auto before = [](float s1, uint64_t r1, float s2, uint64_t r2) {
        return (s1 > s2) || (s1 == s2 && r1 < r2);
    };
    const int n = p.n_sort;
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < n / 2; i += kRescoreThreads) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool asc = ((lo & size) == 0);
                const float s1 = s_score[lo], s2 = s_score[hi];
                const uint64_t r1 = s_row[lo], r2 = s_row[hi];
                const bool in_order = !before(s2, r2, s1, r1);  // lo already not after hi
                if (asc != in_order) {
                    s_score[lo] = s2;
                    s_score[hi] = s1;
                    s_row[lo] = r2;
                    s_row[hi] = r1;
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < p.k; i += kRescoreThreads) {
        const bool ok = i < n && s_row[i] != kSentinelRow;
        p.out_scores[(size_t)q * p.k + i] = ok ? s_score[i] : -CUDART_INF_F;
        p.out_rows[(size_t)q * p.k + i] = ok ? s_row[i] : kSentinelRow;
    }
}

}  // namespace vl
