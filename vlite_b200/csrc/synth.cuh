// Device side of the counter-based synthetic generator (oracle/synth.py is the numpy twin)
// and the integer-pipe micro-benchmarks that give the roofline its POPC denominator.
#pragma once
#include "common.cuh"

namespace vl {

// rows [first, first+n) of a W-byte corpus: words j < W/8 with stride 16, little-endian
__global__ void synth_fill_kernel(uint8_t* __restrict__ rows, int w, uint64_t seed, uint64_t first, uint64_t n,
                                  uint64_t period, uint64_t base_row) {
    const int wpr = w / 8;
    const uint64_t total = n * (uint64_t)wpr;
    uint64_t* dst = reinterpret_cast<uint64_t*>(rows + first * (uint64_t)w);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t row = base_row + first + i / wpr;
        if (period) row %= period;
        dst[i] = synth_word(seed, row, i % wpr, 16);
    }
}

__global__ void synth_docs_kernel(int8_t* __restrict__ docs, int dims, uint64_t seed, uint64_t first_row, uint64_t n) {
    const int wpr = dims / 8;
    const uint64_t total = n * (uint64_t)wpr;
    uint64_t* dst = reinterpret_cast<uint64_t*>(docs);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x)
        dst[i] = synth_word(seed, first_row + i / wpr, i % wpr, 128);
}

// bit(row) = (synth tag of global row) < threshold
__global__ void synth_mask_kernel(uint32_t* __restrict__ mask, uint64_t n_rows, uint64_t n_words, uint64_t seed,
                                  uint64_t base_row, uint64_t threshold) {
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words;
         w += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t bits = 0;
        for (int b = 0; b < 32; ++b) {
            const uint64_t r = w * 32 + b;
            if (r < n_rows) {
                const uint64_t tag = synth_word(seed ^ 0x7461675F73616C74ull, base_row + r, 0, 1) >> 32;
                if (tag < threshold) bits |= 1u << b;
            }
        }
        mask[w] = bits;
    }
}

// ---- integer-pipe micro-benchmarks -------------------------------------------------
// kind 0 POPC, 1 LOP3, 2 IADD, 3 IDP4A, 4 the scan's inner mix (LOP3 xor + POPC + IADD).
constexpr int kMbChains = 8;
template <int KIND>
__global__ void __launch_bounds__(256) microbench_kernel(int iters, uint32_t seed, uint32_t* __restrict__ out) {
    uint32_t a[kMbChains];
    const uint32_t b = seed ^ threadIdx.x, c = seed * 2654435761u + blockIdx.x;
#pragma unroll
    for (int i = 0; i < kMbChains; ++i) a[i] = b + i * 0x9E3779B9u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < kMbChains; ++i) {
            if constexpr (KIND == 0) {
                asm volatile("popc.b32 %0, %0;" : "+r"(a[i]));
            } else if constexpr (KIND == 1) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
            } else if constexpr (KIND == 2) {
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            } else if constexpr (KIND == 3) {
                asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
            } else {
                uint32_t t;
                asm volatile("xor.b32 %0, %1, %2;" : "=r"(t) : "r"(b + it), "r"(c + i));
                asm volatile("popc.b32 %0, %0;" : "+r"(t));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(t));
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < kMbChains; ++i) r ^= a[i];
    if (r == 0x12345678u) out[0] = r;  // keep the chains alive
}

}  // namespace vl
