// The ONE exchange step of the row-sharded path for one-process-per-GPU runs, without
// a collective library call: every rank's k candidates per query are STORED straight into every
// peer's gather buffer over NVLink (CUDA IPC mappings of cudaMalloc'd blocks, opened once), followed
// by a release flag; the final merge kernel of each rank acquires the flags of all ranks and merges.
// Two small launches (push, merge) on the search stream and no host round trip -- the latency of the
// exchange is a peer store + a flag, not a latency-bound ncclAllGather (~25 us for 123 KB).
//
// Buffer of rank r (identical layout on every rank):
//     slots[2][world][slot_bytes]   parity = epoch & 1 (a rank can be at most one exchange ahead of a
//                                   peer: its next push needs its own merge, which needs that peer's
//                                   previous push), slot = [scores u32 x Q*k | rows u64 x Q*k]
//     flags[2][world]               epoch number written by the pushing rank after its data
// The reference has no multi-GPU path; this replaces nothing there.
#pragma once
#include "merge.cuh"

namespace vl {

constexpr int kExchangeMaxWorld = 16;
constexpr int kPushThreads = 1024;

struct PeerPtrs {
    uint8_t* p[kExchangeMaxWorld];
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// grid = world CTAs: CTA r copies this rank's block into peer r's slot, then raises this rank's flag there
__global__ void __launch_bounds__(kPushThreads)
exchange_push_kernel(const uint4* __restrict__ src, uint64_t n16, PeerPtrs peers, uint64_t slot_off, uint64_t flag_off,
                     uint32_t epoch) {
    uint8_t* base = peers.p[blockIdx.x];
    uint4* dst = reinterpret_cast<uint4*>(base + slot_off);
    for (uint64_t i = threadIdx.x; i < n16; i += kPushThreads) dst[i] = src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys(reinterpret_cast<uint32_t*>(base + flag_off), epoch);
}

// the merge of merge.cuh behind an acquire of every rank's flag (spins at most ~2 minutes, then reports)
__global__ void __launch_bounds__(kMergeWarps * 32)
exchange_merge_kernel(const uint32_t* flags, int world, uint32_t epoch, int* timeout_flag_host,
                      const uint32_t* __restrict__ scores, const uint64_t* __restrict__ rows, int n_queries, int k,
                      uint32_t* __restrict__ out_scores, uint64_t* __restrict__ out_rows, size_t score_list_stride,
                      size_t row_list_stride) {
    if (threadIdx.x < world) {
        const long long t0 = clock64();
        while (ld_acquire_sys(flags + threadIdx.x) != epoch) {
            if (clock64() - t0 > 240000000000ll) {   // ~2 minutes: a peer may be busy on its host for seconds
                *timeout_flag_host = 1;
                break;
            }
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * kMergeWarps + warp;
    if (q >= n_queries) return;
    // world <= 16 lists: one head per lane
    uint32_t cs = kSentinelScore;
    uint64_t cr = kSentinelRow;
    int hd = 0;
    auto load = [&]() {
        cs = kSentinelScore;
        cr = kSentinelRow;
        if (lane < world && hd < k) {
            const size_t idx = (size_t)q * k + hd;
            cr = __ldcg(rows + (size_t)lane * row_list_stride + idx);
            cs = __ldcg(scores + (size_t)lane * score_list_stride + idx);
            if (cr == kSentinelRow) cs = kSentinelScore;
        }
    };
    load();
    for (int o = 0; o < k; ++o) {
        uint32_t bs = cs;
        uint64_t br = cr;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const uint32_t s2 = __shfl_xor_sync(kFullMask, bs, off);
            const uint64_t r2 = __shfl_xor_sync(kFullMask, br, off);
            if (key_less(s2, r2, bs, br)) {
                bs = s2;
                br = r2;
            }
        }
        if (cr == br && br != kSentinelRow) {   // global rows are unique: one owner
            hd += 1;
            load();
        }
        if (lane == 0) {
            out_scores[(size_t)q * k + o] = bs;
            out_rows[(size_t)q * k + o] = br;
        }
    }
}

}  // namespace vl
