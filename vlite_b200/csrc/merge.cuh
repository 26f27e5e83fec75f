// Kernel 2b: device-wide merge of sorted candidate lists.
//
// Input: n_lists lists per query, each k_in entries ascending by (score, row), sentinel padded
// (the scan's per-(split, row-slice) lists, or the all-gathered per-shard top-k of SURVEY.md 8e).
// Output: per query the k_out smallest by (score, global row) -- the canonical order that
// stands in for the reference's torch.topk (vlite/model.py:85).
//
// One warp per query: lane l owns lists l, l+32, ...; every round each lane offers the best head
// among its lists, a shuffle min-reduction on the 96-bit key (score, row) picks the winner, the
// winning lane advances that list's head.  No atomics, so the result is deterministic.
#pragma once
#include "common.cuh"

namespace vl {

constexpr int kMergeWarps = 4;

__device__ __forceinline__ bool key_less(uint32_t s1, uint64_t r1, uint32_t s2, uint64_t r2) {
    return (s1 < s2) || (s1 == s2 && r1 < r2);
}

__global__ void __launch_bounds__(kMergeWarps * 32)
merge_topk_kernel(const uint32_t* __restrict__ scores, const uint64_t* __restrict__ rows, int n_lists,
                  int n_queries, int k_in, int k_out, uint32_t* __restrict__ out_scores,
                  uint64_t* __restrict__ out_rows, int out_stride, int out_offset, size_t score_list_stride,
                  size_t row_list_stride) {
    extern __shared__ uint16_t s_heads[];  // [kMergeWarps][n_lists]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * kMergeWarps + warp;
    if (q >= n_queries) return;
    uint16_t* heads = s_heads + (size_t)warp * n_lists;
    for (int l = lane; l < n_lists; l += 32) heads[l] = 0;
    __syncwarp();

    for (int o = 0; o < k_out; ++o) {
        uint32_t bs = kSentinelScore;
        uint64_t br = kSentinelRow;
        int bl = -1;
        for (int l = lane; l < n_lists; l += 32) {
            const int h = heads[l];
            if (h < k_in) {
                const size_t idx = (size_t)q * k_in + h;
                const uint64_t r = rows[(size_t)l * row_list_stride + idx];
                const uint32_t s = scores[(size_t)l * score_list_stride + idx];
                if (r != kSentinelRow && key_less(s, r, bs, br)) {
                    bs = s;
                    br = r;
                    bl = l;
                }
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const uint32_t os = __shfl_xor_sync(kFullMask, bs, off);
            const uint64_t orow = __shfl_xor_sync(kFullMask, br, off);
            const int ol = __shfl_xor_sync(kFullMask, bl, off);
            if (key_less(os, orow, bs, br)) {
                bs = os;
                br = orow;
                bl = ol;
            }
        }
        // all lanes now agree on (bs, br, bl); rows are unique so there is exactly one owner
        if (bl >= 0 && (bl & 31) == lane) heads[bl] += 1;
        if (lane == 0) {
            out_scores[(size_t)q * out_stride + out_offset + o] = bs;
            out_rows[(size_t)q * out_stride + out_offset + o] = br;
        }
        __syncwarp();
    }
}

// cursor for the next pass of a multi-pass (large k) search: the last key already emitted
__global__ void cursor_update_kernel(const uint32_t* __restrict__ out_scores, const uint64_t* __restrict__ out_rows,
                                     int n_queries, int out_stride, int last_pos, uint64_t base_row,
                                     uint32_t* __restrict__ cur_score, uint32_t* __restrict__ cur_row) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_queries) return;
    const uint64_t r = out_rows[(size_t)q * out_stride + last_pos];
    const bool done = (r == kSentinelRow);  // list exhausted: nothing is eligible any more
    cur_score[q] = done ? kSentinelScore : out_scores[(size_t)q * out_stride + last_pos];
    cur_row[q] = done ? 0xFFFFFFFFu : (uint32_t)(r - base_row);
}

}  // namespace vl
