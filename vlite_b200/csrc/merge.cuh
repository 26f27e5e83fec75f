// Kernel 2b: device-wide merge of sorted candidate lists.
//
// Input: n_lists lists per query, each k_in entries ascending by (score, row), sentinel padded
// (the scan's per-(split, row-slice) lists, or the all-gathered per-shard top-k of SURVEY.md 8e).
// Output: per query the k_out smallest by (score, global row) -- the canonical order that
// stands in for the reference's torch.topk (vlite/model.py:85).
//
// One warp merges up to 128 lists of one query: list heads live in registers (4 per lane), every
// round the lane's best head enters a shuffle min-reduction on the 96-bit key (score, row), the
// owning lane advances that list with one load.  More than 128 lists (Q <= 8 produces one list
// per warp of every CTA: up to 2368) are merged level by level: group g of a level writes its
// k_out winners as list g of the next level.  No atomics, so the result is deterministic.
#pragma once
#include "common.cuh"

namespace vl {

constexpr int kMergeWarps = 4;
constexpr int kMergeListsPerLane = 4;
constexpr int kMergeGroup = 32 * kMergeListsPerLane;  // lists one warp merges per level

__device__ __forceinline__ bool key_less(uint32_t s1, uint64_t r1, uint32_t s2, uint64_t r2) {
    return (s1 < s2) || (s1 == s2 && r1 < r2);
}

// grid = (ceil(Q / kMergeWarps), n_groups)
__global__ void __launch_bounds__(kMergeWarps * 32)
merge_topk_kernel(const uint32_t* __restrict__ scores, const uint64_t* __restrict__ rows, int n_lists,
                  int n_queries, int k_in, int k_out, uint32_t* __restrict__ out_scores,
                  uint64_t* __restrict__ out_rows, int out_stride, int out_offset, size_t score_list_stride,
                  size_t row_list_stride, size_t out_group_stride, const int* __restrict__ gate = nullptr) {
    if (gate != nullptr && *gate == 0) return;   // device-side gate, see ScanParams::gate
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * kMergeWarps + warp;
    if (q >= n_queries) return;
    const int group = blockIdx.y;
    const int list0 = group * kMergeGroup;

    uint32_t cs[kMergeListsPerLane];
    uint64_t cr[kMergeListsPerLane];
    int hd[kMergeListsPerLane] = {};
    auto load = [&](int j) {
        const int l = list0 + j * 32 + lane;
        cs[j] = kSentinelScore;
        cr[j] = kSentinelRow;
        if (l < n_lists && hd[j] < k_in) {
            VL_DCHECK(l >= 0 && hd[j] >= 0 && q < n_queries, 301);
            const size_t idx = (size_t)q * k_in + hd[j];
            cr[j] = rows[(size_t)l * row_list_stride + idx];
            cs[j] = scores[(size_t)l * score_list_stride + idx];
            if (cr[j] == kSentinelRow) cs[j] = kSentinelScore;
        }
    };
#pragma unroll
    for (int j = 0; j < kMergeListsPerLane; ++j) {
        hd[j] = 0;
        load(j);
    }
    uint32_t* os = out_scores + (size_t)group * out_group_stride + (size_t)q * out_stride + out_offset;
    uint64_t* orow = out_rows + (size_t)group * out_group_stride + (size_t)q * out_stride + out_offset;
    for (int o = 0; o < k_out; ++o) {
        uint32_t bs = cs[0];
        uint64_t br = cr[0];
        int bj = 0;
#pragma unroll
        for (int j = 1; j < kMergeListsPerLane; ++j)
            if (key_less(cs[j], cr[j], bs, br)) {
                bs = cs[j];
                br = cr[j];
                bj = j;
            }
        int bl = lane;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const uint32_t s2 = __shfl_xor_sync(kFullMask, bs, off);
            const uint64_t r2 = __shfl_xor_sync(kFullMask, br, off);
            const int l2 = __shfl_xor_sync(kFullMask, bl, off);
            if (key_less(s2, r2, bs, br) || (s2 == bs && r2 == br && l2 < bl)) {
                bs = s2;
                br = r2;
                bl = l2;
            }
        }
        // (bs, br, bl) now agree on every lane; rows are unique, so a real key has one owner
        if (bl == lane && br != kSentinelRow) {
#pragma unroll
            for (int j = 0; j < kMergeListsPerLane; ++j)
                if (j == bj) {
                    hd[j] += 1;
                    load(j);
                }
        }
        if (lane == 0) {
            os[o] = bs;
            orow[o] = br;
        }
    }
}

// Local -> global row numbers for a shard that holds several appended segments (base
// offsets are fixed at append time).  Segment i covers local rows [seg_local[i], seg_local[i+1]) and
// global rows seg_global[i] + (local - seg_local[i]); both tables ascend, so a shard's local order is
// its global order and the local top-k stays the global top-k of that shard.
__global__ void map_rows_kernel(uint64_t* __restrict__ rows, uint64_t n, const uint64_t* __restrict__ seg_local,
                                const uint64_t* __restrict__ seg_global, int n_seg) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t r = rows[i];
    if (r == kSentinelRow) return;
    int lo = 0, hi = n_seg - 1;                  // last segment whose start is <= r
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (seg_local[mid] <= r) lo = mid; else hi = mid - 1;
    }
    rows[i] = seg_global[lo] + (r - seg_local[lo]);
}

// cursor for the next pass of a multi-pass (large k) search: the last key already emitted
__global__ void cursor_update_kernel(const uint32_t* __restrict__ out_scores, const uint64_t* __restrict__ out_rows,
                                     int n_queries, int out_stride, int last_pos, uint64_t base_row,
                                     uint32_t* __restrict__ cur_score, uint32_t* __restrict__ cur_row,
                                     const int* __restrict__ gate = nullptr) {
    if (gate != nullptr && *gate == 0) return;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_queries) return;
    const uint64_t r = out_rows[(size_t)q * out_stride + last_pos];
    const bool done = (r == kSentinelRow);  // list exhausted: nothing is eligible any more
    cur_score[q] = done ? kSentinelScore : out_scores[(size_t)q * out_stride + last_pos];
    cur_row[q] = done ? 0xFFFFFFFFu : (uint32_t)(r - base_row);
}

}  // namespace vl
