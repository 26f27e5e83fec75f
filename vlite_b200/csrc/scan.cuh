// Kernel 1+2a: fused scan (XOR + POPC / DP4A) + metadata/tombstone mask + per-warp top-k.
//
// Replaces EmbeddingModel.search's  sum(q.int() ^ E.int(), dim=1) + topk
// (reference vlite/model.py:82-85) for a batch of queries over the HBM-resident
// packed corpus.
//
// Shape of the computation (M rows x Q queries x W/4 words, integer pipe, never
// tensor cores):
//   grid  = (query tiles, row splits).  A CTA owns QT = C*WQ queries (staged in
//           shared memory once) and a contiguous range of 256-row corpus tiles.
//   warps = 8 per CTA, two CTAs per SM.  One elected thread streams tiles global ->
//           shared into a multi-stage mbarrier ring (full/empty barriers): the rows
//           by 2-D tensor loads through a CUtensorMap of the corpus
//           (cp.async.bulk.tensor.2d, SASS UTMALDG, hardware swizzle), the tile's
//           32 B of live mask + 32 B of filter mask by 1-D bulk copies (UBLKCP).
//           No warp touches global memory in the loop.  Warp (wq, wr): query
//           group wq (C queries in a register tile), row slice wr of every tile.
//   lane  = one corpus row per lane (R rows per step), 128-bit LDS.  The TMA
//           swizzle stores logical 16 B chunk j of row r at physical chunk
//           j ^ f(r), so the stride-W row reads are bank-conflict free while all
//           lanes work on the same logical chunk (query reads are broadcasts).
//   score = HAMMING: carry-save adders (LOP3) + POPC + IMAD; XORSUM: LOP3 + IDP4A.
//   top-k = each (warp, query) keeps a sorted k-list in shared memory and its
//           k-th best score in a register; the hot loop compares the minimum of
//           a query's R scores with that threshold.  On a hit the flagged rows
//           enter the list by key (score << 32 | row) -- the (score, row)
//           tie-break whatever the arrival order.  Device-wide per-query bounds
//           (tau, candidate histogram) let every CTA prune rows early.
//   out   = every (split, row slice) list goes to a partial buffer that the
//           merge kernel (merge.cuh) reduces with the same (score, row) order.
#pragma once
#include "common.cuh"

namespace vl {

constexpr int kTileRows = 256;      // rows of one TMA box
constexpr int kMaxStageRows = 1024;  // a pipeline stage holds 1, 2 or 4 boxes; capacity is padded to this
constexpr int kScanWarps = 8;
constexpr int kScanThreads = kScanWarps * 32;
constexpr int kWarpListMaxK = 128;
constexpr uint32_t kMaskedBias = 0x40000000u;
constexpr int kHistFine = 1024, kHistCoarse = 32, kHistWords = kHistFine + kHistCoarse;

struct ScanParams {
    const uint8_t* rows;     // corpus, n_rows x W, 256-byte aligned
    const uint32_t* live;    // live mask, padded to whole tiles, bits >= n_rows are 0
    const uint32_t* filter;  // optional filter mask, same layout (or nullptr)
    const uint8_t* queries;  // n_queries x W
    uint32_t* part_scores;   // [n_lists][n_queries][k]
    uint64_t* part_rows;
    uint64_t base_row;  // global index of local row 0
    uint32_t n_rows;
    uint32_t n_tiles;
    int n_queries;
    int k;
    int stages;
    int evict_first;  // 1: corpus is read once -> stream it past L2
    // optional per-query cursor (multi-pass large k): only rows whose key (score, local row) is
    // strictly greater than (cur_score[q], cur_row[q]) are eligible
    const uint32_t* cur_score;
    const uint32_t* cur_row;
    // per-query upper bound on the final k-th best score, shared by every CTA (init 0xFFFFFFFF).
    // A full list publishes its k-th score with atomicMin; rows scoring above the bound cannot be
    // in the final top-k, so dropping them early changes no result (only prunes insert work).
    uint32_t* tau;
    // per-query score histogram of the rows that entered some list (fine[1024] then coarse[32] per
    // query, bucket = min(score >> hist_shift, 1023)).  Every counted row is a distinct eligible
    // corpus row, so "k rows counted at or below bucket X" proves the final k-th best lies in or
    // below bucket X: a device-wide bound far tighter than any single CTA's own k-th best.
    uint32_t* hist;
    int hist_shift;
    // tile sampling (large-k threshold estimate): the kernel visits tiles 0, stride, 2*stride, ...
    // n_tiles counts the VISITED tiles; 1 = every tile
    uint32_t tile_stride;
    // collect mode (large k): instead of keeping k-lists, every eligible row with score <= fixed
    // per-query threshold collect_thr[q] is appended to collect_keys[q][..] (score << 32 | local row)
    const uint32_t* collect_thr;
    unsigned long long* collect_keys;
    uint32_t* collect_count;
    uint32_t collect_cap;
    uint32_t one;  // == 1, opaque to ptxas: keeps "acc += popc" on IMAD (FMA pipe) instead of IADD3 (ALU pipe)
    // optional device-side gate: when set and *gate == 0 the whole launch is a no-op.  Lets the host
    // ENQUEUE a fallback pass whose need is only known on the device (large k: search_host.cuh).
    const int* gate;
    // fused final merge (1-2 queries, the latency path): every CTA merges its warps' lists, stores one
    // list of keys per query in fuse_keys[split][q][k], takes a ticket; the LAST CTA to finish merges all
    // splits and writes the final top-k -- one launch instead of scan + two merge levels.
    uint32_t poll_ns, poll_steady;   // tuning (0 = defaults): histogram poll interval and how many polls keep it
    int q_offset;                    // tensor-core scan: first query of this launch (big batches go out in several launches)
    int k_bound;                     // tensor-core scan: rank the SHARED admission bound stands for (0 = k).  > k: the caller
                                     // will merge the k-entry lists to the k_bound best of their union (sampled pass of large k)
    unsigned long long* trace;       // optional (tuning): %globaltimer stamps of CTA (0,0) of the tensor-core scan
    unsigned long long* fuse_keys;   // nullptr = off
    unsigned int* fuse_ticket;       // zero before the launch; the last CTA resets it
    uint32_t* out_scores;            // final [q * out_stride + e]
    uint64_t* out_rows;
    int out_stride;
};

// Warp-level k-way merge of sorted key lists (ascending, unique keys, ~0 = exhausted): lane l owns
// lists l, l + 32, ... (LPL of them) with their heads in registers; every round the warp's smallest
// head is emitted through put(o, key) on lane 0 and its owner advances.  get(list, pos) fetches.
template <int LPL, class Get, class Put>
__device__ __forceinline__ void warp_merge_keys(int n_lists, int k, int lane, Get get, Put put) {
    unsigned long long head[LPL];
    int hd[LPL];
#pragma unroll
    for (int j = 0; j < LPL; ++j) {
        hd[j] = 0;
        const int l = lane + 32 * j;
        head[j] = (l < n_lists) ? get(l, 0) : ~0ull;
    }
    for (int o = 0; o < k; ++o) {
        unsigned long long best = head[0];
        int bj = 0;
#pragma unroll
        for (int j = 1; j < LPL; ++j)
            if (head[j] < best) {
                best = head[j];
                bj = j;
            }
        unsigned long long w = best;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) w = min(w, __shfl_xor_sync(kFullMask, w, off));
        if (best == w && w != ~0ull) {   // keys are unique: exactly one lane owns the winner
#pragma unroll
            for (int j = 0; j < LPL; ++j)
                if (j == bj) {
                    hd[j] += 1;
                    head[j] = (hd[j] < k) ? get(lane + 32 * j, hd[j]) : ~0ull;
                }
        }
        if (lane == 0) put(o, w);
    }
}

// ---- warp-wide selection network for k <= 32: keys are (score << 32 | local row), unique per row ----
__device__ __forceinline__ uint64_t warp_sort32(uint64_t key, int lane) {
    // bitonic sort, ascending across the 32 lanes
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const uint64_t other = __shfl_xor_sync(kFullMask, key, stride);
            const bool take_min = (((lane & size) == 0) == ((lane & stride) == 0));
            key = take_min ? min(key, other) : max(key, other);
        }
    }
    return key;
}
__device__ __forceinline__ uint64_t warp_bitonic_merge32(uint64_t key, int lane) {
    // input: a bitonic sequence across the lanes; output: ascending
#pragma unroll
    for (int stride = 16; stride > 0; stride >>= 1) {
        const uint64_t other = __shfl_xor_sync(kFullMask, key, stride);
        key = ((lane & stride) == 0) ? min(key, other) : max(key, other);
    }
    return key;
}
// Merge up to 32 new candidates (one per lane, ~0 = none) into the warp's sorted k-list (k <= 32)
// in one go: sort the newcomers, pair them in reverse with the list (the element-wise minimum is a
// bitonic sequence holding the 32 smallest keys of the union), bitonic-merge, keep the first k.
// Returns the new k-th best score.  Ordering is by the key itself, so ties break on the row index
// no matter in which order rows arrive.
__device__ __noinline__ uint32_t warp_list_merge32(uint32_t* ls, uint32_t* li, int k, uint64_t cand, int lane) {
    uint64_t cur = (lane < k) ? (((uint64_t)ls[lane] << 32) | li[lane]) : ~0ull;
    unsigned have = __ballot_sync(kFullMask, cand != ~0ull);
    if (__popc(have) <= 8) {
        // few newcomers (the steady state is one): slide each into the register-held list --
        // a vote, a shuffle-up and a select per candidate instead of a 20-stage sorting network
        while (have) {
            const int src = __ffs(have) - 1;
            have &= have - 1;
            const uint64_t key = __shfl_sync(kFullMask, cand, src);
            const int pos = __popc(__ballot_sync(kFullMask, cur < key));   // keys are unique
            const uint64_t up = __shfl_up_sync(kFullMask, cur, 1);
            cur = (lane > pos) ? up : (lane == pos ? key : cur);
        }
    } else {
        cand = warp_sort32(cand, lane);
        const uint64_t rev = __shfl_sync(kFullMask, cand, 31 - lane);
        cur = warp_bitonic_merge32(min(cur, rev), lane);
    }
    if (lane < k) {
        ls[lane] = (uint32_t)(cur >> 32);
        li[lane] = (uint32_t)cur;
    }
    __syncwarp();
    return __shfl_sync(kFullMask, (uint32_t)(cur >> 32), k - 1);
}

// Insert (s, idx) into the warp's sorted list (ascending score; an equal score goes after the
// existing equals because rows arrive in ascending index order).  Returns the new threshold.
__device__ __noinline__ uint32_t warp_list_insert(uint32_t* ls, uint32_t* li, int k, uint32_t s, uint32_t idx,
                                                  int lane) {
    int pos = 0;
    for (int base = 0; base < k; base += 32) {
        int e = base + lane;
        bool le = (e < k) && (ls[e] <= s);
        pos += __popc(__ballot_sync(kFullMask, le));
    }
    // entries [pos, k-1) move up one slot, highest chunk first; the last entry drops out
    for (int top = k - 1; top > pos; top -= 32) {
        int e = top - lane;
        bool act = e > pos;
        uint32_t a = 0, b = 0;
        if (act) {
            a = ls[e - 1];
            b = li[e - 1];
        }
        __syncwarp();
        if (act) {
            ls[e] = a;
            li[e] = b;
        }
        __syncwarp();
    }
    VL_DCHECK(pos >= 0 && pos < k, 101);
    if (lane == 0) {
        ls[pos] = s;
        li[pos] = idx;
    }
    __syncwarp();
    return ls[k - 1];
}

template <int W, int C, int WQ, int TPS>
struct ScanCfg {
    static constexpr int WR = kScanWarps / WQ;
    static constexpr int TR = kTileRows * TPS;            // rows per pipeline stage (TPS TMA boxes)
    static constexpr int MASKW = TR / 32;                 // mask words per stage
    static constexpr int RPW = TR / WR;                   // rows of a stage one warp scores
#ifndef VL_RMAX
#define VL_RMAX 4
#endif
    static constexpr int R = RPW >= 32 * VL_RMAX ? VL_RMAX : RPW / 32;  // rows per lane per step
    static constexpr int STEPS = RPW / (32 * R);
    static constexpr int QT = C * WQ;                     // queries per CTA
    static constexpr int CHUNKS = W / 16;
    static constexpr int TILE_BYTES = TR * W;
    static constexpr int STAGE_BYTES = TILE_BYTES + 2 * MASKW * 4;
    static_assert(W == 32 || W == 64 || W == 128 || W == 256, "row width");
    static_assert(WQ == 1 || WQ == 2 || WQ == 4 || WQ == 8, "warps along queries");
    static_assert(TPS == 1 || TPS == 2 || TPS == 4, "boxes per stage");
    static size_t smem_bytes(int stages, int k) {
        return (size_t)stages * STAGE_BYTES + (size_t)QT * W + (size_t)kScanWarps * C * k * 8 +
               (size_t)QT * 8 + (size_t)kScanWarps * C * 4 + (size_t)stages * 16;
    }
};

// 8 words (two 16 B chunks) of one (row, query) pair.
//  HAMMING: carry-save adders (LOP3 0x96 = xor3, 0xE8 = majority) compress 7 of the 8 XOR words
//  into ones/twos/fours words, so the pair costs 4 POPC (XU pipe, 16 lanes/clk/SM) + 16 LOP3
//  (ALU pipe, 64 lanes/clk/SM) instead of 8 POPC + 8 LOP3: both integer pipes end up evenly
//  loaded, twice the pair rate of the plain XOR+POPC loop.  The weighted adds go to IMAD (FMA pipe).
//  XORSUM: one IDP4A per word (byte-value sum), no compression needed.
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
template <int METRIC>
__device__ __forceinline__ uint32_t score_chunk_pair(const uint4& ra, const uint4& rb, const uint4& qa,
                                                     const uint4& qb, uint32_t acc, uint32_t one) {
    if constexpr (METRIC == kHamming) {
        const uint32_t x0 = ra.x ^ qa.x, x1 = ra.y ^ qa.y, x2 = ra.z ^ qa.z, x3 = ra.w ^ qa.w;
        const uint32_t x4 = rb.x ^ qb.x, x5 = rb.y ^ qb.y, x6 = rb.z ^ qb.z, x7 = rb.w ^ qb.w;
        const uint32_t s1 = xor3(x0, x1, x2), c1 = maj3(x0, x1, x2);
        const uint32_t s2 = xor3(x3, x4, x5), c2 = maj3(x3, x4, x5);
        const uint32_t s3 = xor3(s1, s2, x6), c3 = maj3(s1, s2, x6);
        const uint32_t tw = xor3(c1, c2, c3), fo = maj3(c1, c2, c3);
        // all four weighted adds as IMAD: the FMA pipe idles while the ALU pipe (LOP3) is the bound
        acc = mad_u32(__popc(s3), one, acc);
        acc = mad_u32(__popc(x7), one, acc);
        acc = mad_u32(__popc(tw), 2u, acc);
        acc = mad_u32(__popc(fo), 4u, acc);
        return acc;
    } else {
        acc = score_chunk<METRIC>(ra, qa, acc);
        return score_chunk<METRIC>(rb, qb, acc);
    }
}

template <int W, int METRIC, int C, int WQ, int TPS>
#ifndef VL_MINBLOCKS
#define VL_MINBLOCKS 2
#endif
__global__ void __launch_bounds__(kScanThreads, VL_MINBLOCKS)
scan_topk_kernel(const __grid_constant__ CUtensorMap tmap, const ScanParams p) {
    using Cfg = ScanCfg<W, C, WQ, TPS>;
    if (p.gate != nullptr && *p.gate == 0) return;   // uniform over the grid
    constexpr int WR = Cfg::WR, RPW = Cfg::RPW, R = Cfg::R, STEPS = Cfg::STEPS, QT = Cfg::QT;
    constexpr int CHUNKS = Cfg::CHUNKS, TILE_BYTES = Cfg::TILE_BYTES, TR = Cfg::TR, MASKW = Cfg::MASKW;
    static_assert(CHUNKS % 2 == 0, "rows are scored two 16 B chunks at a time");

    extern __shared__ __align__(1024) uint8_t smem[];  // swizzled TMA tiles want 1024 B alignment
    const int stages = p.stages;
    const int k = p.k;
    uint8_t* s_tiles = smem;                                                        // stages x TILE_BYTES
    uint32_t* s_masks = reinterpret_cast<uint32_t*>(smem + (size_t)stages * TILE_BYTES);  // stages x 16 words
    uint8_t* s_queries = reinterpret_cast<uint8_t*>(s_masks + stages * 2 * MASKW);
    uint32_t* s_lscore = reinterpret_cast<uint32_t*>(s_queries + QT * W);           // [8][C][k]
    uint32_t* s_lidx = s_lscore + kScanWarps * C * k;
    uint32_t* s_cursor = s_lidx + kScanWarps * C * k;                                // [QT][2]
    uint32_t* s_thr = s_cursor + QT * 2;                                             // [8][C]
    uint64_t* s_full = reinterpret_cast<uint64_t*>(s_thr + kScanWarps * C);
    uint64_t* s_empty = s_full + stages;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int qtile = blockIdx.x;
    const int split = blockIdx.y;
    const int n_splits = gridDim.y;
    const uint32_t tile_begin = (uint32_t)(((uint64_t)p.n_tiles * split) / n_splits);
    const uint32_t tile_end = (uint32_t)(((uint64_t)p.n_tiles * (split + 1)) / n_splits);
    const uint32_t my_tiles = tile_end - tile_begin;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap);
        for (int s = 0; s < stages; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], kScanWarps);
        }
        mbar_fence_init();
    }
    // stage this CTA's queries (zero-fill past n_queries)
    {
        const int q0 = qtile * QT;
        const uint4* src = reinterpret_cast<const uint4*>(p.queries);
        uint4* dst = reinterpret_cast<uint4*>(s_queries);
        for (int i = threadIdx.x; i < QT * CHUNKS; i += kScanThreads) {
            int q = q0 + i / CHUNKS;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (q < p.n_queries) v = src[(size_t)q * CHUNKS + (i % CHUNKS)];
            dst[i] = v;
        }
        if (p.cur_score != nullptr) {
            for (int i = threadIdx.x; i < QT; i += kScanThreads) {
                const bool in = q0 + i < p.n_queries;
                s_cursor[2 * i] = in ? p.cur_score[q0 + i] : kSentinelScore;
                s_cursor[2 * i + 1] = in ? p.cur_row[q0 + i] : 0xFFFFFFFFu;
            }
        }
    }
    __syncthreads();

    // ---- TMA ring: one elected thread issues the bulk copies of tile `i` into stage i % stages ----
    const uint64_t policy = p.evict_first ? l2_policy_evict_first() : l2_policy_evict_last();
    auto issue_tile = [&](uint32_t i) {
        const int s = i % stages;
        const uint32_t t = (tile_begin + i) * p.tile_stride;
        const uint32_t row0 = t * TR;
        VL_DCHECK(i < my_tiles && (uint64_t)row0 < (uint64_t)p.n_rows + TR, 102);   // never a tile past the shard
        const uint32_t mask_bytes = MASKW * 4;
        // the tile arrives hardware-swizzled (16 B chunk index ^= row bits): rows past n_rows are
        // zero-filled by the TMA unit and the whole box is always counted
        mbar_arrive_expect_tx(&s_full[s], TILE_BYTES + mask_bytes + (p.filter ? mask_bytes : 0));
#pragma unroll
        for (int b = 0; b < TPS; ++b)
            tma_load_2d(s_tiles + (size_t)s * TILE_BYTES + (size_t)b * kTileRows * W, &tmap, 0,
                        (int)(row0 + b * kTileRows), &s_full[s], policy);
        uint32_t* m = s_masks + s * 2 * MASKW;
        bulk_g2s(m, p.live + (size_t)t * MASKW, mask_bytes, &s_full[s]);
        if (p.filter) bulk_g2s(m + MASKW, p.filter + (size_t)t * MASKW, mask_bytes, &s_full[s]);
    };
    if (threadIdx.x == 0) {
        const uint32_t pre = min((uint32_t)stages, my_tiles);
        for (uint32_t i = 0; i < pre; ++i) issue_tile(i);
    }

    const int wq = warp % WQ;
    const int wr = warp / WQ;
    uint32_t* ls = s_lscore + (size_t)warp * C * k;
    uint32_t* li = s_lidx + (size_t)warp * C * k;
    for (int e = lane; e < C * k; e += 32) {
        ls[e] = kSentinelScore;
        li[e] = 0xFFFFFFFFu;
    }
    __syncwarp();

    uint32_t thr[C];
#pragma unroll
    for (int c = 0; c < C; ++c) thr[c] = (qtile * QT + wq * C + c < p.n_queries) ? kSentinelScore : 0u;
    if (lane < C) s_thr[warp * C + lane] = (qtile * QT + wq * C + lane < p.n_queries) ? kSentinelScore : 0u;
    __syncwarp();
    const bool collect = p.collect_thr != nullptr;
    if (collect) {
        // fixed thresholds: "score <= thr" is tested as "score < thr + 1" (thr + 1 saturates)
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int qg = qtile * QT + wq * C + c;
            uint32_t v = 0;
            if (qg < p.n_queries) {
                v = p.collect_thr[qg];
                v = (v == kSentinelScore) ? v : v + 1;
            }
            thr[c] = v;
        }
    }

    // TMA wrote the tile with the 128 B / 64 B / 32 B swizzle (W = 128 / 64 / 32): logical chunk j of
    // row r sits at physical chunk j ^ rot(r), rot = r & 7, (r >> 1) & 3, (r >> 2) & 1.  Lane l owns row l (mod 32), so
    // its physical offset is (j * 16) ^ rot16 -- one row per lane at stride W, conflict-free 128-bit LDS
    constexpr int LANES_PER_128B = (W >= 128) ? 1 : 128 / W;
    constexpr int ROT_MASK = (CHUNKS < 8 ? CHUNKS : 8) - 1;
    const uint32_t rot16 = (uint32_t)(((lane / LANES_PER_128B) & ROT_MASK) * 16);
    const uint32_t q_saddr = smem_u32(s_queries) + (uint32_t)(wq * C * W);
    const uint32_t tiles_saddr = smem_u32(s_tiles);
    const bool has_filter = p.filter != nullptr;
    const bool has_cursor = p.cur_score != nullptr;
    const uint32_t one = p.one;

    // Q <= 2 is HBM-bound with idle integer pipes: pruning buys nothing there, skip the traffic
    constexpr bool USE_TAU = QT > 2;
    const bool tau_mine = USE_TAU && p.collect_thr == nullptr && lane < C && (qtile * QT + wq * C + lane) < p.n_queries;
    const uint32_t* tau_ptr = p.tau + (qtile * QT + wq * C + (tau_mine ? lane : 0));
    uint32_t tau_pref = kSentinelScore;
    bool dirty = false;
    // only where one warp per CTA serves a query (WQ == 8): with row-sliced warps (Q <= 32) thousands
    // of warps would hammer the same few bins and the bound is not worth the atomics (measured)
    const bool use_hist = WQ == 8 && p.hist != nullptr && p.collect_thr == nullptr && k <= 32;

    for (uint32_t i = 0; i < my_tiles; ++i) {
        const uint32_t t = (tile_begin + i) * p.tile_stride;
        const int s = i % stages;
        // warp 0 refills the stage every warp released one tile ago (one tile of slack keeps it
        // from waiting on the slowest warp)
        if (warp == 0 && i >= 1) {
            const uint32_t nxt = i - 1 + stages;
            if (nxt < my_tiles) {
                mbar_wait(&s_empty[(i - 1) % stages], ((i - 1) / stages) & 1);
                if (lane == 0) issue_tile(nxt);
                __syncwarp();
            }
        }
        // refresh the effective thresholds: min(own k-th best, device-wide bound + 1).  The bound
        // was fetched one tile ago, so the L2 round trip hides behind that tile's work.
        if constexpr (USE_TAU) {
            const uint32_t g = (tau_pref == kSentinelScore) ? tau_pref : tau_pref + 1;
#pragma unroll
            for (int c = 0; c < C; ++c) thr[c] = min(thr[c], __shfl_sync(kFullMask, g, c));
            if (tau_mine) tau_pref = __ldcg(tau_ptr);
        }
        // at tiles 1, 2, 4, 8, ... turn the device-wide histogram into a bound: the bucket where the
        // running count of candidate rows reaches k caps the final k-th best score
        if (use_hist && i != 0 && (i & (i - 1)) == 0) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int qg = qtile * QT + wq * C + c;
                if (qg >= p.n_queries) continue;                       // warp-uniform
                const uint32_t* h = p.hist + (size_t)qg * kHistWords;
                uint32_t cum = __ldcg(h + kHistFine + lane);
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const uint32_t up = __shfl_up_sync(kFullMask, cum, off);
                    if (lane >= off) cum += up;
                }
                const unsigned reach = __ballot_sync(kFullMask, cum >= (uint32_t)k);
                if (!reach) continue;
                const int bucket = __ffs(reach) - 1;
                const uint32_t before = bucket ? __shfl_sync(kFullMask, cum, bucket - 1) : 0u;
                uint32_t cf = __ldcg(h + bucket * 32 + lane);
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const uint32_t up = __shfl_up_sync(kFullMask, cf, off);
                    if (lane >= off) cf += up;
                }
                const unsigned reach2 = __ballot_sync(kFullMask, before + cf >= (uint32_t)k);
                if (!reach2) continue;                                 // fine bins lag the coarse ones: skip
                const int x = bucket * 32 + (__ffs(reach2) - 1);
                if (x >= kHistFine - 1) continue;                      // the last bucket is open-ended
                const uint32_t bound = ((uint32_t)(x + 1) << p.hist_shift) - 1u;   // k-th best <= bound
                thr[c] = min(thr[c], bound + 1u);
                if (lane == 0) atomicMin(p.tau + qg, bound);
            }
        }
        mbar_wait(&s_full[s], (i / stages) & 1);
        const uint32_t* m = s_masks + s * 2 * MASKW;
        // ring protocol: the stage being read must not have been handed to a tile `stages` ahead yet --
        // the refill of this stage waits on s_empty, which this warp has not arrived on for tile i
        VL_DCHECK(!mbar_test_wait(&s_empty[s], (i / stages) & 1), 103);

#pragma unroll 1
        for (int st = 0; st < STEPS; ++st) {
            const int row_in_tile0 = wr * RPW + st * 32 * R;  // multiple of 32
            bool valid[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                uint32_t w = m[row_in_tile0 / 32 + r];
                if (has_filter) w &= m[MASKW + row_in_tile0 / 32 + r];
                valid[r] = (w >> lane) & 1u;
            }
            // masked rows start from a bias no real score reaches (max 255 * W), so the hot
            // threshold test below needs no per-row predicate
            uint32_t acc[R][C];
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int c = 0; c < C; ++c) acc[r][c] = valid[r] ? 0u : kMaskedBias;

            const uint32_t row_saddr =
                tiles_saddr + (uint32_t)s * TILE_BYTES + (uint32_t)(row_in_tile0 + lane) * W;
            // keep the unrolled body near 12 KB of SASS so the hot loop stays inside the I-cache
            constexpr int UNROLL = (R * C >= 32) ? 1 : (R * C >= 16 ? 2 : CHUNKS / 2);
#pragma unroll UNROLL
            for (int j = 0; j < CHUNKS; j += 2) {
                const uint32_t coff_a = ((uint32_t)(j * 16)) ^ rot16;
                const uint32_t coff_b = ((uint32_t)((j + 1) * 16)) ^ rot16;
                uint4 ra[R], rb[R];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    ra[r] = lds128(row_saddr + (uint32_t)(r * 32 * W) + coff_a);
                    rb[r] = lds128(row_saddr + (uint32_t)(r * 32 * W) + coff_b);
                }
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    // every lane scores the same LOGICAL chunk: the query read is a broadcast
                    const uint4 qa = lds128(q_saddr + (uint32_t)(c * W + j * 16));
                    const uint4 qb = lds128(q_saddr + (uint32_t)(c * W + (j + 1) * 16));
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        acc[r][c] = score_chunk_pair<METRIC>(ra[r], rb[r], qa, qb, acc[r][c], one);
                }
            }

            bool hit = false;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                uint32_t best = acc[0][c];
#pragma unroll
                for (int r = 1; r < R; ++r) best = min(best, acc[r][c]);
                hit |= best < thr[c];
            }

            if (__any_sync(kFullMask, hit) && !collect && k <= 32) {
                // Rare path of the common case (k-lists with k <= 32): statically indexed, one
                // warp-uniform vote per query, then one list merge per row group that has a hit.
                const uint32_t row_base = t * TR + (uint32_t)row_in_tile0;
                // one warp-wide OR tells which (query, row group) pairs have a candidate at all
                uint32_t hm = 0;
#pragma unroll
                for (int c = 0; c < C; ++c)
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (valid[r] && (acc[r][c] < thr[c])) hm |= 1u << (c * R + r);
                const uint32_t any_hit = __reduce_or_sync(kFullMask, hm);
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    if (((any_hit >> (c * R)) & ((1u << R) - 1u)) == 0) continue;   // warp-uniform
                    uint32_t tc = s_thr[warp * C + c];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if (((any_hit >> (c * R + r)) & 1u) == 0) continue;         // warp-uniform
                        const uint32_t sc = acc[r][c];
                        const uint32_t idx = row_base + (uint32_t)(r * 32 + lane);
                        bool mine = valid[r] && (sc < min(thr[c], tc));
                        if (has_cursor && mine) {
                            const uint32_t cs = s_cursor[2 * (wq * C + c)];
                            const uint32_t cr = s_cursor[2 * (wq * C + c) + 1];
                            mine = (sc > cs) || (sc == cs && idx > cr);
                        }
                        if (__any_sync(kFullMask, mine)) {
                            if (use_hist && mine) {
                                const uint32_t b = min(sc >> p.hist_shift, (uint32_t)(kHistFine - 1));
                                uint32_t* h = p.hist + (size_t)(qtile * QT + wq * C + c) * kHistWords;
                                atomicAdd(h + b, 1u);
                                atomicAdd(h + kHistFine + (b >> 5), 1u);
                            }
                            tc = warp_list_merge32(ls + c * k, li + c * k, k,
                                                   mine ? (((uint64_t)sc << 32) | idx) : ~0ull, lane);
                        }
                    }
                    if (lane == 0) s_thr[warp * C + c] = tc;
                    thr[c] = min(thr[c], tc);
                }
                __syncwarp();
                dirty = true;
            } else if (__any_sync(kFullMask, hit)) {
                // Rare path of the other modes (collect, k > 32), kept small on purpose (a compact loop, not R*C unrolled copies) so it
                // does not evict the hot loop from the instruction cache.  Bit b = c*R + r of `hm`
                // marks a candidate; queries (c) are independent, rows ascend with (r, lane).
                uint32_t hm = 0;
#pragma unroll
                for (int c = 0; c < C; ++c)
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (valid[r] && (acc[r][c] < thr[c])) hm |= 1u << (c * R + r);
                uint32_t any = __reduce_or_sync(kFullMask, hm);
                const uint32_t row_base = t * TR + (uint32_t)row_in_tile0;
                while (any) {
                    const int b = __ffs(any) - 1;
                    any &= any - 1;
                    const int c = b / R, r = b % R;
                    uint32_t v = 0;
#pragma unroll
                    for (int cc = 0; cc < C; ++cc)
#pragma unroll
                        for (int rr = 0; rr < R; ++rr) v = (b == cc * R + rr) ? acc[rr][cc] : v;
                    unsigned mm = __ballot_sync(kFullMask, (hm >> b) & 1u);
                    if (collect) {
                        // append every flagged row: one atomicAdd per (warp, query, row group)
                        const int qg = qtile * QT + wq * C + c;
                        bool mine = (mm >> lane) & 1u;
                        const uint32_t idx = row_base + (uint32_t)(r * 32 + lane);
                        if (mine && has_cursor) {
                            const uint32_t cs = s_cursor[2 * (wq * C + c)];
                            const uint32_t cr = s_cursor[2 * (wq * C + c) + 1];
                            mine = (v > cs) || (v == cs && idx > cr);
                        }
                        const unsigned sel = __ballot_sync(kFullMask, mine);
                        if (sel) {
                            uint32_t base = 0;
                            if (lane == __ffs(sel) - 1) base = atomicAdd(p.collect_count + qg, (uint32_t)__popc(sel));
                            base = __shfl_sync(kFullMask, base, __ffs(sel) - 1);
                            const uint32_t slot = base + (uint32_t)__popc(sel & ((1u << lane) - 1));
                            VL_DCHECK(qg < p.n_queries, 104);
                            if (mine && slot < p.collect_cap)
                                p.collect_keys[(size_t)qg * p.collect_cap + slot] =
                                    ((unsigned long long)v << 32) | idx;
                        }
                        continue;
                    }
                    uint32_t tc = s_thr[warp * C + c];
                    while (mm) {
                        const int src = __ffs(mm) - 1;
                        mm &= mm - 1;
                        const uint32_t sc = __shfl_sync(kFullMask, v, src);
                        if (sc < tc) {
                            const uint32_t idx = row_base + (uint32_t)(r * 32 + src);
                            bool eligible = true;
                            if (has_cursor) {
                                const uint32_t cs = s_cursor[2 * (wq * C + c)];
                                const uint32_t cr = s_cursor[2 * (wq * C + c) + 1];
                                eligible = (sc > cs) || (sc == cs && idx > cr);
                            }
                            if (eligible) tc = warp_list_insert(ls + c * k, li + c * k, k, sc, idx, lane);
                        }
                    }
                    if (lane == 0) s_thr[warp * C + c] = tc;
                    __syncwarp();
                }
                if (!collect) {
                    dirty = true;
#pragma unroll
                    for (int c = 0; c < C; ++c) thr[c] = min(thr[c], s_thr[warp * C + c]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[s]);
        // publish improved k-th scores at most once per tile, and only when they beat the bound
        // this warp last saw (one hot address per query: keep the atomics rare)
        if (dirty) {
            dirty = false;
            if (tau_mine) {
                const uint32_t mine = s_thr[warp * C + lane];
                if (mine < tau_pref) atomicMin(const_cast<uint32_t*>(tau_ptr), mine);
            }
        }
    }

    // -------- write this warp's lists to the partial buffer --------
    if (collect) return;
    if constexpr (WQ == 1 && C <= 2) {
        if (p.fuse_keys != nullptr) {
            // ---- latency path: CTA-level merge, ticket, final merge by the last CTA ----
            __syncthreads();                       // every warp's k-list is final (shared memory)
            // the tile ring is drained (every issued tile was consumed): reuse it as scratch
            unsigned long long (*s_l1)[32] = reinterpret_cast<unsigned long long (*)[32]>(s_tiles);
            volatile uint32_t& s_last = *reinterpret_cast<volatile uint32_t*>(s_thr);
            if (warp < C) {
                const int c = warp;                // query c of this CTA: its 8 row-slice lists
                warp_merge_keys<1>(
                    kScanWarps, k, lane,
                    [&](int l, int e) -> unsigned long long {
                        const uint32_t sc = s_lscore[(size_t)l * C * k + c * k + e];
                        const uint32_t ix = s_lidx[(size_t)l * C * k + c * k + e];
                        return ix == 0xFFFFFFFFu ? ~0ull : (((unsigned long long)sc << 32) | ix);
                    },
                    [&](int o, unsigned long long key) {
                        const int q = qtile * QT + c;
                        if (q < p.n_queries) p.fuse_keys[((size_t)split * p.n_queries + q) * k + o] = key;
                    });
            }
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) s_last = (atomicAdd(p.fuse_ticket, 1u) == gridDim.x * gridDim.y - 1) ? 1u : 0u;
            __syncthreads();
            if (s_last == 0u) return;
            __threadfence();
            const int per = (n_splits + kScanWarps - 1) / kScanWarps;   // <= 128 (host caps the splits at 1024)
            // staging area behind s_l1: each warp's lists are fetched from L2 in ONE batch of independent
            // loads (one latency) instead of k dependent rounds; 8 warps x 1024 keys = 64 KB of the ring
            unsigned long long* s_stage = reinterpret_cast<unsigned long long*>(s_tiles) + kScanWarps * 32 + (size_t)warp * 1024;
            const bool staged = per * k <= 1024 && (size_t)stages * TILE_BYTES >= (size_t)(kScanWarps * 32 + kScanWarps * 1024) * 8;
            for (int q = 0; q < p.n_queries; ++q) {
                // level 1: warp w merges splits [w * per, (w + 1) * per)
                const int l0 = warp * per;
                const int nl = max(0, min(per, n_splits - l0));
                if (staged) {
                    for (int e = lane; e < nl * k; e += 32)
                        s_stage[e] = __ldcg(p.fuse_keys + ((size_t)(l0 + e / k) * p.n_queries + q) * k + (e % k));
                    __syncwarp();
                    warp_merge_keys<4>(
                        nl, k, lane, [&](int l, int e) -> unsigned long long { return s_stage[l * k + e]; },
                        [&](int o, unsigned long long key) { s_l1[warp][o] = key; });
                } else {
                    warp_merge_keys<4>(
                        nl, k, lane,
                        [&](int l, int e) -> unsigned long long {
                            return __ldcg(p.fuse_keys + ((size_t)(l0 + l) * p.n_queries + q) * k + e);
                        },
                        [&](int o, unsigned long long key) { s_l1[warp][o] = key; });
                }
                __syncthreads();
                // level 2: warp 0 merges the 8 partial lists and writes the result
                if (warp == 0) {
                    warp_merge_keys<1>(
                        kScanWarps, k, lane, [&](int l, int e) -> unsigned long long { return s_l1[l][e]; },
                        [&](int o, unsigned long long key) {
                            const bool none = key == ~0ull;
                            p.out_scores[(size_t)q * p.out_stride + o] = none ? kSentinelScore : (uint32_t)(key >> 32);
                            p.out_rows[(size_t)q * p.out_stride + o] = none ? kSentinelRow : p.base_row + (uint32_t)key;
                        });
                }
                __syncthreads();
            }
            if (threadIdx.x == 0) *p.fuse_ticket = 0u;   // ready for the next launch on this stream
            return;
        }
    }
    const int list_id = split * WR + wr;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int q = qtile * QT + wq * C + c;
        if (q >= p.n_queries) continue;
        const size_t o = ((size_t)list_id * p.n_queries + q) * k;
        VL_DCHECK(list_id < n_splits * WR && q < p.n_queries, 105);
        for (int e = lane; e < k; e += 32) {
            const uint32_t idx = li[c * k + e];
            p.part_scores[o + e] = ls[c * k + e];
            p.part_rows[o + e] = (idx == 0xFFFFFFFFu) ? kSentinelRow : p.base_row + idx;
        }
    }
}

// ---- full score vector of one query (parity tests; not the search path) --------
template <int METRIC>
__global__ void scores_kernel(const uint8_t* __restrict__ rows, int w, const uint8_t* __restrict__ query,
                              uint64_t first, uint64_t n, uint32_t* __restrict__ out) {
    extern __shared__ uint32_t s_q[];
    const int words = w / 4;
    for (int i = threadIdx.x; i < words; i += blockDim.x) s_q[i] = reinterpret_cast<const uint32_t*>(query)[i];
    __syncthreads();
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4* r = reinterpret_cast<const uint4*>(rows + (first + i) * (uint64_t)w);
        uint32_t acc = 0;
        for (int j = 0; j < words / 4; ++j) {
            uint4 v = ldg128_stream(r + j);
            uint4 q = make_uint4(s_q[4 * j], s_q[4 * j + 1], s_q[4 * j + 2], s_q[4 * j + 3]);
            acc = score_chunk<METRIC>(v, q, acc);
        }
        out[i] = acc;
    }
}

}  // namespace vl
