// Host side of a search: scan configuration dispatch and grid sizing, the multi-level merge, the
// k <= 128 list path with its multi-pass extension, and the sampled collect/select path for large
// k.  Included by api.cu only, after corpus.cuh.
#pragma once

namespace {

using namespace vl;

// ---- scan dispatch -------------------------------------------------------------

struct ScanPlan {
    int C, WQ, WR, QT, T, S, stages, occ;
    size_t smem;
};

template <int W, int METRIC, int C, int WQ, int TPS>
int launch_scan_cfg(vl_corpus* c, ScanParams& p, int n_queries, int k, ScanPlan* plan_out, bool dry) {
    using Cfg = ScanCfg<W, C, WQ, TPS>;
    auto kern = scan_topk_kernel<W, METRIC, C, WQ, TPS>;
    // stages visited by this launch: every tile_stride-th stage of the shard
    const uint32_t all_stages = (uint32_t)((p.n_rows + Cfg::TR - 1) / Cfg::TR);
    p.n_tiles = (all_stages + p.tile_stride - 1) / p.tile_stride;
    const size_t fixed = Cfg::smem_bytes(0, k);
    const size_t max_smem = 227 * 1024;
    const size_t two_per_sm = 113 * 1024;
    int stages = (int)std::min<size_t>(8, std::max<size_t>(2, (96 * 1024) / Cfg::TILE_BYTES));
    if (Cfg::TILE_BYTES >= 64 * 1024) stages = 3;  // two-box stages: one CTA per SM, three stages in flight
    if (c->tune_stages > 0) stages = c->tune_stages;
    auto total = [&](int s) { return fixed + (size_t)s * (Cfg::STAGE_BYTES + 16); };
    if (c->tune_stages <= 0 && Cfg::TILE_BYTES < 64 * 1024)
        while (stages > 2 && total(stages) > two_per_sm) --stages;
    while (stages > 1 && total(stages) > max_smem) --stages;
    const size_t smem = total(stages);
    if (smem > max_smem) return fail(VL_EINVAL, "k=%d needs %zu B of shared memory per CTA (max %zu)", k, smem, max_smem);
    TRY(allow_max_smem(kern, c->device));
    // the occupancy of this kernel at this shared-memory size is asked for once per thread
    static thread_local size_t cached_smem[64] = {};
    static thread_local int cached_occ[64] = {};
    int occ = 0;
    const int dev_slot = c->device & 63;
    if (cached_smem[dev_slot] == smem && cached_occ[dev_slot] > 0) {
        occ = cached_occ[dev_slot];
    } else {
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kScanThreads, smem));
        if (occ < 1) return fail(VL_ECUDA, "scan kernel does not fit on an SM (smem %zu)", smem);
        cached_smem[dev_slot] = smem;
        cached_occ[dev_slot] = occ;
    }

    const int T = (n_queries + Cfg::QT - 1) / Cfg::QT;
    const long slots = (long)c->sms * occ;
    const long max_splits = std::max<long>(1, (long)p.n_tiles / 4);
    long S = 1;
    if (c->tune_splits > 0) {
        S = c->tune_splits;
    } else {
        double best_eff = -1;
        for (int waves = 1; waves <= 4; ++waves) {
            long s = std::max<long>(1, (waves * slots) / T);
            s = std::min(s, max_splits);
            const long ctas = s * T;
            const double eff = (double)ctas / ((double)((ctas + slots - 1) / slots) * slots);
            // an extra wave re-pays per-CTA prologue, list warm-up and merge work: take it only when
            // it fills the SMs better (measured at 500k x 1024: 2 exact waves of 592 CTAs beat 1 wave
            // of 288 by 2 % now that list maintenance is cheap; 4 waves lose)
            // (XORSUM steps are shorter, so the second wave's fixed costs weigh more: 1 wave of 288
            // beats 2 exact waves by 3 % there)
            if (eff > best_eff + (METRIC == VL_METRIC_HAMMING ? 0.02 : 0.05)) {
                best_eff = eff;
                S = s;
            }
            if (s == max_splits) break;
        }
    }
    S = std::min<long>(S, std::max<long>(1, p.n_tiles));
    S = std::min<long>(S, 65535);
    if (plan_out) *plan_out = ScanPlan{C, WQ, Cfg::WR, Cfg::QT, T, (int)S, stages, occ, smem};
    if (dry) return VL_OK;
    p.stages = stages;
    CUtensorMap tmap;
    TRY(make_corpus_tmap(c, &tmap));
    dim3 grid((unsigned)T, (unsigned)S);
    kern<<<grid, kScanThreads, smem, c->stream>>>(tmap, p);
    LAUNCHED();
    return VL_OK;
}

template <int W, int METRIC>
int launch_scan_wm(vl_corpus* c, ScanParams& p, int nq, int k, ScanPlan* plan, bool dry) {
    // Q = 1 is HBM-bound: one box per stage.  2 <= Q <= 16 has few queries per warp: where two
    // boxes still fit a 32 KB stage (W = 64) a stage carries two, so every warp scores twice the rows
    // per step at unchanged occupancy (+8 %); at W = 128 the lost CTA per SM costs more (-15 %, measured).
    constexpr int VL_MIDQ_TPS = (W <= 32) ? 4 : (W <= 64) ? 2 : 1;   // boxes per stage: 32 KB of rows either way
    constexpr int VL_Q1_TPS = (W <= 32) ? 2 : 1;
    // (two boxes per stage at W <= 64 also for Q = 2: 5.0 -> 6.1 TB/s on 32M x 64 B; Q = 1 gains nothing)
    if (nq <= 1) return launch_scan_cfg<W, METRIC, 1, 1, VL_Q1_TPS>(c, p, nq, k, plan, dry);
    if (nq <= 2) return launch_scan_cfg<W, METRIC, 2, 1, VL_MIDQ_TPS>(c, p, nq, k, plan, dry);
    if (nq <= 4) return launch_scan_cfg<W, METRIC, 4, 1, VL_MIDQ_TPS>(c, p, nq, k, plan, dry);
    if (nq <= 8) return launch_scan_cfg<W, METRIC, 8, 1, VL_MIDQ_TPS>(c, p, nq, k, plan, dry);
    if (nq <= 16) return launch_scan_cfg<W, METRIC, 8, 2, VL_MIDQ_TPS>(c, p, nq, k, plan, dry);
    if (nq <= 32) return launch_scan_cfg<W, METRIC, 8, 4, 1>(c, p, nq, k, plan, dry);
    return launch_scan_cfg<W, METRIC, 8, 8, 1>(c, p, nq, k, plan, dry);
}

// ---- tensor-core scan (mmascan.cuh) ----------------------------------------------

// smallest batch the tensor-core scan takes in automatic mode (tuning override: VLITE_B200_MMA_MINQ;
// a value <= 0 disables the path).  A 128-query tile costs the same whether it holds 16 queries or 128; measured
// crossover of the e2m1 HAMMING form against the integer-pipe kernel (profiles/r2_sweep.md): W = 128: Q = 16
// 0.080 vs 0.100 ms on 500k rows, 2.81 vs 2.70 ms on 32M (Q = 24: 2.81 vs 4.0); W = 64: Q = 16 0.063 vs 0.080 ms.
// XORSUM on shards of more than ~1M rows: its integer-pipe kernel (LOP3 + IDP4A) costs 0.0030 / 0.0017 ns per
// (row, query) at W = 128 / 64 against 0.127 / 0.067 ns per row for a 128-query tensor-core tile, so the crossover
// sits near 43 / 40 queries (32M rows, Q = 32: 2.6 vs 4.1 ms); below ~1M rows the tile kernel's smaller fixed cost
// keeps the lower threshold right (500k rows, Q = 32: 0.104 vs 0.131 ms).  profiles/r1_sweep.md / r2_sweep.md.
int mma_min_q(const vl_corpus* c, int metric) {
    static int v = [] {
        const char* e = getenv("VLITE_B200_MMA_MINQ");
        return e ? atoi(e) : -1;
    }();
    if (v != -1) return v;
    if (metric == VL_METRIC_XORSUM && c->w >= 64 && c->count > (1u << 20)) return c->w >= 128 ? 44 : 40;
    return 16;
}

// 0 = not eligible, 1 = one CTA per query tile, 2 = CTA pairs (cta_group::2)
int mma_path_for(const vl_corpus* c, const ScanParams& p, int nq, int k, int metric) {
    if (p.cur_score != nullptr) return 0;
    const bool collect = p.collect_thr != nullptr;
    if (!collect && k > kMmaListCap) return 0;
    if (c->scan_path == 1) return 0;
    if (c->scan_path == 2 || c->scan_path == 5) return 1;
    if (c->scan_path == 3 || c->scan_path == 6) return 2;
    const int mq = mma_min_q(c, metric);
    if (mq <= 0 || nq < mq) return 0;
    return nq > kMmaQT ? 2 : 1;
}

template <int W, int CG, bool COLLECT, int METRIC, bool F4 = false>
int launch_mma_cfg(vl_corpus* c, ScanParams& p, int nq, int k, ScanPlan* plan_out, bool dry) {
    using Cfg = MmaCfg<W, CG, F4, METRIC>;
    auto kern = mma_scan_kernel<W, CG, COLLECT, METRIC, F4>;
    const uint32_t all_tiles = (uint32_t)((p.n_rows + Cfg::NT - 1) / Cfg::NT);
    p.n_tiles = (all_tiles + p.tile_stride - 1) / p.tile_stride;
    const int T = (nq + kMmaQT * CG - 1) / (kMmaQT * CG);   // clusters along the queries
    const long slots = std::max(1, c->sms / CG);             // clusters resident at once (one CTA per SM)
    // Query tiles per launch.  One launch of T tiles x floor(slots / T) row splits leaves SMs idle when T does not
    // divide the slots well (4096 queries: 16 pairs x 4 splits = 128 of 148 SMs, ncu profiles/r2_collect_scan.md);
    // several launches of fewer tiles each re-stream the packed rows (cheap: the scan is nowhere near HBM-bound
    // at these batch sizes) but fill the machine: 2 x (8 pairs x 9 splits) = 144 SMs.  Only for batches of 8+
    // tiles, whose launches are long; the number of row splits (= lists per query) is the same in every launch.
    int Tl = T;
    if (c->tune_splits <= 0 && T >= 8) {
        double best = (double)1 / (double)std::max<long>(1, slots / T);
        for (int t = T - 1; t >= 4; --t) {
            const double cost = (double)((T + t - 1) / t) / (double)std::max<long>(1, slots / t);   // launches / splits
            if (cost < best * 0.97) {
                best = cost;
                Tl = t;
            }
        }
    }
    long S = c->tune_splits > 0 ? c->tune_splits : std::max<long>(1, slots / Tl);
    S = std::min<long>(S, std::max<long>(1, p.n_tiles));
    S = std::min<long>(S, 65535);
    const size_t smem = Cfg::SMEM;
    if (plan_out) *plan_out = ScanPlan{0, 0, Cfg::EPI_GROUPS, kMmaQT * CG, T, (int)S, Cfg::BSTAGES, 1, smem};   // WR = lists per (split, query)
    if (dry) return VL_OK;
    p.stages = c->tune_stages;   // 0 = the configuration's own ring depth
    TRY(allow_max_smem(kern, c->device));
    CUtensorMap tmap;
    TRY(make_corpus_tmap(c, &tmap, Cfg::ROWS));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(Tl * CG), (unsigned)S);
    cfg.blockDim = dim3(kMmaThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CG;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = CG == 2 ? 1 : 0;
    for (int t0 = 0; t0 < T; t0 += Tl) {
        p.q_offset = t0 * kMmaQT * CG;
        cfg.gridDim.x = (unsigned)(std::min(Tl, T - t0) * CG);
        CU(cudaLaunchKernelEx(&cfg, kern, tmap, p));
        LAUNCHED();
    }
    p.q_offset = 0;
    return VL_OK;
}

// The +-1 contraction runs in 4-bit floats (kind::mxf4, exact; XORSUM with its bit weights as block scales, W >= 64)
// unless VLITE_B200_FP4=0 or the scan path asks for the int8 form (vl_set_scan_path 5 / 6)
bool mma_use_f4(const vl_corpus* c) {
    static const bool env_on = [] { const char* e = getenv("VLITE_B200_FP4"); return !(e && atoi(e) == 0); }();
    return env_on && c->scan_path != 5 && c->scan_path != 6;
}

template <int W, int METRIC>
int launch_mma_wm(vl_corpus* c, ScanParams& p, int nq, int k, int cg, ScanPlan* plan, bool dry) {
    const bool collect = p.collect_thr != nullptr;
    if constexpr (METRIC == kHamming || W >= 64) {   // (XORSUM as e2m1 needs a K = 64 step inside one bit plane: W >= 64)
        if (mma_use_f4(c)) {
            if (cg == 2)
                return collect ? launch_mma_cfg<W, 2, true, METRIC, true>(c, p, nq, k, plan, dry)
                               : launch_mma_cfg<W, 2, false, METRIC, true>(c, p, nq, k, plan, dry);
            return collect ? launch_mma_cfg<W, 1, true, METRIC, true>(c, p, nq, k, plan, dry)
                           : launch_mma_cfg<W, 1, false, METRIC, true>(c, p, nq, k, plan, dry);
        }
    }
    if (cg == 2)
        return collect ? launch_mma_cfg<W, 2, true, METRIC>(c, p, nq, k, plan, dry)
                       : launch_mma_cfg<W, 2, false, METRIC>(c, p, nq, k, plan, dry);
    return collect ? launch_mma_cfg<W, 1, true, METRIC>(c, p, nq, k, plan, dry)
                   : launch_mma_cfg<W, 1, false, METRIC>(c, p, nq, k, plan, dry);
}
template <int W>
int launch_mma_w(vl_corpus* c, ScanParams& p, int nq, int k, int cg, int metric, ScanPlan* plan, bool dry) {
    return metric == VL_METRIC_HAMMING ? launch_mma_wm<W, kHamming>(c, p, nq, k, cg, plan, dry)
                                       : launch_mma_wm<W, kXorSum>(c, p, nq, k, cg, plan, dry);
}

int launch_scan(vl_corpus* c, ScanParams& p, int nq, int k, int metric, ScanPlan* plan, bool dry) {
    if (const int cg = mma_path_for(c, p, nq, k, metric)) {
        switch (c->w) {
            case 32: return launch_mma_w<32>(c, p, nq, k, cg, metric, plan, dry);
            case 64: return launch_mma_w<64>(c, p, nq, k, cg, metric, plan, dry);
            case 128: return launch_mma_w<128>(c, p, nq, k, cg, metric, plan, dry);
            default: break;
        }
    }
#define VL_W(Wv)                                                                          \
    case Wv:                                                                              \
        return metric == VL_METRIC_HAMMING ? launch_scan_wm<Wv, kHamming>(c, p, nq, k, plan, dry) \
                                           : launch_scan_wm<Wv, kXorSum>(c, p, nq, k, plan, dry);
    switch (c->w) {
        VL_W(32)
        VL_W(64)
        VL_W(128)
        default:
            return fail(VL_EINVAL, "unsupported row width %d", c->w);
    }
#undef VL_W
}

int fill_sentinels(cudaStream_t st, void* s_any, void* r_any, size_t n, bool s_dev, bool r_dev, size_t s_elem = 4) {
    // 0xFF bytes are both sentinels (and a NaN/-inf is handled by the rescore caller)
    if (s_dev)
        CU(cudaMemsetAsync(s_any, 0xFF, n * s_elem, st));
    else
        memset(s_any, 0xFF, n * s_elem);
    if (r_dev)
        CU(cudaMemsetAsync(r_any, 0xFF, n * 8, st));
    else
        memset(r_any, 0xFF, n * 8);
    return VL_OK;
}

// stage a filter mask into the handle's padded buffer (whole tiles, 16 B aligned)
int stage_filter(vl_corpus* c, const uint32_t* filter_any, const uint32_t** out_dev) {
    const uint64_t n_stages = (c->count + kMaxStageRows - 1) / kMaxStageRows;
    const size_t padded = n_stages * (kMaxStageRows / 32) * 4;
    const size_t given = ((c->count + 31) / 32) * 4;
    const bool dev = is_device_ptr(filter_any);
    if (dev && given == padded && (reinterpret_cast<uintptr_t>(filter_any) & 15) == 0) {
        *out_dev = filter_any;
        return VL_OK;
    }
    TRY(c->filt_buf.ensure(padded));
    if (padded > given)
        CU(cudaMemsetAsync(static_cast<uint8_t*>(c->filt_buf.p) + given, 0, padded - given, c->stream));
    CU(cudaMemcpyAsync(c->filt_buf.p, filter_any, given, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                       c->stream));
    *out_dev = static_cast<const uint32_t*>(c->filt_buf.p);
    return VL_OK;
}

// Multi-level merge: <= kMergeGroup lists per warp per level.  tmp_* must hold
// merge_tmp_lists(n_lists) lists of n_queries x k_out entries (unused when n_lists <= kMergeGroup).
size_t merge_tmp_lists(size_t n_lists) {
    size_t total = 0;
    while (n_lists > (size_t)kMergeGroup) {
        n_lists = (n_lists + kMergeGroup - 1) / kMergeGroup;
        total += n_lists;
    }
    return total;
}

int merge_levels(cudaStream_t st, const uint32_t* in_s, const uint64_t* in_r, size_t n_lists, size_t s_stride,
                 size_t r_stride, int n_queries, int k_in, int k_out, uint32_t* out_s, uint64_t* out_r,
                 int out_stride, int out_offset, uint32_t* tmp_s, uint64_t* tmp_r, const int* gate = nullptr) {
    const unsigned gx = (unsigned)((n_queries + kMergeWarps - 1) / kMergeWarps);
    const size_t dense = (size_t)n_queries * k_out;
    while (true) {
        const size_t groups = (n_lists + kMergeGroup - 1) / kMergeGroup;
        if (groups <= 1) {
            merge_topk_kernel<<<dim3(gx, 1), kMergeWarps * 32, 0, st>>>(in_s, in_r, (int)n_lists, n_queries, k_in, k_out,
                                                                      out_s, out_r, out_stride, out_offset, s_stride,
                                                                      r_stride, 0, gate);
            LAUNCHED();
            return VL_OK;
        }
        merge_topk_kernel<<<dim3(gx, (unsigned)groups), kMergeWarps * 32, 0, st>>>(
            in_s, in_r, (int)n_lists, n_queries, k_in, k_out, tmp_s, tmp_r, k_out, 0, s_stride, r_stride, dense, gate);
        LAUNCHED();
        in_s = tmp_s;
        in_r = tmp_r;
        tmp_s += groups * dense;
        tmp_r += groups * dense;
        n_lists = groups;
        k_in = k_out;
        s_stride = r_stride = dense;
    }
}

// k <= 128: one scan + merge.  Larger k: ceil(k/128) passes, each returning the next 128 keys after
// a per-query cursor (exact and deterministic for any tie pattern).  `p` arrives with the corpus,
// queries and masks filled in.
// k_merge > k (single pass only): the final merge keeps the k_merge best of the union of the k-entry lists (the
// sampled pass of large k wants a deeper rank than a register list holds)
int search_lists(vl_corpus* c, ScanParams p, int n_queries, int k, int metric, uint32_t* o_s, uint64_t* o_r,
                 int out_stride = 0, const int* gate = nullptr, int k_merge = 0) {
    p.gate = gate;
    cudaStream_t st = c->stream;
    if (k_merge <= k || k > kWarpListMaxK) k_merge = 0;
    if (out_stride == 0) out_stride = k_merge ? k_merge : k;
    p.k_bound = k_merge;
    const int n_pass = (k + kWarpListMaxK - 1) / kWarpListMaxK;
    // tau[Q] | cursor score[Q] | cursor row[Q] | histogram[Q][1056] (k <= 32, not for huge batches)
    // (the histogram is read only by the one-warp-per-query configuration, Q > 32; the shared bound
    // tau only by query tiles of more than two queries: small batches skip both memsets)
    // the tensor-core kernel shares tau AND the histogram at every batch size
    const bool mma = mma_path_for(c, p, n_queries, std::min(k, kWarpListMaxK), metric) != 0;
    const bool want_hist = (mma || (k <= 32 && n_queries > 32)) && n_queries <= 16384;
    const bool want_tau = mma || n_queries > 2;
    const size_t hist_bytes = want_hist ? (size_t)n_queries * kHistWords * 4 : 0;
    const size_t head_bytes = ((size_t)n_queries * 12 + 15) & ~(size_t)15;   // the histogram is read 16 B at a time
    TRY(c->cand_buf.ensure(head_bytes + hist_bytes));
    p.tau = static_cast<uint32_t*>(c->cand_buf.p);
    uint32_t* cur_s = p.tau + n_queries;
    uint32_t* cur_r = cur_s + n_queries;
    p.hist = want_hist ? reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(c->cand_buf.p) + head_bytes) : nullptr;
    {
        const uint32_t max_score = metric == VL_METRIC_HAMMING ? 8u * c->w : 255u * c->w;
        p.hist_shift = 0;
        while ((max_score >> p.hist_shift) > (uint32_t)(kHistFine - 1)) ++p.hist_shift;
        if (metric == VL_METRIC_HAMMING) p.hist_shift = 0;   // 8W <= 1024: only the all-bits-differ score overflows
    }
    for (int pass = 0; pass < n_pass; ++pass) {
        const int k_pass = std::min(kWarpListMaxK, k - pass * kWarpListMaxK);
        p.k = k_pass;
        p.cur_score = pass ? cur_s : nullptr;
        p.cur_row = pass ? cur_r : nullptr;
        ScanPlan plan{};
        TRY(launch_scan(c, p, n_queries, k_pass, metric, &plan, /*dry=*/true));
        // latency path (1-2 queries, k <= 32): the scan's last CTA does the whole merge -- one launch
        const bool fuse = n_pass == 1 && n_queries <= 2 && k <= 32 && plan.C <= 2 && plan.WQ == 1 && plan.S <= 1024 &&
                          p.tile_stride == 1 && gate == nullptr && c->fuse_merge &&
                          c->count * (uint64_t)c->w <= (1ull << 30);   // (beyond ~1 GB the serial tail costs more than two launches)
        if (fuse) {
            TRY(c->part_r.ensure((size_t)plan.S * n_queries * k * 8));
            p.fuse_keys = static_cast<unsigned long long*>(c->part_r.p);
            p.fuse_ticket = reinterpret_cast<unsigned int*>(c->d_err + 2);
            p.out_scores = o_s;
            p.out_rows = o_r;
            p.out_stride = out_stride;
            p.part_scores = nullptr;
            p.part_rows = nullptr;
            p.evict_first = (c->count * (uint64_t)c->w > (96ull << 20)) ? 1 : 0;
            TRY(launch_scan(c, p, n_queries, k_pass, metric, nullptr, /*dry=*/false));
            if (c->timing) CU(cudaEventRecord(c->ev1, st));
            return VL_OK;
        }
        p.fuse_keys = nullptr;
        const size_t n_lists = (size_t)plan.S * plan.WR;
        const size_t dense = (size_t)n_queries * k_pass;
        const int k_out = k_merge ? k_merge : k_pass;
        const size_t n_part = n_lists * dense + merge_tmp_lists(n_lists) * (size_t)n_queries * k_out;
        TRY(c->part_s.ensure(n_part * 4));
        TRY(c->part_r.ensure(n_part * 8));
        p.part_scores = static_cast<uint32_t*>(c->part_s.p);
        p.part_rows = static_cast<uint64_t*>(c->part_r.p);
        // corpus read once and far larger than L2 -> stream it; otherwise let L2 keep it
        p.evict_first = (plan.T == 1 && n_pass == 1 && p.tile_stride == 1 &&
                         c->count * (uint64_t)c->w > (96ull << 20)) ? 1 : 0;
        if (want_tau) CU(cudaMemsetAsync(p.tau, 0xFF, (size_t)n_queries * 4, st));
        if (p.hist) CU(cudaMemsetAsync(p.hist, 0, hist_bytes, st));
        TRY(launch_scan(c, p, n_queries, k_pass, metric, nullptr, /*dry=*/false));
        if (c->timing && pass == 0 && p.tile_stride == 1 && gate == nullptr) CU(cudaEventRecord(c->ev1, st));
        TRY(merge_levels(st, p.part_scores, p.part_rows, n_lists, dense, dense, n_queries, k_pass, k_out, o_s, o_r,
                         out_stride, pass * kWarpListMaxK, p.part_scores + n_lists * dense,
                         p.part_rows + n_lists * dense, gate));
        if (pass + 1 < n_pass) {
            cursor_update_kernel<<<(n_queries + 127) / 128, 128, 0, st>>>(
                o_s, o_r, n_queries, out_stride, pass * kWarpListMaxK + k_pass - 1, c->base_row, cur_s, cur_r, gate);
            LAUNCHED();
        }
    }
    return VL_OK;
}

// above this k the sampled collect/select path replaces the per-warp k-lists (tuning override:
// VLITE_B200_LARGEK_MIN)
int large_k_min(const vl_corpus* c, const ScanParams& p, int n_queries, int metric) {
    static int v = [] {
        const char* e = getenv("VLITE_B200_LARGEK_MIN");
        return e ? atoi(e) : -1;
    }();
    if (v >= 0) return v;
    // integer-pipe kernel (shared-memory lists up to 32): measured crossover: equal at k=32, 1.7x faster at k=100,
    // 3x at k=256.  Where the batch is big enough for the tensor-core kernel, its register lists end at 16
    // and both passes of this path run on it: k = 17..32 at 500k x 1024 queries 2.62-2.72 ms (integer-pipe
    // lists) vs 0.65 ms (sampled + collect on the tensor cores), 4M rows: 19.1 vs 3.09 ms.
    ScanParams pc = p;
    pc.collect_thr = reinterpret_cast<const uint32_t*>(&pc);   // "collect mode" for the eligibility question only
    return mma_path_for(c, pc, n_queries, 1, metric) ? kMmaListCap : 32;
}

// k > large_k_min(): sampled threshold + collect + select (largek.cuh).  *done = false asks the caller to run
// the exact multi-pass path instead (tiny corpus, or a buffer overflowed / came up short).
int search_large_k(vl_corpus* c, ScanParams p, int n_queries, int k, int metric, uint32_t* o_s, uint64_t* o_r,
                   bool* done) {
    *done = false;
    cudaStream_t st = c->stream;
    uint32_t cap = 8192;
    while (cap < 6u * (uint32_t)k) cap <<= 1;                       // expected ~3k candidates per query
    const size_t key_bytes = (size_t)n_queries * cap * 8;
    if (key_bytes > (8ull << 30)) return VL_OK;                      // would need > 8 GB of scratch
    // sample list length: 16..32 keeps the sampled pass on the scan's register-list fast path and lets
    // it visit few tiles; the threshold's relative spread is 1/sqrt(ks) <= 25 %, far inside the margin
    // between the ~3k rows expected at or below it and both failure modes (fewer than k: fallback;
    // more than cap >= 6k: fallback)
    // (16 whenever the tensor-core kernel can take the sampled pass: its register lists hold 16)
    const bool mma_sample = mma_path_for(c, p, n_queries, kMmaListCap, metric) != 0;
    const int ks = mma_sample ? kMmaListCap : std::min(kLargeKSample, std::max(16, k / 4));
    // The threshold's rank in the sample.  A 16-deep estimate is too noisy for a batch: the k-th best row lies
    // above the threshold of one query in ~7000 (P[Poisson(16/3) >= 16]), and ONE such query sends the whole
    // batch to the exact fallback (seen: k = 2048 at 500k x 1024 queries, 0.8 -> 166 ms).  The tensor-core pass
    // therefore merges its 16-entry lists (two per row split) to the 64 best of their union and samples four
    // times as many tiles: relative spread 12.5 %, both failure modes below 1e-10 per query.  (A list that holds
    // more than 16 of those 64 only raises the threshold: more candidates, never fewer.)
    // (k <= 213: the 5 % sampling floor -- every 20th tile -- already puts k / 20 <= 10.7 of the true top k in the
    // sample: rank 32 is as safe there (P[Poisson(10.7) >= 32] ~ 1e-7) and collects 640 rows per query, not 1280)
    const int thr_rank = mma_sample ? ((3 * k) / 32 < 20 ? 32 : kLargeKThrRank) : ks;
    // sample every `stride`-th tile: at least 3k expected candidates, and never more than 5 % of a scan
    const uint32_t stride = std::max<uint32_t>(20u, (3u * (uint32_t)k) / (uint32_t)thr_rank);
    TRY(c->lk_keys.ensure(key_bytes));
    TRY(c->lk_misc.ensure((size_t)n_queries * (4 + 4 + kLargeKThrRank * 12) + 64));
    uint32_t* thr = static_cast<uint32_t*>(c->lk_misc.p);
    uint32_t* cnt = thr + n_queries;
    uint64_t* smp_r = reinterpret_cast<uint64_t*>(cnt + n_queries);  // 8*Q bytes in: 8-byte aligned
    uint32_t* smp_s = reinterpret_cast<uint32_t*>(smp_r + (size_t)n_queries * kLargeKThrRank);
    int* flags = c->d_err + 1;   // the fallback gate (d_err[0] is the repack error flag)
    if (c->count <= cap) {
        // the whole shard fits a buffer: no sampling, collect every eligible row
        CU(cudaMemsetAsync(thr, 0xFF, (size_t)n_queries * 4, st));
    } else {
        ScanParams ps = p;
        ps.tile_stride = stride;
        TRY(search_lists(c, ps, n_queries, ks, metric, smp_s, smp_r, 0, nullptr, thr_rank));
        extract_thr_kernel<<<(n_queries + 127) / 128, 128, 0, st>>>(smp_s, smp_r, n_queries, thr_rank, thr);
        LAUNCHED();
    }
    CU(cudaMemsetAsync(cnt, 0, (size_t)n_queries * 4, st));
    ScanParams pc = p;
    pc.k = 1;
    pc.collect_thr = thr;
    pc.collect_keys = static_cast<unsigned long long*>(c->lk_keys.p);
    pc.collect_count = cnt;
    pc.collect_cap = cap;
    TRY(c->cand_buf.ensure((size_t)n_queries * 12));
    pc.tau = static_cast<uint32_t*>(c->cand_buf.p);
    TRY(c->part_s.ensure(4096));
    TRY(c->part_r.ensure(4096));
    pc.part_scores = static_cast<uint32_t*>(c->part_s.p);
    pc.part_rows = static_cast<uint64_t*>(c->part_r.p);
    pc.evict_first = 0;
    TRY(launch_scan(c, pc, n_queries, 1, metric, nullptr, /*dry=*/false));
    if (c->timing) CU(cudaEventRecord(c->ev1, st));
    const size_t smem = (size_t)cap * 8;
    TRY(allow_max_smem(collect_select_kernel, c->device));
    collect_select_kernel<<<n_queries, kSelectThreads, smem, st>>>(pc.collect_keys, cnt, thr, cap, k, c->base_row, o_s,
                                                                o_r, flags);
    LAUNCHED();
    // Overflow (mass ties at the threshold) or shortfall is only known on the device: the exact
    // multi-pass path is ENQUEUED behind the select kernel with its launches gated on the flag word
    // (every gated kernel returns at once when the flag is 0), then the flag is cleared.  No host
    // round trip: k > 32 is enqueue-only like k <= 32, so it pipelines and can be graph-captured.
    TRY(search_lists(c, p, n_queries, k, metric, o_s, o_r, 0, flags));
    CU(cudaMemsetAsync(flags, 0, sizeof(int), st));
    *done = true;
    return VL_OK;
}

}  // namespace
