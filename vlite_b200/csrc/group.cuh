// vl_shard_group_*: several shards (usually one per GPU) searched as one corpus from one process.
// Included at the end of api.cu (it calls vl_search).
#pragma once

// =============================== shard group ===============================
// Several shards (usually one per GPU of the box) searched as ONE corpus from a single host
// thread: every shard scans on its own device and stream, its k candidates travel to the home
// device by a peer copy (NVLink when the devices are peers), and one merge on the home device
// orders them by (score, global row).  This is the in-process twin of the one-process-per-GPU
// path (vlite_b200/sharded.py: NCCL all-gather + the same merge kernel).
struct vl_shard_group {
    std::vector<vl_corpus*> shards;
    int home = 0;
    cudaStream_t stream = nullptr;       // on the home device
    std::vector<cudaEvent_t> ev;         // per shard, created on the shard's device
    std::vector<DevBuf> q, s, r;         // per shard, on the shard's device
    std::vector<DevBuf> seg;             // per shard: [local starts | global starts] of its appended segments
    std::vector<int> n_seg;
    std::vector<char> direct;            // per shard: its kernels may store straight into home-device memory
    DevBuf gather_s, gather_r, out_s, out_r;  // home device
    void* pinned = nullptr;              // host staging: queries in, results out
    size_t pinned_bytes = 0;
    cudaEvent_t home_ev = nullptr;
    // one persistent host thread per shard: a search's per-shard enqueue work (H2D, memsets, scan,
    // merges: ~20 us of launch calls each) runs on all of them at once instead of back to back
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    uint64_t epoch = 0;
    int pending = 0;
    bool stop = false;
    std::function<int(int)> job;
    std::vector<int> rcs;
    std::vector<std::string> msgs;
};

namespace {
int group_pinned(vl_shard_group* g, size_t bytes) {
    if (bytes <= g->pinned_bytes) return VL_OK;
    if (g->pinned) cudaFreeHost(g->pinned);
    g->pinned = nullptr;
    g->pinned_bytes = 0;
    CU(cudaHostAlloc(&g->pinned, bytes, cudaHostAllocPortable));
    g->pinned_bytes = bytes;
    return VL_OK;
}
}  // namespace

extern "C" int vl_shard_group_create(int n_shards, vl_corpus* const* shards, vl_shard_group** out) {
    if (!out || !shards) return fail(VL_EINVAL, "NULL argument");
    *out = nullptr;
    if (n_shards < 1 || n_shards > 128) return fail(VL_EINVAL, "a group holds 1..128 shards");
    for (int i = 0; i < n_shards; ++i) {
        if (!shards[i]) return fail(VL_EINVAL, "shard %d is NULL", i);
        if (shards[i]->w != shards[0]->w) return fail(VL_EINVAL, "shard %d has a different row width", i);
        for (int j = 0; j < i; ++j)
            if (shards[j] == shards[i]) return fail(VL_EINVAL, "shard %d appears twice", i);
    }
    auto* g = new vl_shard_group();
    g->shards.assign(shards, shards + n_shards);
    g->home = shards[0]->device;
    g->ev.assign(n_shards, nullptr);
    g->q.resize(n_shards);
    g->s.resize(n_shards);
    g->r.resize(n_shards);
    g->seg.resize(n_shards);
    g->n_seg.assign(n_shards, 0);
    g->direct.assign(n_shards, 0);
    int rc = VL_OK;
    {
        DeviceGuard dg(g->home);
        if (cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&g->home_ev, cudaEventDisableTiming) != cudaSuccess)
            rc = fail(VL_ECUDA, "cannot create the group's stream on device %d", g->home);
    }
    for (int i = 0; i < n_shards && rc == VL_OK; ++i) {
        const int dev = shards[i]->device;
        DeviceGuard dg(dev);
        if (cudaEventCreateWithFlags(&g->ev[i], cudaEventDisableTiming) != cudaSuccess)
            rc = fail(VL_ECUDA, "cannot create an event on device %d", dev);
        if (dev == g->home) {
            g->direct[i] = 1;
        } else {  // peers: the shard's merge kernel stores its k candidates straight into the home
                  // device's gather buffer over NVLink; otherwise a peer copy (staged by the driver) does
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, dev, g->home) == cudaSuccess && can) {
                cudaError_t e = cudaDeviceEnablePeerAccess(g->home, 0);
                if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) g->direct[i] = 1;
                if (e != cudaSuccess) cudaGetLastError();
            }
        }
    }
    if (rc != VL_OK) {
        vl_shard_group_destroy(g);
        return rc;
    }
    g->rcs.assign(n_shards, VL_OK);
    g->msgs.assign(n_shards, std::string());
    if (n_shards > 1) {
        for (int i = 0; i < n_shards; ++i)
            g->workers.emplace_back([g, i] {
                uint64_t seen = 0;
                for (;;) {
                    std::unique_lock<std::mutex> lk(g->mu);
                    g->cv_go.wait(lk, [&] { return g->stop || g->epoch != seen; });
                    if (g->stop) return;
                    seen = g->epoch;
                    lk.unlock();
                    const int r = g->job(i);
                    lk.lock();
                    g->rcs[i] = r;
                    if (r != VL_OK) g->msgs[i] = g_err;  // the error text is thread-local
                    if (--g->pending == 0) g->cv_done.notify_one();
                }
            });
    }
    *out = g;
    return VL_OK;
}

extern "C" int vl_shard_group_destroy(vl_shard_group* g) {
    if (!g) return VL_OK;
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->stop = true;
    }
    g->cv_go.notify_all();
    for (auto& t : g->workers) t.join();
    g->workers.clear();
    for (size_t i = 0; i < g->shards.size(); ++i) {
        DeviceGuard dg(g->shards[i]->device);
        if (g->ev[i]) cudaEventDestroy(g->ev[i]);
        g->q[i].release();
        g->s[i].release();
        g->r[i].release();
        g->seg[i].release();
    }
    {
        DeviceGuard dg(g->home);
        if (g->stream) cudaStreamSynchronize(g->stream);
        g->gather_s.release();
        g->gather_r.release();
        g->out_s.release();
        g->out_r.release();
        if (g->home_ev) cudaEventDestroy(g->home_ev);
        if (g->stream) cudaStreamDestroy(g->stream);
        if (g->pinned) cudaFreeHost(g->pinned);
    }
    delete g;
    return VL_OK;
}

extern "C" int vl_shard_group_set_segments(vl_shard_group* g, int shard, const uint64_t* seg_local,
                                           const uint64_t* seg_global, int n_segments) {
    if (!g) return fail(VL_EINVAL, "group is NULL");
    if (shard < 0 || shard >= (int)g->shards.size()) return fail(VL_EINVAL, "shard %d out of range", shard);
    if (n_segments < 0 || (n_segments > 0 && (!seg_local || !seg_global))) return fail(VL_EINVAL, "bad segment table");
    for (int i = 1; i < n_segments; ++i)
        if (seg_local[i] <= seg_local[i - 1] || seg_global[i] <= seg_global[i - 1])
            return fail(VL_EINVAL, "segment starts must ascend");
    vl_corpus* c = g->shards[shard];
    DeviceGuard dg(c->device);
    CU(cudaStreamSynchronize(c->stream));  // no search may still be reading the old table
    g->n_seg[shard] = n_segments;
    if (n_segments == 0) return VL_OK;
    TRY(g->seg[shard].ensure((size_t)n_segments * 16));
    uint64_t* d = static_cast<uint64_t*>(g->seg[shard].p);
    CU(cudaMemcpy(d, seg_local, (size_t)n_segments * 8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d + n_segments, seg_global, (size_t)n_segments * 8, cudaMemcpyHostToDevice));
    return VL_OK;
}

extern "C" int vl_shard_group_rows(const vl_shard_group* g, uint64_t* rows, uint64_t* live_rows) {
    if (!g) return fail(VL_EINVAL, "group is NULL");
    uint64_t n = 0, l = 0;
    for (const vl_corpus* c : g->shards) {
        n += c->count;
        l += c->live_count;
    }
    if (rows) *rows = n;
    if (live_rows) *live_rows = l;
    return VL_OK;
}

extern "C" int vl_shard_group_search(vl_shard_group* g, const uint8_t* queries_host, int n_queries, int k, int metric,
                                     const uint32_t* const* filter_masks_or_null, uint32_t* out_score_host,
                                     uint64_t* out_row_host) {
    if (!g) return fail(VL_EINVAL, "group is NULL");
    if (!queries_host || !out_score_host || !out_row_host) return fail(VL_EINVAL, "NULL argument");
    if (n_queries < 1) return fail(VL_EINVAL, "n_queries must be >= 1");
    if (k < 1 || k > VL_MAX_K) return fail(VL_EINVAL, "k must be in [1, %d]", VL_MAX_K);
    if (metric != VL_METRIC_HAMMING && metric != VL_METRIC_XORSUM) return fail(VL_EINVAL, "bad metric %d", metric);
    if (is_device_ptr(queries_host) || is_device_ptr(out_score_host) || is_device_ptr(out_row_host))
        return fail(VL_EINVAL, "vl_shard_group_search takes host pointers");
    const int n = (int)g->shards.size();
    const int w = g->shards[0]->w;
    const size_t qbytes = (size_t)n_queries * w, n_out = (size_t)n_queries * k;
    // pinned staging: [queries | scores | rows]
    const size_t off_s = (qbytes + 255) & ~(size_t)255, off_r = off_s + ((n_out * 4 + 255) & ~(size_t)255);
    {
        DeviceGuard dg(g->home);
        TRY(group_pinned(g, off_r + n_out * 8));
        TRY(g->gather_s.ensure((size_t)n * n_out * 4));
        TRY(g->gather_r.ensure((size_t)n * n_out * 8));
        TRY(g->out_s.ensure(n_out * 4));
        TRY(g->out_r.ensure(n_out * 8));
        // the previous call's result copy must have left the staging buffer
        CU(cudaStreamSynchronize(g->stream));
    }
    memcpy(g->pinned, queries_host, qbytes);

    // Phase 1: every shard's H2D + scan + local merge is enqueued by that shard's own host thread, so
    // the launch calls of all shards overlap and all devices start together (k > 32 synchronises
    // inside vl_search -- the sampled threshold sizes its collect pass -- which the same threads absorb).
    auto shard_work = [&](int i) -> int {
        vl_corpus* c = g->shards[i];
        DeviceGuard dg(c->device);
        TRY(g->q[i].ensure(qbytes));
        CU(cudaMemcpyAsync(g->q[i].p, g->pinned, qbytes, cudaMemcpyHostToDevice, c->stream));
        uint32_t* slot_s = static_cast<uint32_t*>(g->gather_s.p) + (size_t)i * n_out;   // home-device memory
        uint64_t* slot_r = static_cast<uint64_t*>(g->gather_r.p) + (size_t)i * n_out;
        uint32_t* o_s = slot_s;
        uint64_t* o_r = slot_r;
        if (!g->direct[i]) {
            TRY(g->s[i].ensure(n_out * 4));
            TRY(g->r[i].ensure(n_out * 8));
            o_s = static_cast<uint32_t*>(g->s[i].p);
            o_r = static_cast<uint64_t*>(g->r[i].p);
        }
        // the exchange is the epilogue of the shard's own search: its last merge kernel writes the k
        // candidates per query into the home device's gather slot (peer stores when direct)
        TRY(vl_search(c, static_cast<const uint8_t*>(g->q[i].p), n_queries, k, metric,
                      filter_masks_or_null ? filter_masks_or_null[i] : nullptr, o_s, o_r));
        if (g->n_seg[i] > 0) {  // shard-local rows -> global rows fixed at append time
            const uint64_t* sl = static_cast<const uint64_t*>(g->seg[i].p);
            map_rows_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, c->stream>>>(o_r, n_out, sl, sl + g->n_seg[i],
                                                                                  g->n_seg[i]);
            LAUNCHED();
        }
        if (!g->direct[i]) {
            CU(cudaMemcpyPeerAsync(slot_s, g->home, o_s, c->device, n_out * 4, c->stream));
            CU(cudaMemcpyPeerAsync(slot_r, g->home, o_r, c->device, n_out * 8, c->stream));
        }
        CU(cudaEventRecord(g->ev[i], c->stream));
        return VL_OK;
    };
    if (n == 1) {
        TRY(shard_work(0));
    } else {
        {
            std::lock_guard<std::mutex> lk(g->mu);
            g->job = shard_work;
            g->pending = n;
            ++g->epoch;
        }
        g->cv_go.notify_all();
        {
            std::unique_lock<std::mutex> lk(g->mu);
            g->cv_done.wait(lk, [&] { return g->pending == 0; });
            g->job = nullptr;
        }
        for (int i = 0; i < n; ++i)
            if (g->rcs[i] != VL_OK) return fail(g->rcs[i], "shard %d: %s", i, g->msgs[i].c_str());
    }

    // Phase 2: the home device waits for every shard's candidates, merges, and returns them.
    DeviceGuard dg(g->home);
    for (int i = 0; i < n; ++i) CU(cudaStreamWaitEvent(g->stream, g->ev[i], 0));
    if (n == 1) {
        CU(cudaMemcpyAsync(g->out_s.p, g->gather_s.p, n_out * 4, cudaMemcpyDeviceToDevice, g->stream));
        CU(cudaMemcpyAsync(g->out_r.p, g->gather_r.p, n_out * 8, cudaMemcpyDeviceToDevice, g->stream));
    } else {
        TRY(merge_levels(g->stream, static_cast<const uint32_t*>(g->gather_s.p),
                         static_cast<const uint64_t*>(g->gather_r.p), (size_t)n, n_out, n_out, n_queries, k, k,
                         static_cast<uint32_t*>(g->out_s.p), static_cast<uint64_t*>(g->out_r.p), k, 0, nullptr, nullptr));
    }
    uint8_t* pin = static_cast<uint8_t*>(g->pinned);
    CU(cudaMemcpyAsync(pin + off_s, g->out_s.p, n_out * 4, cudaMemcpyDeviceToHost, g->stream));
    CU(cudaMemcpyAsync(pin + off_r, g->out_r.p, n_out * 8, cudaMemcpyDeviceToHost, g->stream));
    CU(cudaStreamSynchronize(g->stream));
    memcpy(out_score_host, pin + off_s, n_out * 4);
    memcpy(out_row_host, pin + off_r, n_out * 8);
    return VL_OK;
}
