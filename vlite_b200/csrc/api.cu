// C ABI of vlite_b200 (see include/vlite_b200.h): corpus residency, search, rescore, CTX
// embeddings load, measurement helpers.  Host-side orchestration only; the kernels live in
// scan.cuh / merge.cuh / rescore.cuh / repack.cuh / synth.cuh.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/vlite_b200.h"
#include "largek.cuh"
#include "merge.cuh"
#include "repack.cuh"
#include "rescore.cuh"
#include "scan.cuh"
#include "mmascan.cuh"
#include "synth.cuh"
#include "exchange.cuh"

#include "host_util.cuh"
#include "corpus.cuh"
#include "search_host.cuh"

// =============================== library ===============================
extern "C" const char* vl_version(void) { return "vlite_b200 0.1.0 (sm_100a)"; }
extern "C" const char* vl_last_error(void) { return g_err.c_str(); }
extern "C" uint64_t vl_launch_count(void) { return g_launches.load(); }

extern "C" int vl_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, i) == cudaSuccess && prop.major == 10) ++ok;
    }
    return ok;
}

// =============================== corpus ===============================
extern "C" int vl_corpus_create(int device, int width_bytes, uint64_t capacity_rows, vl_corpus** out) {
    if (!out) return fail(VL_EINVAL, "out is NULL");
    *out = nullptr;
    if (width_bytes != 32 && width_bytes != 64 && width_bytes != 128)
        return fail(VL_EINVAL, "width_bytes must be 32, 64 or 128 (got %d)", width_bytes);
    int sms = 0;
    TRY(check_device(device, &sms));
    DeviceGuard g(device);
    if (!g.ok) return fail(VL_ECUDA, "cudaSetDevice(%d) failed", device);
    vl_corpus* c = new vl_corpus();
    c->device = device;
    c->w = width_bytes;
    c->sms = sms;
    auto bail = [&](int rc) {
        vl_corpus_destroy(c);
        return rc;
    };
    if (cudaMalloc(&c->d_err, 4 * sizeof(int)) != cudaSuccess || cudaMalloc(&c->d_counter, 8) != cudaSuccess)
        return bail(fail(VL_ENOMEM, "cudaMalloc failed"));
    cudaMemset(c->d_err, 0, 4 * sizeof(int));   // [0] repack error, [1] large-k gate, [2] fused-merge ticket
    cudaMemset(c->d_counter, 0, 8);
    cudaEventCreate(&c->ev0);
    cudaEventCreate(&c->ev1);
    cudaEventCreate(&c->ev2);
    for (int i = 0; i < 2; ++i) cudaEventCreateWithFlags(&c->stage_ev[i], cudaEventDisableTiming);
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return bail(fail(VL_ECUDA, "cudaMemGetInfo failed"));
    // address space is free: reserve room for a shard as large as the whole device
    int rc = c->rows_arena.init(device, total_b);
    if (rc == VL_OK) rc = c->live_arena.init(device, total_b / ((size_t)width_bytes * 8) + (1 << 21));
    if (rc != VL_OK) return bail(rc);
    c->rows = reinterpret_cast<uint8_t*>(c->rows_arena.base);
    c->live = reinterpret_cast<uint32_t*>(c->live_arena.base);
    rc = capacity_rows ? grow_to(c, capacity_rows) : VL_OK;   // exact reservation; 0 = map on first append
    if (rc != VL_OK) return bail(rc);
    *out = c;
    return VL_OK;
}

extern "C" int vl_corpus_destroy(vl_corpus* c) {
    if (!c) return VL_OK;
    DeviceGuard g(c->device);
    cudaStreamSynchronize(c->stream);
    c->rows_arena.destroy();
    c->live_arena.destroy();
    for (DevBuf* b : {&c->q_buf, &c->filt_buf, &c->part_s, &c->part_r, &c->out_s, &c->out_r, &c->stage[0],
                      &c->stage[1], &c->cand_buf, &c->misc, &c->lk_keys, &c->lk_misc, &c->qx_buf})
        b->release();
    for (int i = 0; i < 2; ++i) {
        if (c->pinned[i]) cudaFreeHost(c->pinned[i]);
        if (c->stage_ev[i]) cudaEventDestroy(c->stage_ev[i]);
    }
    if (c->io_pinned) cudaFreeHost(c->io_pinned);
    if (c->trace_dev) cudaFree(c->trace_dev);
    if (c->d_err) cudaFree(c->d_err);
    if (c->d_counter) cudaFree(c->d_counter);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev2) cudaEventDestroy(c->ev2);
    delete c;
    return VL_OK;
}

extern "C" int vl_corpus_set_stream(vl_corpus* c, void* cuda_stream) {
    if (!c) return fail(VL_EINVAL, "corpus is NULL");
    DeviceGuard g(c->device);
    CU(cudaStreamSynchronize(c->stream));
    c->stream = static_cast<cudaStream_t>(cuda_stream);
    return VL_OK;
}

extern "C" int vl_corpus_reserve(vl_corpus* c, uint64_t capacity_rows) {
    if (!c) return fail(VL_EINVAL, "corpus is NULL");
    DeviceGuard g(c->device);
    if (capacity_rows <= c->capacity) return VL_OK;
    return grow_to(c, capacity_rows);  // exact: a 100+ GB shard must not be doubled
}

extern "C" int vl_corpus_append(vl_corpus* c, const uint8_t* rows_any, uint64_t n, uint64_t* first_row_out) {
    if (!c) return fail(VL_EINVAL, "corpus is NULL");
    if (n && !rows_any) return fail(VL_EINVAL, "rows is NULL");
    DeviceGuard g(c->device);
    const uint64_t first = c->count;
    if (first_row_out) *first_row_out = first;
    if (n == 0) return VL_OK;
    TRY(ensure_capacity(c, first + n));
    const bool dev = is_device_ptr(rows_any);
    CU(cudaMemcpyAsync(c->rows + first * (uint64_t)c->w, rows_any, n * (uint64_t)c->w,
                       dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
    TRY(mark_appended(c, first, n));
    if (!dev) CU(cudaStreamSynchronize(c->stream));
    return VL_OK;
}

extern "C" int vl_corpus_append_f32(vl_corpus* c, const float* vals_any, uint64_t n, int src_kind,
                                    uint64_t* first_row_out) {
    if (!c) return fail(VL_EINVAL, "corpus is NULL");
    if (src_kind != VL_SRC_F32_OFFSET && src_kind != VL_SRC_F32_BITS)
        return fail(VL_EINVAL, "src_kind must be VL_SRC_F32_OFFSET or VL_SRC_F32_BITS");
    if (n && !vals_any) return fail(VL_EINVAL, "vals is NULL");
    DeviceGuard g(c->device);
    const uint64_t first = c->count;
    if (first_row_out) *first_row_out = first;
    if (n == 0) return VL_OK;
    TRY(ensure_capacity(c, first + n));
    const size_t rb = src_row_bytes(c->w, src_kind);
    if (is_device_ptr(vals_any)) {
        if (reinterpret_cast<uintptr_t>(vals_any) & 15) return fail(VL_EINVAL, "device f32 rows must be 16 B aligned");
        TRY(repack_rows(c, vals_any, first, n, src_kind));
    } else {
        const uint64_t chunk_rows = std::max<uint64_t>(1, (32ull << 20) / rb);
        for (uint64_t done = 0, i = 0; done < n; done += chunk_rows, ++i) {
            const uint64_t m = std::min(chunk_rows, n - done);
            DevBuf& sb = c->stage[i & 1];
            TRY(sb.ensure(chunk_rows * rb));
            CU(cudaMemcpyAsync(sb.p, reinterpret_cast<const uint8_t*>(vals_any) + done * rb, m * rb,
                               cudaMemcpyHostToDevice, c->stream));
            TRY(repack_rows(c, static_cast<const float*>(sb.p), first + done, m, src_kind));
        }
    }
    // validate BEFORE the rows go live (like vl_corpus_load_file): on a bad value nothing was
    // appended -- count, live mask and every tombstone stay exactly as they were
    TRY(check_err_flag(c, "vl_corpus_append_f32"));
    TRY(mark_appended(c, first, n));
    return VL_OK;
}

extern "C" int vl_corpus_load_file(vl_corpus* c, const char* path, uint64_t byte_off, uint64_t byte_len,
                                   int src_kind, uint64_t max_rows, uint64_t* first_row_out,
                                   uint64_t* n_rows_out) {
    if (!c || !path) return fail(VL_EINVAL, "NULL argument");
    if (src_kind < VL_SRC_U8 || src_kind > VL_SRC_F32_BITS) return fail(VL_EINVAL, "bad src_kind %d", src_kind);
    DeviceGuard g(c->device);
    const size_t rb = src_row_bytes(c->w, src_kind);
    uint64_t n = byte_len / rb;  // floor, like vlite/ctx.py:115
    if (max_rows && n > max_rows) n = max_rows;
    const uint64_t first = c->count;
    if (first_row_out) *first_row_out = first;
    if (n_rows_out) *n_rows_out = n;
    if (n == 0) return VL_OK;
    int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(VL_EIO, "cannot open %s", path);
    int rc = ensure_capacity(c, first + n);
    const uint64_t chunk_rows = std::max<uint64_t>(1, (32ull << 20) / rb);
    if (rc == VL_OK) rc = ensure_pinned(c, chunk_rows * rb);
    // page cache -> pinned staging with several readers (one pread stream tops out near 4 GB/s).  The
    // readers are started ONCE per call and fed chunk after chunk (a generation counter + two condition
    // variables), instead of spawning and joining a set of threads for every 32 MB chunk.
    static const unsigned io_threads = [] {
        const char* e = getenv("VLITE_B200_IO_THREADS");
        const unsigned want_thr = e ? (unsigned)atoi(e) : 8u;  // measured: 8 readers reach the file rate, 32 lose
        return std::max(1u, std::min(std::min(want_thr, 64u), std::max(1u, std::thread::hardware_concurrency())));
    }();
    const unsigned n_thr = io_threads;
    struct {
        std::mutex mu;
        std::condition_variable go, done;
        uint64_t gen = 0;
        unsigned pending = 0;
        bool quit = false;
        uint8_t* dst = nullptr;
        size_t want = 0;
        uint64_t file_off = 0;
    } job;
    std::vector<size_t> got_part(n_thr, 0);
    std::vector<std::thread> pool;
    if (rc == VL_OK) {
        for (unsigned t = 0; t < n_thr; ++t) {
            pool.emplace_back([&, t] {
                uint64_t seen = 0;
                for (;;) {
                    uint8_t* dst;
                    size_t want;
                    uint64_t off;
                    {
                        std::unique_lock<std::mutex> lk(job.mu);
                        job.go.wait(lk, [&] { return job.quit || job.gen != seen; });
                        if (job.quit) return;
                        seen = job.gen;
                        dst = job.dst;
                        want = job.want;
                        off = job.file_off;
                    }
                    const size_t slice = ((want + n_thr - 1) / n_thr + 4095) & ~(size_t)4095;
                    const size_t lo = std::min(want, (size_t)t * slice), hi = std::min(want, lo + slice);
                    size_t g = lo;
                    while (g < hi) {
                        ssize_t r = pread(fd, dst + g, hi - g, off + g);
                        if (r <= 0) break;
                        g += (size_t)r;
                    }
                    got_part[t] = g - lo;
                    {
                        std::lock_guard<std::mutex> lk(job.mu);
                        if (--job.pending == 0) job.done.notify_one();
                    }
                }
            });
        }
    }
    for (uint64_t done = 0, i = 0; rc == VL_OK && done < n; done += chunk_rows, ++i) {
        const int b = (int)(i & 1);
        const uint64_t m = std::min(chunk_rows, n - done);
        if (i >= 2 && cudaEventSynchronize(c->stage_ev[b]) != cudaSuccess) {
            rc = fail(VL_ECUDA, "event sync failed");
            break;
        }
        const size_t want = m * rb;
        {
            std::unique_lock<std::mutex> lk(job.mu);
            job.dst = static_cast<uint8_t*>(c->pinned[b]);
            job.want = want;
            job.file_off = byte_off + done * rb;
            job.pending = n_thr;
            ++job.gen;
            job.go.notify_all();
            job.done.wait(lk, [&] { return job.pending == 0; });
        }
        size_t got = 0;
        for (size_t g : got_part) got += g;
        if (got != want) {
            rc = fail(VL_EIO, "short read from %s", path);
            break;
        }
        cudaError_t e;
        if (src_kind == VL_SRC_U8) {
            e = cudaMemcpyAsync(c->rows + (first + done) * (uint64_t)c->w, c->pinned[b], want, cudaMemcpyHostToDevice,
                                c->stream);
        } else {
            rc = c->stage[b].ensure(chunk_rows * rb);
            if (rc != VL_OK) break;
            e = cudaMemcpyAsync(c->stage[b].p, c->pinned[b], want, cudaMemcpyHostToDevice, c->stream);
            if (e == cudaSuccess) rc = repack_rows(c, static_cast<const float*>(c->stage[b].p), first + done, m, src_kind);
        }
        if (e != cudaSuccess) {
            rc = fail(VL_ECUDA, "H2D copy failed: %s", cudaGetErrorString(e));
            break;
        }
        cudaEventRecord(c->stage_ev[b], c->stream);
    }
    {
        std::lock_guard<std::mutex> lk(job.mu);
        job.quit = true;
    }
    job.go.notify_all();
    for (auto& th : pool) th.join();
    close(fd);
    cudaStreamSynchronize(c->stream);
    if (rc != VL_OK) return rc;
    if (src_kind != VL_SRC_U8) TRY(check_err_flag(c, "vl_corpus_load_file"));
    TRY(mark_appended(c, first, n));
    CU(cudaStreamSynchronize(c->stream));
    return VL_OK;
}

extern "C" int vl_corpus_append_sign_i8(vl_corpus* c, const int8_t* docs_i8_dev, uint64_t n, uint64_t* first_row_out) {
    if (!c) return fail(VL_EINVAL, "corpus is NULL");
    if (n && (!docs_i8_dev || !is_device_ptr(docs_i8_dev))) return fail(VL_EINVAL, "docs must be device memory");
    if (reinterpret_cast<uintptr_t>(docs_i8_dev) & 15) return fail(VL_EINVAL, "docs must be 16 B aligned");
    DeviceGuard g(c->device);
    const uint64_t first = c->count;
    if (first_row_out) *first_row_out = first;
    if (n == 0) return VL_OK;
    TRY(ensure_capacity(c, first + n));
    const uint64_t n_words = n * (uint64_t)c->w / 4;
    sign_i8_kernel<<<grid_for(n_words, 256, c->sms), 256, 0, c->stream>>>(
        docs_i8_dev, n_words, reinterpret_cast<uint32_t*>(c->rows + first * (uint64_t)c->w));
    LAUNCHED();
    return mark_appended(c, first, n);
}

extern "C" int vl_corpus_set_row(vl_corpus* c, uint64_t row, const uint8_t* row_any) {
    if (!c || !row_any) return fail(VL_EINVAL, "NULL argument");
    if (row >= c->count) return fail(VL_EINVAL, "row %llu out of range", (unsigned long long)row);
    DeviceGuard g(c->device);
    const bool dev = is_device_ptr(row_any);
    CU(cudaMemcpyAsync(c->rows + row * (uint64_t)c->w, row_any, c->w,
                       dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
    if (!dev) CU(cudaStreamSynchronize(c->stream));
    return VL_OK;
}

extern "C" int vl_corpus_tombstone(vl_corpus* c, const uint64_t* rows_host, uint64_t n) {
    if (!c) return fail(VL_EINVAL, "corpus is NULL");
    if (n == 0) return VL_OK;
    if (!rows_host) return fail(VL_EINVAL, "rows is NULL");
    DeviceGuard g(c->device);
    TRY(c->misc.ensure(n * 8));
    CU(cudaMemsetAsync(c->d_counter, 0, 8, c->stream));
    CU(cudaMemcpyAsync(c->misc.p, rows_host, n * 8, cudaMemcpyHostToDevice, c->stream));
    live_clear_rows_kernel<<<grid_for(n, 256, c->sms), 256, 0, c->stream>>>(
        c->live, static_cast<const uint64_t*>(c->misc.p), n, c->count, c->d_counter);
    LAUNCHED();
    unsigned long long cleared = 0;
    CU(cudaMemcpyAsync(&cleared, c->d_counter, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->live_count -= cleared;
    return VL_OK;
}

extern "C" int vl_corpus_clear(vl_corpus* c) {
    if (!c) return fail(VL_EINVAL, "corpus is NULL");
    DeviceGuard g(c->device);
    if (c->capacity) CU(cudaMemsetAsync(c->live, 0, c->capacity / 8, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->count = 0;
    c->live_count = 0;
    return VL_OK;
}

extern "C" int vl_corpus_rows(const vl_corpus* c, uint64_t* rows, uint64_t* live_rows) {
    if (!c) return fail(VL_EINVAL, "corpus is NULL");
    if (rows) *rows = c->count;
    if (live_rows) *live_rows = c->live_count;
    return VL_OK;
}

extern "C" int vl_corpus_width(const vl_corpus* c) { return c ? c->w : VL_EINVAL; }

extern "C" int vl_corpus_set_base_row(vl_corpus* c, uint64_t base_row) {
    if (!c) return fail(VL_EINVAL, "corpus is NULL");
    c->base_row = base_row;
    return VL_OK;
}

extern "C" int vl_corpus_read(vl_corpus* c, uint64_t first, uint64_t n, uint8_t* out_host) {
    if (!c || (n && !out_host)) return fail(VL_EINVAL, "NULL argument");
    if (first + n > c->count) return fail(VL_EINVAL, "rows [%llu, %llu) out of range", (unsigned long long)first,
                                          (unsigned long long)(first + n));
    DeviceGuard g(c->device);
    if (n == 0) return VL_OK;
    CU(cudaMemcpyAsync(out_host, c->rows + first * (uint64_t)c->w, n * (uint64_t)c->w, cudaMemcpyDeviceToHost,
                       c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return VL_OK;
}

extern "C" int vl_corpus_device_ptrs(vl_corpus* c, void** rows_dev, void** live_mask_dev) {
    if (!c) return fail(VL_EINVAL, "corpus is NULL");
    if (rows_dev) *rows_dev = c->rows;
    if (live_mask_dev) *live_mask_dev = c->live;
    return VL_OK;
}

extern "C" int vl_corpus_fill_synthetic(vl_corpus* c, uint64_t seed, uint64_t first, uint64_t n, uint64_t period) {
    if (!c) return fail(VL_EINVAL, "corpus is NULL");
    if (first > c->count) return fail(VL_EINVAL, "first (%llu) is past the end of the corpus", (unsigned long long)first);
    DeviceGuard g(c->device);
    if (n == 0) return VL_OK;
    TRY(ensure_capacity(c, first + n));
    synth_fill_kernel<<<grid_for(n * (uint64_t)(c->w / 8), 256, c->sms), 256, 0, c->stream>>>(
        c->rows, c->w, seed, first, n, period, c->base_row);
    LAUNCHED();
    if (first + n > c->count) {
        const uint64_t added_first = c->count;
        TRY(mark_appended(c, added_first, first + n - added_first));
    }
    return VL_OK;
}

extern "C" int vl_synth_filter_mask(vl_corpus* c, uint64_t seed, uint64_t threshold, uint32_t* mask_dev) {
    if (!c || !mask_dev) return fail(VL_EINVAL, "NULL argument");
    if (!is_device_ptr(mask_dev)) return fail(VL_EINVAL, "mask must be device memory");
    DeviceGuard g(c->device);
    const uint64_t n_words = (c->count + 31) / 32;
    if (n_words == 0) return VL_OK;
    synth_mask_kernel<<<grid_for(n_words, 128, c->sms), 128, 0, c->stream>>>(mask_dev, c->count, n_words, seed,
                                                                            c->base_row, threshold);
    LAUNCHED();
    return VL_OK;
}

// =============================== search ===============================
extern "C" int vl_search(vl_corpus* c, const uint8_t* queries_any, int n_queries, int k, int metric,
                         const uint32_t* filter_mask_any, uint32_t* out_score_any, uint64_t* out_row_any) {
    if (!c) return fail(VL_EINVAL, "corpus is NULL");
    if (!queries_any || !out_score_any || !out_row_any) return fail(VL_EINVAL, "NULL argument");
    if (n_queries < 1) return fail(VL_EINVAL, "n_queries must be >= 1");
    if (k < 1 || k > VL_MAX_K) return fail(VL_EINVAL, "k must be in [1, %d]", VL_MAX_K);
    if (metric != VL_METRIC_HAMMING && metric != VL_METRIC_XORSUM) return fail(VL_EINVAL, "bad metric %d", metric);
    DeviceGuard g(c->device);
    cudaStream_t st = c->stream;
    const int q_kind = ptr_kind(queries_any), s_kind = ptr_kind(out_score_any), r_kind = ptr_kind(out_row_any);
    const bool q_dev = q_kind == 2, s_dev = s_kind == 2, r_dev = r_kind == 2;
    const size_t n_out = (size_t)n_queries * k;
    c->timed = false;
    if (c->count == 0 || c->live_count == 0) {
        TRY(fill_sentinels(st, out_score_any, out_row_any, n_out, s_dev, r_dev));
        return VL_OK;
    }

    // Pageable host buffers of a small call go through the handle's pinned staging block: one memcpy
    // on the host, truly asynchronous copies on the stream, and ONE device -> host copy for
    // [rows | scores] instead of two staged driver copies (the latency path: 1 query, k = 5).
    const size_t q_bytes = (size_t)n_queries * c->w;
    const bool stage_q = q_kind == 0 && q_bytes <= (4u << 20);
    const bool stage_o = s_kind == 0 && r_kind == 0 && n_out * 12 <= (4u << 20);
    const size_t q_slot = (q_bytes + 63) & ~(size_t)63;
    if (stage_q || stage_o) TRY(ensure_io_pinned(c, q_slot + n_out * 12));

    const uint8_t* q_ptr = queries_any;
    if (!q_dev || (reinterpret_cast<uintptr_t>(queries_any) & 15)) {
        TRY(c->q_buf.ensure(q_bytes));
        const void* src = queries_any;
        if (stage_q) {
            memcpy(c->io_pinned, queries_any, q_bytes);
            src = c->io_pinned;
        }
        CU(cudaMemcpyAsync(c->q_buf.p, src, q_bytes, q_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
        q_ptr = static_cast<const uint8_t*>(c->q_buf.p);
    }
    const uint32_t* filt = nullptr;
    if (filter_mask_any) TRY(stage_filter(c, filter_mask_any, &filt));

    uint32_t* o_s = out_score_any;
    uint64_t* o_r = out_row_any;
    if (stage_o) {
        TRY(c->out_r.ensure(n_out * 12));            // [rows u64 | scores u32]: contiguous for the single copy back
        o_r = static_cast<uint64_t*>(c->out_r.p);
        o_s = reinterpret_cast<uint32_t*>(o_r + n_out);
    } else {
        if (!s_dev) {
            TRY(c->out_s.ensure(n_out * 4));
            o_s = static_cast<uint32_t*>(c->out_s.p);
        }
        if (!r_dev) {
            TRY(c->out_r.ensure(n_out * 8));
            o_r = static_cast<uint64_t*>(c->out_r.p);
        }
    }

    if (c->timing) CU(cudaEventRecord(c->ev0, st));
    ScanParams p{};
    p.rows = c->rows;
    p.live = c->live;
    p.filter = filt;
    p.queries = q_ptr;
    p.base_row = c->base_row;
    p.n_rows = (uint32_t)c->count;
    p.n_queries = n_queries;
    p.tile_stride = 1;
    p.one = 1u;
    p.trace = c->trace_dev;
    {
        static const uint32_t poll_ns = [] { const char* e = getenv("VLITE_B200_POLL_NS"); return e ? (uint32_t)atoi(e) : 0u; }();
        static const uint32_t poll_steady = [] { const char* e = getenv("VLITE_B200_POLL_STEADY"); return e ? (uint32_t)atoi(e) : 0u; }();
        p.poll_ns = poll_ns;
        p.poll_steady = poll_steady;
    }
    bool done = false;
    if (k > large_k_min(c, p, n_queries, metric) && !c->force_multipass) {
        int rc = search_large_k(c, p, n_queries, k, metric, o_s, o_r, &done);
        if (rc != VL_OK) return rc;
    }
    if (!done) TRY(search_lists(c, p, n_queries, k, metric, o_s, o_r));
    if (c->timing) {
        CU(cudaEventRecord(c->ev2, st));
        c->timed = true;
    }
    if (stage_o) {
        uint8_t* host = static_cast<uint8_t*>(c->io_pinned) + q_slot;
        CU(cudaMemcpyAsync(host, o_r, n_out * 12, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        memcpy(out_row_any, host, n_out * 8);
        memcpy(out_score_any, host + n_out * 8, n_out * 4);
        return VL_OK;
    }
    if (!s_dev) CU(cudaMemcpyAsync(out_score_any, o_s, n_out * 4, cudaMemcpyDeviceToHost, st));
    if (!r_dev) CU(cudaMemcpyAsync(out_row_any, o_r, n_out * 8, cudaMemcpyDeviceToHost, st));
    if (!s_dev || !r_dev || !q_dev) CU(cudaStreamSynchronize(st));
    return VL_OK;
}

// float embeddings in, top-k out: quantise on the device, then the ordinary search -- one call, one
// host -> device copy, one copy back (the whole of VLite.retrieve's device work for a query batch)
extern "C" int vl_search_embeddings(vl_corpus* c, const float* x_any, int n_queries, int dims, int k, int metric,
                                    const uint32_t* filter_mask_any, uint32_t* out_score_any, uint64_t* out_row_any) {
    if (!c) return fail(VL_EINVAL, "corpus is NULL");
    if (!x_any) return fail(VL_EINVAL, "NULL argument");
    if (n_queries < 1) return fail(VL_EINVAL, "n_queries must be >= 1");
    if (dims < c->w * 8) return fail(VL_EINVAL, "embeddings have %d features, the corpus rows %d bits", dims, c->w * 8);
    DeviceGuard g(c->device);
    cudaStream_t st = c->stream;
    const size_t x_bytes = (size_t)n_queries * dims * 4;
    const int x_kind = ptr_kind(x_any);
    const float* x_dev = x_any;
    if (x_kind != 2) {
        TRY(c->stage[0].ensure(x_bytes));
        const void* src = x_any;
        if (x_kind == 0 && x_bytes <= (4u << 20)) {
            // (the pinned block is shared with vl_search's result staging: stream order makes that safe --
            // the copy back is enqueued behind the kernels, which are behind this copy in)
            TRY(ensure_io_pinned(c, x_bytes));
            memcpy(c->io_pinned, x_any, x_bytes);
            src = c->io_pinned;
        }
        CU(cudaMemcpyAsync(c->stage[0].p, src, x_bytes, cudaMemcpyHostToDevice, st));
        x_dev = static_cast<const float*>(c->stage[0].p);
    }
    TRY(c->qx_buf.ensure((size_t)n_queries * c->w));
    const int words = c->w / 4;
    sign_pack_rows_kernel<<<grid_for((uint64_t)n_queries * words, 128, c->sms), 128, 0, st>>>(
        x_dev, (uint64_t)n_queries, dims, words, static_cast<uint32_t*>(c->qx_buf.p));
    LAUNCHED();
    return vl_search(c, static_cast<const uint8_t*>(c->qx_buf.p), n_queries, k, metric, filter_mask_any, out_score_any,
                     out_row_any);
}

extern "C" int vl_scores(vl_corpus* c, const uint8_t* query_any, int metric, uint64_t first, uint64_t n,
                         uint32_t* out_any) {
    if (!c || !query_any || !out_any) return fail(VL_EINVAL, "NULL argument");
    if (first + n > c->count) return fail(VL_EINVAL, "rows out of range");
    if (metric != VL_METRIC_HAMMING && metric != VL_METRIC_XORSUM) return fail(VL_EINVAL, "bad metric %d", metric);
    DeviceGuard g(c->device);
    if (n == 0) return VL_OK;
    cudaStream_t st = c->stream;
    TRY(c->q_buf.ensure(c->w));
    CU(cudaMemcpyAsync(c->q_buf.p, query_any, c->w,
                       is_device_ptr(query_any) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    const bool o_dev = is_device_ptr(out_any);
    uint32_t* o = out_any;
    if (!o_dev) {
        TRY(c->out_s.ensure(n * 4));
        o = static_cast<uint32_t*>(c->out_s.p);
    }
    const unsigned gr = grid_for(n, 256, c->sms);
    if (metric == VL_METRIC_HAMMING)
        scores_kernel<kHamming><<<gr, 256, c->w, st>>>(c->rows, c->w, static_cast<const uint8_t*>(c->q_buf.p), first, n, o);
    else
        scores_kernel<kXorSum><<<gr, 256, c->w, st>>>(c->rows, c->w, static_cast<const uint8_t*>(c->q_buf.p), first, n, o);
    LAUNCHED();
    if (!o_dev) {
        CU(cudaMemcpyAsync(out_any, o, n * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return VL_OK;
}

extern "C" int vl_merge_topk(int device, void* cuda_stream, const uint32_t* scores_dev, const uint64_t* rows_dev,
                             int n_lists, int n_queries, int k_in, int k_out, uint64_t score_list_stride,
                             uint64_t row_list_stride, uint32_t* out_score_dev, uint64_t* out_row_dev) {
    if (!scores_dev || !rows_dev || !out_score_dev || !out_row_dev) return fail(VL_EINVAL, "NULL argument");
    if (n_lists < 1 || n_queries < 1 || k_in < 1 || k_out < 1 || k_in > 65535)
        return fail(VL_EINVAL, "bad merge shape");
    TRY(check_device(device, nullptr));
    DeviceGuard g(device);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const size_t ss = score_list_stride ? (size_t)score_list_stride : (size_t)n_queries * k_in;
    const size_t rs = row_list_stride ? (size_t)row_list_stride : (size_t)n_queries * k_in;
    const size_t tmp_lists = merge_tmp_lists((size_t)n_lists);
    uint32_t* tmp_s = nullptr;
    uint64_t* tmp_r = nullptr;
    if (tmp_lists) {  // only beyond 128 lists; stream-ordered scratch
        const size_t n = tmp_lists * (size_t)n_queries * k_out;
        CU(cudaMallocAsync(reinterpret_cast<void**>(&tmp_r), n * 12, st));
        tmp_s = reinterpret_cast<uint32_t*>(tmp_r + n);
    }
    int rc = merge_levels(st, scores_dev, rows_dev, (size_t)n_lists, ss, rs, n_queries, k_in, k_out, out_score_dev,
                          out_row_dev, k_out, 0, tmp_s, tmp_r);
    if (tmp_r) cudaFreeAsync(tmp_r, st);
    return rc;
}

extern "C" int vl_map_rows(int device, void* cuda_stream, uint64_t* rows_dev, uint64_t n, const uint64_t* seg_local_dev,
                           const uint64_t* seg_global_dev, int n_segments) {
    if (!rows_dev || !seg_local_dev || !seg_global_dev) return fail(VL_EINVAL, "NULL argument");
    if (n_segments < 1) return fail(VL_EINVAL, "need at least one segment");
    TRY(check_device(device, nullptr));
    DeviceGuard g(device);
    if (n == 0) return VL_OK;
    map_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        rows_dev, n, seg_local_dev, seg_global_dev, n_segments);
    LAUNCHED();
    return VL_OK;
}

// =============================== rescore ===============================
extern "C" int vl_rescore(vl_corpus* c, const int8_t* docs_i8_dev, uint64_t n_docs, int dims, const void* queries_any,
                          int qtype, const uint64_t* cand_rows_any, int n_queries, int n_cand, int k,
                          float* out_score_any, uint64_t* out_row_any) {
    if (!c || !docs_i8_dev || !queries_any || !cand_rows_any || !out_score_any || !out_row_any)
        return fail(VL_EINVAL, "NULL argument");
    if (!is_device_ptr(docs_i8_dev)) return fail(VL_EINVAL, "docs must be device memory");
    if (dims < 16 || dims % 16) return fail(VL_EINVAL, "dims must be a positive multiple of 16");
    if (qtype != VL_QTYPE_F32 && qtype != VL_QTYPE_I8 && qtype != VL_QTYPE_GATHER_PROBE) return fail(VL_EINVAL, "bad qtype");
    if (n_queries < 1 || n_cand < 1 || n_cand > kRescoreMaxCand || k < 1 || k > n_cand)
        return fail(VL_EINVAL, "need 1 <= k <= n_cand <= %d", kRescoreMaxCand);
    DeviceGuard g(c->device);
    cudaStream_t st = c->stream;
    const size_t qbytes = (size_t)n_queries * dims * (qtype == VL_QTYPE_F32 ? 4 : 1);
    const void* q_ptr = queries_any;
    const bool q_dev = is_device_ptr(queries_any);
    if (!q_dev || (reinterpret_cast<uintptr_t>(queries_any) & 15)) {
        TRY(c->q_buf.ensure(qbytes));
        CU(cudaMemcpyAsync(c->q_buf.p, queries_any, qbytes, q_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
        q_ptr = c->q_buf.p;
    }
    const uint64_t* cand = cand_rows_any;
    const bool c_dev = is_device_ptr(cand_rows_any);
    if (!c_dev) {
        TRY(c->misc.ensure((size_t)n_queries * n_cand * 8));
        CU(cudaMemcpyAsync(c->misc.p, cand_rows_any, (size_t)n_queries * n_cand * 8, cudaMemcpyHostToDevice, st));
        cand = static_cast<const uint64_t*>(c->misc.p);
    }
    const bool s_dev = is_device_ptr(out_score_any), r_dev = is_device_ptr(out_row_any);
    const size_t n_out = (size_t)n_queries * k;
    float* o_s = out_score_any;
    uint64_t* o_r = out_row_any;
    if (!s_dev) {
        TRY(c->out_s.ensure(n_out * 4));
        o_s = static_cast<float*>(c->out_s.p);
    }
    if (!r_dev) {
        TRY(c->out_r.ensure(n_out * 8));
        o_r = static_cast<uint64_t*>(c->out_r.p);
    }
    int n_sort = 4;
    while (n_sort < n_cand) n_sort <<= 1;
    RescoreParams p{};
    p.docs = docs_i8_dev;
    p.queries = q_ptr;
    p.cand = cand;
    p.out_scores = o_s;
    p.out_rows = o_r;
    p.base_row = c->base_row;
    p.n_docs = n_docs;
    p.dims = dims;
    p.n_cand = n_cand;
    p.n_sort = n_sort;
    p.k = k;
    const size_t smem = (size_t)n_sort * 12 + (size_t)dims * (qtype == VL_QTYPE_F32 ? 8 : 1);   // fp32 queries are held as fp64
    c->timed = false;
    if (c->timing) CU(cudaEventRecord(c->ev0, st));
    if (qtype == VL_QTYPE_I8) {
        TRY(allow_max_smem(rescore_kernel<1>, c->device));
        rescore_kernel<1><<<n_queries, kRescoreThreads, smem, st>>>(p);
    } else if (qtype == VL_QTYPE_GATHER_PROBE) {
        TRY(allow_max_smem(rescore_kernel<2>, c->device));
        rescore_kernel<2><<<n_queries, kRescoreThreads, smem, st>>>(p);
    } else {
        TRY(allow_max_smem(rescore_kernel<0>, c->device));
        rescore_kernel<0><<<n_queries, kRescoreThreads, smem, st>>>(p);
    }
    LAUNCHED();
    if (c->timing) {
        CU(cudaEventRecord(c->ev1, st));
        CU(cudaEventRecord(c->ev2, st));
        c->timed = true;
    }
    if (!s_dev) CU(cudaMemcpyAsync(out_score_any, o_s, n_out * 4, cudaMemcpyDeviceToHost, st));
    if (!r_dev) CU(cudaMemcpyAsync(out_row_any, o_r, n_out * 8, cudaMemcpyDeviceToHost, st));
    if (!s_dev || !r_dev || !q_dev || !c_dev) CU(cudaStreamSynchronize(st));
    return VL_OK;
}

extern "C" int vl_quantize_binary(int device, void* cuda_stream, const float* x_any, uint64_t n, int dims,
                                  uint8_t* out_any) {
    if (!x_any || !out_any) return fail(VL_EINVAL, "NULL argument");
    if (dims < 32 || dims % 32) return fail(VL_EINVAL, "dims must be a positive multiple of 32");
    int sms = 0;
    TRY(check_device(device, &sms));
    DeviceGuard g(device);
    if (n == 0) return VL_OK;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const bool x_dev = is_device_ptr(x_any), o_dev = is_device_ptr(out_any);
    const size_t in_bytes = n * (size_t)dims * 4, out_bytes = n * (size_t)dims / 8;
    // grow-only per-device scratch shared by the handle-less entry points (one call at a time per device)
    static std::mutex mu[64];
    static DevBuf scratch_in[64], scratch_out[64];
    std::lock_guard<std::mutex> lock(mu[device & 63]);
    const float* src = x_any;
    uint8_t* dst = out_any;
    if (!x_dev) {
        TRY(scratch_in[device & 63].ensure(in_bytes));
        CU(cudaMemcpyAsync(scratch_in[device & 63].p, x_any, in_bytes, cudaMemcpyHostToDevice, st));
        src = static_cast<const float*>(scratch_in[device & 63].p);
    }
    if (!o_dev) {
        TRY(scratch_out[device & 63].ensure(out_bytes));
        dst = static_cast<uint8_t*>(scratch_out[device & 63].p);
    }
    const uint64_t n_words = out_bytes / 4;
    repack_f32_bits_kernel<true><<<grid_for(n_words, 256, sms), 256, 0, st>>>(src, n_words,
                                                                            reinterpret_cast<uint32_t*>(dst), nullptr);
    LAUNCHED();
    if (!o_dev) CU(cudaMemcpyAsync(out_any, dst, out_bytes, cudaMemcpyDeviceToHost, st));
    if (!x_dev || !o_dev) CU(cudaStreamSynchronize(st));
    return VL_OK;
}

extern "C" int vl_quantize_embeddings(int device, void* cuda_stream, const float* x_dev, uint64_t n, int dims,
                                      int binary_dims, float i8_scale, uint8_t* out_u8_dev, int8_t* out_i8_dev,
                                      float* out_f32_dev) {
    if (!x_dev || !out_u8_dev) return fail(VL_EINVAL, "NULL argument");
    if (dims < 32 || dims % 32 || dims > 8192) return fail(VL_EINVAL, "dims must be a multiple of 32, at most 8192");
    if (binary_dims < 32 || binary_dims % 32 || binary_dims > dims)
        return fail(VL_EINVAL, "binary_dims must be a multiple of 32, at most dims");
    if (out_i8_dev && !(i8_scale > 0.f)) return fail(VL_EINVAL, "i8_scale must be positive");
    int sms = 0;
    TRY(check_device(device, &sms));
    DeviceGuard g(device);
    if (!is_device_ptr(x_dev) || !is_device_ptr(out_u8_dev) || (out_i8_dev && !is_device_ptr(out_i8_dev)) ||
        (out_f32_dev && !is_device_ptr(out_f32_dev)))
        return fail(VL_EINVAL, "vl_quantize_embeddings takes device pointers only");
    if (n == 0) return VL_OK;
    quantize_embeddings_kernel<<<grid_for(n * 32, 256, sms), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        x_dev, n, dims, binary_dims, i8_scale, reinterpret_cast<uint32_t*>(out_u8_dev), out_i8_dev, out_f32_dev);
    LAUNCHED();
    return VL_OK;
}

extern "C" int vl_synth_docs_i8(int device, void* cuda_stream, uint64_t seed, uint64_t first_row, uint64_t n, int dims,
                                int8_t* out_dev) {
    if (!out_dev || !is_device_ptr(out_dev)) return fail(VL_EINVAL, "out must be device memory");
    if (dims < 8 || dims % 8 || dims > 1024) return fail(VL_EINVAL, "dims must be a multiple of 8, at most 1024");
    int sms = 0;
    TRY(check_device(device, &sms));
    DeviceGuard g(device);
    if (n == 0) return VL_OK;
    synth_docs_kernel<<<grid_for(n * (uint64_t)(dims / 8), 256, sms), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        out_dev, dims, seed, first_row, n);
    LAUNCHED();
    return VL_OK;
}

// =============================== measurement ===============================
extern "C" int vl_microbench(int device, int kind, int iters, double* ops_per_sec_out, double* ms_out) {
    if (kind < 0 || kind > 4 || iters < 1) return fail(VL_EINVAL, "bad microbench arguments");
    int sms = 0;
    TRY(check_device(device, &sms));
    DeviceGuard g(device);
    uint32_t* out = nullptr;
    CU(cudaMalloc(&out, 4));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    const int threads = 256, blocks = sms * 8;
    auto run = [&](int it) {
        switch (kind) {
            case 0: microbench_kernel<0><<<blocks, threads>>>(it, 12345u, out); break;
            case 1: microbench_kernel<1><<<blocks, threads>>>(it, 12345u, out); break;
            case 2: microbench_kernel<2><<<blocks, threads>>>(it, 12345u, out); break;
            case 3: microbench_kernel<3><<<blocks, threads>>>(it, 12345u, out); break;
            default: microbench_kernel<4><<<blocks, threads>>>(it, 12345u, out); break;
        }
        g_launches.fetch_add(1);
    };
    run(std::max(1, iters / 10));  // warm-up
    CU(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CU(cudaEventRecord(e0));
        run(iters);
        CU(cudaEventRecord(e1));
        CU(cudaEventSynchronize(e1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
    }
    CU(cudaGetLastError());
    const double ops = (double)blocks * threads * (double)iters * kMbChains;  // primary op count (POPC for kind 4)
    if (ops_per_sec_out) *ops_per_sec_out = ops / (best * 1e-3);
    if (ms_out) *ms_out = best;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return VL_OK;
}

extern "C" int vl_set_timing(vl_corpus* c, int enabled) {
    if (!c) return fail(VL_EINVAL, "corpus is NULL");
    c->timing = enabled != 0;
    c->timed = false;
    return VL_OK;
}

extern "C" int vl_last_timing(vl_corpus* c, float* total_ms, float* scan_ms) {
    if (!c) return fail(VL_EINVAL, "corpus is NULL");
    if (!c->timed) return fail(VL_EINVAL, "no timed call recorded (vl_set_timing first)");
    DeviceGuard g(c->device);
    CU(cudaEventSynchronize(c->ev2));
    float t = 0, s = 0;
    CU(cudaEventElapsedTime(&t, c->ev0, c->ev2));
    CU(cudaEventElapsedTime(&s, c->ev0, c->ev1));
    if (total_ms) *total_ms = t;
    if (scan_ms) *scan_ms = s;
    return VL_OK;
}

extern "C" int vl_set_tuning(vl_corpus* c, int row_splits, int stages) {
    if (!c) return fail(VL_EINVAL, "corpus is NULL");
    c->force_multipass = stages == -1 ? 1 : 0;  // stages = -1: force the exact multi-pass path for large k
    c->fuse_merge = stages == -2 ? 0 : 1;       // stages = -2: separate merge kernels for 1-2 queries (A/B runs)
    if (stages < 0) stages = 0;
    c->tune_splits = row_splits;
    c->tune_stages = stages;
    return VL_OK;
}

extern "C" int vl_debug_checks(int device, uint32_t* failures_out, uint32_t* first_id_out) {
#ifdef VL_DEBUG_BOUNDS
    TRY(check_device(device, nullptr));
    DeviceGuard g(device);
    unsigned int h[2] = {0, 0};
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpyFromSymbol(h, vl::g_vl_dcheck, sizeof h));
    if (failures_out) *failures_out = h[0];
    if (first_id_out) *first_id_out = h[1];
    return VL_OK;
#else
    (void)device;
    if (failures_out) *failures_out = 0;
    if (first_id_out) *first_id_out = 0;
    return fail(VL_EINVAL, "not a debug build: rebuild with -DVL_DEBUG_BOUNDS (python -m vlite_b200.build --debug)");
#endif
}

// tuning aid, not part of the documented ABI surface: %globaltimer stamps (ns) of CTA (0,0) of the last
// tensor-core scan: [0] kernel entry, [1] prologue done, [2] last epilogue done, [3] exit, [4..15] accumulator i ready
extern "C" int vl_scan_trace(vl_corpus* c, uint64_t* out16_host) {
    if (!c) return fail(VL_EINVAL, "corpus is NULL");
    DeviceGuard g(c->device);
    if (!c->trace_dev) {
        CU(cudaMalloc(&c->trace_dev, vl::kTraceWords * 8));
        CU(cudaMemset(c->trace_dev, 0, vl::kTraceWords * 8));
    }
    if (out16_host) {
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaMemcpy(out16_host, c->trace_dev, 16 * 8, cudaMemcpyDeviceToHost));
    }
    return VL_OK;
}
// ... and the per-tile part of the same trace: [16 + i] = epilogue of tile i done (ns), [16 + 256 + i] = insert-loop
// iterations warp 4 ran in tile i, for the first 256 tiles of CTA (0,0)
extern "C" int vl_scan_trace_tiles(vl_corpus* c, uint64_t* out512_host) {
    if (!c || !out512_host) return fail(VL_EINVAL, "NULL argument");
    if (!c->trace_dev) return fail(VL_EINVAL, "tracing is off: call vl_scan_trace(h, NULL) first");
    DeviceGuard g(c->device);
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpy(out512_host, c->trace_dev + 16, 512 * 8, cudaMemcpyDeviceToHost));
    return VL_OK;
}
// ... and the role timeline of one steady-state tile (mmascan.cuh: kTraceTile), 64 words
extern "C" int vl_scan_trace_roles(vl_corpus* c, uint64_t* out64_host) {
    if (!c || !out64_host) return fail(VL_EINVAL, "NULL argument");
    if (!c->trace_dev) return fail(VL_EINVAL, "tracing is off: call vl_scan_trace(h, NULL) first");
    DeviceGuard g(c->device);
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpy(out64_host, c->trace_dev + vl::kTraceRoleBase, 64 * 8, cudaMemcpyDeviceToHost));
    return VL_OK;
}
// ... and entry [i] / exit [296 + i] stamps of every CTA (linear index, first 296)
extern "C" int vl_scan_trace_ctas(vl_corpus* c, uint64_t* out592_host) {
    if (!c || !out592_host) return fail(VL_EINVAL, "NULL argument");
    if (!c->trace_dev) return fail(VL_EINVAL, "tracing is off: call vl_scan_trace(h, NULL) first");
    DeviceGuard g(c->device);
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpy(out592_host, c->trace_dev + 16 + 512, 2 * vl::kTraceCtas * 8, cudaMemcpyDeviceToHost));
    return VL_OK;
}

extern "C" int vl_set_scan_path(vl_corpus* c, int path) {
    if (!c) return fail(VL_EINVAL, "corpus is NULL");
    if (path < 0 || path > 6 || path == 4)
        return fail(VL_EINVAL, "scan path must be 0 (auto), 1 (integer pipes), 2 / 3 (tensor cores, 1 / 2 CTAs per tile), 5 / 6 (the same, HAMMING as int8 instead of e2m1)");
    c->scan_path = path;
    return VL_OK;
}

#include "group.cuh"

// =============================== exchange (one process per GPU) ===============================
struct vl_exchange {
    int device = 0, world = 0, rank = 0;
    size_t slot_bytes = 0, flags_off = 0, total = 0;
    uint8_t* local = nullptr;
    uint8_t* peer[vl::kExchangeMaxWorld] = {};
    bool opened[vl::kExchangeMaxWorld] = {};
    uint32_t epoch = 0;
    int* timeout_host = nullptr;   // pinned, mapped: the merge kernel reports a peer that never arrived
    int* timeout_dev = nullptr;
};

extern "C" int vl_exchange_create(int device, int world, int rank, uint64_t slot_bytes, vl_exchange** out) {
    if (!out) return fail(VL_EINVAL, "out is NULL");
    *out = nullptr;
    if (world < 2 || world > vl::kExchangeMaxWorld || rank < 0 || rank >= world)
        return fail(VL_EINVAL, "need 2 <= world <= %d and 0 <= rank < world", vl::kExchangeMaxWorld);
    if (slot_bytes == 0) return fail(VL_EINVAL, "slot_bytes is 0");
    TRY(check_device(device, nullptr));
    DeviceGuard g(device);
    vl_exchange* x = new vl_exchange();
    x->device = device;
    x->world = world;
    x->rank = rank;
    x->slot_bytes = (slot_bytes + 255) & ~(size_t)255;
    x->flags_off = 2 * (size_t)world * x->slot_bytes;
    x->total = x->flags_off + 2 * (size_t)world * 4 + 256;
    if (cudaMalloc(&x->local, x->total) != cudaSuccess) {
        cudaGetLastError();
        delete x;
        return fail(VL_ENOMEM, "cudaMalloc(%zu) failed", x->total);
    }
    cudaMemset(x->local, 0, x->total);
    cudaDeviceSynchronize();
    if (cudaHostAlloc(&x->timeout_host, sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(&x->timeout_dev, x->timeout_host, 0) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(x->local);
        delete x;
        return fail(VL_ECUDA, "cannot allocate the mapped status word");
    }
    *x->timeout_host = 0;
    x->peer[rank] = x->local;
    *out = x;
    return VL_OK;
}

extern "C" int vl_exchange_handle(vl_exchange* x, void* handle_out_64) {
    if (!x || !handle_out_64) return fail(VL_EINVAL, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DeviceGuard g(x->device);
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, x->local));
    memcpy(handle_out_64, &h, sizeof h);
    return VL_OK;
}

extern "C" int vl_exchange_attach(vl_exchange* x, int peer_rank, const void* handle_64) {
    if (!x || !handle_64) return fail(VL_EINVAL, "NULL argument");
    if (peer_rank < 0 || peer_rank >= x->world) return fail(VL_EINVAL, "peer rank out of range");
    if (peer_rank == x->rank) return VL_OK;
    DeviceGuard g(x->device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_64, sizeof h);
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    x->peer[peer_rank] = static_cast<uint8_t*>(p);
    x->opened[peer_rank] = true;
    return VL_OK;
}

extern "C" int vl_exchange_gather_merge(vl_exchange* x, void* cuda_stream, const void* local_block_dev,
                                        uint64_t block_bytes, int n_queries, int k, uint32_t* out_score_dev,
                                        uint64_t* out_row_dev) {
    if (!x || !local_block_dev || !out_score_dev || !out_row_dev) return fail(VL_EINVAL, "NULL argument");
    const size_t nb = (size_t)n_queries * k;
    if (n_queries < 1 || k < 1 || (nb * 4) % 8 || block_bytes != nb * 12 || block_bytes > x->slot_bytes)
        return fail(VL_EINVAL, "block must be [Q*k u32 | Q*k u64] with Q*k even, at most the slot size");
    if (reinterpret_cast<uintptr_t>(local_block_dev) & 15) return fail(VL_EINVAL, "block must be 16 B aligned");
    for (int r = 0; r < x->world; ++r)
        if (!x->peer[r]) return fail(VL_EINVAL, "rank %d is not attached", r);
    if (*x->timeout_host) return fail(VL_ECUDA, "an earlier exchange timed out waiting for a peer rank");
    DeviceGuard g(x->device);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const uint32_t epoch = ++x->epoch;
    const size_t parity = epoch & 1u;
    vl::PeerPtrs peers{};
    for (int r = 0; r < x->world; ++r) peers.p[r] = x->peer[r];
    const uint64_t slot_off = (parity * x->world + x->rank) * x->slot_bytes;
    const uint64_t flag_off = x->flags_off + (parity * x->world + x->rank) * 4;
    const uint64_t n16 = (block_bytes + 15) / 16;          // (the slot is padded to 256 B: a ragged tail is harmless)
    vl::exchange_push_kernel<<<x->world, vl::kPushThreads, 0, st>>>(static_cast<const uint4*>(local_block_dev), n16,
                                                                   peers, slot_off, flag_off, epoch);
    LAUNCHED();
    const uint8_t* slots = x->local + parity * x->world * x->slot_bytes;
    const uint32_t* flags = reinterpret_cast<const uint32_t*>(x->local + x->flags_off) + parity * x->world;
    vl::exchange_merge_kernel<<<(unsigned)((n_queries + vl::kMergeWarps - 1) / vl::kMergeWarps), vl::kMergeWarps * 32, 0, st>>>(
        flags, x->world, epoch, x->timeout_dev, reinterpret_cast<const uint32_t*>(slots),
        reinterpret_cast<const uint64_t*>(slots + nb * 4), n_queries, k, out_score_dev, out_row_dev, x->slot_bytes / 4,
        x->slot_bytes / 8);
    LAUNCHED();
    return VL_OK;
}

extern "C" int vl_exchange_destroy(vl_exchange* x) {
    if (!x) return VL_OK;
    DeviceGuard g(x->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < x->world; ++r)
        if (x->opened[r]) cudaIpcCloseMemHandle(x->peer[r]);
    if (x->local) cudaFree(x->local);
    if (x->timeout_host) cudaFreeHost(x->timeout_host);
    delete x;
    return VL_OK;
}
