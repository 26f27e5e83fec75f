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
#include "synth.cuh"

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(expr)                                                                                  \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            cudaGetLastError();                                                                   \
            return fail(_e == cudaErrorMemoryAllocation ? VL_ENOMEM : VL_ECUDA, "%s: %s (%s:%d)", \
                        #expr, cudaGetErrorString(_e), __FILE__, __LINE__);                       \
        }                                                                                         \
    } while (0)
#define TRY(expr)              \
    do {                       \
        int _r = (expr);       \
        if (_r != VL_OK) return _r; \
    } while (0)
#define LAUNCHED()                       \
    do {                                 \
        g_launches.fetch_add(1);         \
        CU(cudaGetLastError());          \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return VL_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = std::max(bytes, (size_t)4096);
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(VL_ENOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        }
        cap = want;
        return VL_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

bool is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// cudaGetDeviceProperties costs milliseconds per call: query each device once per process
struct DevInfo {
    std::atomic<int> state{0};  // 0 unknown, 1 ok, -1 unusable
    int major = 0, minor = 0, sms = 0;
};
DevInfo g_dev[64];
std::atomic<int> g_ndev{-1};

int check_device(int device, int* sms_out) {
    int n = g_ndev.load();
    if (n < 0) {
        if (cudaGetDeviceCount(&n) != cudaSuccess) {
            cudaGetLastError();
            n = 0;
        }
        if (n > 0) g_ndev.store(n);
    }
    if (n == 0) return fail(VL_ENODEV, "no CUDA device visible; vlite_b200 has no CPU fallback");
    if (device < 0 || device >= n || device >= 64)
        return fail(VL_EINVAL, "device %d out of range (0..%d)", device, n - 1);
    DevInfo& d = g_dev[device];
    if (d.state.load() == 0) {
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, device));
        d.major = prop.major;
        d.minor = prop.minor;
        d.sms = prop.multiProcessorCount;
        d.state.store(prop.major == 10 ? 1 : -1);
    }
    if (d.state.load() < 0)
        return fail(VL_ENODEV, "device %d is sm_%d%d; vlite_b200 is built for sm_100a (B200) only", device, d.major,
                    d.minor);
    if (sms_out) *sms_out = d.sms;
    return VL_OK;
}

inline unsigned grid_for(uint64_t items, int threads, int sms) {
    uint64_t blocks = (items + threads - 1) / threads;
    uint64_t cap = (uint64_t)sms * 16;
    return (unsigned)std::max<uint64_t>(1, std::min(blocks, cap));
}

}  // namespace

// ---- append-in-place storage: a virtual address range reserved once, physical HBM mapped behind
// it chunk by chunk (CUDA virtual memory management).  Growing a shard never copies a byte and never
// moves the base pointer, so a 100 GB corpus can keep appending without a 2x transient.  The driver
// entry points come from cudaGetDriverEntryPoint: the library does not link libcuda. ----
struct VmmApi {
    CUresult (*reserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*release_va)(CUdeviceptr, size_t) = nullptr;
    CUresult (*create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*set_access)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*granularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    bool ok = false;
};

int vmm_api(const VmmApi** out) {
    static VmmApi api;
    static std::atomic<int> state{0};
    if (state.load() == 0) {
        static std::mutex mu;
        std::lock_guard<std::mutex> lock(mu);
        if (state.load() == 0) {
            auto get = [](const char* name, void** fn) {
                cudaDriverEntryPointQueryResult q;
                return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess &&
                       q == cudaDriverEntryPointSuccess && *fn != nullptr;
            };
            api.ok = get("cuMemAddressReserve", reinterpret_cast<void**>(&api.reserve)) &&
                     get("cuMemAddressFree", reinterpret_cast<void**>(&api.release_va)) &&
                     get("cuMemCreate", reinterpret_cast<void**>(&api.create)) &&
                     get("cuMemRelease", reinterpret_cast<void**>(&api.release)) &&
                     get("cuMemMap", reinterpret_cast<void**>(&api.map)) &&
                     get("cuMemUnmap", reinterpret_cast<void**>(&api.unmap)) &&
                     get("cuMemSetAccess", reinterpret_cast<void**>(&api.set_access)) &&
                     get("cuMemGetAllocationGranularity", reinterpret_cast<void**>(&api.granularity));
            cudaGetLastError();
            state.store(1);
        }
    }
    if (!api.ok) return fail(VL_ECUDA, "the CUDA driver does not export the virtual memory management API");
    *out = &api;
    return VL_OK;
}

struct VmmArena {
    CUdeviceptr base = 0;
    size_t reserved = 0, mapped = 0, gran = 0;
    int device = 0;
    std::vector<std::pair<CUmemGenericAllocationHandle, size_t>> chunks;

    CUmemAllocationProp prop() const {
        CUmemAllocationProp p{};
        p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        p.location.id = device;
        return p;
    }
    int init(int dev, size_t reserve_bytes) {
        const VmmApi* a;
        TRY(vmm_api(&a));
        device = dev;
        CUmemAllocationProp p = prop();
        if (a->granularity(&gran, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0)
            return fail(VL_ECUDA, "cuMemGetAllocationGranularity failed");
        reserved = (reserve_bytes + gran - 1) / gran * gran;
        CUresult r = a->reserve(&base, reserved, gran, 0, 0);
        if (r != CUDA_SUCCESS) return fail(VL_ENOMEM, "cuMemAddressReserve(%zu) failed (CUresult %d)", reserved, (int)r);
        return VL_OK;
    }
    // map physical memory until at least `need` bytes are backed
    int grow(size_t need) {
        if (need <= mapped) return VL_OK;
        if (need > reserved) return fail(VL_ENOMEM, "corpus exceeds the reserved address range (%zu > %zu)", need, reserved);
        const VmmApi* a;
        TRY(vmm_api(&a));
        const size_t add = (need - mapped + gran - 1) / gran * gran;
        CUmemAllocationProp p = prop();
        CUmemGenericAllocationHandle h;
        CUresult r = a->create(&h, add, &p, 0);
        if (r != CUDA_SUCCESS) return fail(VL_ENOMEM, "cuMemCreate(%zu) failed (CUresult %d): out of device memory", add, (int)r);
        r = a->map(base + mapped, add, 0, h, 0);
        if (r != CUDA_SUCCESS) {
            a->release(h);
            return fail(VL_ECUDA, "cuMemMap failed (CUresult %d)", (int)r);
        }
        CUmemAccessDesc acc{};
        acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        acc.location.id = device;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        r = a->set_access(base + mapped, add, &acc, 1);
        if (r != CUDA_SUCCESS) {
            a->unmap(base + mapped, add);
            a->release(h);
            return fail(VL_ECUDA, "cuMemSetAccess failed (CUresult %d)", (int)r);
        }
        chunks.emplace_back(h, add);
        mapped += add;
        return VL_OK;
    }
    void destroy() {
        const VmmApi* a;
        if (vmm_api(&a) != VL_OK) return;
        size_t off = 0;
        for (auto& ch : chunks) {
            a->unmap(base + off, ch.second);
            a->release(ch.first);
            off += ch.second;
        }
        chunks.clear();
        if (base) a->release_va(base, reserved);
        base = 0;
        reserved = mapped = 0;
    }
};

struct vl_corpus {
    int device = 0;
    int w = 0;
    int sms = 148;
    cudaStream_t stream = nullptr;
    VmmArena rows_arena, live_arena;
    uint8_t* rows = nullptr;   // = rows_arena.base: never moves
    uint32_t* live = nullptr;  // = live_arena.base
    size_t live_zeroed = 0;    // bytes of the live mask initialised so far
    uint64_t capacity = 0;  // rows backed by physical memory, multiple of kMaxStageRows
    uint64_t count = 0;
    uint64_t live_count = 0;
    uint64_t base_row = 0;
    DevBuf q_buf, filt_buf, part_s, part_r, out_s, out_r, stage[2], cand_buf, misc, lk_keys, lk_misc;
    int force_multipass = 0;  // tuning/testing: disable the sampled large-k path
    void* pinned[2] = {nullptr, nullptr};
    size_t pinned_bytes = 0;
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    int* d_err = nullptr;
    unsigned long long* d_counter = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    bool timing = false;
    bool timed = false;
    int tune_splits = 0;
    int tune_stages = 0;
};

namespace {

using namespace vl;

// back at least `cap_rows` rows (rounded up to whole tiles) with physical memory -- no copy, no move
int grow_to(vl_corpus* c, uint64_t cap_rows) {
    // local rows are u32 in the kernels and a signed 32-bit TMA coordinate
    if (cap_rows > 0x7FFFFC00ull) return fail(VL_EINVAL, "a corpus shard holds at most 2^31-1024 rows");
    const uint64_t cap = (cap_rows + kMaxStageRows - 1) / kMaxStageRows * kMaxStageRows;
    if (cap <= c->capacity) return VL_OK;
    TRY(c->rows_arena.grow(cap * (uint64_t)c->w));
    TRY(c->live_arena.grow(cap / 8));
    if (c->live_arena.mapped > c->live_zeroed) {  // fresh mask bytes start dead
        CU(cudaMemsetAsync(reinterpret_cast<uint8_t*>(c->live) + c->live_zeroed, 0,
                           c->live_arena.mapped - c->live_zeroed, c->stream));
        c->live_zeroed = c->live_arena.mapped;
    }
    // everything that is mapped is usable capacity (whole tiles only)
    const uint64_t by_rows = c->rows_arena.mapped / (uint64_t)c->w;
    const uint64_t by_live = (uint64_t)c->live_arena.mapped * 8;
    c->capacity = std::min<uint64_t>(std::min(by_rows, by_live), 0x7FFFFC00ull) / kMaxStageRows * kMaxStageRows;
    return VL_OK;
}

// append-in-place growth: map at least 1/2 of the current size (64 MB minimum) so that a stream of
// small appends costs O(log) driver calls; nothing is copied either way
int ensure_capacity(vl_corpus* c, uint64_t need_rows) {
    if (need_rows <= c->capacity) return VL_OK;
    const uint64_t step_rows = std::max<uint64_t>(c->capacity / 2, (64ull << 20) / c->w);
    int rc = grow_to(c, std::max(need_rows, c->capacity + step_rows));
    if (rc == VL_ENOMEM) rc = grow_to(c, need_rows);
    return rc;
}

int mark_appended(vl_corpus* c, uint64_t first, uint64_t n) {
    if (n == 0) return VL_OK;
    live_set_range_kernel<<<grid_for((n + 31) / 32 + 1, 256, c->sms), 256, 0, c->stream>>>(c->live, first, n);
    LAUNCHED();
    c->count = std::max(c->count, first + n);
    c->live_count += n;
    return VL_OK;
}

int ensure_pinned(vl_corpus* c, size_t bytes) {
    if (bytes <= c->pinned_bytes) return VL_OK;
    for (int i = 0; i < 2; ++i) {
        if (c->pinned[i]) cudaFreeHost(c->pinned[i]);
        c->pinned[i] = nullptr;
    }
    c->pinned_bytes = 0;
    for (int i = 0; i < 2; ++i) CU(cudaMallocHost(&c->pinned[i], bytes));
    c->pinned_bytes = bytes;
    return VL_OK;
}

// repack `n` rows of f32 already on the device (src_dev) into corpus rows starting at `first`
int repack_rows(vl_corpus* c, const float* src_dev, uint64_t first, uint64_t n, int kind) {
    const uint64_t n_words = n * (uint64_t)c->w / 4;
    uint32_t* dst = reinterpret_cast<uint32_t*>(c->rows + first * (uint64_t)c->w);
    const unsigned g = grid_for(n_words, 256, c->sms);
    if (kind == VL_SRC_F32_OFFSET)
        repack_f32_offset_kernel<<<g, 256, 0, c->stream>>>(src_dev, n_words, dst, c->d_err);
    else
        repack_f32_bits_kernel<false><<<g, 256, 0, c->stream>>>(src_dev, n_words, dst, c->d_err);
    LAUNCHED();
    return VL_OK;
}

size_t src_row_bytes(int w, int kind) {
    return kind == VL_SRC_U8 ? (size_t)w : kind == VL_SRC_F32_OFFSET ? (size_t)w * 4 : (size_t)w * 32;
}

int check_err_flag(vl_corpus* c, const char* what) {
    int h = 0;
    CU(cudaMemcpyAsync(&h, c->d_err, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (h) {
        CU(cudaMemsetAsync(c->d_err, 0, sizeof(int), c->stream));
        return fail(VL_ERANGE, "%s: stored value is not an integer in the encodable range", what);
    }
    return VL_OK;
}

// ---- TMA tensor map of the corpus: 2-D uint8 tensor {W, rows}, box {W, 256 rows}, hardware swizzle
// matching the row width (128 B or 64 B) so that one-row-per-lane 128-bit LDS is conflict-free ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int make_corpus_tmap(const vl_corpus* c, CUtensorMap* out) {
    static std::atomic<EncodeTiledFn> cached{nullptr};
    EncodeTiledFn fn = cached.load();
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !sym)
            return fail(VL_ECUDA, "the CUDA driver does not export cuTensorMapEncodeTiled");
        fn = reinterpret_cast<EncodeTiledFn>(sym);
        cached.store(fn);
    }
    const cuuint64_t dims[2] = {(cuuint64_t)c->w, (cuuint64_t)c->count};
    const cuuint64_t strides[1] = {(cuuint64_t)c->w};
    const cuuint32_t box[2] = {(cuuint32_t)c->w, (cuuint32_t)kTileRows};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = c->w == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : c->w == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                               : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, c->rows, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(VL_ECUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return VL_OK;
}

// Opt a kernel in to the full 227 KB of dynamic shared memory ONCE per device.  The attribute
// belongs to the (function, context), not to a handle or a thread: setting it to each launch's
// own size would let two handles on different threads shrink it under each other.
int allow_max_smem_fn(const void* kern, int device) {
    static std::mutex mu;
    static std::vector<std::pair<const void*, int>> done;   // (kernel, device) pairs already opted in
    std::lock_guard<std::mutex> lock(mu);
    for (const auto& d : done)
        if (d.first == kern && d.second == device) return VL_OK;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    done.emplace_back(kern, device);
    return VL_OK;
}
template <typename K>
int allow_max_smem(K kern, int device) {
    return allow_max_smem_fn(reinterpret_cast<const void*>(kern), device);
}

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

int launch_scan(vl_corpus* c, ScanParams& p, int nq, int k, int metric, ScanPlan* plan, bool dry) {
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
                 int out_stride, int out_offset, uint32_t* tmp_s, uint64_t* tmp_r) {
    const unsigned gx = (unsigned)((n_queries + kMergeWarps - 1) / kMergeWarps);
    const size_t dense = (size_t)n_queries * k_out;
    while (true) {
        const size_t groups = (n_lists + kMergeGroup - 1) / kMergeGroup;
        if (groups <= 1) {
            merge_topk_kernel<<<dim3(gx, 1), kMergeWarps * 32, 0, st>>>(in_s, in_r, (int)n_lists, n_queries, k_in, k_out,
                                                                      out_s, out_r, out_stride, out_offset, s_stride,
                                                                      r_stride, 0);
            LAUNCHED();
            return VL_OK;
        }
        merge_topk_kernel<<<dim3(gx, (unsigned)groups), kMergeWarps * 32, 0, st>>>(
            in_s, in_r, (int)n_lists, n_queries, k_in, k_out, tmp_s, tmp_r, k_out, 0, s_stride, r_stride, dense);
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
int search_lists(vl_corpus* c, ScanParams p, int n_queries, int k, int metric, uint32_t* o_s, uint64_t* o_r,
                 int out_stride = 0) {
    cudaStream_t st = c->stream;
    if (out_stride == 0) out_stride = k;
    const int n_pass = (k + kWarpListMaxK - 1) / kWarpListMaxK;
    // tau[Q] | cursor score[Q] | cursor row[Q] | histogram[Q][1056] (k <= 32, not for huge batches)
    // (the histogram is read only by the one-warp-per-query configuration, Q > 32; the shared bound
    // tau only by query tiles of more than two queries: small batches skip both memsets)
    const bool want_hist = k <= 32 && n_queries > 32 && n_queries <= 16384;
    const bool want_tau = n_queries > 2;
    const size_t hist_bytes = want_hist ? (size_t)n_queries * kHistWords * 4 : 0;
    TRY(c->cand_buf.ensure((size_t)n_queries * 12 + hist_bytes));
    p.tau = static_cast<uint32_t*>(c->cand_buf.p);
    uint32_t* cur_s = p.tau + n_queries;
    uint32_t* cur_r = cur_s + n_queries;
    p.hist = want_hist ? cur_r + n_queries : nullptr;
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
        const size_t n_lists = (size_t)plan.S * plan.WR;
        const size_t dense = (size_t)n_queries * k_pass;
        const size_t n_part = (n_lists + merge_tmp_lists(n_lists)) * dense;
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
        if (c->timing && pass == 0 && p.tile_stride == 1) CU(cudaEventRecord(c->ev1, st));
        TRY(merge_levels(st, p.part_scores, p.part_rows, n_lists, dense, dense, n_queries, k_pass, k_pass, o_s, o_r,
                         out_stride, pass * kWarpListMaxK, p.part_scores + n_lists * dense,
                         p.part_rows + n_lists * dense));
        if (pass + 1 < n_pass) {
            cursor_update_kernel<<<(n_queries + 127) / 128, 128, 0, st>>>(
                o_s, o_r, n_queries, out_stride, pass * kWarpListMaxK + k_pass - 1, c->base_row, cur_s, cur_r);
            LAUNCHED();
        }
    }
    return VL_OK;
}

// above this k the sampled collect/select path replaces the per-warp k-lists (tuning override:
// VLITE_B200_LARGEK_MIN)
int large_k_min() {
    static int v = [] {
        const char* e = getenv("VLITE_B200_LARGEK_MIN");
        return e ? atoi(e) : 32;  // measured crossover: equal at k=32, 1.7x faster at k=100, 3x at k=256
    }();
    return v;
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
    const int ks = std::min(kLargeKSample, std::max(16, k / 4));
    // sample every `stride`-th tile: at least 3k expected candidates, and never more than 5 % of a scan
    const uint32_t stride = std::max<uint32_t>(20u, (3u * (uint32_t)k) / (uint32_t)ks);
    TRY(c->lk_keys.ensure(key_bytes));
    TRY(c->lk_misc.ensure((size_t)n_queries * (4 + 4 + kLargeKSample * 12) + 64));
    uint32_t* thr = static_cast<uint32_t*>(c->lk_misc.p);
    uint32_t* cnt = thr + n_queries;
    uint64_t* smp_r = reinterpret_cast<uint64_t*>(cnt + n_queries);  // 8*Q bytes in: 8-byte aligned
    uint32_t* smp_s = reinterpret_cast<uint32_t*>(smp_r + (size_t)n_queries * kLargeKSample);
    int* flags = c->d_err;
    if (c->count <= cap) {
        // the whole shard fits a buffer: no sampling, collect every eligible row
        CU(cudaMemsetAsync(thr, 0xFF, (size_t)n_queries * 4, st));
    } else {
        ScanParams ps = p;
        ps.tile_stride = stride;
        TRY(search_lists(c, ps, n_queries, ks, metric, smp_s, smp_r));
        extract_thr_kernel<<<(n_queries + 127) / 128, 128, 0, st>>>(smp_s, smp_r, n_queries, ks, thr);
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
    int h = 0;
    CU(cudaMemcpyAsync(&h, flags, sizeof h, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (h) {
        CU(cudaMemsetAsync(flags, 0, sizeof(int), st));
        return VL_OK;  // overflow / shortfall: let the exact multi-pass path answer
    }
    *done = true;
    return VL_OK;
}

}  // namespace

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
    if (cudaMalloc(&c->d_err, sizeof(int)) != cudaSuccess || cudaMalloc(&c->d_counter, 8) != cudaSuccess)
        return bail(fail(VL_ENOMEM, "cudaMalloc failed"));
    cudaMemset(c->d_err, 0, sizeof(int));
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
                      &c->stage[1], &c->cand_buf, &c->misc, &c->lk_keys, &c->lk_misc})
        b->release();
    for (int i = 0; i < 2; ++i) {
        if (c->pinned[i]) cudaFreeHost(c->pinned[i]);
        if (c->stage_ev[i]) cudaEventDestroy(c->stage_ev[i]);
    }
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
    TRY(mark_appended(c, first, n));
    int rc = check_err_flag(c, "vl_corpus_append_f32");
    if (rc != VL_OK) {  // roll the append back
        c->count = first;
        c->live_count -= n;
        cudaMemsetAsync(c->live, 0, c->capacity / 8, c->stream);  // conservative: rebuild below
        if (first) {
            live_set_range_kernel<<<grid_for(first / 32 + 1, 256, c->sms), 256, 0, c->stream>>>(c->live, 0, first);
            g_launches.fetch_add(1);
        }
        cudaStreamSynchronize(c->stream);
        return fail(VL_ERANGE, "vl_corpus_append_f32: a stored value is outside the encoding (tombstones were reset)");
    }
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
    for (uint64_t done = 0, i = 0; rc == VL_OK && done < n; done += chunk_rows, ++i) {
        const int b = (int)(i & 1);
        const uint64_t m = std::min(chunk_rows, n - done);
        if (i >= 2 && cudaEventSynchronize(c->stage_ev[b]) != cudaSuccess) {
            rc = fail(VL_ECUDA, "event sync failed");
            break;
        }
        // page cache -> pinned staging with several threads (one pread stream tops out near 4 GB/s)
        const size_t want = m * rb;
        static const unsigned io_threads = [] {
            const char* e = getenv("VLITE_B200_IO_THREADS");
            const unsigned want_thr = e ? (unsigned)atoi(e) : 8u;  // measured: 8 = 16 readers (24 GB/s file rate), 32 lose
            return std::max(1u, std::min(std::min(want_thr, 64u), std::max(1u, std::thread::hardware_concurrency())));
        }();
        const unsigned n_thr = io_threads;
        std::vector<size_t> got_part(n_thr, 0);
        {
            std::vector<std::thread> pool;
            const size_t slice = ((want + n_thr - 1) / n_thr + 4095) & ~(size_t)4095;
            for (unsigned t = 0; t < n_thr; ++t) {
                const size_t lo = std::min(want, (size_t)t * slice), hi = std::min(want, lo + slice);
                if (lo >= hi) continue;
                pool.emplace_back([&, t, lo, hi] {
                    size_t g = lo;
                    while (g < hi) {
                        ssize_t r = pread(fd, static_cast<uint8_t*>(c->pinned[b]) + g, hi - g, byte_off + done * rb + g);
                        if (r <= 0) break;
                        g += (size_t)r;
                    }
                    got_part[t] = g - lo;
                });
            }
            for (auto& th : pool) th.join();
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
    const bool q_dev = is_device_ptr(queries_any);
    const bool s_dev = is_device_ptr(out_score_any), r_dev = is_device_ptr(out_row_any);
    const size_t n_out = (size_t)n_queries * k;
    c->timed = false;
    if (c->count == 0 || c->live_count == 0) {
        TRY(fill_sentinels(st, out_score_any, out_row_any, n_out, s_dev, r_dev));
        return VL_OK;
    }

    const uint8_t* q_ptr = queries_any;
    if (!q_dev || (reinterpret_cast<uintptr_t>(queries_any) & 15)) {
        TRY(c->q_buf.ensure((size_t)n_queries * c->w));
        CU(cudaMemcpyAsync(c->q_buf.p, queries_any, (size_t)n_queries * c->w,
                           q_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
        q_ptr = static_cast<const uint8_t*>(c->q_buf.p);
    }
    const uint32_t* filt = nullptr;
    if (filter_mask_any) TRY(stage_filter(c, filter_mask_any, &filt));

    uint32_t* o_s = out_score_any;
    uint64_t* o_r = out_row_any;
    if (!s_dev) {
        TRY(c->out_s.ensure(n_out * 4));
        o_s = static_cast<uint32_t*>(c->out_s.p);
    }
    if (!r_dev) {
        TRY(c->out_r.ensure(n_out * 8));
        o_r = static_cast<uint64_t*>(c->out_r.p);
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
    bool done = false;
    if (k > large_k_min() && !c->force_multipass) {
        int rc = search_large_k(c, p, n_queries, k, metric, o_s, o_r, &done);
        if (rc != VL_OK) return rc;
    }
    if (!done) TRY(search_lists(c, p, n_queries, k, metric, o_s, o_r));
    if (c->timing) {
        CU(cudaEventRecord(c->ev2, st));
        c->timed = true;
    }
    if (!s_dev) CU(cudaMemcpyAsync(out_score_any, o_s, n_out * 4, cudaMemcpyDeviceToHost, st));
    if (!r_dev) CU(cudaMemcpyAsync(out_row_any, o_r, n_out * 8, cudaMemcpyDeviceToHost, st));
    if (!s_dev || !r_dev || !q_dev) CU(cudaStreamSynchronize(st));
    return VL_OK;
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
    if (qtype != VL_QTYPE_F32 && qtype != VL_QTYPE_I8) return fail(VL_EINVAL, "bad qtype");
    if (n_queries < 1 || n_cand < 1 || n_cand > kRescoreMaxCand || k < 1 || k > n_cand)
        return fail(VL_EINVAL, "need 1 <= k <= n_cand <= %d", kRescoreMaxCand);
    DeviceGuard g(c->device);
    cudaStream_t st = c->stream;
    const size_t qbytes = (size_t)n_queries * dims * (qtype == VL_QTYPE_I8 ? 1 : 4);
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
    const size_t smem = (size_t)n_sort * 12 + (size_t)dims * (qtype == VL_QTYPE_I8 ? 1 : 4);
    c->timed = false;
    if (c->timing) CU(cudaEventRecord(c->ev0, st));
    if (qtype == VL_QTYPE_I8) {
        TRY(allow_max_smem(rescore_kernel<true>, c->device));
        rescore_kernel<true><<<n_queries, kRescoreThreads, smem, st>>>(p);
    } else {
        TRY(allow_max_smem(rescore_kernel<false>, c->device));
        rescore_kernel<false><<<n_queries, kRescoreThreads, smem, st>>>(p);
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
    c->force_multipass = stages < 0 ? 1 : 0;  // stages = -1: force the exact multi-pass path for large k
    if (stages < 0) stages = 0;
    c->tune_splits = row_splits;
    c->tune_stages = stages;
    return VL_OK;
}

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
