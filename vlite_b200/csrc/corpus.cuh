// The shard handle behind `vl_corpus*`: append-in-place storage on CUDA virtual memory management
// arenas, the live mask, capacity growth, the TMA tensor map of the rows, filter staging, and the
// once-per-(kernel, device) shared-memory opt-in.  Included by api.cu only.
#pragma once

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
    DevBuf q_buf, qx_buf, filt_buf, part_s, part_r, out_s, out_r, stage[2], cand_buf, misc, lk_keys, lk_misc;
    int force_multipass = 0;  // tuning/testing: disable the sampled large-k path
    void* pinned[2] = {nullptr, nullptr};
    size_t pinned_bytes = 0;
    void* io_pinned = nullptr;   // small pinned staging for pageable host queries / results of vl_search
    size_t io_pinned_bytes = 0;
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    int* d_err = nullptr;
    unsigned long long* d_counter = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    bool timing = false;
    bool timed = false;
    int tune_splits = 0;
    int tune_stages = 0;
    unsigned long long* trace_dev = nullptr;   // vl_scan_trace
    int fuse_merge = 1;  // 1-2 queries: final merge inside the scan launch (0 = separate merge kernels, for A/B runs)
    int scan_path = 0;  // 0 auto, 1 integer-pipe kernel only, 2 / 3 tensor-core kernel with 1 / 2 CTAs per tile
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

int ensure_io_pinned(vl_corpus* c, size_t bytes) {
    if (bytes <= c->io_pinned_bytes) return VL_OK;
    if (c->io_pinned) {
        CU(cudaStreamSynchronize(c->stream));
        cudaFreeHost(c->io_pinned);
    }
    c->io_pinned = nullptr;
    c->io_pinned_bytes = 0;
    const size_t want = std::max<size_t>(bytes, 256 << 10);
    CU(cudaMallocHost(&c->io_pinned, want));
    c->io_pinned_bytes = want;
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
int make_corpus_tmap(const vl_corpus* c, CUtensorMap* out, int box_rows = kTileRows) {
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
    const cuuint32_t box[2] = {(cuuint32_t)c->w, (cuuint32_t)box_rows};
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


}  // namespace
