// Host-side helpers shared by every entry point of the library (included by api.cu only: the
// library is one translation unit): thread-local error text, CUDA error macros, device guard,
// grow-only device buffers, cached device facts, grid sizing.
#pragma once

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

// Grow-only device scratch.  Growing never frees the old block while work that uses it may still
// be in flight on the handle's stream (cudaFree would synchronise the device -- a latency cliff when
// a batch grows -- and enqueue-only calls must stay enqueue-only): the old block is retired and
// freed with the handle.  Growth is geometric below 64 MB, exact above, so retired blocks add up
// to less than the live one.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    std::vector<void*> retired;
    int ensure(size_t bytes) {
        if (bytes <= cap) return VL_OK;
        size_t want = std::max(bytes, (size_t)4096);
        if (want < ((size_t)64 << 20)) want = std::max(want, cap * 2);
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(VL_ENOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        }
        if (p) retired.push_back(p);
        p = q;
        cap = want;
        return VL_OK;
    }
    void release() {
        if (p) cudaFree(p);
        for (void* r : retired) cudaFree(r);
        retired.clear();
        p = nullptr;
        cap = 0;
    }
};

// 0 = pageable host memory, 1 = pinned / registered host memory, 2 = device (or managed) memory
int ptr_kind(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) return 2;
    return a.type == cudaMemoryTypeHost ? 1 : 0;
}

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
