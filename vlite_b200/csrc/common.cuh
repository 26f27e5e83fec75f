// Shared device helpers for the vlite_b200 kernels (sm_100a only).
#pragma once
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "vlite_b200 kernels are written for sm_100a (B200) only"
#endif

namespace vl {

// ---- debug build (-DVL_DEBUG_BOUNDS): device-side bounds / protocol assertions -----------------
// compute-sanitizer is closed on the GPU pool, so the kernels carry their own checks: every index
// into a list, candidate buffer, shared-memory ring, TMEM column range or document store, and the
// pipeline distances the mbarrier rings rely on, are asserted in this build.  A failed check does
// not trap (a trap would take the context down): it bumps a counter and records the id of the
// first failure, which vl_debug_checks() returns to the host.  Off (zero cost) in the shipped build.
#ifdef VL_DEBUG_BOUNDS
__device__ unsigned int g_vl_dcheck[2];   // [0] failures so far, [1] id of the first one
#define VL_DCHECK(cond, id)                                                   \
    do {                                                                      \
        if (!(cond)) {                                                        \
            if (atomicAdd(&::vl::g_vl_dcheck[0], 1u) == 0u) ::vl::g_vl_dcheck[1] = (id); \
        }                                                                     \
    } while (0)
#else
#define VL_DCHECK(cond, id) ((void)0)
#endif

constexpr uint32_t kSentinelScore = 0xFFFFFFFFu;
constexpr uint64_t kSentinelRow = 0xFFFFFFFFFFFFFFFFull;
constexpr unsigned kFullMask = 0xFFFFFFFFu;

enum Metric : int { kHamming = 0, kXorSum = 1 };

// ---- shared-memory / mbarrier / bulk-copy (TMA) PTX wrappers -----------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make barrier inits visible to the async proxy before any bulk copy targets them
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// non-blocking probe (debug assertions only)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// 1-D bulk async copy global -> shared through the TMA unit (SASS: UBLKCP);
// completion is signalled on `bar` as `bytes` of transaction count.
// bytes must be a multiple of 16; src and dst 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// same, with an L2 eviction-priority hint (createpolicy value)
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                              uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
// 2-D tiled TMA load (SASS: UTMALDG) through a tensor map: box {x..x+W bytes, y..y+rows}; rows
// outside the tensor are zero-filled, the full box is always counted on the mbarrier.
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* tmap, int x, int y, uint64_t* bar,
                                            uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_u32(dst_smem)),
        "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
    return v;
}
// streaming 128-bit global load that does not allocate in L1
__device__ __forceinline__ uint4 ldg128_stream(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}

// ---- scoring primitives -------------------------------------------------------
template <int METRIC>
__device__ __forceinline__ uint32_t score_word(uint32_t x, uint32_t acc) {
    if constexpr (METRIC == kHamming) {
        return acc + __popc(x);
    } else {
        return __dp4a(x, 0x01010101u, acc);  // sum of the four byte values
    }
}
template <int METRIC>
__device__ __forceinline__ uint32_t score_chunk(const uint4& a, const uint4& b, uint32_t acc) {
    acc = score_word<METRIC>(a.x ^ b.x, acc);
    acc = score_word<METRIC>(a.y ^ b.y, acc);
    acc = score_word<METRIC>(a.z ^ b.z, acc);
    acc = score_word<METRIC>(a.w ^ b.w, acc);
    return acc;
}

// ---- counter-based generator (must match oracle/synth.py bit for bit) ---------
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z ^= z >> 30;
    z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27;
    z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}
__host__ __device__ __forceinline__ uint64_t synth_word(uint64_t seed, uint64_t row, uint64_t j, uint64_t stride) {
    return mix64((row * stride + j) * 0x9E3779B97F4A7C15ull + seed * 0xD1B54A32D192ED03ull + 0x9E3779B97F4A7C15ull);
}

}  // namespace vl
