// Probe of tcgen05.mma kind::mxf4 (block-scaled e2m1, K = 64 per instruction) against kind::i8 (K = 32):
// (1) are the operand-format assumptions right (two e2m1 per byte in the 128 B-swizzle K-major layout,
//     +1.0 = 0x2 / -1.0 = 0xA, UE8M0 scale 0x7F = 1.0 filled uniformly into the scale-factor columns),
// (2) how fast does each kind issue from fixed shared-memory operands on every SM at once?
// (3) how fast can the accumulators be read back (tcgen05.ld x32 streams from 4 / 8 / 16 warps per SM)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -o umma_probe scripts/umma_probe.cu
// One CTA per SM, cta_group::1, M = 128, N = 256 (i8) / N = 224 or 256 (mxf4).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../vlite_b200/csrc/mmascan.cuh"

using namespace vl;

// rows of A: the first 2 * (m % 32) elements are -1, the rest +1; rows of B: the first 2 * (n % 16) elements -1.
// dot(m, n) = K - 2 * |2 (m % 32) - 2 (n % 16)|   (K elements per slab: 256 for e2m1, 128 for int8)
template <bool F4>
__device__ void fill_rows(uint8_t* slab, int rows, int mod, int tid, int nthr) {
    for (int i = tid; i < rows * 128; i += nthr) {
        const int r = i / 128, byte = i % 128;
        const int neg = 2 * (r % mod);
        uint8_t v;
        if (F4) {
            const int e0 = 2 * byte, e1 = 2 * byte + 1;
            v = (uint8_t)((e0 < neg ? 0xA : 0x2) | ((e1 < neg ? 0xA : 0x2) << 4));
        } else {
            v = (uint8_t)(byte < neg ? 0xFF : 0x01);
        }
        const int chunk = byte / 16, within = byte % 16;
        slab[r * 128 + ((chunk ^ (r & 7)) << 4) + within] = v;
    }
}

// TS: the A operand (128 rows x 256 e2m1 = 32 columns) is copied into TMEM and the MMAs read it from there
__device__ __forceinline__ void umma_mxf4_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate,
                                             uint32_t sfa, uint32_t sfb) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], [%1], %2, %3, [%5], [%6], p;\n\t}" ::"r"(d_tmem),
                 "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa), "r"(sfb)
                 : "memory");
}

template <bool F4, int N, bool TS = false, bool COMMIT_EACH = false>
__global__ void __launch_bounds__(128, 1) probe_kernel(int iters, float* out, unsigned long long* ns_out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                 // 128 rows x 128 B
    uint8_t* sB = smem + 16384;         // N rows x 128 B
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    fill_rows<F4>(sA, 128, 32, threadIdx.x, 128);
    fill_rows<F4>(sB, N, 16, threadIdx.x, 128);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_init(&bar2, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc<1>(&s_tmem, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = s_tmem;
    if (F4) {
        // scale factors: 0x7F (UE8M0 1.0) in every byte of columns [N, N + 32) of every lane
        tmem_fill16(tm + ((uint32_t)(warp * 32) << 16) + N, 0x7F7F7F7Fu);
        tmem_fill16(tm + ((uint32_t)(warp * 32) << 16) + N + 16, 0x7F7F7F7Fu);
        tmem_st_wait();
    }
    if (TS) {
        // row m of A: the first 2 (m % 32) elements are -1 (0xA), the rest +1 (0x2): 8 elements per 32-bit column
        const int m = threadIdx.x, neg = 2 * (m % 32);
        uint32_t av[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            uint32_t wv = 0;
#pragma unroll
            for (int e = 0; e < 8; ++e) wv |= (uint32_t)((8 * c + e) < neg ? 0xA : 0x2) << (4 * e);
            av[c] = wv;
        }
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
            "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
            "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(tm + ((uint32_t)(warp * 32) << 16) + N + 64),
            "r"(av[0]), "r"(av[1]), "r"(av[2]), "r"(av[3]), "r"(av[4]), "r"(av[5]), "r"(av[6]), "r"(av[7]), "r"(av[8]), "r"(av[9]),
            "r"(av[10]), "r"(av[11]), "r"(av[12]), "r"(av[13]), "r"(av[14]), "r"(av[15]), "r"(av[16]), "r"(av[17]), "r"(av[18]),
            "r"(av[19]), "r"(av[20]), "r"(av[21]), "r"(av[22]), "r"(av[23]), "r"(av[24]), "r"(av[25]), "r"(av[26]), "r"(av[27]),
            "r"(av[28]), "r"(av[29]), "r"(av[30]), "r"(av[31])
            : "memory");
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    unsigned long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
        const uint32_t idesc = F4 ? umma_idesc_mxf4(128, N) : umma_idesc_i8(128, N);
        const uint64_t ad = umma_desc_sw128(smem_u32(sA)), bd = umma_desc_sw128(smem_u32(sB));
        t0 = global_ns();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const uint32_t acc = (it | kk) != 0 ? 1u : 0u;
                if (TS) umma_mxf4_ts(tm, tm + N + 64 + kk * 8, bd + (uint64_t)(kk * 2), idesc, acc, tm + N, tm + N + 16);
                else if (F4) umma_mxf4<1>(tm, ad + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2), idesc, acc, tm + N, tm + N + 16);
                else umma_i8<1>(tm, ad + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2), idesc, acc);
            }
            if (COMMIT_EACH) umma_commit<1>(&bar2);   // the scan's per-slab hand-back of a ring slot
        }
        umma_commit<1>(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    tc_fence_after();
    if (threadIdx.x == 0) {
        t1 = global_ns();
        ns_out[blockIdx.x] = t1 - t0;
    }
    // read back: D[lane = m][col = n]
    for (int ch = 0; ch < N / 32; ++ch) {
        uint32_t v[32];
        tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)ch * 32u, v);
        tmem_ld_wait();
        if (blockIdx.x == 0) {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const int m = warp * 32 + lane, n = ch * 32 + c;
                out[m * N + n] = F4 ? __uint_as_float(v[c]) : (float)(int)v[c];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<1>(tm, 512);
}

template <bool F4, int N, bool TS = false, bool COMMIT_EACH = false>
int run(const char* name, int iters) {
    float* out;
    unsigned long long* ns;
    cudaMalloc(&out, 128 * N * sizeof(float));
    cudaMalloc(&ns, 148 * 8);
    auto kern = probe_kernel<F4, N, TS, COMMIT_EACH>;
    const size_t smem = 16384 + N * 128 + 1024;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int bad = 0;
    for (int rep = 0; rep < 3; ++rep) {
        kern<<<148, 128, smem>>>(rep == 0 ? 1 : iters, out, ns);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("%s: CUDA error %s\n", name, cudaGetErrorString(e));
            return 1;
        }
        if (rep == 0) {
            std::vector<float> h(128 * N);
            cudaMemcpy(h.data(), out, h.size() * 4, cudaMemcpyDeviceToHost);
            const int K = F4 ? 256 : 128;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < N; ++n) {
                    const int d = 2 * (m % 32) - 2 * (n % 16);
                    const float want = (float)(K - 2 * (d < 0 ? -d : d));
                    if (h[m * N + n] != want) {
                        if (bad < 5) printf("%s: D[%d][%d] = %g, expected %g\n", name, m, n, h[m * N + n], want);
                        ++bad;
                    }
                }
            printf("%s: one slab (4 MMAs): %d of %d accumulators wrong\n", name, bad, 128 * N);
        } else if (rep == 2) {
            std::vector<unsigned long long> t(148);
            cudaMemcpy(t.data(), ns, 148 * 8, cudaMemcpyDeviceToHost);
            unsigned long long mx = 0;
            for (auto v : t) mx = v > mx ? v : mx;
            const double macs = (double)iters * 4 * 128.0 * N * (F4 ? 64 : 32) * 148;
            printf("%s: %d x 4 MMAs per SM on 148 SMs: %.1f us (slowest SM) = %.1f ns per MMA, %.2f P MAC/s = %.2f POP/s chip-wide\n", name,
                   iters, mx / 1e3, (double)mx / (iters * 4.0), macs / (mx * 1e-9) / 1e15, 2 * macs / (mx * 1e-9) / 1e15);
        }
    }
    cudaFree(out);
    cudaFree(ns);
    return bad;
}

// ---- TMEM read-out bandwidth: NW warps per SM stream the 512 allocated columns with tcgen05.ld x32 ----
// WAIT_EVERY = how many loads are issued before one tcgen05.wait::ld (1 = the scan's epilogue pattern)
template <int NW, int WAIT_EVERY>
__global__ void __launch_bounds__(NW * 32, 1) ldtm_kernel(int iters, unsigned long long* ns_out, uint32_t* sink) {
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc<1>(&s_tmem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = s_tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const unsigned long long t0 = global_ns();
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int ch = 0; ch < 16; ch += WAIT_EVERY) {
            uint32_t v[WAIT_EVERY][32];
#pragma unroll
            for (int u = 0; u < WAIT_EVERY; ++u) tmem_ld32(tm + (uint32_t)(ch + u) * 32u, v[u]);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < WAIT_EVERY; ++u)
#pragma unroll
                for (int c = 0; c < 32; c += 8) acc ^= v[u][c];
        }
    }
    __syncthreads();
    const unsigned long long t1 = global_ns();
    if (threadIdx.x == 0) ns_out[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<1>(s_tmem, 512);
}
template <int NW, int WAIT_EVERY>
void run_ldtm(int iters) {
    unsigned long long* ns;
    uint32_t* sink;
    cudaMalloc(&ns, 148 * 8);
    cudaMalloc(&sink, 4);
    for (int rep = 0; rep < 2; ++rep) ldtm_kernel<NW, WAIT_EVERY><<<148, NW * 32>>>(iters, ns, sink);
    if (cudaDeviceSynchronize() != cudaSuccess) {
        printf("ldtm: CUDA error\n");
        return;
    }
    std::vector<unsigned long long> t(148);
    cudaMemcpy(t.data(), ns, 148 * 8, cudaMemcpyDeviceToHost);
    unsigned long long mx = 0;
    for (auto v : t) mx = v > mx ? v : mx;
    const double bytes = (double)iters * 16 * NW * 32 * 32 * 4;     // per SM
    printf("tcgen05.ld x32, %d warps per SM, %d load(s) per wait: %.1f B/ns per SM = %.1f B/clk at 1.965 GHz; %.2f T scores/s chip-wide\n", NW,
           WAIT_EVERY, bytes / mx, bytes / mx / 1.965, bytes / 4 / mx * 148 / 1e3);
    cudaFree(ns);
    cudaFree(sink);
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 20000;
    int bad = 0;
    bad += run<false, 256>("i8   M128 N256 K32", iters);
    bad += run<true, 256>("mxf4 M128 N256 K64", iters);
    bad += run<true, 224>("mxf4 M128 N224 K64", iters);
    bad += run<true, 224, true>("mxf4 M128 N224 K64, A in TMEM", iters);
    bad += run<true, 160, true>("mxf4 M128 N160 K64, A in TMEM", iters);
    bad += run<true, 160>("mxf4 M128 N160 K64", iters);
    bad += run<true, 224, false, true>("mxf4 M128 N224 K64, tcgen05.commit after every 4 MMAs", iters);
    bad += run<false, 256, false, true>("i8   M128 N256 K32, tcgen05.commit after every 4 MMAs", iters);
    run_ldtm<4, 1>(2000);
    run_ldtm<8, 1>(2000);
    run_ldtm<8, 2>(2000);
    run_ldtm<16, 1>(2000);
    return bad ? 1 : 0;
}
