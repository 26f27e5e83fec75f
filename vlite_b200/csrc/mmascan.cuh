// Kernel 1T: the scan for large query batches on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Same contract as scan_topk_kernel (scan.cuh): HAMMING scores of a batch of queries against the
// packed corpus (reference: EmbeddingModel.hamming_distance, vlite/model.py:71-74, applied to
// unpacked bits) + per-CTA top-k lists by (score, row) (vlite/model.py:85), bit-exact.
//
// Why a second kernel: for Q >= 64 the scan is a dense binary contraction,
//     hamming(q, e) = (8W - <s(q), s(e)>) / 2,   s(bit) = +1 (bit 0) / -1 (bit 1),
// and the LOP3/POPC kernel is bound by the integer pipes (16 POPC/clk/SM) at ~2e11 pairs/s while
// the int8 tensor pipe (tcgen05.mma kind::i8, exact int32 accumulation) does 8192 MAC/clk/SM =
// 1.5e12 pairs/s at W = 128.  The corpus stays PACKED in HBM (the format is the product); rows are
// expanded bit -> int8 in shared memory only, once per 128 (or 256) queries.
//
//   grid    = (query tiles of 128, row splits); CG = 1: one CTA per tile, CG = 2: a CTA pair
//             (cluster of 2, tcgen05 cta_group::2) shares every expanded row tile between its 256
//             queries -- each CTA expands only half of the rows it scores.
//   A (M)   = this CTA's 128 queries, expanded once into shared memory (K-major, 128 B swizzle).
//   B (N)   = corpus rows: TMA streams packed 128-row boxes (hardware swizzle) into a 2-stage ring;
//             8 "expander" warps turn each 16-byte chunk of a row into 128 int8 (one K slab of one
//             row: SHL + PRMT sign-replicate + LOP3) and store it in the UMMA canonical layout
//             into a 3-stage slab ring (generic-proxy stores -> fence.proxy.async -> mbarrier).
//             The K order inside a slab is a fixed permutation applied to rows and queries alike.
//             Masked rows (tombstones, filter, past the end) expand to zeros.
//   D       = TMEM, 128 lanes (queries) x NT columns (rows of the tile) int32, double buffered:
//             one elected thread issues W/4 tcgen05.mma (K = 32 each) per tile and commits to
//             mbarriers; the epilogue of tile i overlaps the MMAs of tile i + 1.
//   epilogue= 4 warps, one query per thread (its TMEM lane): tcgen05.ld 32 columns at a time, a
//             max tree against the thread's own threshold, and -- rarely -- the insert of (score,
//             row) into the thread's register-resident sorted k-list (k <= 16) or the append to
//             the per-query collect buffer (large k).  tau[] shares thresholds between CTAs as in
//             scan.cuh.  Output = one sorted list per (row split, query) for merge.cuh.
#pragma once
#include <type_traits>
#include "scan.cuh"

namespace vl {

constexpr int kMmaQT = 128;          // queries per CTA = TMEM lanes = UMMA M per CTA
constexpr int kMmaRows = 128;        // corpus rows one CTA expands per tile
constexpr int kMmaMaxBStages = 8;
constexpr int kMmaPkStages = 2;      // packed row boxes in flight
constexpr int kMmaThreads = 512;     // warps 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle | 4-7 epilogue | 8-15 expanders (e2m1: 8-11 expanders, 12-15 epilogue)
constexpr int kMmaExpWarps = 8;
constexpr int kMmaListCap = 16;      // k <= 16 per-thread register list
constexpr int kMmaSlabBytes = 128 * 128;
constexpr int kTraceCtas = 296;
constexpr int kTraceTile = 64;      // the steady-state tile whose hand-offs every role of CTA (0,0) stamps
constexpr int kTraceRoleBase = 16 + 512 + 2 * kTraceCtas;   // + 64 words: [0..7] TMA, [8..15] MMA, [16..31] expander warp 8, [32..47] / [48..63] epilogue warps 4 / 12
constexpr int kTraceWords = kTraceRoleBase + 64;   // vl_scan_trace: 16 stamps | per-tile stamps + insert counts of 256 tiles | entry / exit of every CTA | one tile, every role

// F4 = the +-1 contraction in 4-bit floats (tcgen05.mma kind::mxf4: e2m1 operands, +1.0 = 0x2, -1.0 = 0xA, two per
// byte, unit UE8M0 block scales, fp32 accumulation -- exact: every product is +-1 and |sum| <= 1024 < 2^24).  The
// instruction takes K = 64 in the time kind::i8 takes K = 32 (measured: 8.9 vs 4.6 POP/s, scripts/umma_probe.cu),
// an expanded row is 4 W bytes instead of 8 W, and the expansion is 2 instructions per 8 elements instead of
// 3 per 4.  HAMMING only: XORSUM's bit weights do not fit e2m1.  The 32 scale-factor columns need TMEM room
// beside the two accumulators, so a CTA pair scores 224 rows per tile (112 per CTA) instead of 256.
template <int W, int CG, bool F4 = false, int METRIC = kHamming>
struct MmaCfg {
    static constexpr int K = W * 8;                 // elements per row
    static constexpr int SLABS = F4 ? W / 32 : W / 16;   // K slabs: 128 B per row = 256 e2m1 / 128 int8
    static constexpr int ROWS = (F4 && CG == 2) ? 112 : kMmaRows;   // corpus rows one CTA expands per tile
    static constexpr int NT = ROWS * CG;            // rows per tile = UMMA N
    static constexpr int MW = NT / 32;              // mask words per tile
    static constexpr int A_BYTES = SLABS * kMmaSlabBytes;
    static constexpr int PK_BYTES = ROWS * W;
    static constexpr int PK_STRIDE = (PK_BYTES + 1023) & ~1023;   // swizzled boxes start on the swizzle period
    static constexpr int SF_COLS = F4 ? 48 : 0;     // scale factors: 32 columns for A (XORSUM: 4 per bit plane), 16 for B
    // F4: expansion is 2 instructions per 8 elements and the MMAs take half the time, so the accumulator read-out is
    // the critical role: 4 expander warps (one thread per row) and TWO epilogue groups of 4 warps (warps 4-7: even
    // 32-column chunks, warps 12-15: odd ones), each thread with its own k-list -> 2 lists per (split, query)
    static constexpr int EXP_WARPS = F4 ? 4 : kMmaExpWarps;
    static constexpr int EPI_GROUPS = F4 ? 2 : 1;
    static constexpr int TMEM_COLS = F4 ? 512 : 2 * NT;   // two accumulator buffers [| scale factors]; a power of two
    // expanded K slabs in flight.  One slab hand-off (b_empty wake -> expand -> st.shared -> fence.proxy.async
    // -> arrive -> MMA thread wake -> 4 MMAs -> commit) is ~4x longer than the slab's 512 tensor clocks, so
    // the ring must hold that many: 4 is what fits beside the 128 KB query operand at W = 128 (3 stages
    // measured 71 % of the tensor rate in steady state), narrower rows leave room for more.
    static constexpr int BSTAGES = F4 ? 8 : (W >= 128 ? 4 : 6);
    static constexpr size_t SMEM =
        (size_t)A_BYTES + (size_t)BSTAGES * kMmaSlabBytes + (size_t)kMmaPkStages * PK_STRIDE +
        (size_t)kMmaPkStages * 2 * 8 * 4 + 32 * 8 + 16 + kMmaQT * 4 + 1024 /* alignment slack */;
    static_assert(W == 32 || W == 64 || W == 128, "row width");
    static_assert(CG == 1 || CG == 2, "cta group");
    static_assert(NT % 32 == 0 && 2 * NT + SF_COLS <= 512, "tile width / TMEM budget");
    static_assert(SMEM <= 227 * 1024, "shared memory");
};

// ---- tcgen05 / cluster PTX wrappers ------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// arrive on a barrier of another CTA of the cluster.  Default semantics (release at CTA scope) on
// purpose: the data the barrier guards never leaves the arriving CTA's own shared memory (it is read
// by that SM's tensor core), only the signal crosses; a cluster-scope release would cost a
// MEMBAR.ALL.GPU + L1 invalidate per slab.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
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
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait_cluster(bar, parity)) {
    }
}
// generic-proxy shared-memory writes -> visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {   // one full warp
    if constexpr (CG == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    if constexpr (CG == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
    else
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], int8 x int8 -> int32, issued by ONE thread (SASS UTCIMMA)
template <int CG>
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (CG == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc),
                     "r"(idesc), "r"(accumulate)
                     : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc),
                     "r"(idesc), "r"(accumulate)
                     : "memory");
}
// D[tmem] (+)= A * B, e2m1 x e2m1 -> f32 with UE8M0 block scales (one per 32 elements) read from TMEM (SASS UTCOMMA)
template <int CG>
__device__ __forceinline__ void umma_mxf4(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate,
                                          uint32_t sfa_tmem, uint32_t sfb_tmem) {
    if constexpr (CG == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(d_tmem),
                     "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa_tmem), "r"(sfb_tmem)
                     : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(d_tmem),
                     "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa_tmem), "r"(sfb_tmem)
                     : "memory");
}
// one value into 32 lanes x 16 consecutive 32-bit columns of TMEM (lane = thread)
__device__ __forceinline__ void tmem_fill16(uint32_t taddr, uint32_t x) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(x)
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// The same instructions issued from CONVERGED warp code by one elected lane (elect.sync inside the asm).  A tcgen05
// operand must sit in a uniform register: from a divergent `if (lane == 0)` region the compiler moves every operand
// there with an ELECT / R2UR.BROADCAST / BRA.U.ANY loop -- 14 instructions per MMA on the one thread whose
// bookkeeping paces the whole CTA pair; converged code needs a plain R2UR.
template <int CG>
__device__ __forceinline__ void umma_mxf4_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate,
                                                uint32_t sfa_tmem, uint32_t sfb_tmem) {
    if constexpr (CG == 1)
        asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "@q tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(d_tmem),
                     "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa_tmem), "r"(sfb_tmem)
                     : "memory");
    else
        asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "@q tcgen05.mma.cta_group::2.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(d_tmem),
                     "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa_tmem), "r"(sfb_tmem)
                     : "memory");
}
template <int CG>
__device__ __forceinline__ void umma_i8_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (CG == 1)
        asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "@q tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc),
                     "r"(accumulate)
                     : "memory");
    else
        asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "@q tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc),
                     "r"(accumulate)
                     : "memory");
}
template <int CG>
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
    if constexpr (CG == 1)
        asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                     "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
                     : "memory");
    else
        asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                     "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(
                         smem_u32(bar)),
                     "h"((uint16_t)3)
                     : "memory");
}
// arrive on an mbarrier once every tcgen05 operation issued so far by this thread has completed;
// CG = 2: the same barrier offset in both CTAs of the pair
template <int CG>
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    if constexpr (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                         smem_u32(bar)),
                     "h"((uint16_t)3)
                     : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns of TMEM -> 32 registers per thread (lane = thread)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor of a K-major operand slab in the canonical 128 B-swizzle layout:
// row r of the slab at r * 128 B, its 16 B chunk c stored at chunk position c ^ (r & 7); groups of 8
// rows 1024 B apart (stride byte offset), slab base 1024 B aligned (base offset 0).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) /* LBO (unused for swizzled K-major) */ |
           (64ull << 32) /* SBO = 1024 B */ | (1ull << 46) /* descriptor version: sm_100 */ |
           (2ull << 61) /* SWIZZLE_128B */;
}
// instruction descriptor: D = s32, A = B = signed 8 bit, both K-major, shape M x N (K = 32 implied)
__host__ __device__ constexpr uint32_t umma_idesc_i8(int m, int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d);
// block-scaled instruction descriptor: A = B = e2m1 (format 1), both K-major, scale format UE8M0, shape M x N (K = 64)
__host__ __device__ constexpr uint32_t umma_idesc_mxf4(int m, int n) {
    return (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (1u << 23) | ((uint32_t)(m >> 4) << 24);
}

// ---- bit -> e2m1 expansion (F4) ------------------------------------------------------------------
// One packed 32-bit word -> one 16-byte chunk of 32 e2m1 elements: nibble n of output word j is bit 4 n + j
// of the input, as sign (bit 1 -> 0xA = -1.0, bit 0 -> 0x2 = +1.0).  The same map for rows and queries; a
// masked row passes keep = 0 and expands to zeros.  Two instructions per 8 elements (SHL + LOP3).
__device__ __forceinline__ void emit_chunk_f4(uint32_t x, uint32_t keep, uint32_t chunk_saddr) {
    const uint32_t sgn = 0x88888888u & keep, one = 0x22222222u & keep;
    sts128(chunk_saddr, ((x << 3) & sgn) | one, ((x << 2) & sgn) | one, ((x << 1) & sgn) | one, (x & sgn) | one);
}

// XORSUM as e2m1: K order = bit planes (element p * W + byte), so that every 32-element block lies in ONE plane and
// the plane's weight 2^(7 - p) can ride the query operand's UE8M0 block scale (127 + 7 - p); rows and queries are
// plain +-1.  One 16-byte output chunk = plane p of 32 consecutive bytes (8 packed words): output word t takes the
// bytes of words 2 t (even nibbles) and 2 t + 1 (odd nibbles); s = 7 - p selects the bit.
__device__ __forceinline__ uint32_t plane_word_f4(uint32_t xa, uint32_t xb, uint32_t s, uint32_t keep) {
    // bit s of byte i -> the sign bit of nibble 2 i (position 8 i + 3) / 2 i + 1 (8 i + 7): one rotate each -- whatever
    // wraps around lands on positions the masks drop -- then two LOP3 (masks pre-ANDed with keep): 4 instructions
    const uint32_t ra = __funnelshift_r(xa, xa, s - 3u), rb = __funnelshift_r(xb, xb, s - 7u);
    return ((ra & (0x08080808u & keep)) | (0x22222222u & keep)) | (rb & (0x80808080u & keep));
}
__device__ __forceinline__ void emit_plane_chunk_f4(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t x4, uint32_t x5,
                                                    uint32_t x6, uint32_t x7, uint32_t s, uint32_t keep, uint32_t chunk_saddr) {
    sts128(chunk_saddr, plane_word_f4(x0, x1, s, keep), plane_word_f4(x2, x3, s, keep), plane_word_f4(x4, x5, s, keep),
           plane_word_f4(x6, x7, s, keep));
}
// 16 registers -> 32 lanes x 16 consecutive 32-bit columns of TMEM (lane = thread)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}

// ---- bit -> int8 expansion -----------------------------------------------------------------------
// Element order of an expanded row ("bit planes"): e = p * W + byte, p = 0..7 <-> bit (7 - p) of every
// byte -- the same permutation of K for corpus rows and queries, so the contraction is unchanged.
// Output chunk g (16 int8) is plane p = g / (W/16), bytes 16 * (g % (W/16)) .. + 15, i.e. four packed
// words shifted left by p with the sign of each byte replicated (PRMT) -> -1 (bit set) / +1.
// Rows are always +-1.  Queries are +-1 for HAMMING; for XORSUM (score = sum of byte values of q ^ e,
// vlite/model.py:82) they carry the bit's weight: +-2^(7-p) for p >= 1 and +-64 for plane 0, whose
// k-steps the MMA thread issues TWICE (64 + 64 = 128: +128 does not fit an int8).  Then
//     sum_e a_e q_e = sum_bits weight * (+1 equal, -1 different)  =>  score = (255 W - dot) / 2.
template <int P, bool WEIGHTED>
__device__ __forceinline__ uint32_t expand_plane(uint32_t x, uint32_t keep) {
    // prmt selector nibble 8 + i: replicate the sign bit of byte i (the __byte_perm intrinsic masks
    // that mode bit off, hence inline PTX)
    uint32_t m;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(m) : "r"(x << P), "r"(0u), "r"(0xBA98u));
    if constexpr (!WEIGHTED) {
        return (m | 0x01010101u) & keep;
    } else {
        constexpr uint32_t w = P == 0 ? 64u : (1u << (7 - P));
        constexpr uint32_t POS = w * 0x01010101u, NEG = ((256u - w) & 0xFFu) * 0x01010101u;
        return ((POS & ~m) | (NEG & m)) & keep;
    }
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// output chunk G of one row from its packed words pk[W/4] (registers): stored at chunk position
// (G % 8) ^ (row & 7) of the row's 128 B in slab G / 8 (the UMMA 128 B-swizzle layout)
template <int W, int G, bool WEIGHTED>
__device__ __forceinline__ void emit_chunk(const uint32_t (&pk)[W / 4], uint32_t keep, uint32_t slab0_row_saddr, uint32_t r7) {
    constexpr int CHUNKS = W / 16, P = G / CHUNKS, B = G % CHUNKS;
    const uint32_t o0 = expand_plane<P, WEIGHTED>(pk[4 * B + 0], keep), o1 = expand_plane<P, WEIGHTED>(pk[4 * B + 1], keep);
    const uint32_t o2 = expand_plane<P, WEIGHTED>(pk[4 * B + 2], keep), o3 = expand_plane<P, WEIGHTED>(pk[4 * B + 3], keep);
    VL_DCHECK(r7 < 8u && (slab0_row_saddr & 127u) == 0u, 201);   // a 128 B row of a 1024 B-aligned slab
    sts128(slab0_row_saddr + (uint32_t)(G / 8) * kMmaSlabBytes + ((((uint32_t)(G % 8)) ^ r7) << 4), o0, o1, o2, o3);
}
// the four chunks [8 J + 4 H, 8 J + 4 H + 4) of slab J (slab_row_saddr addresses the row inside THAT slab's buffer)
template <int W, int J, int H, bool WEIGHTED>
__device__ __forceinline__ void emit_half_slab(const uint32_t (&pk)[W / 4], uint32_t keep, uint32_t slab_row_saddr, uint32_t r7) {
    const uint32_t base = slab_row_saddr - (uint32_t)J * kMmaSlabBytes;   // emit_chunk adds (G / 8) slabs back
    emit_chunk<W, 8 * J + 4 * H + 0, WEIGHTED>(pk, keep, base, r7);
    emit_chunk<W, 8 * J + 4 * H + 1, WEIGHTED>(pk, keep, base, r7);
    emit_chunk<W, 8 * J + 4 * H + 2, WEIGHTED>(pk, keep, base, r7);
    emit_chunk<W, 8 * J + 4 * H + 3, WEIGHTED>(pk, keep, base, r7);
}
// the same four chunks for corpus ROWS (+-1, no weights) with the slab index a RUNTIME value: the plane
// only enters as a shift amount, so the expanders' slab loop stays rolled -- ~3 KB of SASS instead of
// ~21 KB unrolled, which matters twice: every SM starts with a cold instruction cache, and the hot
// loops of the four warp roles together must fit the 32 KB L1.5 instruction cache
template <int W, int H>
__device__ __forceinline__ void emit_half_slab_rows(const uint32_t (&pk)[W / 4], int j, uint32_t keep, uint32_t slab_row_saddr,
                                                    uint32_t r7) {
    constexpr int CHUNKS = W / 16;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
        // global chunk G = 8 j + 4 H + cc: plane G / CHUNKS (runtime), packed chunk G % CHUNKS (static)
        constexpr int kPlanesPerSlab = 8 / CHUNKS;                    // 1, 2 or 4 bit planes per 128-element slab
        const int B = (4 * H + cc) % CHUNKS;
        const uint32_t sh = (uint32_t)(j * kPlanesPerSlab + (4 * H + cc) / CHUNKS);
        uint32_t o[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            uint32_t m;
            asm("prmt.b32 %0, %1, %2, %3;" : "=r"(m) : "r"(pk[4 * B + t] << sh), "r"(0u), "r"(0xBA98u));
            o[t] = (m | 0x01010101u) & keep;
        }
        sts128(slab_row_saddr + ((((uint32_t)(4 * H + cc)) ^ r7) << 4), o[0], o[1], o[2], o[3]);
    }
}
template <int W, int H, bool WEIGHTED, int J = 0>
__device__ __forceinline__ void emit_half_slab_rt(int j, const uint32_t (&pk)[W / 4], uint32_t keep, uint32_t slab_row_saddr,
                                                  uint32_t r7) {
    if constexpr (J < W / 16) {
        if (j == J) emit_half_slab<W, J, H, WEIGHTED>(pk, keep, slab_row_saddr, r7);
        else emit_half_slab_rt<W, H, WEIGHTED, J + 1>(j, pk, keep, slab_row_saddr, r7);
    }
}

template <int W, int CG, bool COLLECT, int METRIC, bool F4 = false>
__global__ void __launch_bounds__(kMmaThreads, 1)
mma_scan_kernel(const __grid_constant__ CUtensorMap tmap, const ScanParams p) {
    using Cfg = MmaCfg<W, CG, F4, METRIC>;
    if (p.gate != nullptr && *p.gate == 0) return;   // uniform over the grid (and over a cluster)
    constexpr int SLABS = Cfg::SLABS, NT = Cfg::NT, MW = Cfg::MW, ROWS = Cfg::ROWS;
    constexpr int EXPW = Cfg::EXP_WARPS, EG = Cfg::EPI_GROUPS;
    constexpr int CHUNKS = W / 16;
    constexpr bool XS = METRIC == kXorSum;
    static_assert(!(F4 && XS && W < 64), "XORSUM as e2m1: a K = 64 step must lie inside one bit plane (one block scale)");
    constexpr bool F4H = F4 && !XS, F4X = F4 && XS;   // e2m1 forms: HAMMING (word-major K order) / XORSUM (plane-major, scaled)
    using AccT = typename std::conditional<F4, float, int>::type;   // accumulator type as the epilogue compares it
    constexpr int K = XS ? 255 * W : 8 * W;      // score = (K - dot) / 2: the dot product of a perfect match
    const bool tracing = p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
    if (tracing && threadIdx.x == 0) p.trace[0] = global_ns();
    auto stamp = [&](uint32_t i, int slot) {   // role timeline of tile kTraceTile (and the one after it: slot + 4 .. for i + 1 where noted)
        if (tracing && i == (uint32_t)kTraceTile) p.trace[kTraceRoleBase + slot] = global_ns();
    };
    const uint32_t cta_lin = blockIdx.y * gridDim.x + blockIdx.x;
    if (p.trace != nullptr && threadIdx.x == 0 && cta_lin < (uint32_t)kTraceCtas) p.trace[16 + 512 + cta_lin] = global_ns();

    extern __shared__ uint8_t smem_raw[];
    // the dynamic shared window is only 16 B aligned by contract: swizzled operands want 1024 B
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = sA + Cfg::A_BYTES;
    // ring depth: Cfg::BSTAGES unless a tuning run asks for fewer (p.stages in [2, BSTAGES]); the shared
    // memory layout always reserves BSTAGES slabs
    constexpr uint32_t kMmaBStages = (uint32_t)Cfg::BSTAGES;   // (a compile-time ring depth: the MMA thread's per-slab bookkeeping is on the critical path)
    uint8_t* sPk = sB + Cfg::BSTAGES * kMmaSlabBytes;
    uint32_t* sMask = reinterpret_cast<uint32_t*>(sPk + kMmaPkStages * Cfg::PK_STRIDE);   // [stage][live | filter][MW]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sMask + kMmaPkStages * 2 * MW);
    uint64_t* pk_full = bars;                    // [2]  TMA -> expanders
    uint64_t* pk_empty = bars + 2;               // [2]  expanders -> TMA
    uint64_t* acc_full = bars + 4;               // [2]  tcgen05.commit -> epilogue (each CTA its own copy)
    uint64_t* acc_empty = bars + 6;              // [2]  epilogue warps (both CTAs) -> MMA thread of the leader
    uint64_t* b_full = bars + 8;                 // [BSTAGES]  expanders (both CTAs) -> MMA thread of the leader
    uint64_t* b_empty = bars + 8 + kMmaMaxBStages;   // [BSTAGES]  tcgen05.commit -> expanders (each CTA its own copy)
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 8 + 2 * kMmaMaxBStages);
    uint32_t* s_done = s_tmem + 1;                // epilogue warps that have finished (stops the helper warps)
    uint32_t* s_bound = s_tmem + 4;              // [128] per-query admission bound from the device-wide histogram

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    const int q0 = p.q_offset + (int)blockIdx.x * kMmaQT;     // blockIdx.x = pair * CG + rank
    const int split = blockIdx.y;
    const int n_splits = gridDim.y;
    const uint32_t tile_begin = (uint32_t)(((uint64_t)p.n_tiles * split) / n_splits);
    const uint32_t tile_end = (uint32_t)(((uint64_t)p.n_tiles * (split + 1)) / n_splits);
    const uint32_t my_tiles = tile_end - tile_begin;
    const bool has_filter = p.filter != nullptr;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap);
        for (int s = 0; s < kMmaPkStages; ++s) {
            mbar_init(&pk_full[s], 1);
            mbar_init(&pk_empty[s], EXPW);
        }
        for (int s = 0; s < Cfg::BSTAGES; ++s) {
            mbar_init(&b_full[s], EXPW * CG);
            mbar_init(&b_empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 4 * EG * CG);
        }
        mbar_fence_init();
        *s_done = 0;
    }
    if (threadIdx.x < kMmaQT) s_bound[threadIdx.x] = kSentinelScore;
    if (warp == 2) tmem_alloc<CG>(s_tmem, (uint32_t)Cfg::TMEM_COLS);

    // ---- A operand: expand this CTA's queries once (zeros past n_queries; XORSUM: weighted) ----
    if constexpr (F4X) {
        // plane-major: chunk G of the expanded query = plane G / (W / 32) of bytes [32 (G % (W / 32)), + 32)
        constexpr int CPP = W / 32;                 // chunks per plane
        const int ql = threadIdx.x % kMmaQT;
        const int part = threadIdx.x / kMmaQT;
        const int qg = q0 + ql;
        const uint32_t keep = qg < p.n_queries ? 0xFFFFFFFFu : 0u;
        const uint32_t row_saddr = smem_u32(sA) + (uint32_t)ql * 128u;
        const uint32_t r7 = (uint32_t)(ql & 7);
        for (int G = part; G < SLABS * 8; G += kMmaThreads / kMmaQT) {
            const int pl = G / CPP, bb = G % CPP;
            uint4 xa = make_uint4(0u, 0u, 0u, 0u), xb = xa;
            if (keep) {
                const uint4* src = reinterpret_cast<const uint4*>(p.queries + (size_t)qg * W + (size_t)bb * 32);
                xa = src[0];
                xb = src[1];
            }
            emit_plane_chunk_f4(xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w, (uint32_t)(7 - pl), keep,
                                row_saddr + (uint32_t)(G >> 3) * kMmaSlabBytes + ((((uint32_t)(G & 7)) ^ r7) << 4));
        }
        fence_proxy_async_smem();
    } else if constexpr (F4H) {
        const int ql = threadIdx.x % kMmaQT;        // 4 threads per query, each a quarter of the half-slabs
        const int part = threadIdx.x / kMmaQT;
        const int qg = q0 + ql;
        const uint32_t keep = qg < p.n_queries ? 0xFFFFFFFFu : 0u;
        const uint32_t row_saddr = smem_u32(sA) + (uint32_t)ql * 128u;
        const uint32_t r7 = (uint32_t)(ql & 7);
        for (int u = part; u < SLABS * 2; u += kMmaThreads / kMmaQT) {
            // half-slab u = chunks [4 (u & 1), + 4) of slab u / 2 <- packed words [4 u, 4 u + 4) of the query
            uint4 x = make_uint4(0u, 0u, 0u, 0u);
            if (keep) x = *reinterpret_cast<const uint4*>(p.queries + (size_t)qg * W + (size_t)u * 16);
            const uint32_t base = row_saddr + (uint32_t)(u >> 1) * kMmaSlabBytes;
            const uint32_t c0 = (uint32_t)(u & 1) * 4u;
            emit_chunk_f4(x.x, keep, base + (((c0 + 0u) ^ r7) << 4));
            emit_chunk_f4(x.y, keep, base + (((c0 + 1u) ^ r7) << 4));
            emit_chunk_f4(x.z, keep, base + (((c0 + 2u) ^ r7) << 4));
            emit_chunk_f4(x.w, keep, base + (((c0 + 3u) ^ r7) << 4));
        }
        fence_proxy_async_smem();
    } else {
        const int ql = threadIdx.x % kMmaQT;        // 4 threads per query, each a quarter of the half-slabs
        const int part = threadIdx.x / kMmaQT;
        const int qg = q0 + ql;
        uint32_t pk[W / 4];
        uint32_t keep = 0;
#pragma unroll
        for (int i = 0; i < W / 4; ++i) pk[i] = 0;
        if (qg < p.n_queries) {
            const uint4* src = reinterpret_cast<const uint4*>(p.queries + (size_t)qg * W);
#pragma unroll
            for (int i = 0; i < CHUNKS; ++i) {
                const uint4 v = src[i];
                pk[4 * i] = v.x, pk[4 * i + 1] = v.y, pk[4 * i + 2] = v.z, pk[4 * i + 3] = v.w;
            }
            keep = 0xFFFFFFFFu;
        }
        const uint32_t row_saddr = smem_u32(sA) + (uint32_t)ql * 128u;
        const uint32_t r7 = (uint32_t)(ql & 7);
        for (int u = part; u < SLABS * 2; u += kMmaThreads / kMmaQT) {
            const int j = u >> 1;
            if (u & 1) emit_half_slab_rt<W, 1, XS>(j, pk, keep, row_saddr + (uint32_t)j * kMmaSlabBytes, r7);
            else emit_half_slab_rt<W, 0, XS>(j, pk, keep, row_saddr + (uint32_t)j * kMmaSlabBytes, r7);
        }
        fence_proxy_async_smem();
    }
    if constexpr (F4) {
        // unit block scales (UE8M0 0x7F = 2^0) in every byte of the 48 scale-factor columns of all 128 lanes:
        // whatever bytes an MMA reads for its A and B scale vectors, it reads 1.0
        __syncthreads();                               // s_tmem is written
        if (warp >= 4 && warp < 8) {
            const uint32_t sf = *s_tmem + ((uint32_t)((warp & 3) * 32) << 16) + 2u * NT;
            if constexpr (F4X) {
                // the query operand's scales: plane p = 2^(7 - p) in every byte of columns [4 p, 4 p + 4) (a scale
                // address must be 4-column aligned; both 32-element blocks of a K = 64 step lie in one plane)
                uint32_t sv[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) sv[c] = (uint32_t)(127 + 7 - c / 4) * 0x01010101u;
                tmem_st16(sf, sv);
#pragma unroll
                for (int c = 0; c < 16; ++c) sv[c] = (uint32_t)(127 + 3 - c / 4) * 0x01010101u;
                tmem_st16(sf + 16u, sv);
            } else {
                tmem_fill16(sf, 0x7F7F7F7Fu);
                tmem_fill16(sf + 16u, 0x7F7F7F7Fu);
            }
            tmem_fill16(sf + 32u, 0x7F7F7F7Fu);        // the row operand: 1.0
            tmem_st_wait();
        }
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    if (tracing && threadIdx.x == 0) p.trace[1] = global_ns();     // prologue done (TMEM, barriers, A operand)

    if (warp == 0) {
        // ===================== TMA producer: packed row boxes + their mask words =====================
        if (lane == 0) {
            const uint64_t policy = p.evict_first ? l2_policy_evict_first() : l2_policy_evict_last();
            for (uint32_t i = 0; i < my_tiles; ++i) {
                const int s = i % kMmaPkStages;
                stamp(i, 0);
                if (i >= kMmaPkStages) mbar_wait(&pk_empty[s], ((i / kMmaPkStages) - 1) & 1);
                stamp(i, 1);
                const uint32_t t = (tile_begin + i) * p.tile_stride;
                if constexpr (F4) {
                    // (28-byte mask rows cannot ride a bulk copy: the expanders read their mask bits from global)
                    mbar_arrive_expect_tx(&pk_full[s], Cfg::PK_BYTES);
                    tma_load_2d(sPk + (size_t)s * Cfg::PK_STRIDE, &tmap, 0, (int)(t * NT + rank * ROWS), &pk_full[s], policy);
                } else {
                    const uint32_t mask_bytes = MW * 4;
                    mbar_arrive_expect_tx(&pk_full[s], Cfg::PK_BYTES + mask_bytes + (has_filter ? mask_bytes : 0));
                    tma_load_2d(sPk + (size_t)s * Cfg::PK_STRIDE, &tmap, 0, (int)(t * NT + rank * ROWS), &pk_full[s], policy);
                    uint32_t* m = sMask + s * 2 * MW;
                    bulk_g2s(m, p.live + (size_t)t * MW, mask_bytes, &pk_full[s]);
                    if (has_filter) bulk_g2s(m + MW, p.filter + (size_t)t * MW, mask_bytes, &pk_full[s]);
                }
            }
        }
        __syncwarp();   // reconverge: the teardown barrier / cluster barrier is warp-aligned
    } else if (warp == 1) {
        // ===================== MMA issuer: the leader CTA's warp 1, one elected lane per instruction =====================
        if (rank == 0) {
            constexpr uint32_t idesc = F4 ? umma_idesc_mxf4(kMmaQT * CG, NT) : umma_idesc_i8(kMmaQT * CG, NT);
            const uint32_t sfa = tmem_base + 2u * NT, sfb = sfa + 32u;
            const uint32_t a_saddr = smem_u32(sA), b_saddr = smem_u32(sB);
            // One thread issues every MMA of the CTA pair, and its bookkeeping between two slabs -- not the tensor pipe --
            // set the tile time (scripts/mma_trace_roles.py: 384 ns per slab of four 61 ns MMAs): ring slot and parity
            // are carried instead of divided out, descriptors are one add from a base, the slab loop is unrolled.
            const uint64_t ad0 = umma_desc_sw128(a_saddr), bd0 = umma_desc_sw128(b_saddr);
            uint32_t st = 0, par = 0;   // ring slot of the next slab, parity of its generation
            for (uint32_t i = 0; i < my_tiles; ++i) {
                const uint32_t buf = i & 1;
                if (lane == 0) stamp(i, 8);
                if (i >= 2) {
                    if constexpr (CG == 2) mbar_wait_cluster(&acc_empty[buf], ((i >> 1) - 1) & 1);
                    else mbar_wait(&acc_empty[buf], ((i >> 1) - 1) & 1);
                    tc_fence_after();
                }
                if (lane == 0) stamp(i, 9);
                const uint32_t d_tmem = tmem_base + buf * NT;
#pragma unroll
                for (int j = 0; j < SLABS; ++j) {
                    if constexpr (CG == 2) mbar_wait_cluster(&b_full[st], par); else mbar_wait(&b_full[st], par);
                    if (lane == 0) stamp(i, 10 + (j < 4 ? j : 3));
                    tc_fence_after();
                    const uint64_t ad = ad0 + (uint64_t)(j * (kMmaSlabBytes >> 4));
                    const uint64_t bd = bd0 + (uint64_t)(st * (uint32_t)(kMmaSlabBytes >> 4));
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {   // K = 32 int8 (64 e2m1) = 32 B per instruction: +2 in the address field
                        if constexpr (F4) {
                            // XORSUM: this K = 64 step is elements [64 (4 j + kk), + 64) = plane 64 (4 j + kk) / W: its scale column
                            const uint32_t sfa_k = F4X ? sfa + 4u * (uint32_t)((64 * (4 * j + kk)) / W) : sfa;
                            umma_mxf4_elect<CG>(d_tmem, ad + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2), idesc, (j | kk) != 0 ? 1u : 0u, sfa_k, sfb);
                            continue;
                        }
                        umma_i8_elect<CG>(d_tmem, ad + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2), idesc, (j | kk) != 0 ? 1u : 0u);
                        // XORSUM: the elements of plane 0 (bit 7 of every byte, the first W of the row) weigh
                        // 128 = 64 + 64: their k-steps are accumulated twice
                        if (XS && j * 128 + kk * 32 < W)
                            umma_i8_elect<CG>(d_tmem, ad + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2), idesc, 1u);
                    }
                    umma_commit_elect<CG>(&b_empty[st]);
                    if (++st == kMmaBStages) {
                        st = 0;
                        par ^= 1u;
                    }
                }
                umma_commit_elect<CG>(&acc_full[buf]);
                if (lane == 0) stamp(i, 14);
            }
        }
        __syncwarp();
    } else if ((warp >= 4 && warp < 8) || (EG == 2 && warp >= 12)) {
        // ===================== epilogue: one query per thread (and per group) =====================
        const int grp = (warp >= 12) ? 1 : 0;   // which 32-column chunks of every tile: ch % EG == grp
        const uint32_t lane_base = (uint32_t)(warp & 3) * 32u;
        const int ql = (int)lane_base + lane;
        const int qg = q0 + ql;
        const bool q_ok = qg < p.n_queries;
        const int k = p.k;
        // descending list in registers: slot 0 = k-th best (worst kept), slot k-1 = best; slots >= k stay 0.
        // Rows reach a thread in ascending order (tiles, chunks and columns ascend), so a candidate with
        // the score of a kept entry is always the later row: the list order needs score compares only.
        uint32_t ks[kMmaListCap], kr[kMmaListCap];
#pragma unroll
        for (int e = 0; e < kMmaListCap; ++e) {
            ks[e] = (!COLLECT && e < k) ? kSentinelScore : 0u;
            kr[e] = 0xFFFFFFFFu;
        }
        uint32_t thr = kSentinelScore;      // admit rows with score < thr
        if constexpr (COLLECT) {
            if (q_ok) {
                const uint32_t v = p.collect_thr[qg];
                thr = (v == kSentinelScore) ? v : v + 1;
            }
        }
        constexpr int kNever = 1 << 30;
        auto to_dot = [&](uint32_t th) -> AccT {   // score < th  <=>  dot > K - 2 th
            if (!q_ok) return (AccT)kNever;
            return (AccT)((th > (uint32_t)K) ? -kNever : K - 2 * (int)th);   // (K = 8 W or 255 W: the largest score)
        };
        AccT thrd = to_dot(thr);
        const uint32_t n_mask_words = (p.n_rows + 31u) >> 5;
        // (tau = the smallest k-th best of any list: a bound on the k-th best row only.  When the caller wants the
        // k_bound best of the lists' union it would prune rows ranked k + 1 .. k_bound: the histogram bound alone
        // is shared then, with k_bound as its target count)
        const bool use_tau = !COLLECT && q_ok && p.k_bound <= p.k;
        uint32_t* tau_ptr = p.tau + (q_ok ? qg : 0);
        uint32_t tau_pref = kSentinelScore;
        bool dirty = false;
        const bool use_hist = !COLLECT && q_ok && p.hist != nullptr;
        uint32_t* hq = p.hist + (size_t)(q_ok ? qg : 0) * kHistWords;
        const uint32_t acc_empty_remote = (CG == 2) ? mapa_shared(smem_u32(&acc_empty[0]), 0) : 0u;

        for (uint32_t i = 0; i < my_tiles; ++i) {
            const uint32_t buf = i & 1;
            const uint32_t t = (tile_begin + i) * p.tile_stride;
            if (tracing && warp == 4 && lane == 0 && i < 256) p.trace[16 + 256 + i] = 0;
            if (use_tau) {
                const uint32_t gth = (tau_pref == kSentinelScore) ? tau_pref : tau_pref + 1;
                if (gth < thr) {
                    thr = gth;
                    thrd = to_dot(thr);
                }
                tau_pref = __ldcg(tau_ptr);
            }
            // device-wide bound, kept fresh by the helper warps (below): one shared-memory read per tile
            if (use_hist) {
                const uint32_t hb = *reinterpret_cast<volatile uint32_t*>(&s_bound[ql]);
                if (hb < thr) {
                    thr = hb;
                    thrd = to_dot(thr);
                }
            }
            // this tile's eligibility bits (live & filter), one word per lane: fetched before the wait so
            // that the rare path below never touches memory
            uint32_t vw = 0;
            if (lane < MW && (!F4 || t * (uint32_t)MW + (uint32_t)lane < n_mask_words)) {   // (224-row tiles can end past the padded mask)
                vw = __ldg(p.live + (size_t)t * MW + lane);
                if (has_filter) vw &= __ldg(p.filter + (size_t)t * MW + lane);
            }
            if (lane == 0 && (warp == 4 || warp == 12)) stamp(i, (warp == 4 ? 32 : 48) + 0);
            mbar_wait(&acc_full[buf], (i >> 1) & 1);
            tc_fence_after();
            if (lane == 0 && (warp == 4 || warp == 12)) stamp(i, (warp == 4 ? 32 : 48) + 1);
            if (tracing && warp == 4 && lane == 0 && i < 3) p.trace[i == 0 ? 4 : 12 + i] = global_ns();   // accumulator 0 / 1 / 2 ready
#pragma unroll 1
            for (int ch = grp; ch < NT / 32; ch += EG) {
                uint32_t vu[32];
                __syncwarp();   // tcgen05.ld is warp-aligned: reconverge after a divergent insert
                VL_DCHECK(buf * NT + (uint32_t)ch * 32u + 32u <= (uint32_t)Cfg::TMEM_COLS, 202);
                tmem_ld32(tmem_base + (lane_base << 16) + buf * NT + (uint32_t)ch * 32u, vu);
                tmem_ld_wait();
                if (lane == 0 && (warp == 4 || warp == 12)) stamp(i, (warp == 4 ? 32 : 48) + 2 + (ch / EG));
                if (ch + EG >= NT / 32) {
                    // this group's last chunk of the accumulator buffer is in registers now: hand it back to the MMA thread
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (CG == 2) mbar_arrive_cluster(acc_empty_remote + buf * 8u);
                        else mbar_arrive(&acc_empty[buf]);
                    }
                }
                AccT v[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    if constexpr (F4) v[c] = __uint_as_float(vu[c]);   // exact integers in fp32: compared as floats
                    else v[c] = (int)vu[c];
                }
                AccT m = v[0];
#pragma unroll
                for (int c = 1; c < 32; c += 2) m = max(m, max(v[c], (c + 1 < 32) ? v[c + 1] : v[c]));
                if (tracing && warp == 4 && lane == 0 && i == 0 && ch < 8) p.trace[5 + ch] = global_ns();   // tile 0: chunk ch loaded
                if (!__any_sync(kFullMask, m > thrd)) continue;
                if (tracing && warp == 4 && lane == 0 && i < 256) p.trace[16 + 256 + i] += 1u << 16;   // chunks that took the rare path
                // ---- rare: some thread of the warp has a candidate among these 32 rows.  Candidates
                // are taken in ascending row order (so an equal score never displaces an earlier row);
                // every thread works on its own query, the warp only votes on "anything left". ----
                uint32_t hm = 0;
#pragma unroll
                for (int c = 0; c < 32; ++c) hm |= (v[c] > thrd) ? (1u << c) : 0u;
                // a 32-column chunk is one word of the tile's eligibility mask: ineligible rows (tombstone, filter,
                // past the end -- they were expanded to zeros and score K / 2) leave the hit mask here, once
                hm &= __shfl_sync(kFullMask, vw, ch);
                while (__any_sync(kFullMask, hm != 0u)) {
                    if (tracing && warp == 4 && lane == 0 && i < 256) p.trace[16 + 256 + i] += 1u;   // insert-loop iterations
                    const bool mine = hm != 0u;
                    const int pc = mine ? __ffs(hm) - 1 : 0;
                    hm &= hm - 1u;
                    // v[pc] without dynamic register indexing: a 5-level select tree
                    AccT s16[16], s8[8], s4[4], s2[2];
#pragma unroll
                    for (int e = 0; e < 16; ++e) s16[e] = (pc & 16) ? v[e + 16] : v[e];
#pragma unroll
                    for (int e = 0; e < 8; ++e) s8[e] = (pc & 8) ? s16[e + 8] : s16[e];
#pragma unroll
                    for (int e = 0; e < 4; ++e) s4[e] = (pc & 4) ? s8[e + 4] : s8[e];
#pragma unroll
                    for (int e = 0; e < 2; ++e) s2[e] = (pc & 2) ? s4[e + 2] : s4[e];
                    const AccT pv = (pc & 1) ? s2[1] : s2[0];
                    const uint32_t rit = (uint32_t)ch * 32u + (uint32_t)pc;      // row in tile
                    VL_DCHECK(rit < (uint32_t)NT && (rit >> 5) < (uint32_t)MW, 203);
                    if (mine && pv > thrd) {
                        const uint32_t score = (uint32_t)(K - (int)pv) >> 1;
                        const uint32_t lrow = t * (uint32_t)NT + rit;
                        const uint64_t key = ((uint64_t)score << 32) | lrow;
                        VL_DCHECK((int)pv >= -K && (int)pv <= K && ((K - (int)pv) & 1) == 0 && (AccT)(int)pv == pv && lrow < p.n_rows, 204);   // a +-1 dot product of K terms
                        if constexpr (COLLECT) {
                            const uint32_t slot = atomicAdd(p.collect_count + qg, 1u);
                            if (slot < p.collect_cap) p.collect_keys[(size_t)qg * p.collect_cap + slot] = key;
                        } else {
                            // drop slot 0 (the worst), slide the candidate into the descending list
#pragma unroll
                            for (int e = 0; e < kMmaListCap; ++e) {
                                const uint32_t ns = (e + 1 < kMmaListCap) ? ks[e + 1] : 0u;
                                const uint32_t nr = (e + 1 < kMmaListCap) ? kr[e + 1] : 0xFFFFFFFFu;
                                const bool up = ns > score, here = ks[e] > score;
                                kr[e] = up ? nr : (here ? lrow : kr[e]);
                                ks[e] = up ? ns : (here ? score : ks[e]);
                            }
                            thr = min(thr, ks[0]);                       // never looser than a shared bound
                            thrd = to_dot(thr);
                            dirty = true;
                            if (use_hist) {
                                const uint32_t b = min(score >> p.hist_shift, (uint32_t)(kHistFine - 1));
                                atomicAdd(hq + b, 1u);
                                atomicAdd(hq + kHistFine + (b >> 5), 1u);
                            }
                        }
                    }
                }
            }
            if (!COLLECT && dirty) {
                dirty = false;
                if (use_tau && thr != kSentinelScore && thr < tau_pref) atomicMin(tau_ptr, thr);
            }
            if (tracing && warp == 4 && lane == 0 && i < 256) p.trace[16 + i] = global_ns();
            if (lane == 0 && (warp == 4 || warp == 12)) stamp(i, (warp == 4 ? 32 : 48) + 8);
        }
        __syncwarp();
        if (tracing && warp == 4 && lane == 0) p.trace[2] = global_ns();   // last tile's epilogue done
        if (lane == 0) atomicAdd(s_done, 1u);
        if constexpr (!COLLECT) {
            if (q_ok) {
                const size_t o = ((size_t)(split * EG + grp) * p.n_queries + qg) * k;
#pragma unroll
                for (int e = 0; e < kMmaListCap; ++e)
                    if (e < k) {
                        const bool none = kr[e] == 0xFFFFFFFFu;
                        p.part_scores[o + (k - 1 - e)] = none ? kSentinelScore : ks[e];
                        p.part_rows[o + (k - 1 - e)] = none ? kSentinelRow : p.base_row + kr[e];
                    }
            }
        }
    } else if (warp == 2 || warp == 3) {
        // ===================== helpers: device-wide admission bounds =====================
        // Every row that entered ANY CTA's list of a query was counted in that query's histogram
        // (fine[1024] | coarse[32], ScanParams::hist), so the score x at which the running count reaches
        // k caps the final k-th best.  A list's own k-th best only reflects its own row split -- with ~20
        // statistically alike splits per query that leaves ~half of all 32-row chunks with a candidate;
        // the histogram reflects all splits and makes the insert path rare.  Two otherwise idle warps poll
        // it (two L2 round trips per query) and publish bound + 1 to shared memory, off the epilogue's path.
        if (!COLLECT && p.hist != nullptr) {
            const int h = (warp - 2) * 32 + lane;
            const int k = p.k_bound > p.k ? p.k_bound : p.k;   // the rank the published bound stands for
            // the bound tightens like 1 / (rows seen): poll at geometrically growing intervals (8 us .. 64 us).
            // Measured on one box (cfg2 scan): first interval 1 / 8 / 16 us -> 0.4455 / 0.4385 / 0.4361 ms, a
            // fixed 3 us for the first 32 polls -> 0.452 ms: polling more often costs more L2 traffic than
            // the fresher bound saves.
            uint32_t nap = p.poll_ns ? p.poll_ns : 8000u, polls = 0;
            const uint32_t steady_polls = p.poll_steady ? p.poll_steady : 1u;
            while (*reinterpret_cast<volatile uint32_t*>(s_done) < 4u * EG) {
                for (int qq = h; qq < kMmaQT; qq += 64) {
                    const int qg = q0 + qq;
                    if (qg >= p.n_queries) continue;
                    VL_DCHECK(qq < kMmaQT, 206);
                    const uint32_t* hq = p.hist + (size_t)qg * kHistWords;
                    const uint4* coarse = reinterpret_cast<const uint4*>(hq + kHistFine);
                    uint32_t cum = 0, before = 0;
                    int cb = -1;
#pragma unroll
                    for (int j = 0; j < kHistCoarse / 4; ++j) {
                        const uint4 c4 = __ldcg(coarse + j);
                        const uint32_t cw[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (cb < 0 && cum + cw[e] >= (uint32_t)k) {
                                cb = j * 4 + e;
                                before = cum;
                            }
                            cum += cw[e];
                        }
                    }
                    if (cb < 0) continue;
                    const uint4* fine = reinterpret_cast<const uint4*>(hq + cb * 32);
                    uint32_t cf = before;
                    int x = -1;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint4 f4 = __ldcg(fine + j);
                        const uint32_t fw[4] = {f4.x, f4.y, f4.z, f4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (x < 0 && cf + fw[e] >= (uint32_t)k) x = cb * 32 + j * 4 + e;
                            cf += fw[e];
                        }
                    }
                    // (the fine bins can lag the coarse ones: x < 0 then; the last bucket is open-ended)
                    if (x >= 0 && x < kHistFine - 1) s_bound[qq] = ((uint32_t)x + 1u) << p.hist_shift;   // k-th best < this
                }
                // sleep in slices so that the end of the scan is noticed within ~2 us
                for (uint32_t slept = 0; slept < nap && *reinterpret_cast<volatile uint32_t*>(s_done) < 4u * EG; slept += 2000)
                    __nanosleep(min(nap - slept, 2000u));
                if (++polls > steady_polls) nap = min(nap + nap / 2, 64000u);
            }
        }
        __syncwarp();
    } else if (warp >= 8 && warp < 8 + EXPW) {
        // ===================== expanders: packed rows -> int8 / e2m1 K slabs =====================
        const int ew = warp - 8;
        const int rgrp = ew & 3;            // 32 rows of the CTA's 128
        const int half = ew >> 2;           // int8: chunks [4 half, 4 half + 4) of every slab (e2m1: one thread does all 8)
        const int r = rgrp * 32 + lane;
        const bool r_ok = r < ROWS;         // (F4 pairs: 112 rows per CTA, the last 16 row slots idle)
        const uint32_t n_mask_words = (p.n_rows + 31u) >> 5;
        // TMA swizzle of the packed box: logical chunk j of row r at chunk j ^ rot(r)
        constexpr int LANES_PER_128B = (W >= 128) ? 1 : 128 / W;
        constexpr int ROT_MASK = (CHUNKS < 8 ? CHUNKS : 8) - 1;
        const uint32_t rot = (uint32_t)((r / LANES_PER_128B) & ROT_MASK);
        const uint32_t pk_saddr = smem_u32(sPk) + (uint32_t)(r_ok ? r : 0) * W;
        const uint32_t b_saddr = smem_u32(sB) + (uint32_t)r * 128u;
        const uint32_t r7 = (uint32_t)(r & 7);
        const uint32_t b_full_remote = (CG == 2) ? mapa_shared(smem_u32(&b_full[0]), 0) : 0u;
        const int mword = (int)rank * 4 + rgrp;   // this row group's word of the tile's mask
        uint32_t g = 0;
        for (uint32_t i = 0; i < my_tiles; ++i) {
            const int s = i % kMmaPkStages;
            uint32_t keep;
            if constexpr (F4) {
                // this row's eligibility bit straight from global memory, requested before the wait
                const uint32_t bit = (uint32_t)rank * ROWS + (uint32_t)r;
                const uint32_t widx = (tile_begin + i) * p.tile_stride * (uint32_t)MW + (bit >> 5);
                uint32_t wv = 0;
                if (r_ok && widx < n_mask_words) {
                    wv = __ldg(p.live + widx);
                    if (has_filter) wv &= __ldg(p.filter + widx);
                }
                if (warp == 8 && lane == 0) stamp(i, 16);
                mbar_wait(&pk_full[s], (i / kMmaPkStages) & 1);
                if (warp == 8 && lane == 0) stamp(i, 17);
                keep = ((wv >> (bit & 31u)) & 1u) ? 0xFFFFFFFFu : 0u;
            } else {
                mbar_wait(&pk_full[s], (i / kMmaPkStages) & 1);
                const uint32_t* m = sMask + s * 2 * MW;
                uint32_t wv = m[mword];
                if (has_filter) wv &= m[MW + mword];
                keep = ((wv >> lane) & 1u) ? 0xFFFFFFFFu : 0u;
            }
            if constexpr (F4) {
                // Packed chunks are loaded slab by slab (8 live registers instead of the whole row).  HAMMING: slab j <-
                // the row's packed 16-byte chunks 2 j and 2 j + 1 (8 words -> 8 output chunks).  XORSUM (plane-major K):
                // slab j = bit planes [j PPS, + PPS) of ALL the row's bytes: every 32-byte block bb (two packed chunks)
                // is read once per slab and gives one output chunk per plane.  The box goes back to the TMA producer
                // with the last slab's loads -- and only once they have RETURNED: an mbarrier arrive does not wait for
                // the thread's own outstanding shared-memory loads (SASS: LDS.128, then SYNCS.ARRIVE with no scoreboard
                // wait), and with the producer one wake-up away the next box was seen to land before slow wavefronts of
                // those loads had read the old one: rows of tile i carried chunks of tile i + 2 (scripts/mma_stress.py).
                // Making the barrier address depend on the loaded words makes the arrive wait for them.
#pragma unroll 1
                for (int j = 0; j < SLABS; ++j, ++g) {
                    const uint32_t st = g % kMmaBStages;
                    if (g >= kMmaBStages) mbar_wait(&b_empty[st], ((g / kMmaBStages) - 1) & 1);
                    if (warp == 8 && lane == 0) stamp(i, 18 + 2 * (j < 4 ? j : 3));
                    const uint32_t row_saddr = b_saddr + st * kMmaSlabBytes;
                    VL_DCHECK(CG == 2 || !mbar_test_wait(&b_full[st], (g / kMmaBStages) & 1), 207);
                    const uint32_t src = pk_saddr + (uint32_t)s * Cfg::PK_STRIDE;
                    uint32_t h = 0;
                    if constexpr (F4X) {
                        constexpr int CPP = W / 32, PPS = 8 / CPP;   // chunks per plane, planes per slab
#pragma unroll
                        for (int bb = 0; bb < CPP; ++bb) {
                            const uint4 x0 = lds128(src + ((((uint32_t)(2 * bb)) ^ rot) << 4));
                            const uint4 x1 = lds128(src + ((((uint32_t)(2 * bb + 1)) ^ rot) << 4));
                            h ^= x0.x ^ x0.y ^ x0.z ^ x0.w ^ x1.x ^ x1.y ^ x1.z ^ x1.w;
                            if (r_ok) {
#pragma unroll
                                for (int pp = 0; pp < PPS; ++pp)
                                    emit_plane_chunk_f4(x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w, 7u - (uint32_t)(j * PPS + pp), keep,
                                                        row_saddr + ((((uint32_t)(pp * CPP + bb)) ^ r7) << 4));
                            }
                        }
                    } else {
                        const uint4 x0 = lds128(src + ((((uint32_t)(2 * j)) ^ rot) << 4));
                        const uint4 x1 = lds128(src + ((((uint32_t)(2 * j + 1)) ^ rot) << 4));
                        h = x0.x ^ x0.y ^ x0.z ^ x0.w ^ x1.x ^ x1.y ^ x1.z ^ x1.w;
                        if (r_ok) {
                            emit_chunk_f4(x0.x, keep, row_saddr + ((0u ^ r7) << 4));
                            emit_chunk_f4(x0.y, keep, row_saddr + ((1u ^ r7) << 4));
                            emit_chunk_f4(x0.z, keep, row_saddr + ((2u ^ r7) << 4));
                            emit_chunk_f4(x0.w, keep, row_saddr + ((3u ^ r7) << 4));
                            emit_chunk_f4(x1.x, keep, row_saddr + ((4u ^ r7) << 4));
                            emit_chunk_f4(x1.y, keep, row_saddr + ((5u ^ r7) << 4));
                            emit_chunk_f4(x1.z, keep, row_saddr + ((6u ^ r7) << 4));
                            emit_chunk_f4(x1.w, keep, row_saddr + ((7u ^ r7) << 4));
                        }
                    }
                    if (j == SLABS - 1) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&pk_empty[s] + (h & (p.one - 1u)));   // (the offset is 0 at run time)
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (warp == 8 && lane == 0) stamp(i, 19 + 2 * (j < 4 ? j : 3));
                    if (lane == 0) {
                        if constexpr (CG == 2) mbar_arrive_cluster(b_full_remote + st * 8u);
                        else mbar_arrive(&b_full[st]);
                    }
                }
                continue;
            }
            // int8: the whole packed row goes to registers (every slab needs one bit plane of ALL its bytes) and
            // the packed box is handed back to the TMA producer before the expansion starts
            uint32_t pk[W / 4];
            VL_DCHECK(rot < (uint32_t)CHUNKS && r < kMmaRows, 205);
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) {
                const uint4 x = lds128(pk_saddr + (uint32_t)s * Cfg::PK_STRIDE + (((uint32_t)c ^ rot) << 4));
                pk[4 * c] = x.x, pk[4 * c + 1] = x.y, pk[4 * c + 2] = x.z, pk[4 * c + 3] = x.w;
            }
            // (hand-back only once the loads have returned: see the e2m1 branch above)
            uint32_t h = 0;
#pragma unroll
            for (int c = 0; c < W / 4; ++c) h ^= pk[c];
            h &= p.one - 1u;                       // == 0, but only at run time: the arrive below depends on every load
            __syncwarp();
            if (lane == 0) mbar_arrive(&pk_empty[s] + h);
#pragma unroll 1
            for (int j = 0; j < SLABS; ++j, ++g) {
                const uint32_t st = g % kMmaBStages;
                if (g >= kMmaBStages) mbar_wait(&b_empty[st], ((g / kMmaBStages) - 1) & 1);
                const uint32_t row_saddr = b_saddr + st * kMmaSlabBytes;
                // ring protocol: nobody may have declared THIS generation of the slab full before it is written
                VL_DCHECK(CG == 2 || !mbar_test_wait(&b_full[st], (g / kMmaBStages) & 1), 207);
                if (half == 0)
                    emit_half_slab_rows<W, 0>(pk, j, keep, row_saddr, r7);
                else
                    emit_half_slab_rows<W, 1>(pk, j, keep, row_saddr, r7);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (CG == 2) mbar_arrive_cluster(b_full_remote + st * 8u);
                    else mbar_arrive(&b_full[st]);
                }
            }
        }
    }

    // ---- teardown: every tcgen05 operation has completed once the epilogues are through ----
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<CG>(tmem_base, (uint32_t)Cfg::TMEM_COLS);
    }
    if (tracing && threadIdx.x == 0) p.trace[3] = global_ns();
    if (p.trace != nullptr && threadIdx.x == 0 && cta_lin < (uint32_t)kTraceCtas) p.trace[16 + 512 + kTraceCtas + cta_lin] = global_ns();
}

}  // namespace vl
