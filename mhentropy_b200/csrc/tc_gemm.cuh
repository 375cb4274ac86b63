// Blackwell tensor-core GEMM for the coupling MLPs: tcgen05.mma (bf16 operands, fp32 accumulate in TMEM),
// operands staged by TMA (cp.async.bulk.tensor, 128B swizzle) through an mbarrier ring, warp-specialised
// (1 TMA warp, 1 MMA warp, 4 epilogue warps reading the accumulator back with tcgen05.ld).
//
// Precision: fp32 values are carried as SPLIT-bf16 PLANES  x = hi + lo  (hi = bf16(x), lo = bf16(x - hi)).
// A product uses the three tensor-core passes hi*hi + hi*lo + lo*hi accumulated in fp32 ("bf16x3"), which
// keeps ~16 mantissa bits per product — what the parity tolerances need — at the
// bf16 MMA rate.  NPAIR = 1 runs plain bf16.
//
// Operand tensors are 4-D bf16 arrays  [batch][plane][rows][cols]  (cols contiguous) described by a TMA
// tensor map.  An operand is K-major when cols index K (rows index M or N) and MN-major when cols index
// M/N (rows index K); MN-major lets dgrad (dy * W) and wgrad (dy^T * x) read weights and activations in
// their natural layouts, with no transposed copies.  Shared-memory tiles are the canonical UMMA SW128
// layouts: K-major  [rows][64 cols] 128 B per row, 8-row groups 1024 B apart;
//          MN-major boxes of [64 k-rows][64 cols], 8-k groups 1024 B apart, 64-col groups one box apart.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <type_traits>
#include "common.cuh"

namespace mhe {
namespace tc {

constexpr int BM = 128;       // UMMA M (cta_group::1)
constexpr int BK = 64;        // bf16 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kThreads = 192; // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue
constexpr int kThreadsP = 320; // persistent kernel: eight epilogue warps (2-9), two per TMEM lane quadrant

// ---- host: tensor maps ----------------------------------------------------------------------------
struct PlaneTensor {          // [batches][planes][rows][cols] bf16, element strides
    const __nv_bfloat16* base = nullptr;
    int cols = 0, rows = 0, planes = 1, batches = 1;
    long row_pitch = 0, plane_stride = 0, batch_stride = 0;
};

int make_tensor_map(CUtensorMap* map, const PlaneTensor& t, int box_rows);
// 16-bit split planes of fp32 data: x = hi + lo.  f16 = true stores IEEE half (11-bit significands, ~22 bits kept; for
// range-safe data: weights, activations), false stores bfloat16 (8-bit significands, ~16 bits kept; full fp32 range: gradients).
int split_planes(const float* src, long src_ld, long src_batch, int rows, int cols, const float* colscale, __nv_bfloat16* dst, int rows_p,
                 int cols_p, int planes, int batches, bool f16, cudaStream_t stream);

#if defined(__CUDACC__)
template <bool F16>
__device__ __forceinline__ uint16_t to16(float v) {
    if (F16) return __half_as_ushort(__float2half_rn(v));
    return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
template <bool F16>
__device__ __forceinline__ float from16(uint16_t u) {
    if (F16) return __half2float(__ushort_as_half(u));
    return __bfloat162float(__ushort_as_bfloat16(u));
}
#endif

// Epilogue mode kTile8 (optional member of an Epi): the warp's 32 x 32 accumulator chunk is transposed through shared memory and handed
// to Epi::tile8(batch, split, row, col, v[8]) as 8 consecutive columns of one row per lane, 4 lanes per row, 8 rows per pass.  Epilogues
// that read / write ROW-MAJOR 16-bit planes then touch 64 contiguous bytes per row and instruction (8 memory transactions per warp
// instruction) instead of 32 rows x 16 bytes, 1 KB apart (32 transactions): the plane-writing GEMMs of the long-batch flow path are
// bound by exactly that.
template <class E, class = void> struct epi_tile8 : std::false_type {};
template <class E> struct epi_tile8<E, std::void_t<decltype(E::kTile8)>> : std::bool_constant<E::kTile8> {};
// a tile8 epilogue names the register blob of its prefetched global operand: E::Pre, filled by E::pre8(batch, row, col, Pre&)
struct EpiNoPre {};
// ... and may resolve the row-dependent part of that operand's address ONCE per tile (E::RowRef pre_row(batch, row), then
// pre8r(RowRef, col, Pre&) per chunk): the rows a lane prefetches for are the same for every column chunk of a tile, and e.g. the
// conditioning term's `row % images` costs ~35 instructions per evaluation (18 % of the plane-writing epilogue's instructions, ncu)
template <class E, class = void> struct epi_pre_row : std::false_type {};
template <class E> struct epi_pre_row<E, std::void_t<typename E::RowRef>> : std::true_type {};
template <class E, class = void> struct epi_row_ref { typedef int type; };
template <class E> struct epi_row_ref<E, std::void_t<typename E::RowRef>> { typedef typename E::RowRef type; };
template <class E, class = void> struct epi_pre_type { typedef EpiNoPre type; };
template <class E> struct epi_pre_type<E, std::void_t<typename E::Pre>> { typedef typename E::Pre type; };

#if defined(__CUDACC__)
// 8 fp32 -> 8 hi + 8 lo 16-bit values packed as two uint4 (packed two-at-a-time conversions)
template <bool F16>
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float a = v[2 * j], b = v[2 * j + 1];
        if (F16) {
            const __half2 hh = __floats2half2_rn(a, b);
            const float2 hf = __half22float2(hh);
            const __half2 ll = __floats2half2_rn(a - hf.x, b - hf.y);
            h[j] = *reinterpret_cast<const uint32_t*>(&hh);
            l[j] = *reinterpret_cast<const uint32_t*>(&ll);
        } else {
            const __nv_bfloat162 hh = __floats2bfloat162_rn(a, b);
            const float2 hf = __bfloat1622float2(hh);
            const __nv_bfloat162 ll = __floats2bfloat162_rn(a - hf.x, b - hf.y);
            h[j] = *reinterpret_cast<const uint32_t*>(&hh);
            l[j] = *reinterpret_cast<const uint32_t*>(&ll);
        }
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
#endif

// ---- device primitives ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// try_wait may suspend the thread for an implementation-defined time before it re-checks; -DMHE_MBAR_TEST polls with test_wait
#ifdef MHE_MBAR_TEST
#define MHE_MBAR_POLL "mbarrier.test_wait"
#else
#define MHE_MBAR_POLL "mbarrier.try_wait"
#endif
// parity wait with a watchdog: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            MHE_MBAR_POLL ".parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// L2 prefetch of a tensor-map box (no shared memory, no barrier): issued a few k-blocks ahead of the loads proper, it turns the HBM
// latency of streamed operands into an L2 hit by the time the ring slot is free
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), SWIZZLE_128B, version 1
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;   // LayoutType::SWIZZLE_128B
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): bf16 x bf16 -> f32, M = 128
// operand formats: F16 = 0, BF16 = 1 (they may differ between A and B)
__host__ __device__ constexpr uint32_t instr_desc(int n, bool a_mn, bool b_mn, bool a_f16 = false, bool b_f16 = false) {
    return (1u << 4) /*c = f32*/ | ((a_f16 ? 0u : 1u) << 7) | ((b_f16 ? 0u : 1u) << 10) |
           ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// same, with both descriptors given as (low word, shared high word)
__device__ __forceinline__ void umma_bf16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct GemmShape {
    int M, N, K;       // logical extents (rows of A, rows/cols of B, contraction)
    int batches;       // grid.z = batches * ksplit
    int ksplit;
    int a_batch_mul;   // batch coordinate of A = batch * a_batch_mul (0 broadcasts one A to every batch)
    int b_batch_mul;
    int kfold = 1;     // consecutive batches contracted into ONE accumulator per CTA (batches % kfold == 0; the epilogue sees batch / kfold)
    int stages = 0;    // operand ring depth (0 = as deep as shared memory allows).  Short contractions with large outputs run better with a
                       // shallow ring: several CTAs share an SM and one's epilogue overlaps the others' loads and MMAs
    // K-concatenation (persistent kernel, ksplit == kfold == 1): after the K / 64 k-blocks of (A, B), `kb2` more k-blocks are read from a
    // SECOND operand pair (A2, B2: same majorness, planes and element type) into the same accumulator: C = A B^T + A2 B2^T without
    // materialising either product (the conditioning projection folded into the coupling GEMM, flow_tc.cu pass_fwd_rowcond)
    int kb2 = 0;
    int a2_batch_mul = 0;
    int b2_batch_mul = 0;
};

template <int BN, int NPL>
struct SmemPlan {
    static constexpr int kABytes = BM * BK * 2;           // one plane of A per stage (16 KB)
    static constexpr int kBBytes = BN * BK * 2;
    static constexpr int kStageBytes = NPL * (kABytes + kBBytes);
    static constexpr int kStages = (200 * 1024) / kStageBytes >= 4 ? 4 : (200 * 1024) / kStageBytes;
    static constexpr int kBytes = kStages * kStageBytes + 1024;   // + alignment slack
};

// C[batch] (M x N) = sum_pairs A_pa (M x K) * B_pb (K x N).  Epi::operator()(batch, split, row, col0, v[32], shape)
// is called by every epilogue thread once per 32-column chunk with its accumulator row.
template <int BN, bool A_MN, bool B_MN, int NPAIR, bool F16, class Epi>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, GemmShape g, Epi epi) {
    constexpr int NPL = NPAIR == 1 ? 1 : 2;
    using Plan = SmemPlan<BN, NPL>;
    constexpr int NSmax = Plan::kStages;
    const int NS = g.stages;          // launcher: 1 <= stages <= NSmax
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[NSmax], bar_empty[NSmax], bar_accum;
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float s_stage[(Epi::kStaged || epi_tile8<Epi>::value) ? 4 : 1][32][33];   // epilogue transpose buffers, one per epilogue warp

    pdl_launch_dependents();   // let the next kernel of the chain start its prologue; it blocks in pdl_wait()
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int batch = blockIdx.z / g.ksplit, split = blockIdx.z % g.ksplit;
    const int kb_total = (g.K + BK - 1) / BK;
    const int kb_per = (kb_total + g.ksplit - 1) / g.ksplit;
    const int kb_begin = split * kb_per;
    const int kb_end = min(kb_total, kb_begin + kb_per);
    const int nkb1 = max(kb_end - kb_begin, 0);          // k-blocks per batch
    const int nkb = nkb1 * g.kfold;                      // ... per accumulator

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
        mbar_init(smem_u32(&bar_accum), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    }
    if (warp == 1) {   // TMEM: BN fp32 accumulator columns (power of two >= 32)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_d = tmem_base_slot;
    pdl_wait();                // everything above is CTA-private; from here on we touch the predecessor's outputs

    auto stage_a = [&](int s, int p) { return smem0 + s * Plan::kStageBytes + p * Plan::kABytes; };
    auto stage_b = [&](int s, int p) { return smem0 + s * Plan::kStageBytes + NPL * Plan::kABytes + p * Plan::kBBytes; };

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % NS;
                const uint32_t ph = (i / NS) & 1;
                mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1);
                const uint32_t full = smem_u32(&bar_full[s]);
                mbar_expect_tx(full, Plan::kStageBytes);
                const int k0 = (kb_begin + i % nkb1) * BK;
                const int bsrc = batch * g.kfold + i / nkb1;         // source batch of this k-block
#pragma unroll
                for (int p = 0; p < NPL; ++p) {
                    if (!A_MN) tma_load_4d(stage_a(s, p), &mapA, full, k0, m0, p, bsrc * g.a_batch_mul);
                    else {
#pragma unroll
                        for (int j = 0; j < BM / 64; ++j) tma_load_4d(stage_a(s, p) + j * 8192, &mapA, full, m0 + 64 * j, k0, p, bsrc * g.a_batch_mul);
                    }
                    if (!B_MN) tma_load_4d(stage_b(s, p), &mapB, full, k0, n0, p, bsrc * g.b_batch_mul);
                    else {
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j) tma_load_4d(stage_b(s, p) + j * 8192, &mapB, full, n0 + 64 * j, k0, p, bsrc * g.b_batch_mul);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = instr_desc(BN, A_MN, B_MN, F16, F16);   // A and B must share one format (mixing traps)
            // Descriptors differ between MMAs only in the 14-bit start address: build the constant part once and
            // add 16-byte-unit offsets, so the single issuing thread spends a few instructions per MMA.
            constexpr uint32_t kHi = (uint32_t)((1024u >> 4) | (1u << 14) | (2u << 29));          // SBO | version | SWIZZLE_128B
            constexpr uint32_t kLoA = A_MN ? ((8192u >> 4) << 16) : ((16u >> 4) << 16);            // LBO field
            constexpr uint32_t kLoB = B_MN ? ((8192u >> 4) << 16) : ((16u >> 4) << 16);
            constexpr uint32_t kStepA = A_MN ? (2048u >> 4) : (32u >> 4);                          // per UMMA_K step
            constexpr uint32_t kStepB = B_MN ? (2048u >> 4) : (32u >> 4);
            uint32_t accumulate = 0;
            for (int i = 0; i < nkb; ++i) {
                const int s = i % NS;
                const uint32_t ph = (i / NS) & 1;
                mbar_wait(smem_u32(&bar_full[s]), ph);
                tcgen05_fence_after();
                const uint32_t a_lo0 = kLoA | (stage_a(s, 0) >> 4), b_lo0 = kLoB | (stage_b(s, 0) >> 4);
#pragma unroll
                for (int pr = 0; pr < NPAIR; ++pr) {
                    const uint32_t a_lo = a_lo0 + ((pr == 2) ? (uint32_t)(Plan::kABytes >> 4) : 0u);   // pairs: (hi,hi) (hi,lo) (lo,hi)
                    const uint32_t b_lo = b_lo0 + ((pr == 1) ? (uint32_t)(Plan::kBBytes >> 4) : 0u);
#pragma unroll
                    for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                        umma_bf16_lohi(tmem_d, a_lo + ks * kStepA, b_lo + ks * kStepB, kHi, idesc, accumulate);
                        accumulate = 1;
                    }
                }
                tcgen05_commit(smem_u32(&bar_empty[s]));   // frees the stage once these MMAs have read it
            }
            tcgen05_commit(smem_u32(&bar_accum));          // accumulator complete
        }
    } else {
        const int q = warp & 3;                            // TMEM lane quadrant this warp may access
        const int row = m0 + q * 32 + lane;
        if (nkb > 0) {
            mbar_wait(smem_u32(&bar_accum), 0);
            tcgen05_fence_after();
        }
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            float v[32];
            if (nkb > 0) tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + c0, v);
            else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
            if (n0 + c0 >= g.N) continue;                  // warp-uniform
            if constexpr (epi_tile8<Epi>::value) {
                float4* st4 = reinterpret_cast<float4*>(&s_stage[warp - 2][0][0]);      // (swizzled 16-byte chunks: see the persistent kernel)
#pragma unroll
                for (int j = 0; j < 8; ++j) st4[lane * 8 + (j ^ (lane & 7))] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rl = i * 8 + (lane >> 2), cl = (lane & 3) * 8, c4 = (lane & 3) * 2;
                    const float4 w0 = st4[rl * 8 + (c4 ^ (rl & 7))], w1 = st4[rl * 8 + ((c4 + 1) ^ (rl & 7))];
                    float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                    if (m0 + q * 32 + rl < g.M) {
                        typename Epi::Pre pre;
                        epi.pre8(batch, m0 + q * 32 + rl, n0 + c0 + cl, pre);
                        epi.tile8(batch, split, m0 + q * 32 + rl, n0 + c0 + cl, w, pre, g);
                    }
                }
                __syncwarp();
            } else
            if constexpr (Epi::kDirect) {
                // lane = accumulator row: each thread owns 32 consecutive columns (vector stores)
                if (row < g.M) epi(batch, split, row, n0 + c0, v, g);
            }
            if constexpr (Epi::kStaged) {
                // transpose through shared memory so that lanes run along the columns: coalesced read-modify-write /
                // atomics on the output row
                float (*st)[33] = s_stage[warp - 2];
#pragma unroll
                for (int j = 0; j < 32; ++j) st[lane][j] = v[j];
                __syncwarp();
                const int col = n0 + c0 + lane;
                const int rows = min(32, g.M - (m0 + q * 32));
                if constexpr (Epi::kRmw) {
                    // C += acc with every load issued before the first store: a naive `*p += v` loop serialises on
                    // possible aliasing and pays one memory round trip per row
                    long stride;
                    float* p0 = epi.rmw_ptr(batch, m0 + q * 32, col, stride);
                    if (p0 && rows > 0) {
                        if (epi.atomic) {
                            for (int rr = 0; rr < rows; ++rr) atomicAdd(p0 + rr * stride, st[rr][lane]);
                        } else if (epi.overwrite) {   // the destination is known to be zero: store, no read
#pragma unroll
                            for (int rr = 0; rr < 32; ++rr) if (rr < rows) p0[rr * stride] = st[rr][lane];
                        } else {
                            float old[32];
#pragma unroll
                            for (int rr = 0; rr < 32; ++rr) old[rr] = rr < rows ? p0[rr * stride] : 0.f;
#pragma unroll
                            for (int rr = 0; rr < 32; ++rr) if (rr < rows) p0[rr * stride] = old[rr] + st[rr][lane];
                        }
                    }
                } else {
                    if (col < g.N)
                        for (int rr = 0; rr < rows; ++rr) epi.elem(batch, split, m0 + q * 32 + rr, col, st[rr][lane], g);
                }
                __syncwarp();
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(BN) : "memory");
    }
}

const CUtensorMap* cached_map_f32(const float* base, int cols, int rows, int batches, long row_pitch, long batch_stride, int box_rows, int* status);

// ring depth of a launch: the caller's choice, overridden by MHE_TC_STAGES="label=n,label=n" (substring match on the launch label)
int stages_for(const char* what, int requested, int max_stages);

// Persistent variant: a CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... (N tiles fastest), the operand ring keeps running
// across tiles, and the accumulator is DOUBLE-BUFFERED in TMEM: the epilogue of tile i (TMEM -> registers -> global) overlaps the loads
// and MMAs of tile i + 1.  Short contractions with large outputs (K = 64 weight gradients, the MANO skinning / blend GEMMs) are
// dominated by the per-CTA fixed cost (barrier init, TMEM allocation, first TMA round trip, epilogue drain) in the one-tile kernel.
template <int BN, bool A_MN, bool B_MN, int NPAIR, bool F16, class Epi>
__global__ void __launch_bounds__(kThreadsP, 1)
tc_gemm_persistent_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                          const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB2, GemmShape g, Epi epi,
                          int tiles_n, int tiles_m, int tiles_total, int l2_prefetch_distance) {
    constexpr int NPL = NPAIR == 1 ? 1 : 2;
    using Plan = SmemPlan<BN, NPL>;
    constexpr int NSmax = Plan::kStages;
    const int NS = g.stages;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[NSmax], bar_empty[NSmax], bar_acc_full[2], bar_acc_empty[2];
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float s_stage[(Epi::kStaged || epi_tile8<Epi>::value) ? 8 : 1][32][33];     // one transpose buffer per epilogue warp

    pdl_launch_dependents();
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kb_total = (g.K + BK - 1) / BK;
    const int kb_per = (kb_total + g.ksplit - 1) / g.ksplit;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&bar_acc_full[b]), 1); mbar_init(smem_u32(&bar_acc_empty[b]), 8); }   // one elected arrival per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
        if (g.kb2 > 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA2) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB2) : "memory");
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(2 * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_d = tmem_base_slot;
    pdl_wait();

    auto stage_a = [&](int s, int p) { return smem0 + s * Plan::kStageBytes + p * Plan::kABytes; };
    auto stage_b = [&](int s, int p) { return smem0 + s * Plan::kStageBytes + NPL * Plan::kABytes + p * Plan::kBBytes; };
    // tile -> (n0, m0, batch, split, k-block range); every role walks the same sequence
    struct Tile { int n0, m0, batch, split, kb_begin, nkb1, nkb; };
    auto tile_of = [&](int t) {
        Tile T;
        T.n0 = (t % tiles_n) * BN;
        T.m0 = ((t / tiles_n) % tiles_m) * BM;
        const int z = t / (tiles_n * tiles_m);
        T.batch = z / g.ksplit; T.split = z % g.ksplit;
        T.kb_begin = T.split * kb_per;
        const int kb_end = min(kb_total, T.kb_begin + kb_per);
        T.nkb1 = max(kb_end - T.kb_begin, 0);
        T.nkb = T.nkb1 * g.kfold + g.kb2;
        return T;
    };

    if (warp == 0) {
        if (lane == 0) {
            uint32_t q = 0;
            // Optional (MHE_TC_L2_PREFETCH=<k-blocks>, default off): prefetch each k-block into L2 `pf_dist` k-blocks ahead of its ring
            // slot, across tile boundaries.  Measured on the long-batch flow GEMMs (tensor pipe 56 % active): 3-7 % SLOWER at distance
            // 4 and 8 - the operands are not what the MMAs wait for; kept as a diagnostic switch.
            const int pf_dist = l2_prefetch_distance;
            auto prefetch_kb = [&](const Tile& P, int i) {
                const int k0 = (P.kb_begin + i % P.nkb1) * BK;
                const int bsrc = P.batch * g.kfold + i / P.nkb1;
#pragma unroll
                for (int p = 0; p < NPL; ++p) {
                    if (!A_MN) tma_prefetch_4d(&mapA, k0, P.m0, p, bsrc * g.a_batch_mul);
                    else {
#pragma unroll
                        for (int j = 0; j < BM / 64; ++j) tma_prefetch_4d(&mapA, P.m0 + 64 * j, k0, p, bsrc * g.a_batch_mul);
                    }
                    if (!B_MN) tma_prefetch_4d(&mapB, k0, P.n0, p, bsrc * g.b_batch_mul);
                    else {
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j) tma_prefetch_4d(&mapB, P.n0 + 64 * j, k0, p, bsrc * g.b_batch_mul);
                    }
                }
            };
            if (pf_dist > 0 && (int)blockIdx.x < tiles_total) {      // the first tile's leading k-blocks
                const Tile T0 = tile_of(blockIdx.x);
                for (int i = 0; i < pf_dist && i < T0.nkb; ++i) prefetch_kb(T0, i);
            }
            for (int t = blockIdx.x; t < tiles_total; t += gridDim.x) {
                const Tile T = tile_of(t);
                const bool has_next = t + (int)gridDim.x < tiles_total;
                Tile Tn = T;
                if (pf_dist > 0 && has_next) Tn = tile_of(t + gridDim.x);
                for (int i = 0; i < T.nkb; ++i, ++q) {
                    if (pf_dist > 0) {
                        if (i + pf_dist < T.nkb) prefetch_kb(T, i + pf_dist);
                        else if (has_next && i + pf_dist - T.nkb < Tn.nkb) prefetch_kb(Tn, i + pf_dist - T.nkb);
                    }
                    const int s = q % NS;
                    mbar_wait(smem_u32(&bar_empty[s]), ((q / NS) & 1) ^ 1);
                    const uint32_t full = smem_u32(&bar_full[s]);
                    mbar_expect_tx(full, Plan::kStageBytes);
                    const int nmain = T.nkb1 * g.kfold;
                    const bool tail = i >= nmain;                         // k-blocks of the second operand pair (GemmShape::kb2)
                    const int k0 = tail ? (i - nmain) * BK : (T.kb_begin + i % T.nkb1) * BK;
                    const int bsrc = tail ? T.batch : T.batch * g.kfold + i / T.nkb1;
                    const CUtensorMap* mA = tail ? &mapA2 : &mapA;
                    const CUtensorMap* mB = tail ? &mapB2 : &mapB;
                    const int ba = bsrc * (tail ? g.a2_batch_mul : g.a_batch_mul), bb = bsrc * (tail ? g.b2_batch_mul : g.b_batch_mul);
#pragma unroll
                    for (int p = 0; p < NPL; ++p) {
                        if (!A_MN) tma_load_4d(stage_a(s, p), mA, full, k0, T.m0, p, ba);
                        else {
#pragma unroll
                            for (int j = 0; j < BM / 64; ++j) tma_load_4d(stage_a(s, p) + j * 8192, mA, full, T.m0 + 64 * j, k0, p, ba);
                        }
                        if (!B_MN) tma_load_4d(stage_b(s, p), mB, full, k0, T.n0, p, bb);
                        else {
#pragma unroll
                            for (int j = 0; j < BN / 64; ++j) tma_load_4d(stage_b(s, p) + j * 8192, mB, full, T.n0 + 64 * j, k0, p, bb);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = instr_desc(BN, A_MN, B_MN, F16, F16);
            constexpr uint32_t kHi = (uint32_t)((1024u >> 4) | (1u << 14) | (2u << 29));
            constexpr uint32_t kLoA = A_MN ? ((8192u >> 4) << 16) : ((16u >> 4) << 16);
            constexpr uint32_t kLoB = B_MN ? ((8192u >> 4) << 16) : ((16u >> 4) << 16);
            constexpr uint32_t kStepA = A_MN ? (2048u >> 4) : (32u >> 4);
            constexpr uint32_t kStepB = B_MN ? (2048u >> 4) : (32u >> 4);
            uint32_t q = 0, tl = 0;
            for (int t = blockIdx.x; t < tiles_total; t += gridDim.x, ++tl) {
                const Tile T = tile_of(t);
                const uint32_t buf = tl & 1;
                mbar_wait(smem_u32(&bar_acc_empty[buf]), ((tl >> 1) & 1) ^ 1);      // the epilogue has drained this accumulator
                tcgen05_fence_after();
                const uint32_t acc = tmem_d + buf * BN;
                uint32_t accumulate = 0;
                for (int i = 0; i < T.nkb; ++i, ++q) {
                    const int s = q % NS;
                    mbar_wait(smem_u32(&bar_full[s]), (q / NS) & 1);
                    tcgen05_fence_after();
                    const uint32_t a_lo0 = kLoA | (stage_a(s, 0) >> 4), b_lo0 = kLoB | (stage_b(s, 0) >> 4);
#pragma unroll
                    for (int pr = 0; pr < NPAIR; ++pr) {
                        const uint32_t a_lo = a_lo0 + ((pr == 2) ? (uint32_t)(Plan::kABytes >> 4) : 0u);
                        const uint32_t b_lo = b_lo0 + ((pr == 1) ? (uint32_t)(Plan::kBBytes >> 4) : 0u);
#pragma unroll
                        for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                            umma_bf16_lohi(acc, a_lo + ks * kStepA, b_lo + ks * kStepB, kHi, idesc, accumulate);
                            accumulate = 1;
                        }
                    }
                    tcgen05_commit(smem_u32(&bar_empty[s]));
                }
                tcgen05_commit(smem_u32(&bar_acc_full[buf]));
            }
        }
    } else {
        // eight epilogue warps: TMEM lane quadrant = warp id % 4 (hardware rule), column half = the warp's group of four
        const int qd = warp & 3, half = (warp - 2) >> 2;
        constexpr int kCols = BN / 2;
        uint32_t tl = 0;
        for (int t = blockIdx.x; t < tiles_total; t += gridDim.x, ++tl) {
            const Tile T = tile_of(t);
            const uint32_t buf = tl & 1;
            const int row = T.m0 + qd * 32 + lane;
            // tile8 epilogues read a global operand per element (conditioning term, saved activation): its loads are issued one chunk
            // ahead - the first chunk's before the accumulator is even complete - so their L2 latency hides behind the MMAs / the
            // previous chunk instead of stalling every pass (4 passes x ~700 cycles per chunk otherwise: the epilogue, not the tensor
            // pipe, bounded the plane-writing GEMMs)
            [[maybe_unused]] typename epi_pre_type<Epi>::type pre_cur[4], pre_nxt[4];
            [[maybe_unused]] typename epi_row_ref<Epi>::type rref[4];
            if constexpr (epi_tile8<Epi>::value && epi_pre_row<Epi>::value) {
#pragma unroll
                for (int i = 0; i < 4; ++i) rref[i] = epi.pre_row(T.batch, min(T.m0 + qd * 32 + i * 8 + (lane >> 2), g.M - 1));
            }
            [[maybe_unused]] auto issue_pre = [&](int c0, typename epi_pre_type<Epi>::type* pr) {
                if constexpr (epi_tile8<Epi>::value) {
                    if (T.n0 + c0 >= g.N) return;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rr = T.m0 + qd * 32 + i * 8 + (lane >> 2);
                        if constexpr (epi_pre_row<Epi>::value) {
                            if (rr < g.M) epi.pre8r(rref[i], T.n0 + c0 + (lane & 3) * 8, pr[i]);
                        } else {
                            if (rr < g.M) epi.pre8(T.batch, rr, T.n0 + c0 + (lane & 3) * 8, pr[i]);
                        }
                    }
                }
            };
            issue_pre(half * kCols, pre_cur);
            mbar_wait(smem_u32(&bar_acc_full[buf]), (tl >> 1) & 1);
            tcgen05_fence_after();
            const uint32_t acc = tmem_d + buf * BN + ((uint32_t)(qd * 32) << 16);
#pragma unroll 1
            for (int c0 = half * kCols; c0 < (half + 1) * kCols; c0 += 32) {
                float v[32];
                if (c0 + 32 < (half + 1) * kCols) issue_pre(c0 + 32, pre_nxt);
                tmem_ld32(acc + c0, v);
                if (c0 + 32 >= (half + 1) * kCols) {   // last chunk read: hand the accumulator back before the (long) global stores of this chunk
                    // (one arrival per warp: hundreds of arrivals on one mbarrier serialise - ~3 ns each, measured in mano_skin_tc.cu)
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cta(smem_u32(&bar_acc_empty[buf]));
                }
                if (T.n0 + c0 >= g.N) continue;
                if constexpr (epi_tile8<Epi>::value) {
                    // 32 x 32 transposition through shared memory with 16-byte accesses: row l holds its eight 4-float chunks at
                    // positions j ^ (l & 7) (conflict-free per quarter warp on the way in and on the way out) - 8 + 8 instructions per
                    // thread instead of 32 + 32 scalar ones (the scalar version's STS / LDS traffic was the epilogue's largest stall)
                    float4* st4 = reinterpret_cast<float4*>(&s_stage[warp - 2][0][0]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) st4[lane * 8 + (j ^ (lane & 7))] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rl = i * 8 + (lane >> 2), cl = (lane & 3) * 8, c4 = (lane & 3) * 2;
                        const float4 w0 = st4[rl * 8 + (c4 ^ (rl & 7))], w1 = st4[rl * 8 + ((c4 + 1) ^ (rl & 7))];
                        float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                        if (T.m0 + qd * 32 + rl < g.M) epi.tile8(T.batch, T.split, T.m0 + qd * 32 + rl, T.n0 + c0 + cl, w, pre_cur[i], g);
                    }
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 4; ++i) pre_cur[i] = pre_nxt[i];
                } else
                if constexpr (Epi::kDirect) {
                    if (row < g.M) epi(T.batch, T.split, row, T.n0 + c0, v, g);
                }
                if constexpr (Epi::kStaged) {
                    float (*st)[33] = s_stage[warp - 2];
#pragma unroll
                    for (int j = 0; j < 32; ++j) st[lane][j] = v[j];
                    __syncwarp();
                    const int col = T.n0 + c0 + lane;
                    const int rows = min(32, g.M - (T.m0 + qd * 32));
                    if constexpr (Epi::kRmw) {
                        long stride;
                        float* p0 = epi.rmw_ptr(T.batch, T.m0 + qd * 32, col, stride);
                        if (p0 && rows > 0) {
                            if (epi.atomic) {
                                for (int rr = 0; rr < rows; ++rr) atomicAdd(p0 + rr * stride, st[rr][lane]);
                            } else if (epi.overwrite) {
#pragma unroll
                                for (int rr = 0; rr < 32; ++rr) if (rr < rows) p0[rr * stride] = st[rr][lane];
                            } else {
                                float old[32];
#pragma unroll
                                for (int rr = 0; rr < 32; ++rr) old[rr] = rr < rows ? p0[rr * stride] : 0.f;
#pragma unroll
                                for (int rr = 0; rr < 32; ++rr) if (rr < rows) p0[rr * stride] = old[rr] + st[rr][lane];
                            }
                        }
                    } else {
                        if (col < g.N)
                            for (int rr = 0; rr < rows; ++rr) epi.elem(T.batch, T.split, T.m0 + qd * 32 + rr, col, st[rr][lane], g);
                    }
                    __syncwarp();
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(2 * BN) : "memory");
    }
}

bool tc_persistent_enabled(const char* what);

// launch helper: builds (cached) tensor maps and launches.  A: rows = M (K-major) or K (MN-major), etc.
const CUtensorMap* cached_map(const PlaneTensor& t, int box_rows, int* status);

template <int BN, bool A_MN, bool B_MN, int NPAIR, bool F16, class Epi>
inline int launch_tc_gemm(const PlaneTensor& A, const PlaneTensor& B, const GemmShape& g, const Epi& epi, cudaStream_t stream, const char* what,
                          const PlaneTensor* A2 = nullptr, const PlaneTensor* B2 = nullptr) {
    if (g.M <= 0 || g.N <= 0 || g.batches <= 0) return MHE_OK;
    if ((g.kb2 > 0) != (A2 != nullptr && B2 != nullptr) || (g.kb2 > 0 && (g.ksplit != 1 || g.kfold != 1))) {
        set_error("%s: a second operand pair needs kb2 > 0, both tensors, ksplit == kfold == 1", what);
        return MHE_ERR_INVALID_ARG;
    }
    constexpr int NPL = NPAIR == 1 ? 1 : 2;
    using Plan = SmemPlan<BN, NPL>;
    int st = MHE_OK;
    const CUtensorMap* ma = cached_map(A, A_MN ? 64 : BM, &st);
    if (st != MHE_OK) return st;
    const CUtensorMap* mb = cached_map(B, B_MN ? 64 : BN, &st);
    if (st != MHE_OK) return st;
    const CUtensorMap *ma2 = ma, *mb2 = mb;
    if (g.kb2 > 0) {
        ma2 = cached_map(*A2, A_MN ? 64 : BM, &st);
        if (st != MHE_OK) return st;
        mb2 = cached_map(*B2, B_MN ? 64 : BN, &st);
        if (st != MHE_OK) return st;
    }
    auto kern = tc_gemm_kernel<BN, A_MN, B_MN, NPAIR, F16, Epi>;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Plan::kBytes) != cudaSuccess) {
            set_error("%s: cannot raise dynamic shared memory to %d", what, Plan::kBytes);
            return MHE_ERR_CUDA;
        }
        attr_set = true;
    }
    ProbeScope probe(what, stream);
    dim3 grid(cdiv(g.N, BN), cdiv(g.M, BM), (g.batches / (g.kfold > 0 ? g.kfold : 1)) * g.ksplit);
    GemmShape gs = g;
    gs.stages = stages_for(what, g.stages, Plan::kStages);
    {   // a ring deeper than the contraction is wasted shared memory: short contractions leave room for several CTAs per SM
        const int kb_total = cdiv(g.K, BK), kb_per = cdiv(kb_total, g.ksplit > 0 ? g.ksplit : 1) * (g.kfold > 0 ? g.kfold : 1) + g.kb2;
        if (gs.stages > kb_per) gs.stages = kb_per > 0 ? kb_per : 1;
    }
    const size_t smem = (size_t)gs.stages * Plan::kStageBytes + 1024;
    const int kb_check = cdiv(g.K, BK);
    const int tiles_total = (int)(grid.x * grid.y * grid.z);
    if (tc_persistent_enabled(what) && g.K > 0 && tiles_total > 0 && (g.ksplit == 1 || kb_check % g.ksplit == 0)) {   // (no empty k-ranges)
        auto pkern = tc_gemm_persistent_kernel<BN, A_MN, B_MN, NPAIR, F16, Epi>;
        static bool pattr_set = false;
        if (!pattr_set) {
            if (cudaFuncSetAttribute(pkern, cudaFuncAttributeMaxDynamicSharedMemorySize, Plan::kBytes) != cudaSuccess) {
                set_error("%s: cannot raise dynamic shared memory to %d", what, Plan::kBytes);
                return MHE_ERR_CUDA;
            }
            pattr_set = true;
        }
        // CTAs per SM: shared memory (the ring) and TMEM (2 x BN of 512 columns per CTA)
        int per_sm = (int)((220 * 1024) / (smem + 4096));
        if (per_sm > 512 / (2 * BN)) per_sm = 512 / (2 * BN);
        if (per_sm < 1) per_sm = 1;
        const int nctas = tiles_total < 148 * per_sm ? tiles_total : 148 * per_sm;
        static const int pf_env = [] { const char* e = getenv("MHE_TC_L2_PREFETCH"); return e ? atoi(e) : 0; }();
        // (worth it only for contractions long enough to run ahead in; the short ones are epilogue-bound)
        const int pf = (cdiv(g.K, BK) * (g.kfold > 0 ? g.kfold : 1) >= 4 && g.kb2 == 0) ? pf_env : 0;
        if (launch_chain(pkern, dim3(nctas), dim3(kThreadsP), smem, stream, *ma, *mb, *ma2, *mb2, gs, epi, (int)grid.x, (int)grid.y, tiles_total, pf) != cudaSuccess) {
            set_error("%s: launch failed: %s", what, cudaGetErrorString(cudaGetLastError()));
            return MHE_ERR_CUDA;
        }
        return check_launch(what);
    }
    if (g.kb2 > 0) {
        set_error("%s: the second operand pair needs the persistent kernel", what);
        return MHE_ERR_UNSUPPORTED;
    }
    if (launch_chain(kern, grid, dim3(kThreads), smem, stream, *ma, *mb, gs, epi) != cudaSuccess) {
        set_error("%s: launch failed: %s", what, cudaGetErrorString(cudaGetLastError()));
        return MHE_ERR_CUDA;
    }
    return check_launch(what);
}

}  // namespace tc
}  // namespace mhe
