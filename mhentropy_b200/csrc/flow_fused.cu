// Cluster-fused flow passes: one kernel per pass, all coupling layers (see flow_fused.cuh for the mapping).
// Reference: hand/flows.py:75-122 (_nets), :210-217 (forward_p), :219-227 (backward_p).
#include "flow_fused.cuh"

namespace mhe {
namespace fused {
using namespace tc;
typedef __nv_bfloat16 bf16;

constexpr int kWorkers = 256;                  // warps 0-7: epilogues + coupling (warp w: TMEM lane quadrant w % 4, column half w / 4)
constexpr int kThreadsF = kWorkers + 96;       // warp 8: weight (A) producer, warp 9: activation (B) producer, warp 10: MMA issuer
constexpr int kASlots = 4;
constexpr int kABytes = 32768;                 // A block, both planes: 2 x [128][64] 16-bit
constexpr int kAPlane = 16384;
constexpr int kBSlots = 2;
constexpr int kBBytes = 16384;                 // B block, both planes: 2 x [64][64] 16-bit
constexpr int kBPlane = 8192;
constexpr int kXaBytes = 32768;                // xm planes (2 x 8 KB) / a0, a1 slice planes (2 x 16 KB) / partial staging
constexpr int kXs = 49;                        // row stride of the fp32 x tile (odd: conflict-free column walks)
constexpr int kActMax = 24;                    // active (transformed) dims per layer the in-cluster exchange is sized for
constexpr int kRecvBytes = 2 * kCluster * kActMax * 8 * 4;   // [parity][source CTA][active dim][8 rows] fp32
constexpr int kSmemBytes = kASlots * kABytes + kBSlots * kBBytes + kXaBytes + NT * kXs * 4 + kRecvBytes + 1024 /*align*/;
// the backward keeps only the 8 gradient rows a CTA owns (not the 64-row x tile): the room buys a third B slot
constexpr int kBSlotsBwd = 3;
constexpr int kSmemBytesBwd = kASlots * kABytes + kBSlotsBwd * kBBytes + kXaBytes + 8 * kXs * 4 + kRecvBytes + 1024 /*align*/;
static_assert(kSmemBytesBwd <= 232448, "shared memory of the fused backward kernel");
constexpr int kTmemCols = 512;                 // three accumulators of 128 columns: [0,64) = Ah.Bh + Al.Bh, [64,128) = Ah.Bl (summed in the epilogue)
constexpr int kAcc0 = 0, kAcc1 = 128, kAcc2 = 256;

bool supported(const FlowLayout& L, int R) {
    static int max_rows = -1;
    if (max_rows < 0) {
        const char* e = getenv("MHE_FUSED_MAX_ROWS");
        max_rows = e ? atoi(e) : 4096;
    }
    return L.D <= 46 && L.D >= 4 && L.H == 512 && L.C % 8 == 0 && L.L <= 16 && L.max_split <= kActMax && R <= max_rows;
}

// ---- cluster / barrier primitives -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t cta) {   // release at cluster scope
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(local_bar), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t raddr, float a, float b, float c, float d) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t raddr, float a) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(raddr), "f"(a) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {        // acquire at cluster scope
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            MHE_MBAR_POLL ".parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// The remote arrivals are RELEASE operations at cluster scope (mbarrier.arrive.release.cluster): they already order every write that
// happens-before them - this thread's and, through the CTA barrier executed just before, the other threads' - so no separate
// fence.acq_rel.cluster (which also invalidates L1 and costs ~1-2 us here) is issued.  MHE_FUSED_FENCE=1 at build time restores it.
#ifdef MHE_FUSED_FENCE
__device__ __forceinline__ void fence_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
#else
__device__ __forceinline__ void fence_cluster() {}
#endif
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float* v) {   // 64 consecutive columns of this thread's lane
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint32_t* q = r + 32 * h;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]),
              "=r"(q[8]), "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]),
              "=r"(q[16]), "=r"(q[17]), "=r"(q[18]), "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]),
              "=r"(q[24]), "=r"(q[25]), "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
            : "r"(taddr + 32 * h));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float fast_tanh_f(float x) {   // 1 - 2/(e^{2x}+1); same form as the per-GEMM path
    const float e = __expf(2.f * x);
    return 1.f - __fdividef(2.f, e + 1.f);
}

__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define BSTAMP(slot) do { if (p.dbg && blockIdx.x == 0) p.dbg[(p.L - 1 - step) * 64 + (slot)] = clock64(); } while (0)
#define MHE_STAMP(slot) do { if (p.dbg && blockIdx.x == 0) p.dbg[step * 64 + (slot)] = clock64(); } while (0)

// Copy the CTA's slice (128 features x 64 rows, split planes, in xa: [k-block][hi | lo][64 k-rows][128 B], 16-byte chunks swizzled)
// to the transposed global planes [plane][H][Rp]: 8 lanes write one 128-byte feature row, a warp four of them per instruction.
__device__ __forceinline__ void copy_slice_to_global(uint32_t xa, bf16* gplanes, int H, int Rp, int f0, int r0, int t) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int idx = i * kWorkers + t;
        const int pl = idx >> 10, row = (idx >> 3) & 127, cc = idx & 7;
        const uint32_t sa = xa + (uint32_t)(row >> 6) * 16384u + (uint32_t)pl * 8192u + (uint32_t)(row & 63) * 128u + (uint32_t)((cc ^ (row & 7)) << 4);
        uint4 v;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sa));
        *reinterpret_cast<uint4*>(gplanes + ((size_t)pl * H + f0 + row) * Rp + r0 + cc * 8) = v;
    }
}

struct FwdArgs {
    const float* params; const float* mask; const float* cp; const float* in;
    float* out; float* logdet;
    float* saved_x; float* saved_st;           // NULL when nothing is saved
    bf16* a0T; bf16* a1T;                        // a0T: [(L or 1)][2][2][H][Rp]; a1T: saved only
    float* partial;
    long long* dbg;                               // optional phase timestamps of CTA 0 (MHE_FUSED_DEBUG)
    int R, Rp, B, D, H, L, direction, save, tiles, two_mma;
    long cp_ld;
    size_t blk, ob2;
};

// MMA issue helper: one 64-deep k-block = 4 UMMA_K steps.  Split precision in TWO instructions per step instead of three:
//   acc[:, 0:128] += A_hi . [B_hi | B_lo]   (N = 128: the two B planes are one operand, the planes b_plane bytes apart)
//   acc[:, 0:64]  += A_lo . B_hi            (N = 64)
// so A_hi is read from shared memory once, not twice (the MMAs are shared-memory-bandwidth bound at N = 64).
//   K-major operand: +32 B per step; MN-major operand: +2048 B per step (descriptor units of 16 B)
template <bool A_MN, bool B_MN, bool F16>
__device__ __forceinline__ void issue_kblock(uint32_t tmem_d, uint32_t a_addr, uint32_t a_plane, uint32_t b_addr, uint32_t b_plane,
                                             uint32_t& accumulate, int two_mma) {
    constexpr uint32_t idesc128 = instr_desc(2 * NT, A_MN, B_MN, F16, F16), idesc64 = instr_desc(NT, A_MN, B_MN, F16, F16);
    constexpr uint32_t kHi = (uint32_t)((1024u >> 4) | (1u << 14) | (2u << 29));   // SBO = 1024 | version 1 | SWIZZLE_128B
    // MN-major A: two 64-row groups 8192 B apart (LBO), +2048 B per k-step; K-major A: LBO unused, +32 B per k-step
    constexpr uint32_t kLoA = A_MN ? ((8192u >> 4) << 16) : ((16u >> 4) << 16);
    constexpr uint32_t kStepA = A_MN ? (2048u >> 4) : (32u >> 4);
    constexpr uint32_t kStepB = B_MN ? (2048u >> 4) : (32u >> 4);
    // MN-major B: LBO = distance between the 64-column groups = between the hi and lo planes.  K-major B: rows 64..127 of the
    // N = 128 operand are the lo plane, which must sit 8 * 1024 B after the hi plane (b_plane == 8192).
    const uint32_t kLoB = B_MN ? ((b_plane >> 4) << 16) : ((16u >> 4) << 16);
    // In a cluster launch the 32-bit shared address carries the CTA rank in its high bits (shared::cluster window): keep only the
    // 14-bit start-address field, or the rank lands in the LBO field of the descriptor.
    const uint32_t a0 = kLoA | ((a_addr >> 4) & 0x3FFFu), a1 = a0 + (a_plane >> 4), b0 = kLoB | ((b_addr >> 4) & 0x3FFFu);
    if (!two_mma) {   // reference schedule: three N = 64 passes (hi,hi) (hi,lo) (lo,hi) into columns [0,64)
        const uint32_t b1 = b0 + (b_plane >> 4);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            umma_bf16_lohi(tmem_d, a0 + ks * kStepA, b0 + ks * kStepB, kHi, idesc64, accumulate);
            umma_bf16_lohi(tmem_d, a0 + ks * kStepA, b1 + ks * kStepB, kHi, idesc64, 1u);
            umma_bf16_lohi(tmem_d, a1 + ks * kStepA, b0 + ks * kStepB, kHi, idesc64, 1u);
            accumulate = 1;
        }
        return;
    }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        umma_bf16_lohi(tmem_d, a0 + ks * kStepA, b0 + ks * kStepB, kHi, idesc128, accumulate);
        umma_bf16_lohi(tmem_d, a1 + ks * kStepA, b0 + ks * kStepB, kHi, idesc64, 1u);
        accumulate = 1;
    }
}

__global__ void __launch_bounds__(kThreadsF, 1)
flow_fwd_fused_kernel(const __grid_constant__ CUtensorMap mapW0, const __grid_constant__ CUtensorMap mapW1,
                      const __grid_constant__ CUtensorMap mapW2, const __grid_constant__ CUtensorMap mapA0, FwdArgs p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_fullA[kASlots], bar_emptyA[kASlots], bar_fullB[kBSlots], bar_emptyB[kBSlots], bar_acc[3];
    __shared__ __align__(8) uint64_t bar_xm, bar_own, bar_a1, bar_a0, bar_part, bar_x;
    __shared__ uint8_t actd[16][kActMax];       // per layer: the active (transformed) dims in ascending order
    __shared__ int nact_s[16];
    __shared__ uint32_t tmem_slot;
    __shared__ float lds[NT];
    __shared__ uint64_t mbits[64];

    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t ringA = smem0, ringB = ringA + kASlots * kABytes, xa = ringB + kBSlots * kBBytes;
    float* xs = reinterpret_cast<float*>(smem_raw + (xa - smem_u32(smem_raw)) + kXaBytes);   // [NT][kXs]
    float* recv = xs + NT * kXs;                // [2 parities][8 source CTAs][kActMax][8 rows]: partial head outputs of this CTA's 8 rows
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int net = rank >> 2, j = rank & 3;
    const int tile = blockIdx.x / kCluster;
    const int r0 = tile * NT;
    const int nkb = p.H / 64;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kASlots; ++s) { mbar_init(smem_u32(&bar_fullA[s]), 1); mbar_init(smem_u32(&bar_emptyA[s]), 1); }
        for (int s = 0; s < kBSlots; ++s) { mbar_init(smem_u32(&bar_fullB[s]), 1); mbar_init(smem_u32(&bar_emptyB[s]), 1); }
        for (int s = 0; s < 3; ++s) mbar_init(smem_u32(&bar_acc[s]), 1);
        mbar_init(smem_u32(&bar_xm), kWorkers / 32);     // one elected arrival per worker warp (256 arrivals on one barrier serialise)
        mbar_init(smem_u32(&bar_a1), kWorkers / 32);
        mbar_init(smem_u32(&bar_own), 1);
        mbar_init(smem_u32(&bar_a0), 4);
        mbar_init(smem_u32(&bar_part), kCluster);
        mbar_init(smem_u32(&bar_x), kCluster);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW2) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA0) : "memory");
    }
    if (warp == 10) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    cluster_sync_all();            // every CTA's barriers are initialised before any remote arrival
    const uint32_t tmem = tmem_slot;

    if (warp == 8) {
        // ===================== weight (A) producer: W0, W1 k-blocks (own two first), W2 — runs ahead by kASlots blocks
        if (lane == 0) {
            uint32_t q = 0;
            auto acquire = [&]() -> uint32_t {
                const uint32_t s = q % kASlots, use = q / kASlots;
                mbar_wait(smem_u32(&bar_emptyA[s]), (use & 1) ^ 1);
                mbar_expect_tx(smem_u32(&bar_fullA[s]), kABytes);
                ++q;
                return s;
            };
            for (int step = 0; step < p.L; ++step) {
                const int layer = p.direction == 0 ? step : p.L - 1 - step;
                const int wb = layer * 2 + net;
                {   // W0 slice [128 features][64 dims]
                    const uint32_t s = acquire(), full = smem_u32(&bar_fullA[s]), base = ringA + s * kABytes;
                    MHE_STAMP(56);
                    tma_load_4d(base, &mapW0, full, 0, j * FS, 0, wb);
                    tma_load_4d(base + kAPlane, &mapW0, full, 0, j * FS, 1, wb);
                }
                for (int i = 0; i < nkb; ++i) {   // W1 [128 out features][64 in features], k-blocks starting at the CTA's own
                    const int kb = (2 * j + i) % nkb;
                    const uint32_t s = acquire(), full = smem_u32(&bar_fullA[s]), base = ringA + s * kABytes;
                    MHE_STAMP(48 + i);
                    tma_load_4d(base, &mapW1, full, kb * 64, j * FS, 0, wb);
                    tma_load_4d(base + kAPlane, &mapW1, full, kb * 64, j * FS, 1, wb);
                }
                {   // W2 [64 dims][128 feature slice] as two k-blocks per plane
                    const uint32_t s = acquire(), full = smem_u32(&bar_fullA[s]), base = ringA + s * kABytes;
                    MHE_STAMP(57);
#pragma unroll
                    for (int pl = 0; pl < 2; ++pl)
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk)
                            tma_load_4d(base + pl * kAPlane + kk * 8192, &mapW2, full, j * FS + kk * 64, 0, pl, wb);
                }
            }
        }
    } else if (warp == 9) {
        // ===================== activation (B) producer: the three peers' a0T k-blocks from the exchange buffer in global memory
        if (lane == 0) {
            uint32_t q = 0;
            for (int step = 0; step < p.L; ++step) {
                const int ab = (p.save ? step * 2 : 0) + net;
                MHE_STAMP(16);
                mbar_wait_cluster(smem_u32(&bar_a0), step & 1);       // all four slices of this net's a0T are in global memory
                fence_proxy_async();
                MHE_STAMP(17);
                for (int i = 2; i < nkb; ++i) {
                    const int kb = (2 * j + i) % nkb;
                    const uint32_t s = q % kBSlots, use = q / kBSlots;
                    ++q;
                    mbar_wait(smem_u32(&bar_emptyB[s]), (use & 1) ^ 1);
                    const uint32_t full = smem_u32(&bar_fullB[s]), base = ringB + s * kBBytes;
                    mbar_expect_tx(full, kBBytes);
                    tma_load_4d(base, &mapA0, full, r0, kb * 64, 0, ab);
                    tma_load_4d(base + kBPlane, &mapA0, full, r0, kb * 64, 1, ab);
                }
            }
        }
    } else if (warp == 10) {
        // ===================== MMA issuer
        if (lane == 0) {
            uint32_t qa = 0, qb = 0;
            auto wait_a = [&]() -> uint32_t {
                const uint32_t s = qa % kASlots, use = qa / kASlots;
                mbar_wait(smem_u32(&bar_fullA[s]), use & 1);
                ++qa;
                return s;
            };
            for (int step = 0; step < p.L; ++step) {
                const uint32_t par = step & 1;
                {   // G0: acc0 = W0 slice . xm^T
                    const uint32_t s = wait_a();
                    MHE_STAMP(20);
                    mbar_wait(smem_u32(&bar_xm), par);
                    tcgen05_fence_after();
                    MHE_STAMP(21);
                    uint32_t acc = 0;
                    issue_kblock<false, false, true>(tmem + kAcc0, ringA + s * kABytes, kAPlane, xa, 8192, acc, p.two_mma & 1);
                    tcgen05_commit(smem_u32(&bar_emptyA[s]));
                    tcgen05_commit(smem_u32(&bar_acc[0]));
                }
                {   // G1: acc1 = W1 slice . a0T — own two k-blocks straight from shared memory, the peers' through the B ring
                    uint32_t acc = 0;
                    for (int i = 0; i < nkb; ++i) {
                        const uint32_t s = wait_a();
                        MHE_STAMP(32 + i);
                        const uint32_t abase = ringA + s * kABytes;
                        if (i < 2) {
                            if (i == 0) { mbar_wait(smem_u32(&bar_own), par); MHE_STAMP(22); }
                            tcgen05_fence_after();
                            issue_kblock<false, true, true>(tmem + kAcc1, abase, kAPlane, xa + i * 16384, 8192, acc, p.two_mma & 2);
                        } else {
                            const uint32_t sb = qb % kBSlots, useb = qb / kBSlots;
                            ++qb;
                            mbar_wait(smem_u32(&bar_fullB[sb]), useb & 1);
                            tcgen05_fence_after();
                            MHE_STAMP(40 + i);
                            issue_kblock<false, true, true>(tmem + kAcc1, abase, kAPlane, ringB + sb * kBBytes, kBPlane, acc, p.two_mma & 2);
                            tcgen05_commit(smem_u32(&bar_emptyB[sb]));
                        }
                        tcgen05_commit(smem_u32(&bar_emptyA[s]));
                    }
                    tcgen05_commit(smem_u32(&bar_acc[1]));
                }
                {   // G2: acc2 = W2[:, slice] . a1T slice (partial head outputs)
                    const uint32_t s = wait_a();
                    MHE_STAMP(24);
                    mbar_wait(smem_u32(&bar_a1), par);
                    tcgen05_fence_after();
                    MHE_STAMP(25);
                    uint32_t acc = 0;
                    const uint32_t base = ringA + s * kABytes;
                    issue_kblock<false, true, true>(tmem + kAcc2, base, kAPlane, xa, 8192, acc, p.two_mma & 8);
                    issue_kblock<false, true, true>(tmem + kAcc2, base + 8192, kAPlane, xa + 16384, 8192, acc, p.two_mma & 8);
                    tcgen05_commit(smem_u32(&bar_emptyA[s]));
                    tcgen05_commit(smem_u32(&bar_acc[2]));
                }
            }
        }
    } else {
        // ===================== workers (256 threads)
        const int t = threadIdx.x;
        const int lq = warp & 3, ch = warp >> 2;                   // TMEM lane quadrant, column half
        const int fl = lq * 32 + lane;                             // feature inside the slice = TMEM lane
        const int f = j * FS + fl;                                 // feature inside the net
        const uint32_t tm_lane = (uint32_t)(lq * 32) << 16;
        const int D = p.D;
        const int xn = t & 63, xq = t >> 6;                        // xm writer: row, quarter of the 64 padded dims
        const uint32_t rowoff = (uint32_t)(fl >> 6) * 16384u + (uint32_t)(fl & 63) * 128u;  // k-row fl of the slice: [k-block][hi | lo][64 k-rows][128 B]

        // one thread polls the mbarrier, the others block on the named barrier: 256 spinning threads would fight the tensor core for the
        // shared-memory pipe while it streams MMA operands
        auto worker_wait = [&](uint64_t* bar, uint32_t parity, bool cluster_scope) {
            if (t == 0) { if (cluster_scope) mbar_wait_cluster(smem_u32(bar), parity); else mbar_wait(smem_u32(bar), parity); }
            worker_sync();
        };
        // accumulator columns [col, col+32) of both halves (Ah.Bh + Al.Bh | Ah.Bl), summed
        auto load_acc = [&](uint32_t acc_col, float* v, int both) {
            float w[32];
            tmem_ld32(tmem + tm_lane + acc_col + 32 * ch, v);
            if (both) {
                tmem_ld32(tmem + tm_lane + acc_col + 64 + 32 * ch, w);
#pragma unroll
                for (int n = 0; n < 32; ++n) v[n] += w[n];
            }
        };
        auto write_xm = [&](uint64_t mb) {   // xm = mask * x as K-major split planes [64 rows][64 dims]; lo plane 8 KB after hi
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int cc = xq * 2 + c;
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int d = cc * 8 + e;
                    v[e] = (d < D && ((mb >> d) & 1)) ? xs[xn * kXs + d] : 0.f;
                }
                uint4 hi, lo;
                split8<true>(v, hi, lo);
                const uint32_t off = (uint32_t)(xn >> 3) * 1024u + (uint32_t)(xn & 7) * 128u + (uint32_t)((cc ^ (xn & 7)) << 4);
                st_shared_v4(xa + off, hi);
                st_shared_v4(xa + 8192 + off, lo);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_xm));
        };
        auto load_cp = [&](float* c, int which, int layer) {   // cp[image of row][layer, net, which][f] for this thread's 32 rows
            const float* cpb = p.cp + (size_t)(layer * 4 + net * 2 + which) * p.H + f;
            int img = (r0 + 32 * ch) % p.B;
#pragma unroll
            for (int n = 0; n < 32; ++n) { c[n] = __ldg(cpb + (size_t)img * p.cp_ld); if (++img == p.B) img = 0; }
        };
        // this thread's 32 activations -> its half of k-row fl of the MN-major slice planes in xa (and of the global planes)
        auto store_slice = [&](const float* v, uint4* ghi, uint4* glo) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int cc = ch * 4 + i;
                uint4 hi, lo;
                split8<true>(v + 8 * i, hi, lo);
                const uint32_t off = rowoff + (uint32_t)((cc ^ (fl & 7)) << 4);
                st_shared_v4(xa + off, hi);
                st_shared_v4(xa + 8192 + off, lo);
                if (ghi) { ghi[cc] = hi; glo[cc] = lo; }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        };

        // coupling masks as bit sets (bit d = mask[layer][d] != 0)
        if (t < p.L) {
            uint64_t mb = 0;
            for (int d = 0; d < D; ++d) mb |= (uint64_t)(p.mask[(size_t)t * D + d] != 0.f) << d;
            mbits[t] = mb;
            int na = 0;
            for (int d = 0; d < D; ++d) if (!((mb >> d) & 1) && na < kActMax) actd[t][na++] = (uint8_t)d;
            nact_s[t] = na;
        }
        // load the row tile
        for (int i = t; i < NT * D; i += kWorkers) {
            const int n = i / D, d = i - n * D;
            xs[n * kXs + d] = (r0 + n < p.R) ? p.in[(size_t)(r0 + n) * D + d] : 0.f;
        }
        if (t < NT) lds[t] = 0.f;
        worker_sync();
        if (p.save && rank == 0) {
            const int nvalid = min(NT, p.R - r0) * D;
            for (int i = t; i < nvalid; i += kWorkers) { const int n = i / D, d = i - n * D; p.saved_x[(size_t)r0 * D + i] = xs[n * kXs + d]; }
        }
        float v[32], c[32];
        load_cp(c, 0, p.direction == 0 ? 0 : p.L - 1);
        write_xm(mbits[p.direction == 0 ? 0 : p.L - 1]);

        for (int step = 0; step < p.L; ++step) {
            const int layer = p.direction == 0 ? step : p.L - 1 - step;
            const uint32_t par = step & 1;
            const uint64_t mb = mbits[layer];
            const int ab = (p.save ? step * 2 : 0) + net;           // batch index into a0T / a1T
            // head biases of this thread's coupling element (used in U; loaded now, off the critical path)
            float b2s_pre = 0.f, b2t_pre = 0.f;
            if ((t >> 3) < nact_s[layer]) {
                const int d = actd[layer][t >> 3];
                b2s_pre = __ldg(p.params + (size_t)(layer * 2 + 0) * p.blk + p.ob2 + d);
                b2t_pre = __ldg(p.params + (size_t)(layer * 2 + 1) * p.blk + p.ob2 + d);
            }
            if (t == 0) { MHE_STAMP(0); if (p.dbg && blockIdx.x == 0) p.dbg[step * 64 + 60] = gtime(); }
            // ---------------- E0: a0 = lrelu(acc0 + cp0) -> shared slice (own k-blocks of G1) + global exchange buffer (peers' k-blocks)
            {
                worker_wait(&bar_acc[0], par, false);
                tcgen05_fence_after();
                if (t == 0) MHE_STAMP(2);
                load_acc(kAcc0, v, p.two_mma & 1);
#pragma unroll
                for (int n = 0; n < 32; ++n) v[n] = lrelu(v[n] + c[n]);
                store_slice(v, nullptr, nullptr);
                tcgen05_fence_before();
                worker_sync();
                if (t == 0) { MHE_STAMP(3); mbar_arrive(smem_u32(&bar_own)); }   // the MMA warp may consume the own slice
                copy_slice_to_global(xa, p.a0T + (size_t)ab * 2 * p.H * p.Rp, p.H, p.Rp, j * FS, r0, t);
                worker_sync();
                if (t < 4) {   // publish the slice: the CTA barrier ordered every thread's stores before this cumulative fence
                    fence_proxy_async();
                    fence_cluster();
                    mbar_arrive_remote(smem_u32(&bar_a0), net * 4 + t);
                    if (t == 0) MHE_STAMP(4);
                }
                load_cp(c, 1, layer);                                 // in flight while G1 runs
            }
            // ---------------- E1: a1 = lrelu(acc1 + cp1) -> shared (B operand of G2, MN-major) [+ global when saving]
            {
                worker_wait(&bar_acc[1], par, false);
                tcgen05_fence_after();
                if (t == 0) MHE_STAMP(6);
                load_acc(kAcc1, v, p.two_mma & 2);
#pragma unroll
                for (int n = 0; n < 32; ++n) v[n] = lrelu(v[n] + c[n]);
                store_slice(v, nullptr, nullptr);
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bar_a1));
                if (p.save) {   // saved for the backward; off the critical path (G2 only reads xa).  (Storing the planes straight from
                    // the registers instead - 32 scattered 16-byte segments per instruction - measured slower than this coalesced copy.)
                    worker_sync();
                    copy_slice_to_global(xa, p.a1T + (size_t)ab * 2 * p.H * p.Rp, p.H, p.Rp, j * FS, r0, t);
                }
                if (t == 0) MHE_STAMP(7);
            }
            // ---------------- E2: partial head outputs of this CTA's feature slice -> the CTA that owns the rows (DSMEM)
            // Row ownership for the coupling: CTA c updates rows [8c, 8c+8) of the tile.  Thread (dim d, column half ch) holds rows
            // 32 ch .. 32 ch + 31 of partial[d]: four 8-row groups for CTAs 4 ch .. 4 ch + 3.
            {
                worker_wait(&bar_acc[2], par, false);
                tcgen05_fence_after();
                if (t == 0) MHE_STAMP(8);
                if (lq < 2) {   // accumulator rows = flow dims: lanes 0..63
                    load_acc(kAcc2, v, p.two_mma & 8);
                    if (fl < D && !((mb >> fl) & 1)) {
                        const int ai = __popcll(~mb & ((1ull << fl) - 1ull));                 // index of this dim among the active ones
                        if (ai < kActMax) {
                            const uint32_t slot = smem_u32(recv) + (uint32_t)((((par * kCluster + rank) * kActMax + ai) * 8) * 4);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint32_t ra = map_to_cta(slot, 4 * ch + q);
                                st_cluster_v4(ra, v[8 * q], v[8 * q + 1], v[8 * q + 2], v[8 * q + 3]);
                                st_cluster_v4(ra + 16, v[8 * q + 4], v[8 * q + 5], v[8 * q + 6], v[8 * q + 7]);
                            }
                        }
                    }
                }
                tcgen05_fence_before();
                worker_sync();
                if (t < kCluster) { fence_cluster(); mbar_arrive_remote(smem_u32(&bar_part), t); }
                if (t == 0) MHE_STAMP(9);
            }
            // ---------------- U: affine coupling of this CTA's 8 rows, results broadcast to every CTA's x tile (DSMEM)
            {
                const bool last = step == p.L - 1;
                const int next_layer = p.direction == 0 ? layer + 1 : layer - 1;
                worker_wait(&bar_part, par, true);
                if (t == 0) MHE_STAMP(10);
                const int urow = t & 7, uai = t >> 3;          // row inside the CTA's group, active-dim index
                if (uai < nact_s[layer]) {
                    const int d = actd[layer][uai];
                    const int n = 8 * (int)rank + urow;
                    const float* rv = recv + ((size_t)(par * kCluster) * kActMax + uai) * 8 + urow;
                    float sp = b2s_pre, tp = b2t_pre;
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) { sp += rv[(size_t)c4 * kActMax * 8]; tp += rv[(size_t)(4 + c4) * kActMax * 8]; }
                    const float sv = fast_tanh_f(sp);
                    const float xv = xs[n * kXs + d];
                    float y;
                    if (p.direction == 0) { y = fmaf(xv, expf(sv), tp); atomicAdd(&lds[n], sv); }
                    else { y = (xv - tp) * expf(-sv); atomicAdd(&lds[n], -sv); }
                    const uint32_t xaddr = smem_u32(xs + n * kXs + d);
#pragma unroll
                    for (int c8 = 0; c8 < kCluster; ++c8) st_cluster_f32(map_to_cta(xaddr, c8), y);
                    if (p.save && r0 + n < p.R) {
                        p.saved_st[((size_t)(step * 2 + 0) * p.R + r0 + n) * D + d] = sv;
                        p.saved_st[((size_t)(step * 2 + 1) * p.R + r0 + n) * D + d] = tp;
                    }
                }
                worker_sync();
                if (t < kCluster) { fence_cluster(); mbar_arrive_remote(smem_u32(&bar_x), t); }
                worker_wait(&bar_x, par, true);                // every CTA's rows of the new x have landed in this CTA's tile
                if (t == 0) MHE_STAMP(11);
                if (!last) {
                    write_xm(mbits[next_layer]);
                    load_cp(c, 0, next_layer);                 // next layer's conditioning, in flight during G0
                }
                if (p.save && (int)rank == ((step + 1) & 7)) {
                    const int nvalid = min(NT, p.R - r0) * D;
                    float* dst = p.saved_x + (size_t)(step + 1) * p.R * D + (size_t)r0 * D;
                    for (int i = t; i < nvalid; i += kWorkers) { const int n = i / D, d = i - n * D; dst[i] = xs[n * kXs + d]; }
                }
                if (last) {
                    if (rank == 0) {
                        const int nvalid = min(NT, p.R - r0) * D;
                        for (int i = t; i < nvalid; i += kWorkers) { const int n = i / D, d = i - n * D; p.out[(size_t)r0 * D + i] = xs[n * kXs + d]; }
                    }
                    // every CTA holds the log-determinant of its own 8 rows
                    if (p.logdet && t < 8 && r0 + 8 * (int)rank + t < p.R) p.logdet[r0 + 8 * rank + t] = lds[8 * rank + t];
                }
            }
        }
    }
    // teardown: nobody leaves while a peer may still signal it
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 10) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

// =====================================================================================================================
// Backward pass (data gradients) of all L layers for one row tile; same cluster mapping and exchange scheme as the forward.
// Per layer (reverse order), with g = dL/d(layer output) held in shared memory:
//   bU  coupling backward (reference flows.py:213-216 / 222-226 differentiated): dpre (head pre-activation gradient of the CTA's
//       net) -> shared, K-major; direct path of dL/dx -> g in place
//   bG2 dh1T[slice] = W2^T[slice, :] . dpre^T         then  * lrelu'(a1)   -> shared slice + global (exchange, weight gradients)
//   bG1 dh0T[slice] = W1^T[slice, :] . dh1T           then  * lrelu'(a0)   -> shared slice + global (weight gradients)
//   bG0 part[d]     = W0^T[:, slice] . dh0T[slice]    partial input gradients, summed by every CTA after the exchange
//   g[n][d] += mask[d] * sum_c part_c[d][n]
// Gradients travel as bfloat16 split planes (full fp32 range); the weights are read from their bfloat16 planes, MN-major.
struct BwdArgs {
    const float* mask; const float* saved_x; const float* saved_st;
    const bf16* a0T; const bf16* a1T;           // saved half planes: only their signs are read here
    const float* dout; const float* dlogdet; float dlogdet_scale;
    float* din; float* dparams;
    bf16* dh1T; bf16* dh0T; bf16* dpreT;        // [L][2 nets][2 planes][H][Rp] x2, [L][2][2][64][Rp]
    float* partial;
    long long* dbg;
    int R, Rp, D, H, L, direction, tiles, two_mma;
    int s_lo, s_hi;                             // the steps [s_lo, s_hi) this launch runs (descending): the pass may be cut into chunks
    int zero_dpreT;                             // the kernel zeroes its rows of dpreT itself (no memset in front of it)
    size_t blk, ob2;
};

__global__ void __launch_bounds__(kThreadsF, 1)
flow_bwd_fused_kernel(const __grid_constant__ CUtensorMap mapW0, const __grid_constant__ CUtensorMap mapW1,
                      const __grid_constant__ CUtensorMap mapW2, const __grid_constant__ CUtensorMap mapDh1, BwdArgs p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_fullA[kASlots], bar_emptyA[kASlots], bar_fullB[kBSlotsBwd], bar_emptyB[kBSlotsBwd], bar_acc[3];
    __shared__ __align__(8) uint64_t bar_xm, bar_own, bar_a1, bar_a0, bar_part, bar_dpre;
    __shared__ uint32_t tmem_slot;
    __shared__ float gls[8];
    __shared__ uint64_t mbits[64];
    __shared__ uint8_t actd[16][kActMax], pasd[16][kActMax];   // per layer: transformed / conditioning dims, ascending
    __shared__ int nact_s[16], npas_s[16];

    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t ringA = smem0, ringB = ringA + kASlots * kABytes, xa = ringB + kBSlotsBwd * kBBytes;
    float* gs = reinterpret_cast<float*>(smem_raw + (xa - smem_u32(smem_raw)) + kXaBytes);   // [8][kXs]: gradient rows this CTA owns (rows 8 rank ..)
    float* recv = gs + 8 * kXs;                // [2 parities][8 source CTAs][kActMax][8 rows]: partial input gradients of the owned rows
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int net = rank >> 2, j = rank & 3;
    const int tile = blockIdx.x / kCluster;
    const int r0 = tile * NT;
    const int nkb = p.H / 64;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kASlots; ++s) { mbar_init(smem_u32(&bar_fullA[s]), 1); mbar_init(smem_u32(&bar_emptyA[s]), 1); }
        for (int s = 0; s < kBSlotsBwd; ++s) { mbar_init(smem_u32(&bar_fullB[s]), 1); mbar_init(smem_u32(&bar_emptyB[s]), 1); }
        for (int s = 0; s < 3; ++s) mbar_init(smem_u32(&bar_acc[s]), 1);
        mbar_init(smem_u32(&bar_xm), 1);
        mbar_init(smem_u32(&bar_a1), kWorkers / 32);
        mbar_init(smem_u32(&bar_own), 1);
        mbar_init(smem_u32(&bar_a0), 4);
        mbar_init(smem_u32(&bar_part), kCluster);
        mbar_init(smem_u32(&bar_dpre), kCluster);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW2) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapDh1) : "memory");
    }
    if (warp == 10) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    cluster_sync_all();
    const uint32_t tmem = tmem_slot;

    if (warp == 8) {
        // ===================== weight (A) producer, all blocks MN-major: [64 k-rows][64 m-cols] boxes, two per plane
        if (lane == 0) {
            uint32_t q = 0;
            auto acquire = [&]() -> uint32_t {
                const uint32_t s = q % kASlots, use = q / kASlots;
                mbar_wait(smem_u32(&bar_emptyA[s]), (use & 1) ^ 1);
                mbar_expect_tx(smem_u32(&bar_fullA[s]), kABytes);
                ++q;
                return s;
            };
            for (int step = p.s_hi - 1; step >= p.s_lo; --step) {
                const int layer = p.direction == 0 ? step : p.L - 1 - step;
                const int wb = layer * 2 + net;
                {   // W2^T slice: k = flow dim (64 rows of W2), m = feature slice (two 64-column groups)
                    const uint32_t s = acquire(), full = smem_u32(&bar_fullA[s]), base = ringA + s * kABytes;
#pragma unroll
                    for (int pl = 0; pl < 2; ++pl)
#pragma unroll
                        for (int g = 0; g < 2; ++g) tma_load_4d(base + pl * kAPlane + g * 8192, &mapW2, full, j * FS + g * 64, 0, pl, wb);
                }
                for (int i = 0; i < nkb; ++i) {   // W1^T: k = out-feature block (rows of W1), m = in-feature slice
                    const int kb = (2 * j + i) % nkb;
                    const uint32_t s = acquire(), full = smem_u32(&bar_fullA[s]), base = ringA + s * kABytes;
#pragma unroll
                    for (int pl = 0; pl < 2; ++pl)
#pragma unroll
                        for (int g = 0; g < 2; ++g) tma_load_4d(base + pl * kAPlane + g * 8192, &mapW1, full, j * FS + g * 64, kb * 64, pl, wb);
                }
                {   // W0^T: k = feature slice (two 64-row blocks of W0), m = flow dim (one 64-column group)
                    const uint32_t s = acquire(), full = smem_u32(&bar_fullA[s]), base = ringA + s * kABytes;
#pragma unroll
                    for (int pl = 0; pl < 2; ++pl)
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) tma_load_4d(base + pl * kAPlane + kk * 8192, &mapW0, full, 0, j * FS + kk * 64, pl, wb);
                }
            }
        }
    } else if (warp == 9) {
        // ===================== gradient (B) producer: the three peers' dh1T k-blocks from global memory
        if (lane == 0) {
            uint32_t q = 0;
            for (int step = p.s_hi - 1; step >= p.s_lo; --step) {
                const int it = p.s_hi - 1 - step;
                const int ab = (p.direction == 0 ? step : p.L - 1 - step) * 2 + net;
                mbar_wait_cluster(smem_u32(&bar_a0), it & 1);
                fence_proxy_async();
                for (int i = 2; i < nkb; ++i) {
                    const int kb = (2 * j + i) % nkb;
                    const uint32_t s = q % kBSlotsBwd, use = q / kBSlotsBwd;
                    ++q;
                    mbar_wait(smem_u32(&bar_emptyB[s]), (use & 1) ^ 1);
                    const uint32_t full = smem_u32(&bar_fullB[s]), base = ringB + s * kBBytes;
                    mbar_expect_tx(full, kBBytes);
                    tma_load_4d(base, &mapDh1, full, r0, kb * 64, 0, ab);
                    tma_load_4d(base + kBPlane, &mapDh1, full, r0, kb * 64, 1, ab);
                }
            }
        }
    } else if (warp == 10) {
        // ===================== MMA issuer
        if (lane == 0) {
            uint32_t qa = 0, qb = 0;
            auto wait_a = [&]() -> uint32_t {
                const uint32_t s = qa % kASlots, use = qa / kASlots;
                mbar_wait(smem_u32(&bar_fullA[s]), use & 1);
                ++qa;
                return s;
            };
            for (int step = p.s_hi - 1; step >= p.s_lo; --step) {
                const uint32_t par = (p.s_hi - 1 - step) & 1;
                {   // bG2: acc0 = W2^T slice . dpre^T   (B K-major in xa)
                    const uint32_t s = wait_a();
                    mbar_wait(smem_u32(&bar_xm), par);
                    tcgen05_fence_after();
                    uint32_t acc = 0;
                    issue_kblock<true, false, false>(tmem + kAcc0, ringA + s * kABytes, kAPlane, xa, 8192, acc, p.two_mma);
                    tcgen05_commit(smem_u32(&bar_emptyA[s]));
                    tcgen05_commit(smem_u32(&bar_acc[0]));
                }
                {   // bG1: acc1 = W1^T slice . dh1T
                    uint32_t acc = 0;
                    for (int i = 0; i < nkb; ++i) {
                        const uint32_t s = wait_a();
                        const uint32_t abase = ringA + s * kABytes;
                        if (i < 2) {
                            if (i == 0) mbar_wait(smem_u32(&bar_own), par);
                            tcgen05_fence_after();
                            issue_kblock<true, true, false>(tmem + kAcc1, abase, kAPlane, xa + i * 16384, 8192, acc, p.two_mma);
                        } else {
                            const uint32_t sb = qb % kBSlotsBwd, useb = qb / kBSlotsBwd;
                            ++qb;
                            mbar_wait(smem_u32(&bar_fullB[sb]), useb & 1);
                            tcgen05_fence_after();
                            issue_kblock<true, true, false>(tmem + kAcc1, abase, kAPlane, ringB + sb * kBBytes, kBPlane, acc, p.two_mma);
                            tcgen05_commit(smem_u32(&bar_emptyB[sb]));
                        }
                        tcgen05_commit(smem_u32(&bar_emptyA[s]));
                    }
                    tcgen05_commit(smem_u32(&bar_acc[1]));
                }
                {   // bG0: acc2 = W0^T[:, slice] . dh0T slice
                    const uint32_t s = wait_a();
                    mbar_wait(smem_u32(&bar_a1), par);
                    tcgen05_fence_after();
                    uint32_t acc = 0;
                    const uint32_t base = ringA + s * kABytes;
                    issue_kblock<true, true, false>(tmem + kAcc2, base, kAPlane, xa, 8192, acc, p.two_mma);
                    issue_kblock<true, true, false>(tmem + kAcc2, base + 8192, kAPlane, xa + 16384, 8192, acc, p.two_mma);
                    tcgen05_commit(smem_u32(&bar_emptyA[s]));
                    tcgen05_commit(smem_u32(&bar_acc[2]));
                }
            }
        }
    } else {
        // ===================== workers (256 threads)
        // Row ownership: CTA c owns rows [8c, 8c+8) of the tile.  The owner keeps their gradient g, sums the partial input gradients the
        // eight CTAs send it (DSMEM), runs the coupling backward for them and writes the head gradients straight into the shared-memory
        // B tiles of the CTAs of each net (DSMEM), so no CTA repeats the elementwise work and nothing goes through global memory.
        const int t = threadIdx.x;
        const int lq = warp & 3, ch = warp >> 2;
        const int fl = lq * 32 + lane;
        const int f = j * FS + fl;
        const uint32_t tm_lane = (uint32_t)(lq * 32) << 16;
        const int D = p.D;
        const int urow = t & 7, uk = t >> 3;                       // owned row, index into the active / passive dim lists
        const int own_n = 8 * (int)rank + urow, own_r = r0 + own_n;
        const uint32_t rowoff = (uint32_t)(fl >> 6) * 16384u + (uint32_t)(fl & 63) * 128u;

        auto worker_wait = [&](uint64_t* bar, uint32_t parity, bool cluster_scope) {
            if (t == 0) { if (cluster_scope) mbar_wait_cluster(smem_u32(bar), parity); else mbar_wait(smem_u32(bar), parity); }
            worker_sync();
        };
        auto load_acc = [&](uint32_t acc_col, float* v) {
            tmem_ld32(tmem + tm_lane + acc_col + 32 * ch, v);
            if (p.two_mma) {
                float w[32];
                tmem_ld32(tmem + tm_lane + acc_col + 64 + 32 * ch, w);
#pragma unroll
                for (int n = 0; n < 32; ++n) v[n] += w[n];
            }
        };
        // 32 gradients of feature fl (rows 32 ch ..) masked by lrelu'(saved activation), as bfloat16 planes -> xa slice
        auto store_grad_slice = [&](float* v, const uint4* sg, uint4* ghi, uint4* glo) {
            uint32_t sw[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) { sw[4 * i] = sg[i].x; sw[4 * i + 1] = sg[i].y; sw[4 * i + 2] = sg[i].z; sw[4 * i + 3] = sg[i].w; }
#pragma unroll
            for (int n = 0; n < 32; ++n) {
                const uint32_t e = (sw[n >> 1] >> ((n & 1) * 16)) & 0xFFFFu;      // saved half activation: > 0 <=> sign clear and non-zero
                v[n] *= ((e & 0x8000u) == 0 && (e & 0x7FFFu) != 0) ? 1.f : kLeakySlope;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int cc = ch * 4 + i;
                uint4 hi, lo;
                split8<false>(v + 8 * i, hi, lo);
                const uint32_t off = rowoff + (uint32_t)((cc ^ (fl & 7)) << 4);
                st_shared_v4(xa + off, hi);
                st_shared_v4(xa + 8192 + off, lo);
                if (ghi) { ghi[cc] = hi; glo[cc] = lo; }      // straight from the registers (consumed only by the weight gradients)
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        };
        auto zero_dpre_tile = [&]() {   // K-major dpre tile (both planes, 16 KB): dims / rows nobody writes stay zero
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int i = 0; i < 4; ++i) st_shared_v4(xa + (uint32_t)(t * 4 + i) * 16u, z);
        };
        // coupling backward of the owned row for the `step` about to run: g -> direct path in place, head gradients -> the nets' CTAs
        // the saved layer input / head outputs the owner's coupling backward of `step` reads: issued a layer ahead, they are global
        // loads whose latency would otherwise sit between the partial sums and the next layer's first GEMM
        float pf_x = 0.f, pf_s = 0.f, pf_t = 0.f;
        auto owner_prefetch = [&](int step, int layer) {
            if (uk < nact_s[layer] && own_r < p.R) {
                const int d = actd[layer][uk];
                pf_x = __ldg(p.saved_x + ((size_t)step * p.R + own_r) * D + d);
                pf_s = __ldg(p.saved_st + ((size_t)(step * 2 + 0) * p.R + own_r) * D + d);
                pf_t = __ldg(p.saved_st + ((size_t)(step * 2 + 1) * p.R + own_r) * D + d);
            }
        };
        auto owner_coupling = [&](int step, int layer) {
            const bool on = uk < nact_s[layer];
            const int d = on ? actd[layer][uk] : 0;
            float ds = 0.f, dt = 0.f;
            if (on && own_r < p.R) {
                const float g = gs[urow * kXs + d];
                const float xv = pf_x, s = pf_s, tt = pf_t;
                const float gl = gls[urow];
                float dx;
                if (p.direction == 0) { const float e = expf(s); dx = g * e; dt = g; ds = g * xv * e + gl; }
                else { const float e = expf(-s); dx = g * e; dt = -g * e; ds = -g * (xv - tt) * e - gl; }
                ds *= (1.f - s * s);
                gs[urow * kXs + d] = dx;
            }
            // element (row own_n, dim d) of the K-major [64 rows][64 dims] bfloat16 planes of every CTA of the net
            const uint32_t off = (uint32_t)(own_n >> 3) * 1024u + (uint32_t)(own_n & 7) * 128u + (uint32_t)(((d >> 3) ^ (own_n & 7)) << 4) + (uint32_t)(d & 7) * 2u;
#pragma unroll
            for (int nn = 0; nn < 2; ++nn) {
                const float dv = nn == 0 ? ds : dt;
                const uint16_t h = to16<false>(dv), l = to16<false>(dv - from16<false>(h));
                if (on) {
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const uint32_t ra = map_to_cta(xa + off, nn * 4 + c4);
                        asm volatile("st.shared::cluster.b16 [%0], %1;" ::"r"(ra), "h"(h) : "memory");
                        asm volatile("st.shared::cluster.b16 [%0], %1;" ::"r"(ra + 8192), "h"(l) : "memory");
                    }
                    // dpre^T planes for the W2 weight gradient
                    uint16_t* gp = reinterpret_cast<uint16_t*>(p.dpreT) + ((size_t)((layer * 2 + nn) * 2) * kDp + d) * p.Rp + own_r;
                    gp[0] = h;
                    gp[(size_t)kDp * p.Rp] = l;
                }
                // bias gradient: sum over the 8 owned rows = 8 adjacent lanes (all lanes take part in the shuffles)
                float sum = dv;
                sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                sum += __shfl_xor_sync(0xffffffffu, sum, 4);
                if (on && urow == 0) atomicAdd(p.dparams + (size_t)(layer * 2 + nn) * p.blk + p.ob2 + d, sum);
            }
            worker_sync();
            if (t < kCluster) { fence_cluster(); mbar_arrive_remote(smem_u32(&bar_dpre), t); }
        };

        if (t < p.L) {
            uint64_t mb = 0;
            for (int d = 0; d < D; ++d) mb |= (uint64_t)(p.mask[(size_t)t * D + d] != 0.f) << d;
            mbits[t] = mb;
            int na = 0, np = 0;
            for (int d = 0; d < D; ++d) {
                if ((mb >> d) & 1) { if (np < kActMax) pasd[t][np++] = (uint8_t)d; }
                else if (na < kActMax) actd[t][na++] = (uint8_t)d;
            }
            nact_s[t] = na; npas_s[t] = np;
        }
        for (int i = t; i < 8 * D; i += kWorkers) {
            const int n = i / D, d = i - n * D;
            const int r = r0 + 8 * (int)rank + n;
            gs[n * kXs + d] = r < p.R ? p.dout[(size_t)r * D + d] : 0.f;
        }
        if (t < 8) { const int r = r0 + 8 * (int)rank + t; gls[t] = (p.dlogdet && r < p.R) ? p.dlogdet_scale * p.dlogdet[r] : 0.f; }
        if (p.zero_dpreT) {
            // dpreT [layer][net][plane][64 dims][Rp]: the owners only write the transformed dims of their rows; this CTA zeroes its 8 rows
            // of every dim / plane / net of the launch's layers once (16 bytes each), instead of a memset of the whole buffer in front
            // of the kernel.  Same-CTA stores to the same addresses follow many barriers later.
            const int nl = p.s_hi - p.s_lo;
            const int l_first = p.direction == 0 ? p.s_lo : p.L - p.s_hi;
            const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
            for (int i = t; i < nl * 4 * kDp; i += kWorkers) {      // (layer, net, plane) x dim
                uint16_t* q = reinterpret_cast<uint16_t*>(p.dpreT) + ((size_t)l_first * 4 * kDp + i) * p.Rp + r0 + 8 * (int)rank;
                *reinterpret_cast<uint4*>(q) = z4;
            }
        }
        zero_dpre_tile();
        worker_sync();
        // every CTA's dpre tile is zeroed before any owner writes into it: cluster-wide rendezvous of the worker warps via bar_part
        if (t < kCluster) { fence_cluster(); mbar_arrive_remote(smem_u32(&bar_part), t); }
        worker_wait(&bar_part, 0, true);
        owner_prefetch(p.s_hi - 1, p.direction == 0 ? p.s_hi - 1 : p.L - p.s_hi);
        owner_coupling(p.s_hi - 1, p.direction == 0 ? p.s_hi - 1 : p.L - p.s_hi);

        float v[32];
        for (int step = p.s_hi - 1; step >= p.s_lo; --step) {
            const int layer = p.direction == 0 ? step : p.L - 1 - step;
            const int it = p.s_hi - 1 - step;
            const uint32_t par = it & 1;
            const uint64_t mb = mbits[layer];
            const size_t sbatch = (size_t)(step * 2 + net) * 2;    // plane index base of the saved activations (indexed by step)
            const size_t gbatch = (size_t)(layer * 2 + net) * 2;   // ... of the gradient planes (indexed by layer, like the parameters)
            if (t == 0) BSTAMP(0);
            if (step > p.s_lo) owner_prefetch(step - 1, p.direction == 0 ? step - 1 : p.L - step);
            // ---------------- the head gradients of all 64 rows have been written into xa by their owners
            {
                uint4 sg[4];
                const uint4* sp = reinterpret_cast<const uint4*>(p.a1T + ((sbatch + 0) * p.H + f) * p.Rp + r0 + 32 * ch);
#pragma unroll
                for (int i = 0; i < 4; ++i) sg[i] = __ldg(sp + i);
                if (t == 0) {
                    mbar_wait_cluster(smem_u32(&bar_dpre), par);
                    BSTAMP(1);
                    fence_proxy_async();                               // the tile was written through the generic proxy (by peers)
                    mbar_arrive(smem_u32(&bar_xm));
                }
                // ---------------- bE2: dh1 = acc0 * lrelu'(a1) -> xa slice + global dh1T (exchange, weight gradients)
                worker_wait(&bar_acc[0], par, false);
                tcgen05_fence_after();
                if (t == 0) BSTAMP(2);
                load_acc(kAcc0, v);
                store_grad_slice(v, sg, nullptr, nullptr);
                tcgen05_fence_before();
                worker_sync();
                if (t == 0) { BSTAMP(3); mbar_arrive(smem_u32(&bar_own)); }
                copy_slice_to_global(xa, p.dh1T + gbatch * p.H * p.Rp, p.H, p.Rp, j * FS, r0, t);
                worker_sync();
                if (t < 4) {
                    fence_proxy_async();
                    fence_cluster();
                    mbar_arrive_remote(smem_u32(&bar_a0), net * 4 + t);
                    if (t == 0) BSTAMP(4);
                }
            }
            // ---------------- bE1: dh0 = acc1 * lrelu'(a0) -> xa slice (B operand of bG0) + global dh0T (weight gradients)
            {
                uint4 sg[4];
                const uint4* sp = reinterpret_cast<const uint4*>(p.a0T + ((sbatch + 0) * p.H + f) * p.Rp + r0 + 32 * ch);
#pragma unroll
                for (int i = 0; i < 4; ++i) sg[i] = __ldg(sp + i);
                worker_wait(&bar_acc[1], par, false);
                tcgen05_fence_after();
                if (t == 0) BSTAMP(6);
                load_acc(kAcc1, v);
                store_grad_slice(v, sg, nullptr, nullptr);
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bar_a1));
                worker_sync();      // kept for the weight gradients; off the critical path (bG0 only reads xa)
                copy_slice_to_global(xa, p.dh0T + gbatch * p.H * p.Rp, p.H, p.Rp, j * FS, r0, t);
                if (t == 0) BSTAMP(7);
            }
            // ---------------- bE0: partial input gradients of this CTA's feature slice -> the CTAs that own the rows (DSMEM)
            {
                worker_wait(&bar_acc[2], par, false);
                tcgen05_fence_after();
                if (t == 0) BSTAMP(8);
                if (lq < 2) {   // accumulator rows = flow dims
                    load_acc(kAcc2, v);
                    if (fl < D && ((mb >> fl) & 1)) {   // only the dims the nets read (mask = 1) receive this gradient
                        const int pi = __popcll(mb & ((1ull << fl) - 1ull));
                        if (pi < kActMax) {
                            const uint32_t slot = smem_u32(recv) + (uint32_t)(((((1 - (int)par) * kCluster + rank) * kActMax + pi) * 8) * 4);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint32_t ra = map_to_cta(slot, 4 * ch + q);
                                st_cluster_v4(ra, v[8 * q], v[8 * q + 1], v[8 * q + 2], v[8 * q + 3]);
                                st_cluster_v4(ra + 16, v[8 * q + 4], v[8 * q + 5], v[8 * q + 6], v[8 * q + 7]);
                            }
                        }
                    }
                }
                tcgen05_fence_before();
                zero_dpre_tile();                                      // bG0 is done with xa; the owners refill the tile for the next layer
                worker_sync();
                if (t < kCluster) { fence_cluster(); mbar_arrive_remote(smem_u32(&bar_part), t); }
                if (t == 0) BSTAMP(9);
            }
            // ---------------- owner: g[row][d] += sum over the 8 CTAs of the partials (conditioning dims), then the next layer's coupling
            {
                worker_wait(&bar_part, 1 - par, true);
                if (t == 0) BSTAMP(10);                                // completion it + 1 of bar_part (completion 0 was the start rendezvous)
                if (uk < npas_s[layer]) {
                    const int d = pasd[layer][uk];
                    const float* rv = recv + ((size_t)((1 - (int)par) * kCluster) * kActMax + uk) * 8 + urow;
                    float a = 0.f;
#pragma unroll
                    for (int c8 = 0; c8 < kCluster; ++c8) a += rv[(size_t)c8 * kActMax * 8];
                    gs[urow * kXs + d] += a;
                }
                worker_sync();
                if (step > p.s_lo) owner_coupling(step - 1, p.direction == 0 ? step - 1 : p.L - step);
                if (t == 0) BSTAMP(11);
            }
        }
        for (int i = t; i < 8 * D; i += kWorkers) {   // every CTA writes the rows it owns
            const int n = i / D, d = i - n * D;
            const int r = r0 + 8 * (int)rank + n;
            if (r < p.R) p.din[(size_t)r * D + d] = gs[n * kXs + d];
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 10) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

long long* debug_buffer();

// ---- host -----------------------------------------------------------------------------------------------------
// MHE_FUSED_DEBUG=1: a small device buffer receives %globaltimer stamps of CTA 0's phases; read it with mhe_fused_debug_read
static long long* g_dbg = nullptr;
long long* debug_buffer() {
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* e = getenv("MHE_FUSED_DEBUG");
        if (e && atoi(e) && cudaMalloc(&g_dbg, 64 * 64 * sizeof(long long)) == cudaSuccess) cudaMemset(g_dbg, 0, 64 * 64 * sizeof(long long));
        else g_dbg = nullptr;
    }
    return g_dbg;
}

static PlaneTensor pt4(const bf16* base, int cols, int rows, int batches) {
    PlaneTensor t;
    t.base = base; t.cols = cols; t.rows = rows; t.planes = 2; t.batches = batches;
    t.row_pitch = cols; t.plane_stride = (long)rows * cols; t.batch_stride = (long)2 * rows * cols;
    return t;
}

int pass_fwd(const FlowLayout& L, const float* params, const void* packed, const float* mask, const float* cp, const float* in, int R, int B,
             int direction, float* out, float* logdet, float* saved, void* workspace, cudaStream_t stream) {
    tcflow::Packed P(L, (bf16*)packed);
    FWs ws(workspace, L, R);
    const int Rp = padded_rows(R), tiles = tiles_of(R);
    FwdArgs a{};
    a.params = params; a.mask = mask; a.cp = cp; a.in = in; a.out = out; a.logdet = logdet;
    a.partial = ws.partial;
    { const char* e = getenv("MHE_FUSED_TWO_MMA"); a.two_mma = e ? atoi(e) : 15; }
    a.dbg = debug_buffer();
    a.R = R; a.Rp = Rp; a.B = B; a.D = L.D; a.H = L.H; a.L = L.L; a.direction = direction; a.save = saved ? 1 : 0; a.tiles = tiles;
    a.cp_ld = (long)L.L * 4 * L.H; a.blk = L.blk; a.ob2 = L.ob2;
    if (saved) {
        FSaved S(saved, L, R);
        a.saved_x = S.x_; a.saved_st = S.st_; a.a0T = S.a0_; a.a1T = S.a1_;
    } else {
        a.a0T = ws.a0T;
    }
    int st = MHE_OK;
    const CUtensorMap* mW0 = cached_map(pt4(P.w0, tcflow::kDp, L.H, L.L * 2), FS, &st);
    if (st != MHE_OK) return st;
    const CUtensorMap* mW1 = cached_map(pt4(P.w1, L.H, L.H, L.L * 2), FS, &st);
    if (st != MHE_OK) return st;
    const CUtensorMap* mW2 = cached_map(pt4(P.w2, L.H, tcflow::kDp, L.L * 2), 64, &st);
    if (st != MHE_OK) return st;
    const CUtensorMap* mA0 = cached_map(pt4(a.a0T, Rp, L.H, saved ? L.L * 2 : 2), 64, &st);
    if (st != MHE_OK) return st;

    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(flow_fwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess) {
            set_error("fused flow fwd: cannot raise dynamic shared memory to %d", kSmemBytes);
            return MHE_ERR_CUDA;
        }
        attr_set = true;
    }
    ProbeScope probe("fused flow fwd", stream);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(tiles * kCluster); cfg.blockDim = dim3(kThreadsF); cfg.dynamicSmemBytes = kSmemBytes; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, flow_fwd_fused_kernel, *mW0, *mW1, *mW2, *mA0, a) != cudaSuccess) {
        set_error("fused flow fwd: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return MHE_ERR_CUDA;
    }
    return check_launch("fused flow fwd");
}


// ---- small kernels around the backward ---------------------------------------------------------------------------------
// saved half planes [step][net][2][n] -> bfloat16 planes [layer][net][2][n] of the same values (operands of the weight-gradient GEMMs)
// for the steps [step0, step0 + gridDim.y / 2); 8 values per thread (16-byte loads and stores; n is a multiple of 8)
__global__ void __launch_bounds__(256) replane_by_layer_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, long n, int L, int direction, int step0) {
    const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i >= n) return;
    const int step = step0 + (blockIdx.y >> 1), net = blockIdx.y & 1;
    const int layer = direction == 0 ? step : L - 1 - step;
    const uint16_t* s16 = reinterpret_cast<const uint16_t*>(src) + (long)(step * 2 + net) * 2 * n;
    uint16_t* d16 = reinterpret_cast<uint16_t*>(dst) + (long)(layer * 2 + net) * 2 * n;
    const uint4 h4 = *reinterpret_cast<const uint4*>(s16 + i), l4 = *reinterpret_cast<const uint4*>(s16 + n + i);
    const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
    uint32_t oh[4], ol[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hw[k])), c = __half22float2(*reinterpret_cast<const __half2*>(&lw[k]));
        const float v0 = a.x + c.x, v1 = a.y + c.y;
        const __nv_bfloat162 hh = __floats2bfloat162_rn(v0, v1);
        const float2 hf = __bfloat1622float2(hh);
        const __nv_bfloat162 ll = __floats2bfloat162_rn(v0 - hf.x, v1 - hf.y);
        oh[k] = *reinterpret_cast<const uint32_t*>(&hh);
        ol[k] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    *reinterpret_cast<uint4*>(d16 + i) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
    *reinterpret_cast<uint4*>(d16 + n + i) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
}
// xmT[layer][net][plane][d][r] = bfloat16 planes of mask[layer][d] * x_step[r][d]  (zero for d >= D, r >= R)
__global__ void xm_transposed_kernel(const float* __restrict__ saved_x, const float* __restrict__ mask, int R, int Rp, int D, int L, int direction,
                                     bf16* __restrict__ xmT, int layer0) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int d = blockIdx.y, layer = layer0 + blockIdx.z;
    if (r >= Rp) return;
    const int step = direction == 0 ? layer : L - 1 - layer;
    float v = 0.f;
    if (r < R && d < D) v = saved_x[((size_t)step * R + r) * D + d] * mask[(size_t)layer * D + d];
    const uint16_t h = to16<false>(v), l = to16<false>(v - from16<false>(h));
    uint16_t* o = reinterpret_cast<uint16_t*>(xmT);
#pragma unroll
    for (int net = 0; net < 2; ++net) {
        const size_t base = ((size_t)(layer * 2 + net) * 2) * kDp * Rp + (size_t)d * Rp + r;
        o[base] = h;
        o[base + (size_t)kDp * Rp] = l;
    }
}
// dcp[b][(layer*4 + net*2 + jj)*H + f] += sum_s dh_jj[layer][net][f][s*B + b]   (jj = 0: dh0, 1: dh1; planes hi + lo)
// block: 32 features x 64 images; thread = (feature, 8 consecutive images), 16-byte plane loads; transposed through shared memory
// so that the rows of dcp receive 32 consecutive features (128 B) per store
__global__ void __launch_bounds__(256) dcp_from_planes_kernel(const bf16* __restrict__ dh0T, const bf16* __restrict__ dh1T, int R, int Rp, int B, int H,
                                                               float* __restrict__ dcp, long cp_ld, int ln0) {
    __shared__ float tile[32][65];
    const int f0 = blockIdx.x * 32, ln = ln0 + blockIdx.y, jj = blockIdx.z & 1, b0 = (blockIdx.z >> 1) * 64;       // ln = layer * 2 + net
    const uint16_t* src = reinterpret_cast<const uint16_t*>(jj == 0 ? dh0T : dh1T) + (size_t)ln * 2 * H * Rp;
    const int layer = ln >> 1, net = ln & 1;
    const int fl = threadIdx.x >> 3, bg = threadIdx.x & 7;
    const uint16_t* hi = src + (size_t)(f0 + fl) * Rp, *lo = hi + (size_t)H * Rp;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int b = b0 + bg * 8;
    if ((B & 7) == 0) {
        if (b < B)
            for (int r = b; r < R; r += B) {
                const uint4 h4 = *reinterpret_cast<const uint4*>(hi + r), l4 = *reinterpret_cast<const uint4*>(lo + r);
                const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    acc[k] += from16<false>((uint16_t)(hw[k >> 1] >> ((k & 1) * 16))) + from16<false>((uint16_t)(lw[k >> 1] >> ((k & 1) * 16)));
            }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (b + k < B)
                for (int r = b + k; r < R; r += B) acc[k] += from16<false>(hi[r]) + from16<false>(lo[r]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) tile[fl][bg * 8 + k] = acc[k];
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int bl = w * 8 + q;
        if (b0 + bl < B) dcp[(size_t)(b0 + bl) * cp_ld + (size_t)(layer * 4 + net * 2 + jj) * H + f0 + lane] += tile[lane][bl];
    }
}

static PlaneTensor ptk(const bf16* base, int cols, int rows, int batches) {   // [batches][2 planes][rows][cols]
    PlaneTensor t;
    t.base = base; t.cols = cols; t.rows = rows; t.planes = 2; t.batches = batches;
    t.row_pitch = cols; t.plane_stride = (long)rows * cols; t.batch_stride = (long)2 * rows * cols;
    return t;
}

static int cuda_ok(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return MHE_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return MHE_ERR_CUDA;
}

constexpr int kMaxChunks = 4;
// Keeps the weight-gradient GEMMs of a chunk off the SMs for a few microseconds after the chunk's data-gradient kernel ends, so that
// the NEXT chunk's cluster kernel (which needs 80 entirely free SMs) is placed first; otherwise it waits for the GEMM CTAs that won
// the race to drain (~12 us measured).
__global__ void head_start_kernel(unsigned ns) {
    const long long t0 = gtime();
    while (gtime() - t0 < (long long)ns) __nanosleep(256);
}
struct BwdAux {
    cudaStream_t stream[7] = {};     // [0..2]: weight gradients (W1, W0 + dcp, W2), [3]: re-planes, [4..6]: conditioning backward
    cudaEvent_t start = nullptr, replaned[kMaxChunks] = {}, chunk_done[kMaxChunks] = {}, dcp_done[kMaxChunks] = {}, gate[kMaxChunks] = {},
                wgrad_done[6] = {}, dcpp_ready[kMaxChunks] = {}, chunk_grads[kMaxChunks][6] = {};
    bool has_pending = false;
    int pending_n = 3;               // how many of wgrad_done[] the pending pass recorded (6 with the conditioning backward)
    int last_chunks = 0, last_L = 0, last_direction = 0;   // geometry of the last pass (join_chunk / chunk_layers)
    bool ok = false;
};
static BwdAux& bwd_aux() {   // one context per device: streams and events belong to the device that was current when they were created
    constexpr int kMaxDevices = 16;
    static BwdAux ctx[kMaxDevices];
    static bool tried_dev[kMaxDevices] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
    BwdAux& a = ctx[dev];
    bool& tried = tried_dev[dev];
    if (!tried) {
        tried = true;
        bool ok = true;
        auto ev = [&](cudaEvent_t* e) { ok = ok && cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess; };
        for (int i = 0; i < 7 && ok; ++i) ok = cudaStreamCreateWithFlags(&a.stream[i], cudaStreamNonBlocking) == cudaSuccess;
        for (int i = 0; i < 6; ++i) ev(&a.wgrad_done[i]);
        for (int i = 0; i < kMaxChunks; ++i) { ev(&a.replaned[i]); ev(&a.chunk_done[i]); ev(&a.dcp_done[i]); ev(&a.gate[i]); ev(&a.dcpp_ready[i]); for (int k = 0; k < 6; ++k) ev(&a.chunk_grads[i][k]); }
        ev(&a.start);
        a.ok = ok;
    }
    return a;
}
static int g_async_wgrad = 0;
void set_async_wgrad(int on) { g_async_wgrad = on; }
static bool async_wgrad() { return g_async_wgrad != 0; }
static int join_pending_on(cudaStream_t stream) {
    BwdAux& ax = bwd_aux();
    for (int i = 0; i < ax.pending_n; ++i)
        if (cudaStreamWaitEvent(stream, ax.wgrad_done[i], 0) != cudaSuccess) { set_error("join weight gradients: %s", cudaGetErrorString(cudaGetLastError())); return MHE_ERR_CUDA; }
    ax.has_pending = false;
    return MHE_OK;
}
int join(cudaStream_t stream) {
    BwdAux& ax = bwd_aux();
    if (!ax.ok || !ax.has_pending) return MHE_OK;
    return join_pending_on(stream);
}
// The pass is cut into chunks of consecutive layers (MHE_FUSED_BWD_CHUNKS, default 2): the data-gradient kernel of chunk c + 1 runs
// on 80 of the 148 SMs while the weight-gradient GEMMs and the dcp sums of chunk c fill the others, instead of all of them
// queueing up behind the last layer.
static int bwd_chunks(int L) {
    static int n = -1;
    if (n < 0) {
        const char* e = getenv("MHE_FUSED_BWD_CHUNKS");
        n = e ? atoi(e) : 2;
        if (n < 1) n = 1;
        if (n > kMaxChunks) n = kMaxChunks;
    }
    return n < L ? n : L;
}

int chunk_count(int L) { return bwd_chunks(L); }
// layers [l0, l0 + nl) of chunk c (the pass walks the steps downwards: chunk 0 holds the LAST steps)
void chunk_layers(int L, int direction, int c, int* l0, int* nl) {
    const int n = bwd_chunks(L);
    const int lo = L - (L * (c + 1)) / n, hi = L - (L * c) / n;
    *l0 = direction == 0 ? lo : L - hi;
    *nl = hi - lo;
}
int join_chunk(cudaStream_t stream, int c) {
    BwdAux& ax = bwd_aux();
    if (!ax.ok || c < 0 || c >= ax.last_chunks) { set_error("join_chunk: no such chunk in the last pass"); return MHE_ERR_INVALID_ARG; }
    for (int i = 0; i < ax.pending_n; ++i)
        if (cudaStreamWaitEvent(stream, ax.chunk_grads[c][i], 0) != cudaSuccess) { set_error("join chunk: %s", cudaGetErrorString(cudaGetLastError())); return MHE_ERR_CUDA; }
    return MHE_OK;
}

static int g_prepared = 0;
void set_wgrad_operands_prepared(int on) { g_prepared = on; }

// The operands of the weight gradients that only depend on the forward pass (bfloat16 planes of the saved activations by layer, the
// masked layer inputs, the zeroed dpreT planes), enqueued on `stream`: callable any time after the forward pass, e.g. on a side stream
// while the loss is computed.  The caller orders it before pass_bwd, which then skips them (mhe_flow_set_async bit 3).
int pass_bwd_prepare(const FlowLayout& L, const float* mask, const float* saved, int R, int direction, void* workspace, cudaStream_t stream) {
    BWs ws(workspace, L, R);
    FSaved S(const_cast<float*>(saved), L, R);
    const int Rp = padded_rows(R);
    const long nact = (long)L.H * Rp;
    MHE_TRY(cuda_ok(cudaMemsetAsync(ws.dpreT, 0, (size_t)L.L * 4 * kDp * Rp * 2, stream), "memset dpreT"));
    dim3 grid(cdiv((int)(nact / 8), 256), L.L * 2);
    replane_by_layer_kernel<<<grid, 256, 0, stream>>>(S.a0_, ws.a0b, nact, L.L, direction, 0);
    MHE_TRY(check_launch("replane a0"));
    replane_by_layer_kernel<<<grid, 256, 0, stream>>>(S.a1_, ws.a1b, nact, L.L, direction, 0);
    MHE_TRY(check_launch("replane a1"));
    dim3 gx(cdiv(Rp, 128), kDp, L.L);
    xm_transposed_kernel<<<gx, 128, 0, stream>>>(S.x_, mask, R, Rp, L.D, L.L, direction, ws.xmT, 0);
    return check_launch("xm transposed");
}

static unsigned head_start_ns() {
    static int ns = -1;
    if (ns < 0) { const char* e = getenv("MHE_FUSED_HEAD_START_NS"); ns = e ? atoi(e) : 6000; if (ns < 0) ns = 0; }
    return (unsigned)ns;
}

int pass_bwd(const FlowLayout& L, const float* params, const void* packed, const float* mask, const float* saved, int R, int B, int direction,
             const float* dout, const float* dlogdet, float dlogdet_scale, float* din, float* dparams, float* dcp, void* workspace,
             cudaStream_t stream, const CondBwd* cond) {
    (void)params;
    tcflow::Packed P(L, (bf16*)packed);
    BWs ws(workspace, L, R);
    FSaved S(const_cast<float*>(saved), L, R);
    const int Rp = padded_rows(R), tiles = tiles_of(R);
    const long nact = (long)L.H * Rp;
    BwdArgs a{};
    a.mask = mask; a.saved_x = S.x_; a.saved_st = S.st_; a.a0T = S.a0_; a.a1T = S.a1_;
    a.dout = dout; a.dlogdet = dlogdet; a.dlogdet_scale = dlogdet_scale; a.din = din; a.dparams = dparams;
    a.dh1T = ws.dh1T; a.dh0T = ws.dh0T; a.dpreT = ws.dpreT; a.partial = ws.partial; a.dbg = getenv("MHE_FUSED_DEBUG_BWD") ? debug_buffer() : nullptr;
    a.R = R; a.Rp = Rp; a.D = L.D; a.H = L.H; a.L = L.L; a.direction = direction; a.tiles = tiles;
    { const char* e = getenv("MHE_FUSED_TWO_MMA"); a.two_mma = e ? (atoi(e) != 0) : 1; }
    a.blk = L.blk; a.ob2 = L.ob2;
    int st = MHE_OK;
    const CUtensorMap* mW0 = cached_map(pt4(P.w0b, tcflow::kDp, L.H, L.L * 2), 64, &st);
    if (st != MHE_OK) return st;
    const CUtensorMap* mW1 = cached_map(pt4(P.w1b, L.H, L.H, L.L * 2), 64, &st);
    if (st != MHE_OK) return st;
    const CUtensorMap* mW2 = cached_map(pt4(P.w2b, L.H, tcflow::kDp, L.L * 2), 64, &st);
    if (st != MHE_OK) return st;
    const CUtensorMap* mDh1 = cached_map(pt4(ws.dh1T, Rp, L.H, L.L * 2), 64, &st);
    if (st != MHE_OK) return st;
    // only the active dims of dpreT are written by the kernel
    const bool prepared = g_prepared != 0;            // pass_bwd_prepare already ran on this saved block / workspace (caller-ordered)
    a.zero_dpreT = prepared ? 0 : 1;                  // (pass_bwd_prepare memsets the whole buffer)

    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(flow_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytesBwd) != cudaSuccess) {
            set_error("fused flow bwd: cannot raise dynamic shared memory to %d", kSmemBytesBwd);
            return MHE_ERR_CUDA;
        }
        attr_set = true;
    }
    BwdAux& ax = bwd_aux();
    const bool par = ax.ok;
    const int nchunk = bwd_chunks(L.L);
    cudaStream_t s0 = par ? ax.stream[0] : stream, s1 = par ? ax.stream[1] : stream, s2 = par ? ax.stream[2] : stream,
                 sr = par ? ax.stream[3] : stream;
    // steps [lo, hi) of chunk c (the pass walks the steps downwards); its layers are consecutive too
    auto chunk_lo = [&](int c) { return L.L - (L.L * (c + 1)) / nchunk; };
    auto chunk_hi = [&](int c) { return L.L - (L.L * c) / nchunk; };
    auto first_layer = [&](int c) { return direction == 0 ? chunk_lo(c) : L.L - chunk_hi(c); };
    auto launch_chunk = [&](int c) -> int {
        BwdArgs ac = a;
        ac.s_lo = chunk_lo(c); ac.s_hi = chunk_hi(c);
        if (c > 0) ac.dout = din;                      // every CTA reads the rows it owns at the start and writes them at the end: in place
        ProbeScope probe("fused flow bwd", stream);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(tiles * kCluster); cfg.blockDim = dim3(kThreadsF); cfg.dynamicSmemBytes = kSmemBytesBwd; cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        if (cudaLaunchKernelEx(&cfg, flow_bwd_fused_kernel, *mW0, *mW1, *mW2, *mDh1, ac) != cudaSuccess) {
            set_error("fused flow bwd: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            return MHE_ERR_CUDA;
        }
        return check_launch("fused flow bwd");
    };
    // operands of the weight gradients (bfloat16 planes of the saved activations by layer, masked inputs) of chunk c: they only
    // depend on the forward pass, so they run on their own stream while the data-gradient kernels run
    auto launch_replane = [&](int c) -> int {
        const int lo = chunk_lo(c), n = chunk_hi(c) - lo;
        dim3 grid(cdiv((int)(nact / 8), 256), n * 2);
        replane_by_layer_kernel<<<grid, 256, 0, sr>>>(S.a0_, ws.a0b, nact, L.L, direction, lo);
        MHE_TRY(check_launch("replane a0"));
        replane_by_layer_kernel<<<grid, 256, 0, sr>>>(S.a1_, ws.a1b, nact, L.L, direction, lo);
        MHE_TRY(check_launch("replane a1"));
        dim3 gx(cdiv(Rp, 128), kDp, n);
        xm_transposed_kernel<<<gx, 128, 0, sr>>>(S.x_, mask, R, Rp, L.D, L.L, direction, ws.xmT, first_layer(c));
        MHE_TRY(check_launch("xm transposed"));
        if (par) MHE_TRY(cuda_ok(cudaEventRecord(ax.replaned[c], sr), "replaned"));
        return MHE_OK;
    };
    // weight gradients and dcp sums of chunk c: mutually independent, off the critical path of the data gradients
    auto launch_wgrads = [&](int c) -> int {
        const int l0 = first_layer(c), nb = (chunk_hi(c) - chunk_lo(c)) * 2;
        const size_t po = (size_t)l0 * 2 * L.blk;                                  // parameter blocks of the chunk's first layer
        const size_t ao = (size_t)l0 * 4 * L.H * Rp, so = (size_t)l0 * 4 * kDp * Rp;   // plane offsets (elements) by layer
        if (par) {
            MHE_TRY(cuda_ok(cudaEventRecord(ax.chunk_done[c], stream), "fork"));
            cudaEvent_t go = ax.chunk_done[c];
            if (c + 1 < nchunk && head_start_ns() > 0) {   // another cluster kernel follows: let it take its SMs first
                MHE_TRY(cuda_ok(cudaStreamWaitEvent(sr, ax.chunk_done[c], 0), "fork"));
                head_start_kernel<<<1, 1, 0, sr>>>(head_start_ns());
                MHE_TRY(check_launch("head start"));
                MHE_TRY(cuda_ok(cudaEventRecord(ax.gate[c], sr), "fork"));
                go = ax.gate[c];
            }
            for (int i = 0; i < 3; ++i) {
                MHE_TRY(cuda_ok(cudaStreamWaitEvent(ax.stream[i], go, 0), "fork"));
                MHE_TRY(cuda_ok(cudaStreamWaitEvent(ax.stream[i], ax.replaned[c], 0), "fork"));
            }
        }
        {   // dcp += sums over the hypotheses of dh0 / dh1 (the conditioning backward waits for it)
            dim3 grid(L.H / 32, nb, 2 * cdiv(B, 64));
            dcp_from_planes_kernel<<<grid, 256, 0, s1>>>(ws.dh0T, ws.dh1T, R, Rp, B, L.H, dcp, (long)L.L * 4 * L.H, l0 * 2);
            MHE_TRY(check_launch("dcp from planes"));
            if (par) MHE_TRY(cuda_ok(cudaEventRecord(ax.dcp_done[c], s1), "dcp done"));
        }
        if (cond && par) {   // conditioning backward of the chunk's layers: only needs their dcp
            const int nl = chunk_hi(c) - chunk_lo(c);
            MHE_TRY(cuda_ok(cudaStreamWaitEvent(ax.stream[4], ax.dcp_done[c], 0), "fork cond"));
            MHE_TRY(tcflow::cond_bwd_dcp_planes(L, dcp, B, cond->ws, l0, nl, ax.stream[4]));
            MHE_TRY(cuda_ok(cudaEventRecord(ax.dcpp_ready[c], ax.stream[4]), "fork cond"));
            MHE_TRY(cuda_ok(cudaStreamWaitEvent(ax.stream[5], ax.dcpp_ready[c], 0), "fork cond"));
            MHE_TRY(cuda_ok(cudaStreamWaitEvent(ax.stream[6], ax.dcp_done[c], 0), "fork cond"));
            MHE_TRY(tcflow::cond_bwd_layers(L, packed, dcp, B, dparams, cond->dfeat, cond->ws, l0, nl, ax.stream[6], ax.stream[4], ax.stream[5]));
        }
        {   // dW1 [out][in] += dh1T . a0T^T  (contraction over the rows)
            GemmShape g{L.H, L.H, Rp, nb, 1, 1, 1};
            MHE_TRY(tcflow::wgrad_kmajor(ptk(ws.dh1T + ao, Rp, L.H, nb), ptk(ws.a0b + ao, Rp, L.H, nb), g, dparams + po + L.oW1, L.H, (long)L.blk, L.H, 0, s0,
                                         "fused wgrad W1"));
        }
        {   // dW0 [feat][d] += dh0T . xmT^T
            GemmShape g{L.H, kDp, Rp, nb, 1, 1, 1};
            MHE_TRY(tcflow::wgrad_kmajor(ptk(ws.dh0T + ao, Rp, L.H, nb), ptk(ws.xmT + so, Rp, kDp, nb), g, dparams + po + L.oW0, L.D, (long)L.blk, L.D, 0, s1,
                                         "fused wgrad W0"));
        }
        {   // dW2 [d][feat] += dpreT . a1T^T, computed as (a1T . dpreT^T)[feat][d] and stored transposed
            GemmShape g{L.H, kDp, Rp, nb, 1, 1, 1};
            MHE_TRY(tcflow::wgrad_kmajor(ptk(ws.a1b + ao, Rp, L.H, nb), ptk(ws.dpreT + so, Rp, kDp, nb), g, dparams + po + L.oW2, L.H, (long)L.blk, L.D, 1, s2,
                                         "fused wgrad W2"));
        }
        if (par) {   // "every gradient of this chunk's layers is complete" (join_chunk: bucketed all-reduce)
            for (int i = 0; i < 3; ++i) MHE_TRY(cuda_ok(cudaEventRecord(ax.chunk_grads[c][i], ax.stream[i]), "chunk grads"));
            if (cond) for (int i = 3; i < 6; ++i) MHE_TRY(cuda_ok(cudaEventRecord(ax.chunk_grads[c][i], ax.stream[i + 1]), "chunk grads"));
        }
        return MHE_OK;
    };

    if (par) MHE_TRY(cuda_ok(cudaEventRecord(ax.start, stream), "fork replane"));
    if (cond && !par) { set_error("fused pass + conditioning backward: internal streams unavailable"); return MHE_ERR_CUDA; }
    MHE_TRY(launch_chunk(0));
    if (cond) {
        MHE_TRY(cuda_ok(cudaStreamWaitEvent(ax.stream[4], ax.start, 0), "fork cond"));
        MHE_TRY(tcflow::cond_bwd_feat_planes(L, cond->feat, B, cond->ws, ax.stream[4]));
        if (cond->dfeat && !tcflow::dfeat_is_zero()) {
            MHE_TRY(cuda_ok(cudaStreamWaitEvent(ax.stream[5], ax.start, 0), "fork cond"));
            MHE_TRY(cuda_ok(cudaMemsetAsync(cond->dfeat, 0, (size_t)B * L.C * sizeof(float), ax.stream[5]), "memset dfeat"));
        }
    }
    // (launched after the first kernel so that its clusters' CTAs get their SMs first; the re-planes fill the remaining ones)
    if (par) MHE_TRY(cuda_ok(cudaStreamWaitEvent(sr, ax.start, 0), "fork replane"));
    for (int c = 0; c < nchunk; ++c) {
        if (!prepared) MHE_TRY(launch_replane(c));
        else if (par) MHE_TRY(cuda_ok(cudaEventRecord(ax.replaned[c], sr), "replaned"));
    }
    for (int c = 0; c < nchunk; ++c) {
        if (c > 0) MHE_TRY(launch_chunk(c));
        MHE_TRY(launch_wgrads(c));
    }
    if (par) {
        for (int c = 0; c < nchunk; ++c) MHE_TRY(cuda_ok(cudaStreamWaitEvent(stream, ax.dcp_done[c], 0), "join dcp"));
        for (int i = 0; i < 3; ++i) MHE_TRY(cuda_ok(cudaEventRecord(ax.wgrad_done[i], ax.stream[i]), "join"));
        if (cond) for (int i = 3; i < 6; ++i) MHE_TRY(cuda_ok(cudaEventRecord(ax.wgrad_done[i], ax.stream[i + 1]), "join"));
        ax.pending_n = cond ? 6 : 3;
        ax.last_chunks = nchunk; ax.last_L = L.L; ax.last_direction = direction;
        // the re-plane stream is joined through the weight-gradient streams (they waited for replaned[c])
        if (async_wgrad()) ax.has_pending = true;    // the caller joins with mhe_flow_join() before reading dparams
        else MHE_TRY(join_pending_on(stream));
    }
    return MHE_OK;
}

}  // namespace fused
}  // namespace mhe

extern "C" int mhe_fused_debug_read(long long* host, int n) {
    if (!mhe::fused::g_dbg) return MHE_ERR_UNSUPPORTED;
    if (n > 64 * 64) n = 64 * 64;
    return cudaMemcpy(host, mhe::fused::g_dbg, n * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? MHE_OK : MHE_ERR_CUDA;
}
