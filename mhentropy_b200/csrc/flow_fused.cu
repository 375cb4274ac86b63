// Cluster-fused flow passes: one kernel per pass, all coupling layers (see flow_fused.cuh for the mapping).
// Reference: hand/flows.py:75-122 (_nets), :210-217 (forward_p), :219-227 (backward_p).
#include "flow_fused.cuh"

namespace mhe {
namespace fused {
using namespace tc;
typedef __nv_bfloat16 bf16;

constexpr int kThreadsF = 192;                 // warps 0-3: workers (epilogues, coupling), warp 4: TMA producer, warp 5: MMA issuer
constexpr int kSlots = 3;
constexpr int kSlotA = 16384;                  // one plane of an A block: [128][64] 16-bit
constexpr int kSlotB = 8192;                   // one plane of a B block: [64][64] 16-bit
constexpr int kSlotBytes = 2 * kSlotA + 2 * kSlotB;   // 48 KB
constexpr int kXaBytes = 32768;                // xm planes (2 x 8 KB) or a1 planes (2 x 16 KB)
constexpr int kXs = 65;                        // row stride of the fp32 x tile (odd: conflict-free column walks)
constexpr int kSmemBytes = kSlots * kSlotBytes + kXaBytes + NT * kXs * 4 + 1024 /*align*/ + 512 /*barriers, logdet*/;
constexpr int kTmemCols = 256;                 // acc0 [0,64) acc1 [64,128) acc2 [128,192)

bool supported(const FlowLayout& L, int R) {
    static int max_rows = -1;
    if (max_rows < 0) {
        const char* e = getenv("MHE_FUSED_MAX_ROWS");
        max_rows = e ? atoi(e) : 4096;
    }
    return L.D <= 64 && L.H == 512 && L.C % 8 == 0 && R <= max_rows;
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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {        // acquire at cluster scope
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

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

// 8 fp32 -> 8 hi + 8 lo 16-bit values packed as two uint4
template <bool F16>
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint16_t h0 = to16<F16>(v[2 * j]), h1 = to16<F16>(v[2 * j + 1]);
        const uint16_t l0 = to16<F16>(v[2 * j] - from16<F16>(h0)), l1 = to16<F16>(v[2 * j + 1] - from16<F16>(h1));
        h[j] = (uint32_t)h0 | ((uint32_t)h1 << 16);
        l[j] = (uint32_t)l0 | ((uint32_t)l1 << 16);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float fast_tanh_f(float x) {   // 1 - 2/(e^{2x}+1); same form as the per-GEMM path
    const float e = __expf(2.f * x);
    return 1.f - __fdividef(2.f, e + 1.f);
}

struct FwdArgs {
    const float* params; const float* mask; const float* cp; const float* in;
    float* out; float* logdet;
    float* saved_x; float* saved_st;           // NULL when nothing is saved
    bf16* a0T; bf16* a1T;                        // a0T: [(L or 1)][2][2][H][Rp]; a1T: saved only
    float* partial;
    int R, Rp, B, D, H, L, direction, save, tiles;
    long cp_ld;
    size_t blk, ob2;
};

// MMA issue helpers: one 64-deep k-block = 4 UMMA_K steps x 3 plane pairs
//   K-major operand: +32 B per step; MN-major operand: +2048 B per step (descriptor units of 16 B)
template <bool B_MN>
__device__ __forceinline__ void issue_kblock(uint32_t tmem_d, uint32_t a_addr, uint32_t a_plane, uint32_t b_addr, uint32_t b_plane,
                                             uint32_t idesc, uint32_t& accumulate) {
    constexpr uint32_t kHi = (uint32_t)((1024u >> 4) | (1u << 14) | (2u << 29));   // SBO = 1024 | version 1 | SWIZZLE_128B
    constexpr uint32_t kLoA = (16u >> 4) << 16;                                    // K-major: LBO unused
    constexpr uint32_t kLoB = B_MN ? ((8192u >> 4) << 16) : ((16u >> 4) << 16);
    constexpr uint32_t kStepB = B_MN ? (2048u >> 4) : (32u >> 4);
    const uint32_t a0 = kLoA | (a_addr >> 4), b0 = kLoB | (b_addr >> 4);
#pragma unroll
    for (int pr = 0; pr < 3; ++pr) {   // (hi,hi) (hi,lo) (lo,hi)
        const uint32_t a = a0 + (pr == 2 ? (a_plane >> 4) : 0u);
        const uint32_t b = b0 + (pr == 1 ? (b_plane >> 4) : 0u);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            umma_bf16_lohi(tmem_d, a + ks * 2u, b + ks * kStepB, kHi, idesc, accumulate);
            accumulate = 1;
        }
    }
}

__global__ void __launch_bounds__(kThreadsF, 1)
flow_fwd_fused_kernel(const __grid_constant__ CUtensorMap mapW0, const __grid_constant__ CUtensorMap mapW1,
                      const __grid_constant__ CUtensorMap mapW2, const __grid_constant__ CUtensorMap mapA0, FwdArgs p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[kSlots], bar_empty[kSlots], bar_acc[3], bar_xm, bar_a1, bar_a0, bar_part;
    __shared__ uint32_t tmem_slot;
    __shared__ float lds[NT];

    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t ring = smem0, xa = smem0 + kSlots * kSlotBytes;
    float* xs = reinterpret_cast<float*>(smem_raw + (smem0 - smem_u32(smem_raw)) + kSlots * kSlotBytes + kXaBytes);   // [NT][kXs]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int net = rank >> 2, j = rank & 3;
    const int tile = blockIdx.x / kCluster;
    const int r0 = tile * NT;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kSlots; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
        for (int s = 0; s < 3; ++s) mbar_init(smem_u32(&bar_acc[s]), 1);
        mbar_init(smem_u32(&bar_xm), 128);
        mbar_init(smem_u32(&bar_a1), 128);
        mbar_init(smem_u32(&bar_a0), 4);
        mbar_init(smem_u32(&bar_part), kCluster);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW2) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA0) : "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    cluster_sync_all();            // every CTA's barriers are initialised before any remote arrival
    const uint32_t tmem = tmem_slot;

    if (warp == 4) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t q = 0;        // job counter: slot = q % kSlots, use = q / kSlots
            auto acquire = [&](uint32_t bytes) -> uint32_t {   // wait for the slot, arm its full barrier
                const uint32_t s = q % kSlots, use = q / kSlots;
                mbar_wait(smem_u32(&bar_empty[s]), (use & 1) ^ 1);
                mbar_expect_tx(smem_u32(&bar_full[s]), bytes);
                ++q;
                return s;
            };
            for (int step = 0; step < p.L; ++step) {
                const int layer = p.direction == 0 ? step : p.L - 1 - step;
                const int wb = layer * 2 + net;                       // batch index into the packed weights
                const int ab = (p.save ? step * 2 : 0) + net;         // batch index into a0T
                {   // W0 slice [128][64], both planes
                    const uint32_t s = acquire(2 * kSlotA), full = smem_u32(&bar_full[s]), base = ring + s * kSlotBytes;
                    tma_load_4d(base, &mapW0, full, 0, j * FS, 0, wb);
                    tma_load_4d(base + kSlotA, &mapW0, full, 0, j * FS, 1, wb);
                }
                uint32_t slot_kb[2];
                for (int kb = 0; kb < 2; ++kb) {   // weights of the first two k-blocks do not wait for the activations
                    const uint32_t s = acquire(kSlotBytes), full = smem_u32(&bar_full[s]), base = ring + s * kSlotBytes;
                    slot_kb[kb] = s;
                    tma_load_4d(base, &mapW1, full, kb * 64, j * FS, 0, wb);
                    tma_load_4d(base + kSlotA, &mapW1, full, kb * 64, j * FS, 1, wb);
                }
                mbar_wait_cluster(smem_u32(&bar_a0), step & 1);       // the 4 slices of a0T are in global memory
                fence_proxy_async();
                for (int kb = 0; kb < 2; ++kb) {
                    const uint32_t s = slot_kb[kb], full = smem_u32(&bar_full[s]), base = ring + s * kSlotBytes + 2 * kSlotA;
                    tma_load_4d(base, &mapA0, full, r0, kb * 64, 0, ab);
                    tma_load_4d(base + kSlotB, &mapA0, full, r0, kb * 64, 1, ab);
                }
                for (int kb = 2; kb < p.H / 64; ++kb) {
                    const uint32_t s = acquire(kSlotBytes), full = smem_u32(&bar_full[s]), base = ring + s * kSlotBytes;
                    tma_load_4d(base, &mapW1, full, kb * 64, j * FS, 0, wb);
                    tma_load_4d(base + kSlotA, &mapW1, full, kb * 64, j * FS, 1, wb);
                    tma_load_4d(base + 2 * kSlotA, &mapA0, full, r0, kb * 64, 0, ab);
                    tma_load_4d(base + 2 * kSlotA + kSlotB, &mapA0, full, r0, kb * 64, 1, ab);
                }
                {   // W2 [64 d][128 feature slice] as two k-blocks per plane
                    const uint32_t s = acquire(2 * kSlotA), full = smem_u32(&bar_full[s]), base = ring + s * kSlotBytes;
#pragma unroll
                    for (int pl = 0; pl < 2; ++pl)
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk)
                            tma_load_4d(base + pl * kSlotA + kk * 8192, &mapW2, full, j * FS + kk * 64, 0, pl, wb);
                }
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idescK = instr_desc(NT, false, false, true, true);    // B K-major (xm)
            constexpr uint32_t idescMN = instr_desc(NT, false, true, true, true);    // B MN-major (a0T blocks, a1T)
            uint32_t q = 0;
            auto wait_full = [&]() -> uint32_t {
                const uint32_t s = q % kSlots, use = q / kSlots;
                mbar_wait(smem_u32(&bar_full[s]), use & 1);
                ++q;
                return s;
            };
            for (int step = 0; step < p.L; ++step) {
                const uint32_t par = step & 1;
                {   // G0
                    const uint32_t s = wait_full();
                    mbar_wait(smem_u32(&bar_xm), par);
                    tcgen05_fence_after();
                    uint32_t acc = 0;
                    issue_kblock<false>(tmem + 0, ring + s * kSlotBytes, kSlotA, xa, 8192, idescK, acc);
                    tcgen05_commit(smem_u32(&bar_empty[s]));
                    tcgen05_commit(smem_u32(&bar_acc[0]));
                }
                {   // G1
                    uint32_t acc = 0;
                    for (int kb = 0; kb < p.H / 64; ++kb) {
                        const uint32_t s = wait_full();
                        tcgen05_fence_after();
                        const uint32_t base = ring + s * kSlotBytes;
                        issue_kblock<true>(tmem + 64, base, kSlotA, base + 2 * kSlotA, kSlotB, idescMN, acc);
                        tcgen05_commit(smem_u32(&bar_empty[s]));
                    }
                    tcgen05_commit(smem_u32(&bar_acc[1]));
                }
                {   // G2
                    const uint32_t s = wait_full();
                    mbar_wait(smem_u32(&bar_a1), par);
                    tcgen05_fence_after();
                    uint32_t acc = 0;
                    const uint32_t base = ring + s * kSlotBytes;
                    issue_kblock<true>(tmem + 128, base, kSlotA, xa, 16384, idescMN, acc);
                    issue_kblock<true>(tmem + 128, base + 8192, kSlotA, xa + 8192, 16384, idescMN, acc);
                    tcgen05_commit(smem_u32(&bar_empty[s]));
                    tcgen05_commit(smem_u32(&bar_acc[2]));
                }
            }
        }
    } else {
        // ===================== workers (128 threads) =====================
        const int t = threadIdx.x;
        const int f = j * FS + t;                                  // feature inside the net (TMEM lane = t)
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        const int D = p.D;
        const int xn = t & 63, xh = t >> 6;                        // xm writer: row, half of the 64 padded dims
        const int uq = t & 15, ug = t >> 4;                        // coupling: rows 4*uq.., dims ug + 8 i

        auto write_xm = [&](const float* mrow) {   // xm = mask * x as K-major split planes [64 rows][64 dims]
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int cc = xh * 4 + c;
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int d = cc * 8 + e;
                    v[e] = (d < D && mrow) ? xs[xn * kXs + d] * __ldg(mrow + d) : 0.f;
                }
                uint4 hi, lo;
                split8<true>(v, hi, lo);
                const uint32_t off = (uint32_t)(xn >> 3) * 1024u + (uint32_t)(xn & 7) * 128u + (uint32_t)((cc ^ (xn & 7)) << 4);
                st_shared_v4(xa + off, hi);
                st_shared_v4(xa + 8192 + off, lo);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(smem_u32(&bar_xm));
        };

        // load the row tile
        for (int i = t; i < NT * D; i += 128) {
            const int n = i / D, d = i - n * D;
            xs[n * kXs + d] = (r0 + n < p.R) ? p.in[(size_t)(r0 + n) * D + d] : 0.f;
        }
        if (t < NT) lds[t] = 0.f;
        worker_sync();
        if (p.save && rank == 0) {
            const int nvalid = min(NT, p.R - r0) * D;
            for (int i = t; i < nvalid; i += 128) { const int n = i / D, d = i - n * D; p.saved_x[(size_t)r0 * D + i] = xs[n * kXs + d]; }
        }
        write_xm(p.mask + (size_t)(p.direction == 0 ? 0 : p.L - 1) * D);

        for (int step = 0; step < p.L; ++step) {
            const int layer = p.direction == 0 ? step : p.L - 1 - step;
            const uint32_t par = step & 1;
            const float* mrow = p.mask + (size_t)layer * D;
            const size_t abatch = (size_t)((p.save ? step * 2 : 0) + net) * 2;   // plane index base into a0T / a1T
            float v[64], c[64];
            // ---------------- E0: a0 = lrelu(acc0 + cp0) -> global a0T (gathered by the net's CTAs)
            {
                const float* cpb = p.cp + (size_t)(layer * 4 + net * 2 + 0) * p.H + f;
                int img = r0 % p.B;
#pragma unroll
                for (int n = 0; n < NT; ++n) { c[n] = __ldg(cpb + (size_t)img * p.cp_ld); if (++img == p.B) img = 0; }
                mbar_wait(smem_u32(&bar_acc[0]), par);
                tcgen05_fence_after();
                tmem_ld64(tmem + lane_base + 0, v);
#pragma unroll
                for (int n = 0; n < NT; ++n) v[n] = lrelu(v[n] + c[n]);
                uint4* ghi = reinterpret_cast<uint4*>(p.a0T + ((abatch + 0) * p.H + f) * p.Rp + r0);
                uint4* glo = reinterpret_cast<uint4*>(p.a0T + ((abatch + 1) * p.H + f) * p.Rp + r0);
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    uint4 hi, lo;
                    split8<true>(v + 8 * cc, hi, lo);
                    ghi[cc] = hi;
                    glo[cc] = lo;
                }
                fence_proxy_async();
                fence_cluster();
                worker_sync();
                if (t < 4) mbar_arrive_remote(smem_u32(&bar_a0), net * 4 + t);
            }
            // ---------------- E1: a1 = lrelu(acc1 + cp1) -> shared (B operand of G2, MN-major) [+ global when saving]
            {
                const float* cpb = p.cp + (size_t)(layer * 4 + net * 2 + 1) * p.H + f;
                int img = r0 % p.B;
#pragma unroll
                for (int n = 0; n < NT; ++n) { c[n] = __ldg(cpb + (size_t)img * p.cp_ld); if (++img == p.B) img = 0; }
                mbar_wait(smem_u32(&bar_acc[1]), par);
                tcgen05_fence_after();
                tmem_ld64(tmem + lane_base + 64, v);
#pragma unroll
                for (int n = 0; n < NT; ++n) v[n] = lrelu(v[n] + c[n]);
                uint4* ghi = p.save ? reinterpret_cast<uint4*>(p.a1T + ((abatch + 0) * p.H + f) * p.Rp + r0) : nullptr;
                uint4* glo = p.save ? reinterpret_cast<uint4*>(p.a1T + ((abatch + 1) * p.H + f) * p.Rp + r0) : nullptr;
                const uint32_t rowoff = (uint32_t)(t >> 6) * 8192u + (uint32_t)(t & 63) * 128u;   // k-row t of the MN-major tile
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    uint4 hi, lo;
                    split8<true>(v + 8 * cc, hi, lo);
                    const uint32_t off = rowoff + (uint32_t)((cc ^ (t & 7)) << 4);
                    st_shared_v4(xa + off, hi);
                    st_shared_v4(xa + 16384 + off, lo);
                    if (ghi) { ghi[cc] = hi; glo[cc] = lo; }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                tcgen05_fence_before();
                mbar_arrive(smem_u32(&bar_a1));
            }
            // ---------------- E2: partial head outputs of this CTA's feature slice -> global exchange buffer
            {
                mbar_wait(smem_u32(&bar_acc[2]), par);
                tcgen05_fence_after();
                if (warp < 2) {
                    tmem_ld64(tmem + lane_base + 128, v);
                    if (t < D && __ldg(mrow + t) == 0.f) {
                        float4* dst = reinterpret_cast<float4*>(p.partial + ((((size_t)par * p.tiles + tile) * kCluster + rank) * kDp + t) * NT);
#pragma unroll
                        for (int i = 0; i < NT / 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    }
                }
                fence_cluster();
                tcgen05_fence_before();
                worker_sync();
                if (t < kCluster) mbar_arrive_remote(smem_u32(&bar_part), t);
            }
            // ---------------- U: sum the partials, affine coupling, next layer's masked input
            {
                mbar_wait_cluster(smem_u32(&bar_part), par);
                const float* pbase = p.partial + (((size_t)par * p.tiles + tile) * kCluster) * kDp * NT + uq * 4;
                const float* b2s = p.params + (size_t)(layer * 2 + 0) * p.blk + p.ob2;
                const float* b2t = p.params + (size_t)(layer * 2 + 1) * p.blk + p.ob2;
                float ld4[4] = {0.f, 0.f, 0.f, 0.f};
                for (int d = ug; d < D; d += 8) {
                    if (__ldg(mrow + d) != 0.f) continue;
                    float4 acc[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int cta = 0; cta < 4; ++cta) {
                            const float4 q4 = __ldcg(reinterpret_cast<const float4*>(pbase + ((size_t)(h * 4 + cta) * kDp + d) * NT));
                            s4.x += q4.x; s4.y += q4.y; s4.z += q4.z; s4.w += q4.w;
                        }
                        acc[h] = s4;
                    }
                    const float bs = __ldg(b2s + d), bt = __ldg(b2t + d);
                    const float sv[4] = {acc[0].x + bs, acc[0].y + bs, acc[0].z + bs, acc[0].w + bs};
                    const float tv[4] = {acc[1].x + bt, acc[1].y + bt, acc[1].z + bt, acc[1].w + bt};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int n = uq * 4 + k;
                        const float s = fast_tanh_f(sv[k]);
                        const float xv = xs[n * kXs + d];
                        float y;
                        if (p.direction == 0) { y = fmaf(xv, expf(s), tv[k]); ld4[k] += s; }
                        else { y = (xv - tv[k]) * expf(-s); ld4[k] -= s; }
                        xs[n * kXs + d] = y;
                        if (p.save && (int)rank == ug && r0 + n < p.R) {
                            p.saved_st[((size_t)(step * 2 + 0) * p.R + r0 + n) * D + d] = s;
                            p.saved_st[((size_t)(step * 2 + 1) * p.R + r0 + n) * D + d] = tv[k];
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) atomicAdd(&lds[uq * 4 + k], ld4[k]);
                worker_sync();
                const bool last = step == p.L - 1;
                if (p.save && (int)rank == ((step + 1) & 7)) {
                    const int nvalid = min(NT, p.R - r0) * D;
                    float* dst = p.saved_x + (size_t)(step + 1) * p.R * D + (size_t)r0 * D;
                    for (int i = t; i < nvalid; i += 128) { const int n = i / D, d = i - n * D; dst[i] = xs[n * kXs + d]; }
                }
                if (!last) {
                    const int next_layer = p.direction == 0 ? layer + 1 : layer - 1;
                    write_xm(p.mask + (size_t)next_layer * D);
                } else if (rank == 0) {
                    const int nvalid = min(NT, p.R - r0) * D;
                    for (int i = t; i < nvalid; i += 128) { const int n = i / D, d = i - n * D; p.out[(size_t)r0 * D + i] = xs[n * kXs + d]; }
                    if (p.logdet && t < NT && r0 + t < p.R) p.logdet[r0 + t] = lds[t];
                }
            }
        }
    }
    // teardown: nobody leaves while a peer may still signal it
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

// ---- host -----------------------------------------------------------------------------------------------------
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

}  // namespace fused
}  // namespace mhe
