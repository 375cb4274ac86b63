// Conditioning projections straight from the fp32 weights (reference hand/flows.py:107-109, hoisted: one GEMM per step).
//   cp[b][idx*H + h] = sum_c feat[b][c] * Cw[idx][h][c] + Cb[idx][h] + b_j[h]
// At the training batch (64 images) this GEMM is a stream over the 50 MB of conditioning weights, which change every step: converting
// them to split half planes first costs a second pass over them (read 50 MB, write 50 MB, read 50 MB again).  Here the weights are
// read ONCE: TMA stages fp32 tiles in shared memory, eight converter warps split them into hi / lo half planes directly in the
// tcgen05 operand layout, one thread issues the 3-pass split-precision MMAs (weights on M, images on N), the same warps run the
// epilogue.  Bound: HBM (one pass over Cw).
#include "flow_tc.cuh"

#define CDSTAMP(k) do { if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0) p.dbg[i * 8 + (k)] = clock64(); } while (0)
namespace mhe {
namespace tcflow {
using namespace tc;
typedef __nv_bfloat16 bf16;

constexpr int kCdRows = 128;                    // weight rows per CTA (MMA M)
constexpr int kCdConv = 256;                    // converter / epilogue threads (warps 2..9)
constexpr int kCdThreads = 64 + kCdConv;        // warp 0: TMA producer, warp 1: MMA issuer
constexpr int kCdAf32 = kCdRows * BK * 4;       // 32 KB: one k-block of fp32 weights (two 32-float boxes)
constexpr int kCdAop = 2 * kCdRows * BK * 2;    // 32 KB: its hi | lo half planes (operand layout)

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void conv_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

struct CondDirectArgs {
    float* cp; long cp_ld;
    const float* params; size_t cb_base, cb_stride, blk, ob0, ob1;
    int B, Bp, H, C, stages;
    long long* dbg;   // MHE_CD_DEBUG: clock stamps of CTA (0,0)
};

// grid: (H / 128, L*4).  Shared memory: stages x [A fp32 32 KB | B planes 2 x Bp x 128 B] + 2 x A operand planes 32 KB.
__global__ void __launch_bounds__(kCdThreads, 1)
cond_fwd_direct_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapF, CondDirectArgs p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[4], bar_empty[4], bar_op_full[2], bar_op_empty[2], bar_acc;
    __shared__ uint32_t tmem_slot;
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NS = p.stages;
    const uint32_t b_bytes = (uint32_t)p.Bp * 128u;                  // one plane of the image operand per k-block
    const uint32_t stage_bytes = kCdAf32 + 2u * b_bytes;
    const uint32_t aop = smem0 + (uint32_t)NS * stage_bytes;         // two operand buffers
    const int idx = blockIdx.y, h0 = blockIdx.x * kCdRows;
    const int nkb = p.C / BK;
    const uint32_t tmem_cols = p.Bp <= 32 ? 32u : (p.Bp <= 64 ? 64u : 128u);

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
        for (int j = 0; j < 2; ++j) { mbar_init(smem_u32(&bar_op_full[j]), 1); mbar_init(smem_u32(&bar_op_empty[j]), 1); }
        mbar_init(smem_u32(&bar_acc), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapF) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_d = tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % NS;
                mbar_wait(smem_u32(&bar_empty[s]), ((i / NS) & 1) ^ 1);
                const uint32_t full = smem_u32(&bar_full[s]), base = smem0 + (uint32_t)s * stage_bytes;
                mbar_expect_tx(full, stage_bytes);
                tma_load_3d(base, &mapW, full, i * BK, h0, idx);
                tma_load_3d(base + kCdAf32 / 2, &mapW, full, i * BK + 32, h0, idx);
                tma_load_4d(base + kCdAf32, &mapF, full, i * BK, 0, 0, 0);
                tma_load_4d(base + kCdAf32 + b_bytes, &mapF, full, i * BK, 0, 1, 0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = instr_desc(p.Bp, false, false, true, true);
            constexpr uint32_t kHi = (uint32_t)((1024u >> 4) | (1u << 14) | (2u << 29));   // SBO | version | SWIZZLE_128B
            constexpr uint32_t kLo = (16u >> 4) << 16;
            uint32_t accumulate = 0;
            for (int i = 0; i < nkb; ++i) {
                const int s = i % NS, j = i & 1;
                mbar_wait(smem_u32(&bar_full[s]), (i / NS) & 1);            // the image planes of this stage (TMA)
                mbar_wait(smem_u32(&bar_op_full[j]), (i >> 1) & 1);        // the converted weight planes
                tcgen05_fence_after();
                CDSTAMP(4);
                const uint32_t a_hi = kLo | ((aop + (uint32_t)j * kCdAop) >> 4), a_lo = a_hi + ((kCdAop / 2) >> 4);
                const uint32_t b_hi = kLo | ((smem0 + (uint32_t)s * stage_bytes + kCdAf32) >> 4), b_lo = b_hi + (b_bytes >> 4);
#pragma unroll
                for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                    umma_bf16_lohi(tmem_d, a_hi + ks * 2, b_hi + ks * 2, kHi, idesc, accumulate);
                    umma_bf16_lohi(tmem_d, a_hi + ks * 2, b_lo + ks * 2, kHi, idesc, 1u);
                    umma_bf16_lohi(tmem_d, a_lo + ks * 2, b_hi + ks * 2, kHi, idesc, 1u);
                    accumulate = 1;
                }
                tcgen05_commit(smem_u32(&bar_op_empty[j]));
                tcgen05_commit(smem_u32(&bar_empty[s]));
                CDSTAMP(5);
            }
            tcgen05_commit(smem_u32(&bar_acc));
        }
    } else {
        // ---- converters: thread = (weight row, half of the k-block)
        const int t = threadIdx.x - 64;
        const int r = t & 127, kh = t >> 7;
        const uint32_t rsw = (uint32_t)(r & 7);
        const uint32_t src_row = (uint32_t)kh * (kCdAf32 / 2) + (uint32_t)r * 128u;
        const uint32_t dst_row = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
        for (int i = 0; i < nkb; ++i) {
            const int s = i % NS, j = i & 1;
            if (t == 0) {
                CDSTAMP(0);
                mbar_wait(smem_u32(&bar_full[s]), (i / NS) & 1);
                CDSTAMP(1);
                mbar_wait(smem_u32(&bar_op_empty[j]), ((i >> 1) & 1) ^ 1);
                CDSTAMP(2);
            }
            conv_sync();
            const uint32_t src = smem0 + (uint32_t)s * stage_bytes + src_row, dst = aop + (uint32_t)j * kCdAop + dst_row;
#pragma unroll
            for (int q = 0; q < 4; ++q) {          // 8 floats -> one 16-byte chunk of each plane
                float v[8];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const uint32_t c = (uint32_t)(2 * q + e);
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(v[4 * e]), "=f"(v[4 * e + 1]), "=f"(v[4 * e + 2]), "=f"(v[4 * e + 3]) : "r"(src + ((c ^ rsw) << 4)));
                }
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const __half2 hh = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
                    const float2 hf = __half22float2(hh);
                    const __half2 ll = __floats2half2_rn(v[2 * k] - hf.x, v[2 * k + 1] - hf.y);
                    hw[k] = *reinterpret_cast<const uint32_t*>(&hh);
                    lw[k] = *reinterpret_cast<const uint32_t*>(&ll);
                }
                const uint32_t cc = (uint32_t)(kh * 4 + q);
                const uint32_t d = dst + ((cc ^ rsw) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d), "r"(hw[0]), "r"(hw[1]), "r"(hw[2]), "r"(hw[3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d + kCdAop / 2), "r"(lw[0]), "r"(lw[1]), "r"(lw[2]), "r"(lw[3]) : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            conv_sync();
            if (t == 0) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar_op_full[j])) : "memory"); CDSTAMP(3); }
        }
        // ---- epilogue: TMEM lane = weight row, columns = images; for a fixed image the 32 lanes of a warp store 32 consecutive floats
        const int cw = warp - 2;                            // 0..7
        const int q = warp & 3;                             // TMEM lane quadrant this warp may access (hardware: warp id % 4)
        const int h = h0 + q * 32 + lane;
        const float bias = __ldg(p.params + p.cb_base + (size_t)idx * p.cb_stride + h) +
                           __ldg(p.params + (size_t)(idx >> 1) * p.blk + ((idx & 1) ? p.ob1 : p.ob0) + h);
        if (t == 0) { mbar_wait(smem_u32(&bar_acc), 0); const int i = nkb; CDSTAMP(0); }
        conv_sync();
        tcgen05_fence_after();
        for (int c0 = (cw >> 2) * 32; c0 < p.Bp; c0 += 64) {
            float v[32];
            tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            float* o = p.cp + (long)idx * p.H + h;
#pragma unroll
            for (int b = 0; b < 32; ++b)
                if (c0 + b < p.B) o[(long)(c0 + b) * p.cp_ld] = v[b] + bias;
        }
        if (t == 0) { const int i = nkb; CDSTAMP(1); }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
}

bool cond_direct_supported(const FlowLayout& L, int B) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("MHE_COND_DIRECT"); on = e ? atoi(e) : 1; }
    return on && B >= 1 && B <= 128 && L.H % kCdRows == 0 && L.C % BK == 0 && L.cw_stride % 4 == 0;
}

// feat [B][C] -> cp [B][L*4][H]; ws: the half planes of feat [2][B][C]
int cond_fwd_direct(const FlowLayout& L, const float* params, const float* feat, int B, float* cp, void* ws_, cudaStream_t stream) {
    bf16* featp = (bf16*)ws_;
    MHE_TRY(split_planes(feat, L.C, 0, B, L.C, nullptr, featp, B, L.C, 2, 1, true, stream));
    const int Bp = (B + 15) / 16 * 16;
    int st = MHE_OK;
    const CUtensorMap* mW = cached_map_f32(params + L.cw_base, L.C, L.H, L.L * 4, L.C, (long)L.cw_stride, kCdRows, &st);
    if (st != MHE_OK) return st;
    PlaneTensor F;
    F.base = featp; F.cols = L.C; F.rows = B; F.planes = 2; F.batches = 1;
    F.row_pitch = L.C; F.plane_stride = (long)B * L.C; F.batch_stride = (long)2 * B * L.C;
    const CUtensorMap* mF = cached_map(F, Bp, &st);
    if (st != MHE_OK) return st;
    CondDirectArgs a{};
    a.cp = cp; a.cp_ld = (long)L.L * 4 * L.H; a.params = params; a.cb_base = L.cb_base; a.cb_stride = L.cb_stride; a.blk = L.blk;
    a.ob0 = L.ob0; a.ob1 = L.ob1; a.B = B; a.Bp = Bp; a.H = L.H; a.C = L.C;
    const size_t stage = (size_t)kCdAf32 + 2 * (size_t)Bp * 128;
    a.stages = (int)((200 * 1024 - 2 * (size_t)kCdAop) / stage);
    if (a.stages > 4) a.stages = 4;
    if (const char* e = getenv("MHE_CD_STAGES")) { const int n = atoi(e); if (n >= 1 && n < a.stages) a.stages = n; }
    if (a.stages > L.C / BK) a.stages = L.C / BK;
    const size_t smem = (size_t)a.stages * stage + 2 * (size_t)kCdAop + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(cond_fwd_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) {
            set_error("cond fwd direct: cannot raise dynamic shared memory");
            return MHE_ERR_CUDA;
        }
        attr_set = true;
    }
    static long long* dbg = nullptr;
    static bool dbg_on = getenv("MHE_CD_DEBUG") != nullptr;
    if (dbg_on && !dbg) { cudaMalloc(&dbg, 64 * 8 * sizeof(long long)); cudaMemset(dbg, 0, 64 * 8 * sizeof(long long)); }
    a.dbg = dbg;
    ProbeScope probe("tc cond fwd direct", stream);
    cond_fwd_direct_kernel<<<dim3(L.H / kCdRows, L.L * 4), kCdThreads, smem, stream>>>(*mW, *mF, a);
    MHE_TRY(check_launch("tc cond fwd direct"));
    if (dbg_on) {
        long long h[64 * 8];
        cudaDeviceSynchronize();
        cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
        const long long t0 = h[0];
        for (int i = 0; i <= L.C / BK; ++i) {
            fprintf(stderr, "kb %d:", i);
            for (int k = 0; k < 6; ++k) fprintf(stderr, " %8lld", h[i * 8 + k] ? h[i * 8 + k] - t0 : -1);
            fprintf(stderr, "\n");
        }
    }
    return MHE_OK;
}

}  // namespace tcflow
}  // namespace mhe
