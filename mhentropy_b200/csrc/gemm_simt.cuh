// fp32 CUDA-core GEMM with fused epilogues: the exact-precision building block of the flow path.
//
//   C(m,n) = sum_k A(m,k) * B(k,n)        per batch z, optionally split along K
//
// Operand layouts are strided views so one kernel serves the three shapes of an MLP layer:
//   forward  y = x W^T     : A K-major [M][K],  B K-major  [N][K]   (nn.Linear weight as stored)
//   dgrad    dx = dy W     : A K-major [M][K],  B MN-major [K][N]
//   wgrad    dW = dy^T x   : A MN-major [K][M], B MN-major [K][N]
// The epilogue functor receives every valid (batch, m, n, acc) and does the store, so bias /
// conditioning adds, leaky-ReLU, tanh, activation-gradient masks and accumulation are all fused.
#pragma once
#include "common.cuh"

namespace mhe {

enum class Major { K, MN };

struct GemmArgs {
    const float* A = nullptr;
    const float* B = nullptr;
    int M = 0, N = 0, K = 0;
    long lda = 0, ldb = 0;          // stride of the non-contiguous index
    long strideA = 0, strideB = 0;  // per-batch element strides
    const float* a_kscale = nullptr;  // optional multiplier per k applied to A (coupling mask)
    int batches = 1;
    int ksplit = 1;
};

template <int BM, int BN, int BK, int RM, int RN, Major AM, Major BMaj, class Epi>
__global__ void __launch_bounds__((BM / (4 * RM)) * (BN / (4 * RN)))
sgemm_kernel(GemmArgs g, Epi epi) {
    constexpr int TX = BN / (4 * RN);
    constexpr int TY = BM / (4 * RM);
    constexpr int NT = TX * TY;
    constexpr int LDA_S = BM + 4;
    constexpr int LDB_S = BN + 4;
    constexpr int A_PER_T = (BM * BK + NT - 1) / NT;
    constexpr int B_PER_T = (BN * BK + NT - 1) / NT;
    static_assert((BM * BK) % NT == 0 && (BN * BK) % NT == 0, "tile must divide evenly over threads");

    __shared__ __align__(16) float As[2][BK][LDA_S];
    __shared__ __align__(16) float Bs[2][BK][LDB_S];

    const int tid = threadIdx.x;
    const int tx = tid % TX;
    const int ty = tid / TX;
    const int m0 = blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;
    const int batch = blockIdx.z / g.ksplit;
    const int split = blockIdx.z % g.ksplit;

    int kc = (g.K + g.ksplit - 1) / g.ksplit;
    kc = (kc + BK - 1) / BK * BK;
    const int k_begin = split * kc;
    const int k_end = min(g.K, k_begin + kc);
    const int nk = (k_end > k_begin) ? (k_end - k_begin + BK - 1) / BK : 0;

    const float* __restrict__ A = g.A + (long)batch * g.strideA;
    const float* __restrict__ B = g.B + (long)batch * g.strideB;

    float ra[A_PER_T], rb[B_PER_T];

    auto load_regs = [&](int kt) {
        const int kb = k_begin + kt * BK;
#pragma unroll
        for (int i = 0; i < A_PER_T; ++i) {
            const int idx = tid + i * NT;
            int m, k;
            if (AM == Major::K) { k = idx % BK; m = idx / BK; } else { m = idx % BM; k = idx / BM; }
            const int gm = m0 + m, gk = kb + k;
            float v = 0.f;
            if (gm < g.M && gk < k_end) {
                v = (AM == Major::K) ? __ldg(A + (long)gm * g.lda + gk) : __ldg(A + (long)gk * g.lda + gm);
                if (g.a_kscale) v *= __ldg(g.a_kscale + gk);
            }
            ra[i] = v;
        }
#pragma unroll
        for (int i = 0; i < B_PER_T; ++i) {
            const int idx = tid + i * NT;
            int n, k;
            if (BMaj == Major::K) { k = idx % BK; n = idx / BK; } else { n = idx % BN; k = idx / BN; }
            const int gn = n0 + n, gk = kb + k;
            float v = 0.f;
            if (gn < g.N && gk < k_end)
                v = (BMaj == Major::K) ? __ldg(B + (long)gn * g.ldb + gk) : __ldg(B + (long)gk * g.ldb + gn);
            rb[i] = v;
        }
    };
    auto store_smem = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_PER_T; ++i) {
            const int idx = tid + i * NT;
            int m, k;
            if (AM == Major::K) { k = idx % BK; m = idx / BK; } else { m = idx % BM; k = idx / BM; }
            As[buf][k][m] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < B_PER_T; ++i) {
            const int idx = tid + i * NT;
            int n, k;
            if (BMaj == Major::K) { k = idx % BK; n = idx / BK; } else { n = idx % BN; k = idx / BN; }
            Bs[buf][k][n] = rb[i];
        }
    };

    float acc[4 * RM][4 * RN];
#pragma unroll
    for (int i = 0; i < 4 * RM; ++i)
#pragma unroll
        for (int j = 0; j < 4 * RN; ++j) acc[i][j] = 0.f;

    if (nk > 0) {
        load_regs(0);
        store_smem(0);
    }
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nk) load_regs(kt + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[4 * RM], b[4 * RN];
#pragma unroll
            for (int r = 0; r < RM; ++r) {
                const float4 v = *reinterpret_cast<const float4*>(&As[cur][k][r * (BM / RM) + ty * 4]);
                a[r * 4 + 0] = v.x; a[r * 4 + 1] = v.y; a[r * 4 + 2] = v.z; a[r * 4 + 3] = v.w;
            }
#pragma unroll
            for (int r = 0; r < RN; ++r) {
                const float4 v = *reinterpret_cast<const float4*>(&Bs[cur][k][r * (BN / RN) + tx * 4]);
                b[r * 4 + 0] = v.x; b[r * 4 + 1] = v.y; b[r * 4 + 2] = v.z; b[r * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < 4 * RM; ++i)
#pragma unroll
                for (int j = 0; j < 4 * RN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) store_smem(cur ^ 1);
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4 * RM; ++i) {
        const int m = m0 + (i / 4) * (BM / RM) + ty * 4 + (i % 4);
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4 * RN; ++j) {
            const int n = n0 + (j / 4) * (BN / RN) + tx * 4 + (j % 4);
            if (n < g.N) epi(batch, split, m, n, acc[i][j]);
        }
    }
}

// Tile choice: 128x128 (8x8 per thread) once the grid fills the 148 SMs, 64x64 (4x4) otherwise.
template <Major AM, Major BMaj, class Epi>
inline int launch_sgemm(const GemmArgs& g, const Epi& epi, cudaStream_t stream, const char* what) {
    if (g.M <= 0 || g.N <= 0 || g.batches <= 0) return MHE_OK;
    const int z = g.batches * g.ksplit;
    ProbeScope probe(what, stream);
    const long big_ctas = (long)cdiv(g.M, 128) * cdiv(g.N, 128) * z;
    if (big_ctas >= 2 * 148) {
        dim3 grid(cdiv(g.N, 128), cdiv(g.M, 128), z);
        sgemm_kernel<128, 128, 16, 2, 2, AM, BMaj, Epi><<<grid, 256, 0, stream>>>(g, epi);
    } else {
        dim3 grid(cdiv(g.N, 64), cdiv(g.M, 64), z);
        sgemm_kernel<64, 64, 16, 1, 1, AM, BMaj, Epi><<<grid, 256, 0, stream>>>(g, epi);
    }
    return check_launch(what);
}

// ---- generic epilogues --------------------------------------------------------------------
struct EpiStore {  // C[batch][m][n] = acc
    float* C; long ldc; long strideC;
    __device__ void operator()(int b, int, int m, int n, float acc) const { C[(long)b * strideC + (long)m * ldc + n] = acc; }
};
struct EpiAccumulate {  // C (+)= acc * nscale[n]; atomic when K is split
    float* C; long ldc; long strideC; const float* nscale; int atomic;
    __device__ void operator()(int b, int, int m, int n, float acc) const {
        if (nscale) acc *= __ldg(nscale + n);
        float* p = C + (long)b * strideC + (long)m * ldc + n;
        if (atomic) atomicAdd(p, acc); else *p += acc;
    }
};

}  // namespace mhe
