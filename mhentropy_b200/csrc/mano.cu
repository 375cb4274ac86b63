// MANO layer kernels: pose/chain (thread per row), blend shapes + linear blend skinning (thread per
// vertex, row-tiled so each posedirs element is reused across the tile), regressed joint set, and
// the exact backward.  Reference: hand/manopth/manolayer.py:110-274, hand/ManoLayer.py:45-60,141-148.
#include "gemm_simt.cuh"
#include "loss_rows.cuh"
#include "mano_math.cuh"
#include "tc_gemm.cuh"

namespace mhe {
namespace skin {   // mano_skin_tc.cu: linear blend skinning with the transform blend on the tensor cores
int launch(const uint16_t* wplanes, const float* A, const float* cen, const float* vp, int ld_vp, float* verts, float* jtr, int R, int order,
           cudaStream_t stream);
}
using namespace mano;

constexpr int kV = MHE_MANO_VERTS;   // 778
constexpr int kVC = kV * 3;          // 2334
constexpr int kNJ = MHE_MANO_JOINTS; // 21
constexpr float kMM = 1000.f;        // metres -> millimetres (manolayer.py:272-273)

// source of output joint i: <16 chain joint, >=16 tip (index-16).  [order][i]
// order 0: manolayer.py:260; order 1: additionally utils.py:15 FreiHand2RHD.
__constant__ int c_jtr_src[2][kNJ] = {
    {0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20},
    {0, 16, 15, 14, 13, 17, 3, 2, 1, 18, 6, 5, 4, 19, 12, 11, 10, 20, 9, 8, 7}};
// tip vertices of the inner layer (manolayer.py:250) in tip order 16..20
__constant__ int c_tip_vert[5] = {745, 317, 444, 556, 673};
// wrapper's regressed set (ManoLayer.py:109-138): >=0 regress MANO joint id, <0 vertex -(v)-1.  [order][i]
__constant__ int c_j2_src[2][kNJ] = {
    {0, 13, 14, 15, -745, 1, 2, 3, -321, 4, 5, 6, -444, 10, 11, 12, -556, 7, 8, 9, -673},
    {0, -745, 15, 14, 13, -321, 3, 2, 1, -444, 6, 5, 4, -556, 12, 11, 10, -673, 9, 8, 7}};

// per-row workspace: pm [136], A [192], center [4]
constexpr int kWsPm = 136, kWsA = 192, kWsCen = 4;
constexpr int kWsRowFwd = kWsPm + kWsA + kWsCen;            // 332
// backward adds: dA [192], dpm [136], dbv [12], dcen [4]
constexpr int kWsRowBwd = 192 + 136 + 12 + 4;               // 344
constexpr int kWsRowMesh = 3 * 2336;                        // vp, dvt, dvp

// tensor-core pose blend: pose-map planes [2][R][kPmK] (half), pose offsets [R][kOffLd] fp32
constexpr int kPmK = 192;            // 135 pose-map entries padded to three 64-deep k-blocks
constexpr int kOffLd = 2336;         // 2334 vertex coordinates padded to a 16-byte multiple of 16-bit elements
constexpr int kLbsN = 16;            // the 12 entries of a 3 x 4 transform padded to 16 GEMM columns per row
constexpr int kVp = 896;             // 778 vertices padded to 7 tiles of 128
constexpr int kBlendRows = kPoseMap + kShape + 1;   // contraction of the blend GEMM: 135 pose-map entries | 10 betas | 1 (template)
struct ManoWs {
    float *pm, *A, *cen, *dA, *dpm, *dbv, *dcen, *vp, *dvt, *dvp, *pmp, *poff, *lbsb;
    ManoWs(float* base, int R, bool mesh) {
        auto take = [&](size_t n) { float* p = base; base += (n + 63) / 64 * 64; return p; };
        pm = take((size_t)R * kWsPm); A = take((size_t)R * kWsA); cen = take((size_t)R * kWsCen);
        dA = take((size_t)R * 192); dpm = take((size_t)R * 136); dbv = take((size_t)R * 12); dcen = take((size_t)R * 4);
        vp = dvt = dvp = nullptr;
        if (mesh) { vp = take((size_t)R * 2336); dvt = take((size_t)R * 2336); dvp = take((size_t)R * 2336); }
        pmp = take((size_t)R * kPmK); poff = take((size_t)R * kOffLd);      // (2 half planes of kPmK = kPmK floats per row)
        lbsb = take((size_t)R * kLbsN * 64);                                // 2 half planes [R * 16][64]
    }
    static size_t floats(int R, bool mesh) {
        return (size_t)R * (kWsRowFwd + kWsRowBwd + (mesh ? kWsRowMesh : 0) + kPmK + kOffLd + kLbsN * 64) + 13 * 64;
    }
};

// ---- warp-per-row pose state ---------------------------------------------------------------------------
// One warp owns one row; the per-row state lives in shared memory and the lanes split the work:
// PCA outputs / joint regression over lanes, one Rodrigues per lane, the kinematic chain level by level with
// one lane per finger.  The primitives are the host-checked ones of mano_math.cuh.
struct WarpPose {
    float pose[kPose];
    float R[kJ][9], J[kJ][3], Gr[kJ][9], Gt[kJ][3], A[kJ][12];
    float dGr[kJ][9], dGt[kJ][3], dJ[kJ][3], dR[kJ][9], dpose[kPose];
    float dA[kJ][12], dpm[kWsPm], dGt_out[kJ * 3], T[12], red[8];
    float th[48], be[12], dj[64];
};
constexpr int kPoseWarps = 4;

// Small constants of the pose / tip path, staged once per block in shared memory: the per-row work is a long chain of short
// dependent steps, and with a handful of warps per SM every global-memory round trip in it is exposed latency.
struct __align__(16) PoseTables {
    float comps[45 * 45], hands_mean[48], jt[kJ * 3], js[kJ * 3 * kShape];
    float ptip[5][kPoseMap][3], stip[5][3][kShape], vtip[5][4], wtip[5][kJ];
};
static_assert(sizeof(PoseTables) % 16 == 0, "PoseTables is copied in 16-byte pieces");
__device__ __forceinline__ void stage_pose_tables(const mhe_mano_consts& c, PoseTables& T, bool tips) {
    const int t = threadIdx.x, n = blockDim.x;
    if (c.pose_tables) {   // pre-gathered: one coalesced copy
        const float4* src = reinterpret_cast<const float4*>(c.pose_tables);
        float4* dst = reinterpret_cast<float4*>(&T);
        for (int i = t; i < (int)(sizeof(PoseTables) / 16); i += n) dst[i] = __ldg(src + i);
        return;
    }
    for (int i = t; i < 45 * 45; i += n) T.comps[i] = __ldg(c.comps + i);
    for (int i = t; i < 45; i += n) T.hands_mean[i] = __ldg(c.hands_mean + i);
    for (int i = t; i < kJ * 3; i += n) T.jt[i] = __ldg(c.jt + i);
    for (int i = t; i < kJ * 3 * kShape; i += n) T.js[i] = __ldg(c.js + i);
    if (tips) {
        for (int i = t; i < 5 * kPoseMap * 3; i += n) {
            const int tip = i / (kPoseMap * 3), k = (i / 3) % kPoseMap, cc = i % 3;
            T.ptip[tip][k][cc] = __ldg(c.posedirs_t + (long)k * kVC + c_tip_vert[tip] * 3 + cc);
        }
        for (int i = t; i < 5 * 3 * kShape; i += n) {
            const int tip = i / (3 * kShape), cc = (i / kShape) % 3, b = i % kShape;
            T.stip[tip][cc][b] = __ldg(c.shapedirs + (c_tip_vert[tip] * 3 + cc) * kShape + b);
        }
        for (int i = t; i < 15; i += n) T.vtip[i / 3][i % 3] = __ldg(c.v_template + c_tip_vert[i / 3] * 3 + i % 3);
        for (int i = t; i < 5 * kJ; i += n) T.wtip[i / kJ][i % kJ] = __ldg(c.weights + c_tip_vert[i / kJ] * kJ + i % kJ);
    }
}

__device__ __forceinline__ void pose_fwd_warp(const PoseTables& T, const float* __restrict__ theta_g, const float* __restrict__ beta_g,
                                              WarpPose& W, int lane) {
    for (int i = lane; i < 48; i += 32) W.th[i] = theta_g[i];
    if (lane < kShape) W.be[lane] = beta_g[lane];
    __syncwarp();
    const float* theta = W.th;
    const float* beta = W.be;
    for (int j = lane; j < 45; j += 32) {
        float acc = T.hands_mean[j];
        for (int k = 0; k < 45; ++k) acc = fmaf(theta[3 + k], T.comps[k * 45 + j], acc);
        W.pose[3 + j] = acc;
    }
    if (lane < 3) W.pose[lane] = theta[lane];
    for (int i = lane; i < kJ * 3; i += 32) {
        float acc = T.jt[i];
        for (int b = 0; b < kShape; ++b) acc = fmaf(T.js[i * kShape + b], beta[b], acc);
        W.J[i / 3][i % 3] = acc;
    }
    __syncwarp();
    if (lane < kJ) rodrigues_fwd(&W.pose[3 * lane], W.R[lane]);
    __syncwarp();
    if (lane == 0) {
        for (int i = 0; i < 9; ++i) W.Gr[0][i] = W.R[0][i];
        for (int cc = 0; cc < 3; ++cc) W.Gt[0][cc] = W.J[0][cc];
    }
    __syncwarp();
    for (int level = 0; level < 3; ++level) {
        if (lane < 5) {
            const int k = 1 + 3 * lane + level, p = parent_of(k);
            mat3_mul(W.Gr[p], W.R[k], W.Gr[k]);
            const float d[3] = {W.J[k][0] - W.J[p][0], W.J[k][1] - W.J[p][1], W.J[k][2] - W.J[p][2]};
            float o[3];
            mat3_vec(W.Gr[p], d, o);
            for (int cc = 0; cc < 3; ++cc) W.Gt[k][cc] = o[cc] + W.Gt[p][cc];
        }
        __syncwarp();
    }
    if (lane < kJ) {
        for (int i = 0; i < 9; ++i) W.A[lane][i] = W.Gr[lane][i];
        float o[3];
        mat3_vec(W.Gr[lane], W.J[lane], o);
        for (int cc = 0; cc < 3; ++cc) W.A[lane][9 + cc] = W.Gt[lane][cc] - o[cc];
    }
    __syncwarp();
}

// blend shapes + LBS of tip vertex `tip` by a whole warp: vp (3) and T (12, in W.T) are left in every lane / smem
__device__ __forceinline__ void tip_skin_warp(const PoseTables& T, int tip, WarpPose& W, int lane, float* vp) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int k = lane; k < kPoseMap; k += 32) {
        const float p = W.R[1 + k / 9][k % 9] - ((k % 9) % 4 == 0 ? 1.f : 0.f);
        a0 = fmaf(T.ptip[tip][k][0], p, a0);
        a1 = fmaf(T.ptip[tip][k][1], p, a1);
        a2 = fmaf(T.ptip[tip][k][2], p, a2);
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    float sh[3];
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) {
        float acc = T.vtip[tip][cc];
        for (int b = 0; b < kShape; ++b) acc = fmaf(T.stip[tip][cc][b], W.be[b], acc);
        sh[cc] = acc;
    }
    vp[0] = sh[0] + a0; vp[1] = sh[1] + a1; vp[2] = sh[2] + a2;
    if (lane < 12) {
        float t = 0.f;
        for (int k = 0; k < kJ; ++k) t = fmaf(T.wtip[tip][k], W.A[k][lane], t);
        W.T[lane] = t;
    }
    __syncwarp();
}

__global__ void mano_pack_pose_tables_kernel(mhe_mano_consts c, float* out) {
    __shared__ PoseTables T;
    stage_pose_tables(c, T, true);
    __syncthreads();
    const float* src = reinterpret_cast<const float*>(&T);
    for (int i = threadIdx.x; i < (int)(sizeof(PoseTables) / sizeof(float)); i += blockDim.x) out[i] = src[i];
}

// ---- forward ------------------------------------------------------------------------------------
// warp per row: pose -> pm / A / centre for the skinning kernel (when requested) and the 16 chain joints;
// with `tips` also the five tip vertices, so the joints-only path is this single kernel.
__global__ void __launch_bounds__(kPoseWarps * 32) mano_pose_fwd_kernel(mhe_mano_consts c, const float* __restrict__ theta, int ld_theta,
                                     const float* __restrict__ beta, int ld_beta, int R, int order, int tips,
                                     float* __restrict__ pm, float* __restrict__ A, float* __restrict__ cen,
                                     float* __restrict__ jtr) {
    __shared__ WarpPose s_w[kPoseWarps];
    __shared__ PoseTables s_t;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * kPoseWarps + warp;
    stage_pose_tables(c, s_t, tips != 0);
    __syncthreads();
    if (r >= R) return;
    WarpPose& W = s_w[warp];
    const float* be = beta + (long)r * ld_beta;
    pose_fwd_warp(s_t, theta + (long)r * ld_theta, be, W, lane);
    if (pm) for (int k = lane; k < kPoseMap; k += 32) pm[(long)r * kWsPm + k] = W.R[1 + k / 9][k % 9] - ((k % 9) % 4 == 0 ? 1.f : 0.f);
    if (A) for (int i = lane; i < kWsA; i += 32) A[(long)r * kWsA + i] = W.A[i / 12][i % 12];
    if (cen && lane < 3) cen[(long)r * kWsCen + lane] = W.Gt[kCenterJoint][lane];
    if (jtr) {
        for (int i = lane; i < kNJ * 3; i += 32) {
            const int src = c_jtr_src[order][i / 3];
            if (src < kJ) jtr[(long)r * kNJ * 3 + i] = (W.Gt[src][i % 3] - W.Gt[kCenterJoint][i % 3]) * kMM;
        }
        if (tips) {
            for (int t = 0; t < 5; ++t) {
                float vp[3];
                tip_skin_warp(s_t, t, W, lane, vp);
                if (lane < 3) {
                    const float o = (W.T[lane * 3 + 0] * vp[0] + W.T[lane * 3 + 1] * vp[1] + W.T[lane * 3 + 2] * vp[2] + W.T[9 + lane] - W.Gt[kCenterJoint][lane]) * kMM;
                    for (int i = 0; i < kNJ; ++i) if (c_jtr_src[order][i] == kJ + t) jtr[((long)r * kNJ + i) * 3 + lane] = o;
                }
                __syncwarp();
            }
        }
    }
}

// blend shapes + LBS of one vertex for one row. pm/A/beta/cen point at this row's data.
__device__ __forceinline__ void skin_vertex(const mhe_mano_consts& c, int v, const float* pm, const float* A, const float* beta,
                                            float* vp_out, float* T /*12*/) {
    float vp[3];
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) {
        float acc = __ldg(c.v_template + v * 3 + cc);
        for (int b = 0; b < kShape; ++b) acc = fmaf(__ldg(c.shapedirs + (v * 3 + cc) * kShape + b), beta[b], acc);
        vp[cc] = acc;
    }
    for (int k = 0; k < kPoseMap; ++k) {
        const float p = pm[k];
        vp[0] = fmaf(__ldg(c.posedirs_t + (long)k * kVC + v * 3 + 0), p, vp[0]);
        vp[1] = fmaf(__ldg(c.posedirs_t + (long)k * kVC + v * 3 + 1), p, vp[1]);
        vp[2] = fmaf(__ldg(c.posedirs_t + (long)k * kVC + v * 3 + 2), p, vp[2]);
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) T[i] = 0.f;
    for (int k = 0; k < kJ; ++k) {
        const float w = __ldg(c.weights + v * kJ + k);
        if (w != 0.f)
#pragma unroll
            for (int i = 0; i < 12; ++i) T[i] = fmaf(w, A[k * 12 + i], T[i]);
    }
    vp_out[0] = vp[0]; vp_out[1] = vp[1]; vp_out[2] = vp[2];
}

// grid (vertex chunks, row tiles); 128 threads = 128 vertices; RT rows share every posedirs load
template <int RT>
__global__ void __launch_bounds__(128) mano_skin_fwd_kernel(mhe_mano_consts c, const float* __restrict__ beta, int ld_beta, int R, int order,
                                                            const float* __restrict__ pm_g, const float* __restrict__ A_g, const float* __restrict__ cen_g,
                                                            float* __restrict__ verts, float* __restrict__ jtr, float* __restrict__ vp_out,
                                                            const float* __restrict__ pose_off) {
    __shared__ float s_pm[RT][kWsPm];
    __shared__ __align__(16) float s_A[RT][kWsA];
    __shared__ float s_beta[RT][12];
    __shared__ float s_cen[RT][4];
    const int r0 = blockIdx.y * RT;
    const int nr = min(RT, R - r0);
    if (!pose_off)   // (the pose map is only read by the in-kernel pose blend)
        for (int i = threadIdx.x; i < RT * kWsPm; i += blockDim.x) { const int rr = i / kWsPm; s_pm[rr][i % kWsPm] = rr < nr ? pm_g[(long)(r0 + rr) * kWsPm + i % kWsPm] : 0.f; }
    {   // the tile's joint transforms are one contiguous block of nr * 192 floats: 16-byte copies, no index arithmetic
        const float4* src = reinterpret_cast<const float4*>(A_g + (long)r0 * kWsA);
        float4* dst = reinterpret_cast<float4*>(&s_A[0][0]);
        for (int i = threadIdx.x; i < RT * kWsA / 4; i += blockDim.x) dst[i] = i < nr * kWsA / 4 ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i = threadIdx.x; i < RT * kShape; i += blockDim.x) { const int rr = i / kShape; s_beta[rr][i % kShape] = rr < nr ? beta[(long)(r0 + rr) * ld_beta + i % kShape] : 0.f; }
    for (int i = threadIdx.x; i < RT * 3; i += blockDim.x) { const int rr = i / 3; s_cen[rr][i % 3] = rr < nr ? cen_g[(long)(r0 + rr) * kWsCen + i % 3] : 0.f; }
    __syncthreads();
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= kV) return;

    float vp[RT][3];
    if (pose_off) {   // template + shape blend + pose blend of every row: the tensor-core blend GEMM's output
#pragma unroll
        for (int rr = 0; rr < RT; ++rr) {
            const float* po = pose_off + (long)(r0 + (rr < nr ? rr : 0)) * kOffLd + v * 3;
            vp[rr][0] = po[0]; vp[rr][1] = po[1]; vp[rr][2] = po[2];
        }
    } else {
#pragma unroll
    for (int rr = 0; rr < RT; ++rr)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
            float acc = __ldg(c.v_template + v * 3 + cc);
            for (int b = 0; b < kShape; ++b) acc = fmaf(__ldg(c.shapedirs + (v * 3 + cc) * kShape + b), s_beta[rr][b], acc);
            vp[rr][cc] = acc;
        }
    for (int k = 0; k < kPoseMap; ++k) {
        const float p0 = __ldg(c.posedirs_t + (long)k * kVC + v * 3 + 0);
        const float p1 = __ldg(c.posedirs_t + (long)k * kVC + v * 3 + 1);
        const float p2 = __ldg(c.posedirs_t + (long)k * kVC + v * 3 + 2);
#pragma unroll
        for (int rr = 0; rr < RT; ++rr) {
            const float p = s_pm[rr][k];
            vp[rr][0] = fmaf(p0, p, vp[rr][0]);
            vp[rr][1] = fmaf(p1, p, vp[rr][1]);
            vp[rr][2] = fmaf(p2, p, vp[rr][2]);
        }
    }
    }
    float w[kJ];
#pragma unroll
    for (int k = 0; k < kJ; k += 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(c.weights + v * kJ + k));
        w[k] = t.x; w[k + 1] = t.y; w[k + 2] = t.z; w[k + 3] = t.w;
    }
    int tip = -1;
#pragma unroll
    for (int t = 0; t < 5; ++t) if (c_tip_vert[t] == v) tip = t;
#pragma unroll
    for (int rr = 0; rr < RT; ++rr) {
        if (rr >= nr) break;
        float T[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) T[i] = 0.f;
#pragma unroll
        for (int k = 0; k < kJ; ++k) {
            if (w[k] != 0.f) {   // real MANO skinning weights have <= 4 non-zeros per vertex   // (the joint's 3 x 4 transform as three 16-byte shared-memory loads: this loop is LDS-bound)
                const float4* a4 = reinterpret_cast<const float4*>(&s_A[rr][k * 12]);
                const float4 a0 = a4[0], a1 = a4[1], a2 = a4[2];
                T[0] = fmaf(w[k], a0.x, T[0]); T[1] = fmaf(w[k], a0.y, T[1]); T[2] = fmaf(w[k], a0.z, T[2]); T[3] = fmaf(w[k], a0.w, T[3]);
                T[4] = fmaf(w[k], a1.x, T[4]); T[5] = fmaf(w[k], a1.y, T[5]); T[6] = fmaf(w[k], a1.z, T[6]); T[7] = fmaf(w[k], a1.w, T[7]);
                T[8] = fmaf(w[k], a2.x, T[8]); T[9] = fmaf(w[k], a2.y, T[9]); T[10] = fmaf(w[k], a2.z, T[10]); T[11] = fmaf(w[k], a2.w, T[11]);
            }
        }
        float o[3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
            o[i] = (T[i * 3 + 0] * vp[rr][0] + T[i * 3 + 1] * vp[rr][1] + T[i * 3 + 2] * vp[rr][2] + T[9 + i] - s_cen[rr][i]) * kMM;
        const long r = r0 + rr;
        if (verts) { float* d = verts + (r * kV + v) * 3; d[0] = o[0]; d[1] = o[1]; d[2] = o[2]; }
        if (vp_out) { float* d = vp_out + r * 2336 + v * 3; d[0] = vp[rr][0]; d[1] = vp[rr][1]; d[2] = vp[rr][2]; }
        if (tip >= 0 && jtr) {
            for (int i = 0; i < kNJ; ++i)
                if (c_jtr_src[order][i] == kJ + tip) { float* d = jtr + (r * kNJ + i) * 3; d[0] = o[0]; d[1] = o[1]; d[2] = o[2]; }
        }
    }
}

// wrapper's regressed joints: joints2[r][i] = sum_v jreg[k][v] verts[r][v] or a tip vertex. block per row, 8 warps.
__global__ void __launch_bounds__(256) mano_joints2_fwd_kernel(mhe_mano_consts c, const float* __restrict__ verts, int R, int order, float* __restrict__ joints2) {
    const int r = blockIdx.x;
    const int warp = threadIdx.x / 32, lane = threadIdx.x & 31;
    const float* vr = verts + (long)r * kVC;
    for (int i = warp; i < kNJ; i += 8) {
        const int src = c_j2_src[order][i];
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        if (src >= 0) {
            for (int v = lane; v < kV; v += 32) {
                const float w = __ldg(c.jreg + src * kV + v);
                if (w != 0.f) { a0 = fmaf(w, vr[v * 3], a0); a1 = fmaf(w, vr[v * 3 + 1], a1); a2 = fmaf(w, vr[v * 3 + 2], a2); }
            }
            a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
        } else {
            const int v = -src - 1;
            a0 = vr[v * 3]; a1 = vr[v * 3 + 1]; a2 = vr[v * 3 + 2];
        }
        if (lane == 0) { float* d = joints2 + ((long)r * kNJ + i) * 3; d[0] = a0; d[1] = a1; d[2] = a2; }
    }
}

// ---- backward -----------------------------------------------------------------------------------
// dvt[r][v][c] = dverts + jreg^T djoints2 (+ tip slots of djoints2 and djtr): every gradient that lands on a vertex
__global__ void mano_dverts_total_kernel(mhe_mano_consts c, const float* __restrict__ dverts, const float* __restrict__ djtr,
                                         const float* __restrict__ dj2, int R, int order, float* __restrict__ dvt) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)R * kV) return;
    const int r = (int)(idx / kV), v = (int)(idx % kV);
    float g[3] = {0.f, 0.f, 0.f};
    if (dverts) for (int cc = 0; cc < 3; ++cc) g[cc] = dverts[idx * 3 + cc];
    if (dj2) {
        for (int i = 0; i < kNJ; ++i) {
            const int src = c_j2_src[order][i];
            float w = 0.f;
            if (src >= 0) w = __ldg(c.jreg + src * kV + v); else if (-src - 1 == v) w = 1.f;
            if (w != 0.f) for (int cc = 0; cc < 3; ++cc) g[cc] = fmaf(w, dj2[((long)r * kNJ + i) * 3 + cc], g[cc]);
        }
    }
    if (djtr) {
        for (int t = 0; t < 5; ++t)
            if (c_tip_vert[t] == v)
                for (int i = 0; i < kNJ; ++i)
                    if (c_jtr_src[order][i] == kJ + t) for (int cc = 0; cc < 3; ++cc) g[cc] += djtr[((long)r * kNJ + i) * 3 + cc];
    }
    for (int cc = 0; cc < 3; ++cc) dvt[(long)r * 2336 + v * 3 + cc] = g[cc];
}

// mesh backward, block per row (192 threads): dA[k][e], dcen, and dvp[v] = T_rot[v]^T dv[v]
__global__ void __launch_bounds__(192) mano_skin_bwd_kernel(mhe_mano_consts c, int R, const float* __restrict__ A_g,
                                                            const float* __restrict__ vp_g, const float* __restrict__ dvt,
                                                            float* __restrict__ dA, float* __restrict__ dcen, float* __restrict__ dvp) {
    __shared__ float s_dv[kVC];
    __shared__ float s_vp[kVC];
    __shared__ float s_A[kWsA];
    const int r = blockIdx.x;
    for (int i = threadIdx.x; i < kVC; i += blockDim.x) { s_dv[i] = dvt[(long)r * 2336 + i] * kMM; s_vp[i] = vp_g[(long)r * 2336 + i]; }
    for (int i = threadIdx.x; i < kWsA; i += blockDim.x) s_A[i] = A_g[(long)r * kWsA + i];
    __syncthreads();
    {   // dA[k][e]
        const int k = threadIdx.x / 12, e = threadIdx.x % 12;
        float acc = 0.f;
        if (e < 9) {
            const int i = e / 3, j = e % 3;
            for (int v = 0; v < kV; ++v) { const float w = __ldg(c.weights + v * kJ + k); if (w != 0.f) acc = fmaf(w * s_dv[v * 3 + i], s_vp[v * 3 + j], acc); }
        } else {
            const int i = e - 9;
            for (int v = 0; v < kV; ++v) { const float w = __ldg(c.weights + v * kJ + k); if (w != 0.f) acc = fmaf(w, s_dv[v * 3 + i], acc); }
        }
        dA[(long)r * 192 + threadIdx.x] = acc;
    }
    if (threadIdx.x < 3) {
        float acc = 0.f;
        for (int v = 0; v < kV; ++v) acc += s_dv[v * 3 + threadIdx.x];
        dcen[(long)r * 4 + threadIdx.x] = -acc;
    }
    for (int v = threadIdx.x; v < kV; v += blockDim.x) {
        float Tr[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) Tr[i] = 0.f;
        for (int k = 0; k < kJ; ++k) {
            const float w = __ldg(c.weights + v * kJ + k);
            if (w != 0.f)
#pragma unroll
                for (int i = 0; i < 9; ++i) Tr[i] = fmaf(w, s_A[k * 12 + i], Tr[i]);
        }
        float o[3];
        mat3t_vec(Tr, &s_dv[v * 3], o);
        dvp[(long)r * 2336 + v * 3 + 0] = o[0]; dvp[(long)r * 2336 + v * 3 + 1] = o[1]; dvp[(long)r * 2336 + v * 3 + 2] = o[2];
    }
}

// one step of the chain backward for joint k with parent p (see mano_math.cuh pose_bwd)
__device__ __forceinline__ void chain_bwd_step(WarpPose& W, int k, int p) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float a = 0.f, b = 0.f;
            for (int l = 0; l < 3; ++l) {
                a = fmaf(W.Gr[p][l * 3 + i], W.dGr[k][l * 3 + j], a);
                b = fmaf(W.dGr[k][i * 3 + l], W.R[k][j * 3 + l], b);
            }
            W.dR[k][i * 3 + j] += a;
            W.dGr[p][i * 3 + j] += b;
        }
    const float d[3] = {W.J[k][0] - W.J[p][0], W.J[k][1] - W.J[p][1], W.J[k][2] - W.J[p][2]};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) W.dGr[p][i * 3 + j] = fmaf(W.dGt[k][i], d[j], W.dGr[p][i * 3 + j]);
    float o[3];
    mat3t_vec(W.Gr[p], W.dGt[k], o);
    for (int cc = 0; cc < 3; ++cc) { W.dJ[k][cc] += o[cc]; W.dJ[p][cc] -= o[cc]; W.dGt[p][cc] += W.dGt[k][cc]; }
}

// Backward of the pose / chain for the row a warp owns.  On entry W holds the forward state (pose_fwd_warp), W.dj the joint
// gradients (already scaled to metres; read when have_dj), W.dA / W.dpm the vertex-side seeds (zero on the joints-only path); db
// (lane b < 10: dbeta seed) and dc (lane cc < 3: centre gradient seed) are per-lane.  With tips_in_kernel the five tip vertices'
// gradients are formed here from W.dj.  Leaves W.dpose (the 48 axis-angle gradients) and returns this lane's dbeta.
__device__ __forceinline__ float pose_bwd_warp(const PoseTables& s_t, WarpPose& W, int lane, int order, bool tips_in_kernel, bool have_dj,
                                               float db, float dc) {
    if (tips_in_kernel && have_dj) {
        for (int t = 0; t < 5; ++t) {
            int slot = 0;
            for (int i = 0; i < kNJ; ++i) if (c_jtr_src[order][i] == kJ + t) slot = i;
            const float dv0 = W.dj[slot * 3 + 0], dv1 = W.dj[slot * 3 + 1], dv2 = W.dj[slot * 3 + 2];
            const float dv[3] = {dv0, dv1, dv2};
            float vp[3];
            tip_skin_warp(s_t, t, W, lane, vp);
            for (int e = lane; e < kWsA; e += 32) {
                const int k = e / 12, ee = e % 12;
                const float w = s_t.wtip[t][k];
                W.dA[k][ee] += w * (ee < 9 ? dv[ee / 3] * vp[ee % 3] : dv[ee - 9]);
            }
            float dvp[3];
            mat3t_vec(W.T, dv, dvp);
            for (int k = lane; k < kPoseMap; k += 32)
                W.dpm[k] += s_t.ptip[t][k][0] * dvp[0] + s_t.ptip[t][k][1] * dvp[1] + s_t.ptip[t][k][2] * dvp[2];
            if (lane < kShape)
                db += s_t.stip[t][0][lane] * dvp[0] + s_t.stip[t][1][lane] * dvp[1] + s_t.stip[t][2][lane] * dvp[2];
            if (lane < 3) dc -= dv[lane];
            __syncwarp();
        }
    }
    // joint-position gradients: chain joints of djtr, and the centring (every output is relative to joint 4)
    for (int i = lane; i < kJ * 3; i += 32) W.dGt_out[i] = 0.f;
    __syncwarp();
    if (have_dj && lane < 3) {
        for (int i = 0; i < kNJ; ++i) {
            const int src = c_jtr_src[order][i];
            if (src < kJ) {
                const float g = W.dj[i * 3 + lane];
                W.dGt_out[src * 3 + lane] += g;
                dc -= g;                 // the tips' share of the centring is already in dc
            }
        }
    }
    __syncwarp();
    if (lane < 3) W.dGt_out[kCenterJoint * 3 + lane] += dc;
    __syncwarp();
    // chain backward
    if (lane < kJ) {
        const int k = lane;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) W.dGr[k][i * 3 + j] = W.dA[k][i * 3 + j] - W.dA[k][9 + i] * W.J[k][j];
        float o[3];
        mat3t_vec(W.Gr[k], &W.dA[k][9], o);
        for (int cc = 0; cc < 3; ++cc) { W.dGt[k][cc] = W.dGt_out[k * 3 + cc] + W.dA[k][9 + cc]; W.dJ[k][cc] = -o[cc]; }
        for (int i = 0; i < 9; ++i) W.dR[k][i] = (k >= 1) ? W.dpm[(k - 1) * 9 + i] : 0.f;
    }
    __syncwarp();
    for (int level = 2; level >= 1; --level) {      // children of distinct parents: one lane per finger
        if (lane < 5) { const int k = 1 + 3 * lane + level; chain_bwd_step(W, k, k - 1); }
        __syncwarp();
    }
    if (lane == 0) {                                 // the five level-0 joints share parent 0: serial
        for (int f = 0; f < 5; ++f) chain_bwd_step(W, 1 + 3 * f, 0);
        for (int i = 0; i < 9; ++i) W.dR[0][i] += W.dGr[0][i];
        for (int cc = 0; cc < 3; ++cc) W.dJ[0][cc] += W.dGt[0][cc];
    }
    __syncwarp();
    if (lane < kShape)
        for (int i = 0; i < kJ * 3; ++i) db = fmaf(s_t.js[i * kShape + lane], W.dJ[i / 3][i % 3], db);
    if (lane < kJ) rodrigues_bwd(&W.pose[3 * lane], W.dR[lane], &W.dpose[3 * lane]);
    __syncwarp();
    return db;
}
// dtheta[i] from W.dpose: the root rotation directly, the 45 PCA coefficients through the component matrix
__device__ __forceinline__ float pose_dtheta(const PoseTables& s_t, const WarpPose& W, int i) {
    if (i < 3) return W.dpose[i];
    float g = 0.f;
    for (int j = 0; j < 45; ++j) g = fmaf(s_t.comps[(i - 3) * 45 + j], W.dpose[3 + j], g);
    return g;
}

// Warp per row.  Vertex gradients arrive either through the workspace (dA, dpm, dbv, dcen written by the mesh kernels) or, with
// tips_in_kernel, are formed here from djtr for the five tip vertices (the joints-only training path: one kernel for the whole
// MANO backward).
__global__ void __launch_bounds__(kPoseWarps * 32) mano_pose_bwd_kernel(mhe_mano_consts c, const float* __restrict__ theta, int ld_theta,
                                     const float* __restrict__ beta, int ld_beta, int R, int order, int tips_in_kernel,
                                     const float* __restrict__ djtr, const float* __restrict__ dA_g, const float* __restrict__ dpm_g,
                                     const float* __restrict__ dbv_g, const float* __restrict__ dcen_g,
                                     float* __restrict__ dtheta, int ld_dtheta, float* __restrict__ dbeta, int ld_dbeta, int accumulate) {
    __shared__ WarpPose s_w[kPoseWarps];
    __shared__ PoseTables s_t;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * kPoseWarps + warp;
    stage_pose_tables(c, s_t, tips_in_kernel != 0);
    __syncthreads();
    if (r >= R) return;
    WarpPose& W = s_w[warp];
    const float* be = beta + (long)r * ld_beta;
    if (djtr) for (int i = lane; i < kNJ * 3; i += 32) W.dj[i] = djtr[(long)r * kNJ * 3 + i] * kMM;   // this row's joint gradients, scaled
    pose_fwd_warp(s_t, theta + (long)r * ld_theta, be, W, lane);

    // vertex-side gradients -> W.dA, W.dpm, dbeta seed, centre gradient
    float db = 0.f;                       // lane b < 10 owns dbeta[b]
    float dc = 0.f;                       // lane cc < 3 owns the centre gradient component
    for (int i = lane; i < kWsA; i += 32) W.dA[i / 12][i % 12] = (dA_g && !tips_in_kernel) ? dA_g[(long)r * 192 + i] : 0.f;
    for (int k = lane; k < kWsPm; k += 32) W.dpm[k] = (dpm_g && !tips_in_kernel) ? dpm_g[(long)r * 136 + k] : 0.f;
    if (!tips_in_kernel) {
        if (dbv_g && lane < kShape) db = dbv_g[(long)r * 12 + lane];
        if (dcen_g && lane < 3) dc = dcen_g[(long)r * 4 + lane];
    }
    __syncwarp();
    db = pose_bwd_warp(s_t, W, lane, order, tips_in_kernel != 0, djtr != nullptr, db, dc);
    if (lane < kShape) {
        float* d = dbeta + (long)r * ld_dbeta + lane;
        *d = accumulate ? *d + db : db;
    }
    for (int i = lane; i < kPose; i += 32) {
        const float g = pose_dtheta(s_t, W, i);
        float* d = dtheta + (long)r * ld_dtheta + i;
        *d = accumulate ? *d + g : g;
    }
}

// The joints-only training path of one hypothesis in ONE launch (warp per row): MANO pose / chain / tip vertices forward, root / bone
// normalisation + orthographic projection + Laplace(visible) + priors, and their backward down to dz.  The loss is linear in the row
// terms, loss = -mean_b mean_n (row_log_p - log_q) (network.py:793-808, criteria.py:55,173), so with dloss = 1 every row's gradient
// seed is the constant -1 / (B N) and nothing of the backward waits for a reduction.  z [R][61] = th3 | th45 | bt | logs | t.
__global__ void __launch_bounds__(kPoseWarps * 32) hypothesis_rows_kernel(mhe_mano_consts c, mhe_loss_cfg cfg, const float* __restrict__ z,
                                     const float* __restrict__ x_flow, const float* __restrict__ z_det,
                                     const float* __restrict__ crop_uv, const float* __restrict__ vis, int R, int B, int order, float dloss,
                                     const float* __restrict__ dlog_p,
                                     float* __restrict__ jtr, float* __restrict__ uv, float* __restrict__ row_lp, float* __restrict__ dz,
                                     float* __restrict__ dx_flow, float* __restrict__ dlog_q) {
    __shared__ WarpPose s_w[kPoseWarps];
    __shared__ PoseTables s_t;
    __shared__ float s_jo[kPoseWarps][64], s_dz[kPoseWarps][64], s_z[kPoseWarps][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * kPoseWarps + warp;
    stage_pose_tables(c, s_t, true);
    __syncthreads();
    if (r >= R) return;
    WarpPose& W = s_w[warp];
    float* jo = s_jo[warp];
    float* dzr = s_dz[warp];
    const int b = r % B;
    const float* zz;
    if (z) zz = z + (long)r * loss::kZ;
    else {   // z = th3 | th45 (flow) | bt | logs | t assembled here (network.py:703-717): no separate z kernel in front of this one
        float* zr = s_z[warp];
        for (int i = lane; i < loss::kZ; i += 32)
            zr[i] = i < 3 ? z_det[b * 16 + i] : (i < 48 ? x_flow[(long)r * 45 + i - 3] : z_det[b * 16 + i - 45]);
        __syncwarp();
        zz = zr;
    }
    pose_fwd_warp(s_t, zz, zz + 48, W, lane);
    // the 21 output joints (chain joints + tip vertices, reordered, centred, millimetres)
    for (int i = lane; i < kNJ * 3; i += 32) {
        const int src = c_jtr_src[order][i / 3];
        if (src < kJ) jo[i] = (W.Gt[src][i % 3] - W.Gt[kCenterJoint][i % 3]) * kMM;
    }
    for (int t = 0; t < 5; ++t) {
        float vp[3];
        tip_skin_warp(s_t, t, W, lane, vp);
        if (lane < 3) {
            const float o = (W.T[lane * 3 + 0] * vp[0] + W.T[lane * 3 + 1] * vp[1] + W.T[lane * 3 + 2] * vp[2] + W.T[9 + lane] - W.Gt[kCenterJoint][lane]) * kMM;
            for (int i = 0; i < kNJ; ++i) if (c_jtr_src[order][i] == kJ + t) jo[i * 3 + lane] = o;
        }
        __syncwarp();
    }
    if (jtr) for (int i = lane; i < kNJ * 3; i += 32) jtr[(long)r * kNJ * 3 + i] = jo[i];
    // reprojection + likelihood + priors, and their gradient with the constant seed
    const loss::RowGeom g = loss::row_geom(cfg, jo, zz, lane);
    const float lp = loss::reproj_row_fwd(cfg, g, zz, crop_uv + b * 42, vis + b * kNJ, lane, uv ? uv + (long)r * 42 : nullptr);
    // gradient seed of the row: dL/d row_log_p = dL/dlog_p[b] / N.  With the criterion's loss = -mean_b log_p (criteria.py:55,173) that is
    // the constant -dloss / (B N); a caller differentiating something else passes its own dL/dlog_p per image.
    const float gr = dlog_p ? __ldg(dlog_p + b) / (float)(R / B) : -dloss / (float)R;
    if (lane == 0) { row_lp[r] = lp; if (dlog_q) dlog_q[r] = -gr; }
    loss::reproj_row_bwd(cfg, g, zz, crop_uv + b * 42, vis + b * kNJ, gr, lane, W.dj, dzr);
    __syncwarp();
    for (int i = lane; i < kNJ * 3; i += 32) W.dj[i] *= kMM;
    for (int i = lane; i < kWsA; i += 32) W.dA[i / 12][i % 12] = 0.f;
    for (int k = lane; k < kWsPm; k += 32) W.dpm[k] = 0.f;
    __syncwarp();
    const float db = pose_bwd_warp(s_t, W, lane, order, true, true, 0.f, 0.f);
    if (lane < kShape) dzr[48 + lane] += db;
    for (int i = lane; i < kPose; i += 32) dzr[i] += pose_dtheta(s_t, W, i);
    __syncwarp();
    for (int i = lane; i < loss::kZ; i += 32) dz[(long)r * loss::kZ + i] = dzr[i];
    if (dx_flow) for (int i = lane; i < 45; i += 32) dx_flow[(long)r * 45 + i] = dzr[3 + i];     // the flow's share of dz (network.py:703-717)
}

// ---- tensor-core mesh path ----------------------------------------------------------------------------------------------
// (1) blend GEMM   vp[r][v*3+c] = sum_k P[r][k] Bd[k][v*3+c],  P = [pose map | beta | 1],  Bd = [posedirs ; shapedirs^T ; v_template]
// (2) skinning GEMM T[v][(r, e)] = sum_j weights[v][j] A[r][j][e]  (e < 12: the 3 x 4 transform), epilogue: verts = T . vp - centre
// Operands are split half planes (3-pass split precision, fp32 accumulate), K padded to the 64-deep k-block of the GEMM kernel.
__global__ void __launch_bounds__(256) mano_pack_blend_planes_kernel(mhe_mano_consts c, uint16_t* __restrict__ bd, uint16_t* __restrict__ wp) {
    const long nbd = (long)kPmK * kOffLd, nw = (long)kVp * 64;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nbd + nw; i += (long)gridDim.x * blockDim.x) {
        float v = 0.f;
        uint16_t* dst;
        long plane;
        if (i < nbd) {
            const int k = (int)(i / kOffLd), col = (int)(i % kOffLd);
            if (col < kVC) {
                if (k < kPoseMap) v = c.posedirs_t[(long)k * kVC + col];
                else if (k < kPoseMap + kShape) v = c.shapedirs[(long)col * kShape + (k - kPoseMap)];
                else if (k == kPoseMap + kShape) v = c.v_template[col];
            }
            dst = bd + i; plane = nbd;
        } else {
            const long j = i - nbd;
            const int vtx = (int)(j / 64), k = (int)(j % 64);
            if (vtx < kV && k < kJ) v = c.weights[vtx * kJ + k];
            dst = wp + j; plane = nw;
        }
        const uint16_t h = tc::to16<true>(v);
        dst[0] = h;
        dst[plane] = tc::to16<true>(v - tc::from16<true>(h));
    }
}
// per-row operands: P planes [2][R][192] and the transposed transforms [2][R*16][64] (element (r*16 + e, j) = A[r][j][e])
__global__ void __launch_bounds__(256) mano_mesh_operands_kernel(const float* __restrict__ pm, const float* __restrict__ A, const float* __restrict__ beta,
                                                                 int ld_beta, int R, uint16_t* __restrict__ pmp, uint16_t* __restrict__ lbsb) {
    const long np = (long)R * kPmK, nl = lbsb ? (long)R * kLbsN * 64 : 0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < np + nl; i += (long)gridDim.x * blockDim.x) {
        float v = 0.f;
        uint16_t* dst;
        long plane;
        if (i < np) {
            const int r = (int)(i / kPmK), k = (int)(i % kPmK);
            if (k < kPoseMap) v = pm[(long)r * kWsPm + k];
            else if (k < kPoseMap + kShape) v = beta[(long)r * ld_beta + k - kPoseMap];
            else if (k == kPoseMap + kShape) v = 1.f;
            dst = pmp + i; plane = np;
        } else {
            const long j = i - np;
            const int k = (int)(j % 64), e = (int)((j / 64) % kLbsN), r = (int)(j / (64 * kLbsN));
            if (k < kJ && e < 12) v = A[(long)r * kWsA + k * 12 + e];
            dst = lbsb + j; plane = nl;
        }
        const uint16_t h = tc::to16<true>(v);
        dst[0] = h;
        dst[plane] = tc::to16<true>(v - tc::from16<true>(h));
    }
}
// epilogue of the skinning GEMM: thread = vertex (TMEM lane), 32 columns = the padded transforms of two consecutive rows
struct EpiSkin {
    static constexpr bool kDirect = true, kStaged = false, kRmw = false;
    const float* vp; const float* cen; float* verts; float* jtr; int R, order;
    __device__ void operator()(int, int, int vtx, int col0, float* t, const tc::GemmShape&) const {
        int tip = -1;
#pragma unroll
        for (int q = 0; q < 5; ++q) if (c_tip_vert[q] == vtx) tip = q;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = (col0 >> 4) + h;
            if (r >= R) continue;
            const float* T = t + 16 * h;
            const float* p = vp + (long)r * kOffLd + vtx * 3;
            const float p0 = p[0], p1 = p[1], p2 = p[2];
            const float* cc = cen + (long)r * kWsCen;
            float o[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) o[i] = (T[i * 3 + 0] * p0 + T[i * 3 + 1] * p1 + T[i * 3 + 2] * p2 + T[9 + i] - cc[i]) * kMM;
            float* d = verts + ((long)r * kV + vtx) * 3;
            d[0] = o[0]; d[1] = o[1]; d[2] = o[2];
            if (tip >= 0 && jtr)
                for (int i = 0; i < kNJ; ++i)
                    if (c_jtr_src[order][i] == kJ + tip) { float* dj = jtr + ((long)r * kNJ + i) * 3; dj[0] = o[0]; dj[1] = o[1]; dj[2] = o[2]; }
        }
    }
    __device__ void elem(int, int, int, int, float, const tc::GemmShape&) const {}
};

// epilogue of the pose-blend GEMM: row = hypothesis (TMEM lane), 32 consecutive vertex coordinates -> pose offsets [R][kOffLd]
struct EpiPoseOffsets {
    static constexpr bool kDirect = true, kStaged = false, kRmw = false, kTile8 = true;
    float* out;
    struct Pre {};
    __device__ void pre8(int, int, int, Pre&) const {}
    __device__ void tile8(int, int, int row, int col, float* v, const Pre&, const tc::GemmShape&) const {   // 128 contiguous bytes per row and 4 lanes
        if (col >= kOffLd) return;
        float4* o = reinterpret_cast<float4*>(out + (long)row * kOffLd + col);
        o[0] = make_float4(v[0], v[1], v[2], v[3]);
        o[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    __device__ void operator()(int, int, int row, int col0, float* v, const tc::GemmShape&) const {
        float4* o = reinterpret_cast<float4*>(out + (long)row * kOffLd + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    __device__ void elem(int, int, int, int, float, const tc::GemmShape&) const {}
};

}  // namespace mhe

using namespace mhe;

extern "C" {

size_t mhe_mano_posedirs_planes_bytes(void) { return ((size_t)2 * kPmK * kOffLd + (size_t)2 * kVp * 64) * 2; }

int mhe_mano_pack_posedirs_planes(const mhe_mano_consts* c, void* posedirs_planes, void* stream) {
    MHE_REQUIRE(c && c->posedirs_t && c->shapedirs && c->v_template && c->weights && posedirs_planes && ((uintptr_t)posedirs_planes & 15) == 0,
                "mano_pack_posedirs_planes: bad args");
    // [posedirs ; shapedirs^T ; v_template] as half planes [2][192][2336] (the MN-major B operand of the blend GEMM), then the skinning
    // weights as half planes [2][896][64] (the A operand of the skinning GEMM), zero padded
    uint16_t* bd = (uint16_t*)posedirs_planes;
    mano_pack_blend_planes_kernel<<<592, 256, 0, (cudaStream_t)stream>>>(*c, bd, bd + (size_t)2 * kPmK * kOffLd);
    return check_launch("mano pack blend planes");
}

size_t mhe_mano_pose_tables_floats(void) { return sizeof(PoseTables) / sizeof(float); }

int mhe_mano_pack_pose_tables(const mhe_mano_consts* c, float* pose_tables, void* stream) {
    MHE_REQUIRE(c && pose_tables && ((uintptr_t)pose_tables & 15) == 0, "mano_pack_pose_tables: bad args");
    mhe_mano_consts src = *c;
    src.pose_tables = nullptr;
    mano_pack_pose_tables_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(src, pose_tables);
    return check_launch("mano pack pose tables");
}

size_t mhe_mano_workspace_bytes(int R, int mesh_grad) { return R < 0 ? 0 : ManoWs::floats(R, mesh_grad != 0) * sizeof(float); }

int mhe_mano_fwd(const mhe_mano_consts* c, const float* theta, int ld_theta, const float* beta, int ld_beta,
                 int R, int joint_order, float* verts, float* jtr, float* joints2,
                 void* workspace, size_t workspace_bytes, void* stream_) {
    if (R == 0) return MHE_OK;
    MHE_REQUIRE(c && theta && beta && jtr && workspace, "mano_fwd: null pointer");
    MHE_REQUIRE(R >= 0 && ld_theta >= 48 && ld_beta >= 10 && joint_order >= 0 && joint_order <= 1, "mano_fwd: bad sizes");
    MHE_REQUIRE(!joints2 || verts, "mano_fwd: joints2 needs verts");
    if (workspace_bytes < mhe_mano_workspace_bytes(R, 0)) { set_error("mano_fwd: workspace too small"); return MHE_ERR_WORKSPACE; }
    if (R == 0) return MHE_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    ManoWs ws((float*)workspace, R, false);
    mano_pose_fwd_kernel<<<cdiv(R, kPoseWarps), kPoseWarps * 32, 0, stream>>>(*c, theta, ld_theta, beta, ld_beta, R, joint_order, verts ? 0 : 1,
                                                                                 verts ? ws.pm : nullptr, verts ? ws.A : nullptr, verts ? ws.cen : nullptr, jtr);
    MHE_TRY(check_launch("mano pose fwd"));
    if (verts) {
        if (c->posedirs_planes) {   // both contractions of the mesh on the tensor cores (see above)
            const uint16_t* bd = (const uint16_t*)c->posedirs_planes;
            const uint16_t* wp = bd + (size_t)2 * kPmK * kOffLd;
            static const bool tc_skin = getenv("MHE_MANO_TC_SKIN") != nullptr;     // (round 1's generic-GEMM skinning: diagnostic only)
            static const bool simt_skin = [] { const char* e = getenv("MHE_MANO_SKIN"); return e && !strcmp(e, "simt"); }();
            const long nel = (long)R * (kPmK + (tc_skin ? kLbsN * 64 : 0));
            const int nblk = (int)((nel + 255) / 256 < 1184 ? (nel + 255) / 256 : 1184);      // grid-stride beyond 8 blocks per SM
            mano_mesh_operands_kernel<<<nblk, 256, 0, stream>>>(ws.pm, ws.A, beta, ld_beta, R, (uint16_t*)ws.pmp, tc_skin ? (uint16_t*)ws.lbsb : nullptr);
            MHE_TRY(check_launch("mano mesh operands"));
            tc::PlaneTensor A, Bp;
            A.base = (const __nv_bfloat16*)ws.pmp; A.cols = kPmK; A.rows = R; A.planes = 2; A.batches = 1;
            A.row_pitch = kPmK; A.plane_stride = (long)R * kPmK; A.batch_stride = (long)2 * R * kPmK;
            Bp.base = (const __nv_bfloat16*)bd; Bp.cols = kOffLd; Bp.rows = kPmK; Bp.planes = 2; Bp.batches = 1;
            Bp.row_pitch = kOffLd; Bp.plane_stride = (long)kPmK * kOffLd; Bp.batch_stride = (long)2 * kPmK * kOffLd;
            tc::GemmShape g{R, kOffLd, kPmK, 1, 1, 1, 1};
            EpiPoseOffsets e{ws.poff};
            MHE_TRY((tc::launch_tc_gemm<128, false, true, 3, true>(A, Bp, g, e, stream, "mano blend")));
            if (!tc_skin && !simt_skin) {   // default: the dedicated tensor-core skinning kernel (mano_skin_tc.cu)
                MHE_TRY(skin::launch(wp, ws.A, ws.cen, ws.poff, kOffLd, verts, jtr, R, joint_order, stream));
            } else
            if (!tc_skin) {   // MHE_MANO_SKIN=simt: skinning on the CUDA cores from the GEMM's vertices (round 1; instruction-bound)
                if (R >= 512) mano_skin_fwd_kernel<8><<<dim3(cdiv(kV, 128), cdiv(R, 8)), 128, 0, stream>>>(*c, beta, ld_beta, R, joint_order, ws.pm, ws.A, ws.cen, verts, jtr, nullptr, ws.poff);
                else mano_skin_fwd_kernel<2><<<dim3(cdiv(kV, 128), cdiv(R, 2)), 128, 0, stream>>>(*c, beta, ld_beta, R, joint_order, ws.pm, ws.A, ws.cen, verts, jtr, nullptr, ws.poff);
                MHE_TRY(check_launch("mano skin fwd"));
            } else {
            tc::PlaneTensor Wt, Tr;
            Wt.base = (const __nv_bfloat16*)wp; Wt.cols = 64; Wt.rows = kVp; Wt.planes = 2; Wt.batches = 1;
            Wt.row_pitch = 64; Wt.plane_stride = (long)kVp * 64; Wt.batch_stride = (long)2 * kVp * 64;
            Tr.base = (const __nv_bfloat16*)ws.lbsb; Tr.cols = 64; Tr.rows = R * kLbsN; Tr.planes = 2; Tr.batches = 1;
            Tr.row_pitch = 64; Tr.plane_stride = (long)R * kLbsN * 64; Tr.batch_stride = (long)2 * R * kLbsN * 64;
            tc::GemmShape gs{kV, R * kLbsN, 64, 1, 1, 1, 1};
            EpiSkin es{ws.poff, ws.cen, verts, jtr, R, joint_order};
            MHE_TRY((tc::launch_tc_gemm<128, false, false, 3, true>(Wt, Tr, gs, es, stream, "mano skinning")));
            }
        } else
        if (R >= 512) {
            dim3 grid(cdiv(kV, 128), cdiv(R, 8));
            mano_skin_fwd_kernel<8><<<grid, 128, 0, stream>>>(*c, beta, ld_beta, R, joint_order, ws.pm, ws.A, ws.cen, verts, jtr, nullptr, nullptr);
            MHE_TRY(check_launch("mano skin fwd"));
        } else {
            dim3 grid(cdiv(kV, 128), cdiv(R, 2));
            mano_skin_fwd_kernel<2><<<grid, 128, 0, stream>>>(*c, beta, ld_beta, R, joint_order, ws.pm, ws.A, ws.cen, verts, jtr, nullptr, nullptr);
            MHE_TRY(check_launch("mano skin fwd"));
        }
        if (joints2) {
            mano_joints2_fwd_kernel<<<R, 256, 0, stream>>>(*c, verts, R, joint_order, joints2);
            MHE_TRY(check_launch("mano joints2 fwd"));
        }
    }
    return MHE_OK;
}

int mhe_mano_bwd(const mhe_mano_consts* c, const float* theta, int ld_theta, const float* beta, int ld_beta,
                 int R, int joint_order, const float* dverts, const float* djtr, const float* djoints2,
                 float* dtheta, int ld_dtheta, float* dbeta, int ld_dbeta, int accumulate,
                 void* workspace, size_t workspace_bytes, void* stream_) {
    if (R == 0) return MHE_OK;
    MHE_REQUIRE(c && theta && beta && dtheta && dbeta && workspace, "mano_bwd: null pointer");
    MHE_REQUIRE(R >= 0 && ld_theta >= 48 && ld_beta >= 10 && ld_dtheta >= 48 && ld_dbeta >= 10 && joint_order >= 0 && joint_order <= 1, "mano_bwd: bad sizes");
    const bool mesh = dverts || djoints2;
    if (workspace_bytes < mhe_mano_workspace_bytes(R, mesh)) { set_error("mano_bwd: workspace too small"); return MHE_ERR_WORKSPACE; }
    if (R == 0) return MHE_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    ManoWs ws((float*)workspace, R, mesh);
    if (mesh) {
        // recompute the forward state the mesh gradient needs
        mano_pose_fwd_kernel<<<cdiv(R, kPoseWarps), kPoseWarps * 32, 0, stream>>>(*c, theta, ld_theta, beta, ld_beta, R, joint_order, 0, ws.pm, ws.A,
                                                                                     ws.cen, nullptr);
        MHE_TRY(check_launch("mano pose recompute"));
        dim3 grid(cdiv(kV, 128), cdiv(R, 2));
        mano_skin_fwd_kernel<2><<<grid, 128, 0, stream>>>(*c, beta, ld_beta, R, joint_order, ws.pm, ws.A, ws.cen, nullptr, nullptr, ws.vp, nullptr);
        MHE_TRY(check_launch("mano vposed recompute"));
        mano_dverts_total_kernel<<<cdiv(R * kV, 256), 256, 0, stream>>>(*c, dverts, djtr, djoints2, R, joint_order, ws.dvt);
        MHE_TRY(check_launch("mano dverts total"));
        mano_skin_bwd_kernel<<<R, 192, 0, stream>>>(*c, R, ws.A, ws.vp, ws.dvt, ws.dA, ws.dcen, ws.dvp);
        MHE_TRY(check_launch("mano skin bwd"));
        {   // dpm [R][135] = dvp [R][2334] posedirs_t^T
            GemmArgs g; g.A = ws.dvp; g.lda = 2336; g.B = c->posedirs_t; g.ldb = kVC; g.M = R; g.N = kPoseMap; g.K = kVC;
            EpiStore e{ws.dpm, 136, 0};
            MHE_TRY((launch_sgemm<Major::K, Major::K>(g, e, stream, "mano dpm")));
        }
        {   // dbv [R][10] = dvp [R][2334] shapedirs [2334][10]
            GemmArgs g; g.A = ws.dvp; g.lda = 2336; g.B = c->shapedirs; g.ldb = kShape; g.M = R; g.N = kShape; g.K = kVC;
            EpiStore e{ws.dbv, 12, 0};
            MHE_TRY((launch_sgemm<Major::K, Major::MN>(g, e, stream, "mano dbv")));
        }
    }
    // mesh path: vertex gradients come through the workspace (tips already folded into dvt);
    // joints-only path: the tips are handled inside the kernel
    mano_pose_bwd_kernel<<<cdiv(R, kPoseWarps), kPoseWarps * 32, 0, stream>>>(*c, theta, ld_theta, beta, ld_beta, R, joint_order, mesh ? 0 : 1, djtr,
                                                                                 mesh ? ws.dA : nullptr, mesh ? ws.dpm : nullptr, mesh ? ws.dbv : nullptr,
                                                                                 mesh ? ws.dcen : nullptr, dtheta, ld_dtheta, dbeta, ld_dbeta, accumulate);
    return check_launch("mano pose bwd");
}

int mhe_hypothesis_rows_fwd_bwd(const mhe_mano_consts* c, const mhe_loss_cfg* cfg, const float* z, const float* x_flow, const float* z_det,
                                const float* crop_uv, const float* vis, int R, int B, int joint_order, float dloss, const float* dlog_p,
                                float* jtr, float* uv, float* row_log_p, float* dz, float* dx_flow, float* dlog_q, void* stream) {
    if (R == 0) return MHE_OK;
    MHE_REQUIRE(c && cfg && (z || (x_flow && z_det)) && crop_uv && vis && row_log_p && dz, "hypothesis_rows_fwd_bwd: null pointer");
    MHE_REQUIRE(R > 0 && B > 0 && R % B == 0 && joint_order >= 0 && joint_order <= 1, "hypothesis_rows_fwd_bwd: bad sizes");
    hypothesis_rows_kernel<<<cdiv(R, kPoseWarps), kPoseWarps * 32, 0, (cudaStream_t)stream>>>(*c, *cfg, z, x_flow, z_det, crop_uv, vis, R, B, joint_order,
                                                                                               dloss, dlog_p, jtr, uv, row_log_p, dz, dx_flow, dlog_q);
    return check_launch("hypothesis rows fwd+bwd");
}

}  // extern "C"
