// MANO layer kernels: pose/chain (thread per row), blend shapes + linear blend skinning (thread per
// vertex, row-tiled so each posedirs element is reused across the tile), regressed joint set, and
// the exact backward.  Reference: hand/manopth/manolayer.py:110-274, hand/ManoLayer.py:45-60,141-148.
#include "gemm_simt.cuh"
#include "mano_math.cuh"

namespace mhe {
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

struct ManoWs {
    float *pm, *A, *cen, *dA, *dpm, *dbv, *dcen, *vp, *dvt, *dvp;
    ManoWs(float* base, int R, bool mesh) {
        auto take = [&](size_t n) { float* p = base; base += (n + 63) / 64 * 64; return p; };
        pm = take((size_t)R * kWsPm); A = take((size_t)R * kWsA); cen = take((size_t)R * kWsCen);
        dA = take((size_t)R * 192); dpm = take((size_t)R * 136); dbv = take((size_t)R * 12); dcen = take((size_t)R * 4);
        vp = dvt = dvp = nullptr;
        if (mesh) { vp = take((size_t)R * 2336); dvt = take((size_t)R * 2336); dvp = take((size_t)R * 2336); }
    }
    static size_t floats(int R, bool mesh) {
        return (size_t)R * (kWsRowFwd + kWsRowBwd + (mesh ? kWsRowMesh : 0)) + 10 * 64;
    }
};

// ---- forward ------------------------------------------------------------------------------------
__global__ void mano_pose_fwd_kernel(mhe_mano_consts c, const float* __restrict__ theta, int ld_theta,
                                     const float* __restrict__ beta, int ld_beta, int R, int order,
                                     float* __restrict__ pm, float* __restrict__ A, float* __restrict__ cen,
                                     float* __restrict__ jtr) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    PoseState st;
    pose_fwd(c.comps, c.hands_mean, c.jt, c.js, theta + (long)r * ld_theta, beta + (long)r * ld_beta, st);
    for (int k = 1; k < kJ; ++k)
        for (int i = 0; i < 9; ++i) pm[(long)r * kWsPm + (k - 1) * 9 + i] = st.R[k][i] - ((i % 4 == 0) ? 1.f : 0.f);
    for (int k = 0; k < kJ; ++k) skin_transform(st, k, A + (long)r * kWsA + k * 12);
    for (int cc = 0; cc < 3; ++cc) cen[(long)r * kWsCen + cc] = st.Gt[kCenterJoint][cc];
    for (int i = 0; i < kNJ; ++i) {
        const int src = c_jtr_src[order][i];
        if (src < kJ)
            for (int cc = 0; cc < 3; ++cc) jtr[((long)r * kNJ + i) * 3 + cc] = (st.Gt[src][cc] - st.Gt[kCenterJoint][cc]) * kMM;
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
                                                            float* __restrict__ verts, float* __restrict__ jtr, float* __restrict__ vp_out) {
    __shared__ float s_pm[RT][kWsPm];
    __shared__ float s_A[RT][kWsA];
    __shared__ float s_beta[RT][12];
    __shared__ float s_cen[RT][4];
    const int r0 = blockIdx.y * RT;
    const int nr = min(RT, R - r0);
    for (int i = threadIdx.x; i < RT * kWsPm; i += blockDim.x) { const int rr = i / kWsPm; s_pm[rr][i % kWsPm] = rr < nr ? pm_g[(long)(r0 + rr) * kWsPm + i % kWsPm] : 0.f; }
    for (int i = threadIdx.x; i < RT * kWsA; i += blockDim.x) { const int rr = i / kWsA; s_A[rr][i % kWsA] = rr < nr ? A_g[(long)(r0 + rr) * kWsA + i % kWsA] : 0.f; }
    for (int i = threadIdx.x; i < RT * kShape; i += blockDim.x) { const int rr = i / kShape; s_beta[rr][i % kShape] = rr < nr ? beta[(long)(r0 + rr) * ld_beta + i % kShape] : 0.f; }
    for (int i = threadIdx.x; i < RT * 3; i += blockDim.x) { const int rr = i / 3; s_cen[rr][i % 3] = rr < nr ? cen_g[(long)(r0 + rr) * kWsCen + i % 3] : 0.f; }
    __syncthreads();
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= kV) return;

    float vp[RT][3];
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
            if (w[k] != 0.f)
#pragma unroll
                for (int i = 0; i < 12; ++i) T[i] = fmaf(w[k], s_A[rr][k * 12 + i], T[i]);
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

// joints-only path: skin just the five tip vertices. thread per (row, tip).
__global__ void mano_tips_fwd_kernel(mhe_mano_consts c, const float* __restrict__ beta, int ld_beta, int R, int order,
                                     const float* __restrict__ pm_g, const float* __restrict__ A_g, const float* __restrict__ cen_g,
                                     float* __restrict__ jtr) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R * 5) return;
    const int r = idx / 5, t = idx % 5;
    const int v = c_tip_vert[t];
    float vp[3], T[12];
    skin_vertex(c, v, pm_g + (long)r * kWsPm, A_g + (long)r * kWsA, beta + (long)r * ld_beta, vp, T);
    for (int i = 0; i < kNJ; ++i)
        if (c_jtr_src[order][i] == kJ + t)
            for (int cc = 0; cc < 3; ++cc)
                jtr[((long)r * kNJ + i) * 3 + cc] = (T[cc * 3 + 0] * vp[0] + T[cc * 3 + 1] * vp[1] + T[cc * 3 + 2] * vp[2] + T[9 + cc] - cen_g[(long)r * kWsCen + cc]) * kMM;
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

// joints-only backward through the five tip vertices. thread per row (5 tips sequentially; tiny).
__global__ void mano_tips_bwd_kernel(mhe_mano_consts c, const float* __restrict__ beta, int ld_beta, int R, int order,
                                     const float* __restrict__ pm_g, const float* __restrict__ A_g, const float* __restrict__ djtr,
                                     float* __restrict__ dA, float* __restrict__ dpm, float* __restrict__ dbv, float* __restrict__ dcen) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    float* dAr = dA + (long)r * 192;
    float* dpmr = dpm + (long)r * 136;
    float* dbr = dbv + (long)r * 12;
    for (int i = 0; i < 192; ++i) dAr[i] = 0.f;
    for (int i = 0; i < 136; ++i) dpmr[i] = 0.f;
    for (int i = 0; i < 12; ++i) dbr[i] = 0.f;
    float dc[3] = {0.f, 0.f, 0.f};
    for (int t = 0; t < 5; ++t) {
        const int v = c_tip_vert[t];
        float dv[3] = {0.f, 0.f, 0.f};
        for (int i = 0; i < kNJ; ++i)
            if (c_jtr_src[order][i] == kJ + t) for (int cc = 0; cc < 3; ++cc) dv[cc] = djtr[((long)r * kNJ + i) * 3 + cc] * kMM;
        float vp[3], T[12];
        skin_vertex(c, v, pm_g + (long)r * kWsPm, A_g + (long)r * kWsA, beta + (long)r * ld_beta, vp, T);
        for (int k = 0; k < kJ; ++k) {
            const float w = __ldg(c.weights + v * kJ + k);
            if (w == 0.f) continue;
            for (int i = 0; i < 3; ++i) {
                for (int j = 0; j < 3; ++j) dAr[k * 12 + i * 3 + j] = fmaf(w * dv[i], vp[j], dAr[k * 12 + i * 3 + j]);
                dAr[k * 12 + 9 + i] = fmaf(w, dv[i], dAr[k * 12 + 9 + i]);
            }
        }
        float dvp[3];
        mat3t_vec(T, dv, dvp);
        for (int k = 0; k < kPoseMap; ++k)
            dpmr[k] += __ldg(c.posedirs_t + (long)k * kVC + v * 3) * dvp[0] + __ldg(c.posedirs_t + (long)k * kVC + v * 3 + 1) * dvp[1] + __ldg(c.posedirs_t + (long)k * kVC + v * 3 + 2) * dvp[2];
        for (int b = 0; b < kShape; ++b)
            dbr[b] += __ldg(c.shapedirs + (v * 3 + 0) * kShape + b) * dvp[0] + __ldg(c.shapedirs + (v * 3 + 1) * kShape + b) * dvp[1] + __ldg(c.shapedirs + (v * 3 + 2) * kShape + b) * dvp[2];
        for (int cc = 0; cc < 3; ++cc) dc[cc] -= dv[cc];
    }
    for (int cc = 0; cc < 3; ++cc) dcen[(long)r * 4 + cc] = dc[cc];
}

// chain backward, thread per row. dA/dpm/dbv/dcen may be NULL (no vertex gradients at all).
__global__ void mano_pose_bwd_kernel(mhe_mano_consts c, const float* __restrict__ theta, int ld_theta, const float* __restrict__ beta, int ld_beta,
                                     int R, int order, const float* __restrict__ djtr, const float* __restrict__ dA, const float* __restrict__ dpm,
                                     const float* __restrict__ dbv, const float* __restrict__ dcen,
                                     float* __restrict__ dtheta, int ld_dtheta, float* __restrict__ dbeta, int ld_dbeta, int accumulate) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    PoseState st;
    pose_fwd(c.comps, c.hands_mean, c.jt, c.js, theta + (long)r * ld_theta, beta + (long)r * ld_beta, st);
    float dGt[kJ * 3];
    for (int i = 0; i < kJ * 3; ++i) dGt[i] = 0.f;
    float dc[3] = {0.f, 0.f, 0.f};
    if (dcen) for (int cc = 0; cc < 3; ++cc) dc[cc] = dcen[(long)r * 4 + cc];
    if (djtr) {
        for (int i = 0; i < kNJ; ++i) {
            const int src = c_jtr_src[order][i];
            for (int cc = 0; cc < 3; ++cc) {
                const float g = djtr[((long)r * kNJ + i) * 3 + cc] * kMM;
                if (src < kJ) { dGt[src * 3 + cc] += g; dc[cc] -= g; }  // tips' centring share is already in dcen
            }
        }
    }
    for (int cc = 0; cc < 3; ++cc) dGt[kCenterJoint * 3 + cc] += dc[cc];
    float dth[kPose], db[kShape];
    for (int i = 0; i < kPose; ++i) dth[i] = 0.f;
    for (int i = 0; i < kShape; ++i) db[i] = dbv ? dbv[(long)r * 12 + i] : 0.f;
    pose_bwd(c.comps, c.js, st, dGt, dA ? dA + (long)r * 192 : nullptr, dpm ? dpm + (long)r * 136 : nullptr, dth, db);
    for (int i = 0; i < kPose; ++i) { float* d = dtheta + (long)r * ld_dtheta + i; *d = accumulate ? *d + dth[i] : dth[i]; }
    for (int i = 0; i < kShape; ++i) { float* d = dbeta + (long)r * ld_dbeta + i; *d = accumulate ? *d + db[i] : db[i]; }
}

}  // namespace mhe

using namespace mhe;

extern "C" {

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
    mano_pose_fwd_kernel<<<cdiv(R, 64), 64, 0, stream>>>(*c, theta, ld_theta, beta, ld_beta, R, joint_order, ws.pm, ws.A, ws.cen, jtr);
    MHE_TRY(check_launch("mano pose fwd"));
    if (verts) {
        if (R >= 8 * 148) {
            dim3 grid(cdiv(kV, 128), cdiv(R, 8));
            mano_skin_fwd_kernel<8><<<grid, 128, 0, stream>>>(*c, beta, ld_beta, R, joint_order, ws.pm, ws.A, ws.cen, verts, jtr, nullptr);
        } else {
            dim3 grid(cdiv(kV, 128), cdiv(R, 2));
            mano_skin_fwd_kernel<2><<<grid, 128, 0, stream>>>(*c, beta, ld_beta, R, joint_order, ws.pm, ws.A, ws.cen, verts, jtr, nullptr);
        }
        MHE_TRY(check_launch("mano skin fwd"));
        if (joints2) {
            mano_joints2_fwd_kernel<<<R, 256, 0, stream>>>(*c, verts, R, joint_order, joints2);
            MHE_TRY(check_launch("mano joints2 fwd"));
        }
    } else {
        mano_tips_fwd_kernel<<<cdiv(R * 5, 128), 128, 0, stream>>>(*c, beta, ld_beta, R, joint_order, ws.pm, ws.A, ws.cen, jtr);
        MHE_TRY(check_launch("mano tips fwd"));
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
    // recompute the forward state the skinning gradient needs (jtr output of the pose kernel is not needed: pass scratch)
    const bool need_skin = mesh || djtr;
    if (need_skin) {
        mano_pose_fwd_kernel<<<cdiv(R, 64), 64, 0, stream>>>(*c, theta, ld_theta, beta, ld_beta, R, joint_order, ws.pm, ws.A, ws.cen, ws.dpm /*scratch >= 63 floats/row*/);
        MHE_TRY(check_launch("mano pose recompute"));
    }
    if (mesh) {
        dim3 grid(cdiv(kV, 128), cdiv(R, 2));
        mano_skin_fwd_kernel<2><<<grid, 128, 0, stream>>>(*c, beta, ld_beta, R, joint_order, ws.pm, ws.A, ws.cen, nullptr, nullptr, ws.vp);
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
    } else if (djtr) {
        mano_tips_bwd_kernel<<<cdiv(R, 64), 64, 0, stream>>>(*c, beta, ld_beta, R, joint_order, ws.pm, ws.A, djtr, ws.dA, ws.dpm, ws.dbv, ws.dcen);
        MHE_TRY(check_launch("mano tips bwd"));
    }
    mano_pose_bwd_kernel<<<cdiv(R, 64), 64, 0, stream>>>(*c, theta, ld_theta, beta, ld_beta, R, joint_order, djtr,
                                                          need_skin ? ws.dA : nullptr, need_skin ? ws.dpm : nullptr,
                                                          need_skin ? ws.dbv : nullptr, need_skin ? ws.dcen : nullptr,
                                                          dtheta, ld_dtheta, dbeta, ld_dbeta, accumulate);
    return check_launch("mano pose bwd");
}

}  // extern "C"
