// Per-row MANO pose mathematics (host + device): PCA pose, Rodrigues, joint regression from betas,
// the 16-joint kinematic chain and their exact gradients.
// Reference: hand/manopth/manolayer.py:131-149 (pose), rodrigues_layer.py:15-54, manolayer.py:181-234 (chain).
// Kept free of CUDA-only constructs so tests/hostmath can compile it with g++ and check the
// derivatives against autograd on the CPU oracle.
#pragma once
#include <math.h>

#ifndef MHE_HD
#if defined(__CUDACC__)
#define MHE_HD __host__ __device__ __forceinline__
#else
#define MHE_HD inline
#endif
#endif

namespace mhe {
namespace mano {

constexpr int kJ = 16;          // articulated joints
constexpr int kPose = 48;       // 3 root + 45 hand
constexpr int kShape = 10;
constexpr int kPoseMap = 135;   // 15 * 9
constexpr int kCenterJoint = 4; // manolayer.py:260 reorder[9] == 4  (center_idx = 9)

// parent of each joint (kintree_table[0]); parents precede children
MHE_HD int parent_of(int k) {
    // {-1,0,1,2,0,4,5,0,7,8,0,10,11,0,13,14}
    return (k == 0) ? -1 : ((k % 3 == 1) ? 0 : k - 1);
}

MHE_HD void mat3_mul(const float* a, const float* b, float* c) {  // c = a b
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            c[i * 3 + j] = a[i * 3 + 0] * b[0 * 3 + j] + a[i * 3 + 1] * b[1 * 3 + j] + a[i * 3 + 2] * b[2 * 3 + j];
}
MHE_HD void mat3_vec(const float* a, const float* v, float* o) {  // o = a v
    for (int i = 0; i < 3; ++i) o[i] = a[i * 3 + 0] * v[0] + a[i * 3 + 1] * v[1] + a[i * 3 + 2] * v[2];
}
MHE_HD void mat3t_vec(const float* a, const float* v, float* o) {  // o = a^T v
    for (int i = 0; i < 3; ++i) o[i] = a[0 * 3 + i] * v[0] + a[1 * 3 + i] * v[1] + a[2 * 3 + i] * v[2];
}

// axis-angle -> rotation, through the half-angle quaternion exactly as the reference does:
// n = |v + 1e-8|, q = (cos n/2, sin(n/2) v / n), q <- q/|q|, R = quat2mat(q)
MHE_HD void rodrigues_fwd(const float* v, float* R) {
    const float e = 1e-8f;
    const float a0 = v[0] + e, a1 = v[1] + e, a2 = v[2] + e;
    const float n = sqrtf(a0 * a0 + a1 * a1 + a2 * a2);
    const float h = 0.5f * n;
    const float c = cosf(h), s = sinf(h);
    float q0 = c, q1 = s * v[0] / n, q2 = s * v[1] / n, q3 = s * v[2] / n;
    const float qn = sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
    q0 /= qn; q1 /= qn; q2 /= qn; q3 /= qn;
    const float w = q0, x = q1, y = q2, z = q3;
    const float w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
    const float wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
    R[0] = w2 + x2 - y2 - z2; R[1] = 2 * xy - 2 * wz;     R[2] = 2 * wy + 2 * xz;
    R[3] = 2 * wz + 2 * xy;   R[4] = w2 - x2 + y2 - z2;   R[5] = 2 * yz - 2 * wx;
    R[6] = 2 * xz - 2 * wy;   R[7] = 2 * wx + 2 * yz;     R[8] = w2 - x2 - y2 + z2;
}

// dL/dR (9) -> dL/dv (3)
MHE_HD void rodrigues_bwd(const float* v, const float* g, float* dv) {
    const float e = 1e-8f;
    const float a0 = v[0] + e, a1 = v[1] + e, a2 = v[2] + e;
    const float n = sqrtf(a0 * a0 + a1 * a1 + a2 * a2);
    const float h = 0.5f * n;
    const float c = cosf(h), s = sinf(h);
    const float u0 = c, u1 = s * v[0] / n, u2 = s * v[1] / n, u3 = s * v[2] / n;  // unnormalised quaternion
    const float qn = sqrtf(u0 * u0 + u1 * u1 + u2 * u2 + u3 * u3);
    const float w = u0 / qn, x = u1 / qn, y = u2 / qn, z = u3 / qn;
    // through quat2mat
    const float dw = 2 * w * (g[0] + g[4] + g[8]) + 2 * (-z * g[1] + y * g[2] + z * g[3] - x * g[5] - y * g[6] + x * g[7]);
    const float dx = 2 * x * (g[0] - g[4] - g[8]) + 2 * (y * g[1] + z * g[2] + y * g[3] - w * g[5] + z * g[6] + w * g[7]);
    const float dy = 2 * y * (-g[0] + g[4] - g[8]) + 2 * (x * g[1] + w * g[2] + x * g[3] + z * g[5] - w * g[6] + z * g[7]);
    const float dz = 2 * z * (-g[0] - g[4] + g[8]) + 2 * (-w * g[1] + x * g[2] + w * g[3] + y * g[5] + x * g[6] + y * g[7]);
    // through the normalisation q = u/|u|
    const float dot = dw * w + dx * x + dy * y + dz * z;
    const float du0 = (dw - w * dot) / qn, du1 = (dx - x * dot) / qn, du2 = (dy - y * dot) / qn, du3 = (dz - z * dot) / qn;
    // u0 = cos h, u_i = sin h * v_i / n
    const float dvec = du1 * v[0] + du2 * v[1] + du3 * v[2];
    const float dc = du0;
    const float ds = dvec / n;
    const float dh = -s * dc + c * ds;
    const float dn = 0.5f * dh - s * dvec / (n * n);
    const float sn = s / n;
    dv[0] = du1 * sn + dn * a0 / n;
    dv[1] = du2 * sn + dn * a1 / n;
    dv[2] = du3 * sn + dn * a2 / n;
}

struct PoseState {
    float pose[kPose];     // full pose: root axis-angle | hands_mean + PCA
    float R[kJ][9];        // local rotations
    float J[kJ][3];        // rest joints for this shape
    float Gr[kJ][9];       // global rotations
    float Gt[kJ][3];       // global joint positions
};

// theta (48), beta (10) -> state. comps [45][45] row k = k-th PCA component, jt [16][3], js [16][3][10].
MHE_HD void pose_fwd(const float* comps, const float* hands_mean, const float* jt, const float* js,
                     const float* theta, const float* beta, PoseState& st) {
    for (int i = 0; i < 3; ++i) st.pose[i] = theta[i];
    for (int j = 0; j < 45; ++j) {
        float acc = hands_mean[j];
        for (int k = 0; k < 45; ++k) acc = fmaf(theta[3 + k], comps[k * 45 + j], acc);
        st.pose[3 + j] = acc;
    }
    for (int k = 0; k < kJ; ++k) {
        rodrigues_fwd(&st.pose[3 * k], st.R[k]);
        for (int c = 0; c < 3; ++c) {
            float acc = jt[k * 3 + c];
            for (int b = 0; b < kShape; ++b) acc = fmaf(js[(k * 3 + c) * kShape + b], beta[b], acc);
            st.J[k][c] = acc;
        }
    }
    for (int i = 0; i < 9; ++i) st.Gr[0][i] = st.R[0][i];
    for (int c = 0; c < 3; ++c) st.Gt[0][c] = st.J[0][c];
    for (int k = 1; k < kJ; ++k) {
        const int p = parent_of(k);
        mat3_mul(st.Gr[p], st.R[k], st.Gr[k]);
        float d[3] = {st.J[k][0] - st.J[p][0], st.J[k][1] - st.J[p][1], st.J[k][2] - st.J[p][2]};
        float o[3];
        mat3_vec(st.Gr[p], d, o);
        for (int c = 0; c < 3; ++c) st.Gt[k][c] = o[c] + st.Gt[p][c];
    }
}

// skinning transform of joint k: A = [Gr | Gt - Gr J]   (manolayer.py:232-234), 12 floats: rot 9, tr 3
MHE_HD void skin_transform(const PoseState& st, int k, float* A) {
    for (int i = 0; i < 9; ++i) A[i] = st.Gr[k][i];
    float o[3];
    mat3_vec(st.Gr[k], st.J[k], o);
    for (int c = 0; c < 3; ++c) A[9 + c] = st.Gt[k][c] - o[c];
}

// Backward of pose_fwd.
//   dGt_out [16][3]: gradient on the joint positions (chain joints, centring included)
//   dA [16][12]: gradient on the skinning transforms (NULL = 0)
//   dpm [135]: gradient on the pose feature R_k - I, k = 1..15 (NULL = 0)
// -> dtheta (48), dbeta (10) are ADDED to.
MHE_HD void pose_bwd(const float* comps, const float* js, const PoseState& st,
                     const float* dGt_out, const float* dA, const float* dpm,
                     float* dtheta, float* dbeta) {
    float dGr[kJ][9], dGt[kJ][3], dJ[kJ][3], dR[kJ][9];
    for (int k = 0; k < kJ; ++k) {
        for (int i = 0; i < 9; ++i) { dGr[k][i] = 0.f; dR[k][i] = 0.f; }
        for (int c = 0; c < 3; ++c) { dGt[k][c] = dGt_out[k * 3 + c]; dJ[k][c] = 0.f; }
        if (dA) {
            const float* a = dA + k * 12;
            // A_rot = Gr, A_tr = Gt - Gr J
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) dGr[k][i * 3 + j] = a[i * 3 + j] - a[9 + i] * st.J[k][j];
            float o[3];
            mat3t_vec(st.Gr[k], a + 9, o);
            for (int c = 0; c < 3; ++c) { dGt[k][c] += a[9 + c]; dJ[k][c] -= o[c]; }
        }
        if (dpm && k >= 1)
            for (int i = 0; i < 9; ++i) dR[k][i] = dpm[(k - 1) * 9 + i];
    }
    for (int k = kJ - 1; k >= 1; --k) {
        const int p = parent_of(k);
        // Gr[k] = Gr[p] R[k]
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                float a = 0.f, b = 0.f;
                for (int l = 0; l < 3; ++l) {
                    a = fmaf(st.Gr[p][l * 3 + i], dGr[k][l * 3 + j], a);   // Gr[p]^T dGr[k]
                    b = fmaf(dGr[k][i * 3 + l], st.R[k][j * 3 + l], b);    // dGr[k] R[k]^T
                }
                dR[k][i * 3 + j] += a;
                dGr[p][i * 3 + j] += b;
            }
        // Gt[k] = Gr[p] (J[k] - J[p]) + Gt[p]
        const float d[3] = {st.J[k][0] - st.J[p][0], st.J[k][1] - st.J[p][1], st.J[k][2] - st.J[p][2]};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) dGr[p][i * 3 + j] = fmaf(dGt[k][i], d[j], dGr[p][i * 3 + j]);
        float o[3];
        mat3t_vec(st.Gr[p], dGt[k], o);
        for (int c = 0; c < 3; ++c) { dJ[k][c] += o[c]; dJ[p][c] -= o[c]; dGt[p][c] += dGt[k][c]; }
    }
    for (int i = 0; i < 9; ++i) dR[0][i] += dGr[0][i];
    for (int c = 0; c < 3; ++c) dJ[0][c] += dGt[0][c];
    for (int b = 0; b < kShape; ++b) {
        float acc = 0.f;
        for (int k = 0; k < kJ; ++k)
            for (int c = 0; c < 3; ++c) acc = fmaf(js[(k * 3 + c) * kShape + b], dJ[k][c], acc);
        dbeta[b] += acc;
    }
    float dpose[kPose];
    for (int k = 0; k < kJ; ++k) rodrigues_bwd(&st.pose[3 * k], dR[k], &dpose[3 * k]);
    for (int i = 0; i < 3; ++i) dtheta[i] += dpose[i];
    for (int k = 0; k < 45; ++k) {
        float acc = 0.f;
        for (int j = 0; j < 45; ++j) acc = fmaf(comps[k * 45 + j], dpose[3 + j], acc);
        dtheta[3 + k] += acc;
    }
}

}  // namespace mano
}  // namespace mhe
