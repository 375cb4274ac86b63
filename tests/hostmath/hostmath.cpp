// TEST INFRASTRUCTURE: compiles the host+device math headers of the CUDA kernels with g++ so the
// per-row formulas (Rodrigues, kinematic chain, loss terms and their hand-derived gradients) can be
// checked against the oracle's autograd on a machine without a GPU.  Never part of the product path.
#include "../../mhentropy_b200/csrc/mano_math.cuh"
#include "../../mhentropy_b200/csrc/loss_math.cuh"

using namespace mhe;

extern "C" {

void hm_rodrigues_fwd(const float* v, float* R) { mano::rodrigues_fwd(v, R); }
void hm_rodrigues_bwd(const float* v, const float* g, float* dv) { mano::rodrigues_bwd(v, g, dv); }

// theta(48), beta(10) -> A[16*12], Gt[16*3], pm[135]
void hm_pose_fwd(const float* comps, const float* hands_mean, const float* jt, const float* js,
                 const float* theta, const float* beta, float* A, float* Gt, float* pm) {
    mano::PoseState st;
    mano::pose_fwd(comps, hands_mean, jt, js, theta, beta, st);
    for (int k = 0; k < 16; ++k) {
        mano::skin_transform(st, k, A + k * 12);
        for (int c = 0; c < 3; ++c) Gt[k * 3 + c] = st.Gt[k][c];
        if (k >= 1) for (int i = 0; i < 9; ++i) pm[(k - 1) * 9 + i] = st.R[k][i] - ((i % 4 == 0) ? 1.f : 0.f);
    }
}

void hm_pose_bwd(const float* comps, const float* hands_mean, const float* jt, const float* js,
                 const float* theta, const float* beta, const float* dGt, const float* dA, const float* dpm,
                 float* dtheta, float* dbeta) {
    mano::PoseState st;
    mano::pose_fwd(comps, hands_mean, jt, js, theta, beta, st);
    for (int i = 0; i < 48; ++i) dtheta[i] = 0.f;
    for (int i = 0; i < 10; ++i) dbeta[i] = 0.f;
    mano::pose_bwd(comps, js, st, dGt, dA, dpm, dtheta, dbeta);
}

float hm_row_log_p(const mhe_loss_cfg* cfg, const float* j, const float* z, const float* crop_uv, const float* vis, float* uv) {
    return loss::row_log_p(*cfg, j, z, crop_uv, vis, uv);
}
void hm_row_log_p_bwd(const mhe_loss_cfg* cfg, const float* j, const float* z, const float* crop_uv, const float* vis, float g,
                      float* dj, float* dz) {
    loss::row_log_p_bwd(*cfg, j, z, crop_uv, vis, g, dj, dz);
}
}
