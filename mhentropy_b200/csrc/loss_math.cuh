// Per-row reprojection + likelihood/prior mathematics (host + device) and its gradient.
// Reference: utils.py:46-66 (root/bone normalise), network.py:497-514 + ManoLayer.py:162-165 (orthographic
// projection), network.py:255-257 (Laplace, const b), network.py:155-165 (box / ball priors), network.py:646-662.
#pragma once
#include <math.h>
#include "../../include/mhentropy_b200.h"

#ifndef MHE_HD
#if defined(__CUDACC__)
#define MHE_HD __host__ __device__ __forceinline__
#else
#define MHE_HD inline
#endif
#endif

namespace mhe {
namespace loss {

constexpr int kNJ = 21;
constexpr int kZ = 61;          // th3 3 | th45 45 | bt 10 | logs 1 | t 2   (network.py:367-369)
constexpr float kLapEps = 1e-4f;  // network.py:257

MHE_HD float relu(float v) { return v > 0.f ? v : 0.f; }
MHE_HD float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

// joints [21][3] -> xyz [21][3] (normalised, root-relative); returns the bone length
MHE_HD float normalize_joints(const mhe_loss_cfg& cfg, const float* j, float* xyz) {
    const float* root = j + cfg.root_idx * 3;
    const float bx = j[cfg.norm_idx * 3 + 0] - root[0], by = j[cfg.norm_idx * 3 + 1] - root[1], bz = j[cfg.norm_idx * 3 + 2] - root[2];
    const float bone = sqrtf(bx * bx + by * by + bz * bz);
    for (int k = 0; k < kNJ; ++k)
        for (int c = 0; c < 3; ++c) xyz[k * 3 + c] = (j[k * 3 + c] - root[c]) / bone;
    return bone;
}

// row log-probability: Laplace on the visible 2D keypoints + priors on th3 / th45 / bt. uv (42) is written.
MHE_HD float row_log_p(const mhe_loss_cfg& cfg, const float* j, const float* z, const float* crop_uv, const float* vis, float* uv) {
    float xyz[kNJ * 3];
    normalize_joints(cfg, j, xyz);
    const float s = expf(z[58]);
    const float log2b = logf(2.f * cfg.laplace_b);
    float lp = 0.f;
    for (int k = 0; k < kNJ; ++k)
        for (int d = 0; d < 2; ++d) {
            const float mu = s * xyz[k * 3 + d] + z[59 + d];
            uv[k * 2 + d] = mu;
            if (vis[k] == 1.f) lp += -(relu(fabsf(crop_uv[k * 2 + d] - mu) - kLapEps) + kLapEps) / cfg.laplace_b - log2b;
        }
    {   // th3: ball of radius th3_radius around 0
        const float r = sqrtf(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]);
        const float u = relu(r / cfg.th3_radius - 1.f);
        lp -= cfg.th3_alpha * u * u;
    }
    for (int i = 0; i < 45; ++i) { const float u = relu(fabsf(z[3 + i]) / cfg.th45_box - 1.f); lp -= cfg.th45_alpha * u * u; }
    for (int i = 0; i < 10; ++i) { const float u = relu(fabsf(z[48 + i]) / cfg.bt_box - 1.f); lp -= cfg.bt_alpha * u * u; }
    return lp;
}

// g = dL/d(row_log_p) -> dj (63), dz (61); both overwritten
MHE_HD void row_log_p_bwd(const mhe_loss_cfg& cfg, const float* j, const float* z, const float* crop_uv, const float* vis, float g,
                          float* dj, float* dz) {
    float xyz[kNJ * 3];
    const float bone = normalize_joints(cfg, j, xyz);
    const float s = expf(z[58]);
    for (int i = 0; i < kZ; ++i) dz[i] = 0.f;
    float drel[kNJ * 3];
    float ds = 0.f, dbone = 0.f;
    for (int k = 0; k < kNJ; ++k) {
        for (int d = 0; d < 2; ++d) {
            float dmu = 0.f;
            if (vis[k] == 1.f) {
                const float diff = crop_uv[k * 2 + d] - (s * xyz[k * 3 + d] + z[59 + d]);
                if (fabsf(diff) - kLapEps > 0.f) dmu = g * sgn(diff) / cfg.laplace_b;
            }
            dz[59 + d] += dmu;
            ds += dmu * xyz[k * 3 + d];
            const float dxyz = dmu * s;
            drel[k * 3 + d] = dxyz / bone;
            dbone -= dxyz * xyz[k * 3 + d] / bone;   // rel / bone^2 = xyz / bone
        }
        drel[k * 3 + 2] = 0.f;
    }
    dz[58] = ds * s;
    for (int c = 0; c < 3; ++c) drel[cfg.norm_idx * 3 + c] += dbone * xyz[cfg.norm_idx * 3 + c];  // d|v|/dv = v/|v| = xyz[norm]
    float droot[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < kNJ; ++k)
        for (int c = 0; c < 3; ++c) { dj[k * 3 + c] = drel[k * 3 + c]; droot[c] += drel[k * 3 + c]; }
    for (int c = 0; c < 3; ++c) dj[cfg.root_idx * 3 + c] -= droot[c];
    {
        const float r = sqrtf(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]);
        const float u = r / cfg.th3_radius - 1.f;
        if (u > 0.f) for (int i = 0; i < 3; ++i) dz[i] = -g * cfg.th3_alpha * 2.f * u / cfg.th3_radius * z[i] / r;
    }
    for (int i = 0; i < 45; ++i) { const float u = fabsf(z[3 + i]) / cfg.th45_box - 1.f; if (u > 0.f) dz[3 + i] = -g * cfg.th45_alpha * 2.f * u * sgn(z[3 + i]) / cfg.th45_box; }
    for (int i = 0; i < 10; ++i) { const float u = fabsf(z[48 + i]) / cfg.bt_box - 1.f; if (u > 0.f) dz[48 + i] = -g * cfg.bt_alpha * 2.f * u * sgn(z[48 + i]) / cfg.bt_box; }
}

}  // namespace loss
}  // namespace mhe
