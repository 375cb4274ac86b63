// Warp-per-row reprojection / likelihood / prior terms and their gradient (device): lane k < 21 owns joint k (root / bone
// broadcast by shuffles), the 58 prior terms are spread over the lanes, sums are warp reductions.  Same formulas as
// loss_math.cuh (host-checked there).  Shared by loss.cu (the stand-alone kernels) and mano.cu (the fused per-row kernel).
// Reference: utils.py:46-66, network.py:497-514, :233-258, :155-165, :612-667.
#pragma once
#include "common.cuh"
#include "loss_math.cuh"

namespace mhe {
namespace loss {

struct RowGeom { float x0, x1, x2, bone, s, mu0, mu1; };

// j: this row's 21 joints (global or shared memory), z: this row's 61 latents
__device__ __forceinline__ RowGeom row_geom(const mhe_loss_cfg& cfg, const float* __restrict__ j, const float* __restrict__ z, int lane) {
    const bool on = lane < kNJ;
    const float jx = on ? j[lane * 3] : 0.f, jy = on ? j[lane * 3 + 1] : 0.f, jz = on ? j[lane * 3 + 2] : 0.f;
    const float rx = __shfl_sync(0xffffffffu, jx, cfg.root_idx), ry = __shfl_sync(0xffffffffu, jy, cfg.root_idx), rz = __shfl_sync(0xffffffffu, jz, cfg.root_idx);
    const float bx = __shfl_sync(0xffffffffu, jx, cfg.norm_idx) - rx, by = __shfl_sync(0xffffffffu, jy, cfg.norm_idx) - ry, bz = __shfl_sync(0xffffffffu, jz, cfg.norm_idx) - rz;
    RowGeom g;
    g.bone = sqrtf(bx * bx + by * by + bz * bz);
    g.x0 = (jx - rx) / g.bone; g.x1 = (jy - ry) / g.bone; g.x2 = (jz - rz) / g.bone;
    g.s = expf(z[58]);
    g.mu0 = g.s * g.x0 + z[59]; g.mu1 = g.s * g.x1 + z[60];
    return g;
}

// row log-probability (Laplace on the visible keypoints + th3 / th45 / bt priors), warp-summed: every lane returns it.
// cu / vs: crop_uv (42) and vis (21) of the row's image; uv (42) receives the projection when not NULL.
__device__ __forceinline__ float reproj_row_fwd(const mhe_loss_cfg& cfg, const RowGeom& g, const float* __restrict__ zz, const float* __restrict__ cu,
                                                const float* __restrict__ vs, int lane, float* __restrict__ uv) {
    float lp = 0.f;
    if (lane < kNJ) {
        if (uv) { uv[lane * 2] = g.mu0; uv[lane * 2 + 1] = g.mu1; }
        if (vs[lane] == 1.f) {
            const float log2b = logf(2.f * cfg.laplace_b);
            lp = -(relu(fabsf(cu[lane * 2] - g.mu0) - kLapEps) + kLapEps) / cfg.laplace_b - log2b
                 - (relu(fabsf(cu[lane * 2 + 1] - g.mu1) - kLapEps) + kLapEps) / cfg.laplace_b - log2b;
        }
    }
    const float t3 = lane < 3 ? zz[lane] : 0.f;
    const float r3 = sqrtf(warp_sum(t3 * t3));
    if (lane == 0) { const float u = relu(r3 / cfg.th3_radius - 1.f); lp -= cfg.th3_alpha * u * u; }
    for (int i = lane; i < 55; i += 32) {
        const float box = i < 45 ? cfg.th45_box : cfg.bt_box, alpha = i < 45 ? cfg.th45_alpha : cfg.bt_alpha;
        const float u = relu(fabsf(zz[3 + i]) / box - 1.f);
        lp -= alpha * u * u;
    }
    return warp_sum(lp);
}

// gradient of gr * row_log_p: dj (63; joint gradients), dzr (61; every element written)
__device__ __forceinline__ void reproj_row_bwd(const mhe_loss_cfg& cfg, const RowGeom& g, const float* __restrict__ zz, const float* __restrict__ cu,
                                               const float* __restrict__ vs, float gr, int lane, float* __restrict__ dj, float* __restrict__ dzr) {
    float dmu0 = 0.f, dmu1 = 0.f;
    if (lane < kNJ && vs[lane] == 1.f) {
        const float d0 = cu[lane * 2] - g.mu0, d1 = cu[lane * 2 + 1] - g.mu1;
        if (fabsf(d0) - kLapEps > 0.f) dmu0 = gr * sgn(d0) / cfg.laplace_b;
        if (fabsf(d1) - kLapEps > 0.f) dmu1 = gr * sgn(d1) / cfg.laplace_b;
    }
    const float dt0 = warp_sum(dmu0), dt1 = warp_sum(dmu1), ds = warp_sum(dmu0 * g.x0 + dmu1 * g.x1);
    const float dx0 = dmu0 * g.s, dx1 = dmu1 * g.s;
    const float dbone = -warp_sum(dx0 * g.x0 + dx1 * g.x1) / g.bone;
    float dr0 = dx0 / g.bone, dr1 = dx1 / g.bone, dr2 = 0.f;
    if (lane == cfg.norm_idx) { dr0 += dbone * g.x0; dr1 += dbone * g.x1; dr2 += dbone * g.x2; }
    const float s0 = warp_sum(dr0), s1 = warp_sum(dr1), s2 = warp_sum(dr2);
    if (lane == cfg.root_idx) { dr0 -= s0; dr1 -= s1; dr2 -= s2; }
    if (lane < kNJ) { dj[lane * 3] = dr0; dj[lane * 3 + 1] = dr1; dj[lane * 3 + 2] = dr2; }
    const float t3 = lane < 3 ? zz[lane] : 0.f;
    const float r3 = sqrtf(warp_sum(t3 * t3));
    if (lane < 3) {
        const float u = r3 / cfg.th3_radius - 1.f;
        dzr[lane] = u > 0.f ? -gr * cfg.th3_alpha * 2.f * u / cfg.th3_radius * t3 / r3 : 0.f;
    }
    for (int i = lane; i < 55; i += 32) {
        const float box = i < 45 ? cfg.th45_box : cfg.bt_box, alpha = i < 45 ? cfg.th45_alpha : cfg.bt_alpha;
        const float v = zz[3 + i];
        const float u = fabsf(v) / box - 1.f;
        dzr[3 + i] = u > 0.f ? -gr * alpha * 2.f * u * sgn(v) / box : 0.f;
    }
    if (lane == 0) { dzr[58] = ds * g.s; dzr[59] = dt0; dzr[60] = dt1; }
}

}  // namespace loss
}  // namespace mhe
