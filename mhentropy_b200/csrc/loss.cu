// z assembly, visible-keypoint reprojection, Laplace / prior / entropy reductions (reference network.py:455-831).
#include "loss_rows.cuh"

namespace mhe {
using namespace loss;

__global__ void combine_z_fwd_kernel(const float* __restrict__ x, const float* __restrict__ zd, int R, int B, float* __restrict__ z) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)R * kZ) return;
    const int r = (int)(i / kZ), c = (int)(i % kZ), b = r % B;
    float v;
    if (c < 3) v = zd[b * 16 + c];                 // th3
    else if (c < 48) v = x[(long)r * 45 + c - 3];  // th45 <- flow
    else v = zd[b * 16 + c - 45];                  // bt | logs | t  (zd cols 3..15)
    z[i] = v;
}

__global__ void combine_z_bwd_kernel(const float* __restrict__ dz, int R, int B, float* __restrict__ dx, float* __restrict__ dzd) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (dx && i < (long)R * 45) {
        const int r = (int)(i / 45), c = (int)(i % 45);
        dx[i] = dz[(long)r * kZ + 3 + c];
    }
    if (i < (long)B * 16) {
        const int b = (int)(i / 16), c = (int)(i % 16);
        const int zc = c < 3 ? c : c + 45;
        float acc = 0.f;
        for (int r = b; r < R; r += B) acc += dz[(long)r * kZ + zc];
        dzd[i] = acc;
    }
}

// Warp per row (loss_rows.cuh).
__global__ void __launch_bounds__(128) reproj_rows_fwd_kernel(mhe_loss_cfg cfg, const float* __restrict__ joints, const float* __restrict__ z,
                                       const float* __restrict__ crop_uv, const float* __restrict__ vis, int R, int B,
                                       float* __restrict__ uv, float* __restrict__ row_lp) {
    const int r = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= R) return;
    const int b = r % B;
    const float* zz = z + (long)r * kZ;
    const RowGeom g = row_geom(cfg, joints + (long)r * 63, zz, lane);
    const float lp = reproj_row_fwd(cfg, g, zz, crop_uv + b * 42, vis + b * kNJ, lane, uv ? uv + (long)r * 42 : nullptr);
    if (lane == 0) row_lp[r] = lp;
}

__global__ void image_reduce_kernel(const float* __restrict__ row_lp, const float* __restrict__ log_q, int R, int B,
                                    float* __restrict__ log_p, float* __restrict__ h, float* __restrict__ qlp) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int N = R / B;
    float a = 0.f, q = 0.f;
    for (int r = b; r < R; r += B) { a += row_lp[r]; q += log_q[r]; }
    const float hh = -q / N, ql = a / N;
    if (h) h[b] = hh;
    if (qlp) qlp[b] = ql;
    if (log_p) log_p[b] = hh + ql;
}

// one block: loss = -mean_b log_p[b], pairwise within the block (deterministic)
__global__ void __launch_bounds__(256) loss_reduce_kernel(const float* __restrict__ log_p, int B, float* __restrict__ loss) {
    __shared__ float s[256];
    float acc = 0.f;
    for (int b = threadIdx.x; b < B; b += 256) acc += log_p[b];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) loss[0] = -s[0] / B;
}

__global__ void __launch_bounds__(128) reproj_rows_bwd_kernel(mhe_loss_cfg cfg, const float* __restrict__ joints, const float* __restrict__ z,
                                       const float* __restrict__ crop_uv, const float* __restrict__ vis, int R, int B,
                                       const float* __restrict__ dlog_p, const float* __restrict__ dloss,
                                       float* __restrict__ djoints, float* __restrict__ dz, float* __restrict__ dlog_q) {
    const int r = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= R) return;
    const int b = r % B;
    const int N = R / B;
    const float gb = dlog_p ? dlog_p[b] : -(dloss ? dloss[0] : 1.f) / B;   // loss = -mean_b log_p
    const float gr = gb / N;
    const float* zz = z + (long)r * kZ;
    const RowGeom g = row_geom(cfg, joints + (long)r * 63, zz, lane);
    reproj_row_bwd(cfg, g, zz, crop_uv + b * 42, vis + b * kNJ, gr, lane, djoints + (long)r * 63, dz + (long)r * kZ);
    if (lane == 0 && dlog_q) dlog_q[r] = -gr;
}

// MHEnt.sample epilogue: xyz, normalised verts and (pixel) uv
__global__ void normalize_project_kernel(mhe_loss_cfg cfg, const float* __restrict__ joints, const float* __restrict__ verts,
                                         const float* __restrict__ z, int ld_z, int R, int inv_norm, int image_size,
                                         float* __restrict__ xyz, float* __restrict__ verts_n, float* __restrict__ uv) {
    const int r = blockIdx.x;
    __shared__ float s_root[3];
    __shared__ float s_bone;
    const float* j = joints + (long)r * 63;
    if (threadIdx.x == 0) {
        const float bx = j[cfg.norm_idx * 3] - j[cfg.root_idx * 3], by = j[cfg.norm_idx * 3 + 1] - j[cfg.root_idx * 3 + 1], bz = j[cfg.norm_idx * 3 + 2] - j[cfg.root_idx * 3 + 2];
        s_bone = sqrtf(bx * bx + by * by + bz * bz);
        s_root[0] = j[cfg.root_idx * 3]; s_root[1] = j[cfg.root_idx * 3 + 1]; s_root[2] = j[cfg.root_idx * 3 + 2];
    }
    __syncthreads();
    const float bone = s_bone;
    const float s = expf(z[(long)r * ld_z + 58]);
    for (int i = threadIdx.x; i < 63; i += blockDim.x) {
        const int c = i % 3, k = i / 3;
        const float v = (j[i] - s_root[c]) / bone;
        if (xyz) xyz[(long)r * 63 + i] = v;
        if (uv && c < 2) {
            float u = s * v + z[(long)r * ld_z + 59 + c];
            if (inv_norm) u = (u + 1.f) / 2.f * image_size;
            uv[(long)r * 42 + k * 2 + c] = u;
        }
    }
    if (verts && verts_n)
        for (int i = threadIdx.x; i < MHE_MANO_VERTS * 3; i += blockDim.x)
            verts_n[(long)r * MHE_MANO_VERTS * 3 + i] = (verts[(long)r * MHE_MANO_VERTS * 3 + i] - s_root[i % 3]) / bone;
}

}  // namespace mhe

using namespace mhe;

extern "C" {

int mhe_combine_z_fwd(const float* x_flow, const float* z_det, int R, int B, float* z, void* stream) {
    if (R == 0) return MHE_OK;
    MHE_REQUIRE(x_flow && z_det && z && R >= 0 && B > 0 && R % B == 0, "combine_z_fwd: bad args");
    if (R == 0) return MHE_OK;
    combine_z_fwd_kernel<<<cdiv((int)((long)R * kZ), 256), 256, 0, (cudaStream_t)stream>>>(x_flow, z_det, R, B, z);
    return check_launch("combine z fwd");
}

int mhe_combine_z_bwd(const float* dz, int R, int B, float* dx_flow, float* dz_det, void* stream) {
    if (R == 0) return MHE_OK;
    MHE_REQUIRE(dz && dz_det && R >= 0 && B > 0 && R % B == 0, "combine_z_bwd: bad args");
    if (R == 0) return MHE_OK;
    const long n = (dx_flow && (long)R * 45 > (long)B * 16) ? (long)R * 45 : (long)B * 16;
    combine_z_bwd_kernel<<<cdiv((int)n, 256), 256, 0, (cudaStream_t)stream>>>(dz, R, B, dx_flow, dz_det);
    return check_launch("combine z bwd");
}

int mhe_reproj_loss_fwd(const mhe_loss_cfg* cfg, const float* joints, const float* z, const float* crop_uv,
                        const float* vis, const float* log_q, int R, int B,
                        float* uv, float* row_log_p, float* log_p, float* h, float* q_log_p, float* loss, void* stream_) {
    MHE_REQUIRE(cfg && joints && z && crop_uv && vis && row_log_p, "reproj_loss_fwd: null pointer");
    MHE_REQUIRE(R >= 0 && B > 0 && R % B == 0, "reproj_loss_fwd: R must be a multiple of B");
    MHE_REQUIRE(!(log_p || h || q_log_p || loss) || log_q, "reproj_loss_fwd: image-level outputs need log_q");
    MHE_REQUIRE(!loss || log_p, "reproj_loss_fwd: loss needs log_p");
    if (R == 0) return MHE_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    reproj_rows_fwd_kernel<<<cdiv(R, 4), 128, 0, stream>>>(*cfg, joints, z, crop_uv, vis, R, B, uv, row_log_p);
    MHE_TRY(check_launch("reproj rows fwd"));
    if (log_p || h || q_log_p) {
        image_reduce_kernel<<<cdiv(B, 64), 64, 0, stream>>>(row_log_p, log_q, R, B, log_p, h, q_log_p);
        MHE_TRY(check_launch("image reduce"));
    }
    if (loss) {
        loss_reduce_kernel<<<1, 256, 0, stream>>>(log_p, B, loss);
        MHE_TRY(check_launch("loss reduce"));
    }
    return MHE_OK;
}

int mhe_image_loss_reduce(const float* row_log_p, const float* log_q, int R, int B, float* log_p, float* h, float* q_log_p, float* loss,
                          void* stream_) {
    MHE_REQUIRE(row_log_p && log_q && R >= 0 && B > 0 && R % B == 0, "image_loss_reduce: bad args");
    MHE_REQUIRE(!loss || log_p, "image_loss_reduce: loss needs log_p");
    if (R == 0) return MHE_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (log_p || h || q_log_p) {
        image_reduce_kernel<<<cdiv(B, 64), 64, 0, stream>>>(row_log_p, log_q, R, B, log_p, h, q_log_p);
        MHE_TRY(check_launch("image reduce"));
    }
    if (loss) {
        loss_reduce_kernel<<<1, 256, 0, stream>>>(log_p, B, loss);
        MHE_TRY(check_launch("loss reduce"));
    }
    return MHE_OK;
}

int mhe_reproj_loss_bwd(const mhe_loss_cfg* cfg, const float* joints, const float* z, const float* crop_uv,
                        const float* vis, int R, int B, const float* dlog_p, const float* dloss,
                        float* djoints, float* dz, float* dlog_q, void* stream) {
    MHE_REQUIRE(cfg && joints && z && crop_uv && vis && djoints && dz, "reproj_loss_bwd: null pointer");
    MHE_REQUIRE(R >= 0 && B > 0 && R % B == 0, "reproj_loss_bwd: R must be a multiple of B");
    if (R == 0) return MHE_OK;
    reproj_rows_bwd_kernel<<<cdiv(R, 4), 128, 0, (cudaStream_t)stream>>>(*cfg, joints, z, crop_uv, vis, R, B, dlog_p, dloss, djoints, dz, dlog_q);
    return check_launch("reproj rows bwd");
}

int mhe_normalize_project(const mhe_loss_cfg* cfg, const float* joints, const float* verts, const float* z,
                          int ld_z, int R, int inv_norm, int image_size,
                          float* xyz, float* verts_n, float* uv, void* stream) {
    if (R == 0) return MHE_OK;
    MHE_REQUIRE(cfg && joints && z && R >= 0 && ld_z >= 61, "normalize_project: bad args");
    if (R == 0) return MHE_OK;
    normalize_project_kernel<<<R, 128, 0, (cudaStream_t)stream>>>(*cfg, joints, verts, z, ld_z, R, inv_norm, image_size, xyz, verts_n, uv);
    return check_launch("normalize project");
}

}  // extern "C"
