// Multi-hypothesis evaluation metrics and top-k hypothesis selection.
// Reference: hand/criteria.py:91-168 (MHEntLoss metrics, aligned = False), hand/utils.py:21-30 (meanEuclideanLoss),
// hand/network.py:866-871 (top-k of log q).  One block per image reduces over its N hypotheses, so the (N, B, .) outputs of
// MHEnt.sample are read exactly once.
#include "common.cuh"

namespace mhe {

constexpr int kMK = 21;          // joints
constexpr int kMetricRows = 14;  // 2 spaces x (sample, sample_std, vis, vis_std, vis_mean, invis, invis_std)
constexpr int kMWarps = 16;      // warps per image: the N hypotheses are strided over them

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// metrics[row][b] WITHOUT the B / num_valid factor of _group_stats (criteria.py:127-130); valid[g][b] = image b has a joint in group g
__global__ void __launch_bounds__(kMWarps * 32) hyp_metrics_kernel(const float* __restrict__ xyz, const float* __restrict__ uv,
                                                                     const float* __restrict__ pose3d, const float* __restrict__ scale,
                                                                     const float* __restrict__ crop_uv, const float* __restrict__ vis, int N, int B,
                                                                     int root_idx, float image_size, float* __restrict__ metrics,
                                                                     int* __restrict__ valid) {
    __shared__ float s_best[kMWarps][6];          // per warp: best (min / max) group means, [space*3 + group]
    __shared__ float s_sum[kMWarps][kMK][7];      // per warp, per joint: sum_n of euc3, euc2, and of the 5 coordinates
    __shared__ float s_sq[kMWarps][kMK][5];       // per warp, per joint: sum_n (coord - mean)^2
    __shared__ float s_mean[kMK][5];
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool jt = lane < kMK;
    const int k = jt ? lane : 0;
    const float sc = scale[b];
    const float g3[3] = {pose3d[(size_t)b * 63 + k * 3], pose3d[(size_t)b * 63 + k * 3 + 1], pose3d[(size_t)b * 63 + k * 3 + 2]};
    const float g2[2] = {(crop_uv[(size_t)b * 42 + k * 2] + 1.f) / 2.f * image_size, (crop_uv[(size_t)b * 42 + k * 2 + 1] + 1.f) / 2.f * image_size};
    const float v = vis[(size_t)b * kMK + k];
    // group weights: all joints / visible / occluded, root excluded from the latter two (criteria.py:107-114)
    const float w[3] = {jt ? 1.f : 0.f, (jt && v == 1.f && k != root_idx) ? 1.f : 0.f, (jt && v != 1.f && k != root_idx) ? 1.f : 0.f};
    float num[3];
#pragma unroll
    for (int g = 0; g < 3; ++g) num[g] = warp_sum(w[g]);

    float best[6] = {INFINITY, INFINITY, INFINITY, INFINITY, -INFINITY, INFINITY};   // 2d-vis keeps the WORST hypothesis (criteria.py:148-150)
    float sum[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int n = warp; n < N; n += kMWarps) {
        const float* px = xyz + ((size_t)n * B + b) * 63 + k * 3;
        const float* pu = uv + ((size_t)n * B + b) * 42 + k * 2;
        const float x0 = px[0], x1 = px[1], x2 = px[2], u0 = pu[0], u1 = pu[1];
        const float d0 = x0 - g3[0], d1 = x1 - g3[1], d2 = x2 - g3[2];
        const float e3 = sqrtf(d0 * d0 + d1 * d1 + d2 * d2) * sc;           // utils.py:25-26
        const float f0 = u0 - g2[0], f1 = u1 - g2[1];
        const float e2 = sqrtf(f0 * f0 + f1 * f1);                          // criteria.py:105
        sum[0] += e3; sum[1] += e2;
        sum[2] += x0 * sc; sum[3] += x1 * sc; sum[4] += x2 * sc; sum[5] += u0; sum[6] += u1;
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            const float m3 = warp_sum(e3 * w[g]) / (num[g] + 1e-16f), m2 = warp_sum(e2 * w[g]) / (num[g] + 1e-16f);
            best[g] = fminf(best[g], m3);
            best[3 + g] = (g == 1) ? fmaxf(best[3 + g], m2) : fminf(best[3 + g], m2);
        }
    }
    if (lane == 0)
        for (int i = 0; i < 6; ++i) s_best[warp][i] = best[i];
    if (jt)
        for (int i = 0; i < 7; ++i) s_sum[warp][k][i] = sum[i];
    __syncthreads();
    // per-joint means over the hypotheses (coordinates), then the centred second pass
    float tot[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        float a = 0.f;
#pragma unroll
        for (int ww = 0; ww < kMWarps; ++ww) a += s_sum[ww][k][i];
        tot[i] = a;
    }
    if (warp == 0 && jt)
        for (int i = 0; i < 5; ++i) s_mean[k][i] = tot[2 + i] / (float)N;
    __syncthreads();
    float sq[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (N > 1)
        for (int n = warp; n < N; n += kMWarps) {
            const float* px = xyz + ((size_t)n * B + b) * 63 + k * 3;
            const float* pu = uv + ((size_t)n * B + b) * 42 + k * 2;
            const float c[5] = {px[0] * sc - s_mean[k][0], px[1] * sc - s_mean[k][1], px[2] * sc - s_mean[k][2], pu[0] - s_mean[k][3], pu[1] - s_mean[k][4]};
#pragma unroll
            for (int i = 0; i < 5; ++i) sq[i] += c[i] * c[i];
        }
    if (jt)
        for (int i = 0; i < 5; ++i) s_sq[warp][k][i] = sq[i];
    __syncthreads();
    if (warp != 0) return;
    float sd[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        float a = 0.f;
#pragma unroll
        for (int ww = 0; ww < kMWarps; ++ww) a += s_sq[ww][k][i];
        sd[i] = N > 1 ? sqrtf(a / (float)(N - 1)) : 0.f;   // torch.std: unbiased
    }
    // spread of the hypotheses per joint: geometric mean of the per-axis std times sqrt(D) (criteria.py:153-160)
    const float sp3 = N > 1 ? cbrtf(sd[0] * sd[1] * sd[2]) * sqrtf(3.f) : 0.f;
    const float sp2 = N > 1 ? sqrtf(sd[3] * sd[4]) * sqrtf(2.f) : 0.f;
    const float mean3 = tot[0] / (float)N, mean2 = tot[1] / (float)N;       // criteria.py:163
    float bst[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        float x = s_best[0][i];
        for (int ww = 1; ww < kMWarps; ++ww) x = (i == 4) ? fmaxf(x, s_best[ww][i]) : fminf(x, s_best[ww][i]);
        bst[i] = x;
    }
    float out[kMetricRows];
#pragma unroll
    for (int sp = 0; sp < 2; ++sp) {
        const float spread = sp == 0 ? sp3 : sp2, mean = sp == 0 ? mean3 : mean2;
        float st[3];
#pragma unroll
        for (int g = 0; g < 3; ++g) st[g] = warp_sum(spread * w[g]) / (num[g] + 1e-16f);
        const float mv = warp_sum(mean * w[1]) / (num[1] + 1e-16f);
        float* o = out + sp * 7;
        o[0] = N > 0 ? bst[sp * 3 + 0] : 0.f; o[1] = st[0];
        o[2] = N > 0 ? bst[sp * 3 + 1] : 0.f; o[3] = st[1]; o[4] = mv;
        o[5] = N > 0 ? bst[sp * 3 + 2] : 0.f; o[6] = st[2];
    }
    if (lane == 0) {
        for (int i = 0; i < kMetricRows; ++i) metrics[(size_t)i * B + b] = out[i];
        for (int g = 0; g < 3; ++g) valid[g * B + b] = num[g] > 0.f ? 1 : 0;
    }
}

// the batch-level factor of _group_stats: B / (#images with a joint in the group), or 0 when there is none (criteria.py:127-130)
__global__ void __launch_bounds__(256) hyp_metrics_scale_kernel(float* __restrict__ metrics, const int* __restrict__ valid, int B) {
    __shared__ int s_cnt[3];
    if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    int c[3] = {0, 0, 0};
    for (int b = threadIdx.x; b < B; b += blockDim.x)
        for (int g = 0; g < 3; ++g) c[g] += valid[g * B + b];
    for (int g = 0; g < 3; ++g)
        if (c[g]) atomicAdd(&s_cnt[g], c[g]);
    __syncthreads();
    const int group_of_row[7] = {0, 0, 1, 1, 1, 2, 2};
    for (int i = threadIdx.x; i < kMetricRows * B; i += blockDim.x) {
        const int nv = s_cnt[group_of_row[(i / B) % 7]];
        metrics[i] = nv ? metrics[i] * (float)B / ((float)nv + 1e-16f) : metrics[i] * 0.f;
    }
}

// idx[j][b] = index of the j-th most likely hypothesis of image b (descending log q; ties: lower index first)
__global__ void topk_hypotheses_kernel(const float* __restrict__ log_q, int N, int B, int k, long long* __restrict__ idx) {
    extern __shared__ float s_lq[];
    const int b = blockIdx.x;
    for (int n = threadIdx.x; n < N; n += blockDim.x) s_lq[n] = log_q[(size_t)n * B + b];
    __syncthreads();
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float v = s_lq[n];
        int rank = 0;
        for (int m = 0; m < N; ++m) rank += (s_lq[m] > v) || (s_lq[m] == v && m < n);
        if (rank < k) idx[(size_t)rank * B + b] = n;
    }
}

}  // namespace mhe

using namespace mhe;

extern "C" {

size_t mhe_hypothesis_metrics_workspace_bytes(int B) { return B < 0 ? 0 : (size_t)3 * B * sizeof(int); }

int mhe_hypothesis_metrics(const float* xyz, const float* uv, const float* pose3d, const float* scale, const float* crop_uv, const float* vis,
                           int N, int B, int root_idx, float image_size, float* metrics, void* workspace, size_t workspace_bytes,
                           void* stream_) {
    if (B == 0) return MHE_OK;
    MHE_REQUIRE(xyz && uv && pose3d && scale && crop_uv && vis && metrics && workspace, "hypothesis_metrics: null pointer");
    MHE_REQUIRE(N >= 1 && B > 0 && root_idx >= 0 && root_idx < kMK, "hypothesis_metrics: bad sizes");
    if (workspace_bytes < mhe_hypothesis_metrics_workspace_bytes(B)) { set_error("hypothesis_metrics: workspace too small"); return MHE_ERR_WORKSPACE; }
    cudaStream_t stream = (cudaStream_t)stream_;
    hyp_metrics_kernel<<<B, kMWarps * 32, 0, stream>>>(xyz, uv, pose3d, scale, crop_uv, vis, N, B, root_idx, image_size, metrics, (int*)workspace);
    MHE_TRY(check_launch("hypothesis metrics"));
    hyp_metrics_scale_kernel<<<1, 256, 0, stream>>>(metrics, (const int*)workspace, B);
    return check_launch("hypothesis metrics scale");
}

int mhe_topk_hypotheses(const float* log_q, int N, int B, int k, long long* idx, void* stream) {
    if (B == 0 || k == 0) return MHE_OK;
    MHE_REQUIRE(log_q && idx && N >= 1 && B > 0 && k >= 1 && k <= N && N <= 12288, "topk_hypotheses: bad args");
    topk_hypotheses_kernel<<<B, 128, (size_t)N * sizeof(float), (cudaStream_t)stream>>>(log_q, N, B, k, idx);
    return check_launch("topk hypotheses");
}

}  // extern "C"
