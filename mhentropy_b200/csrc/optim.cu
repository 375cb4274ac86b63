// Adam on the flat parameter buffer with the split weight planes refreshed in the same pass.
// Reference: hand/CrossModalHand.py:201 (torch.optim.Adam, default betas / eps), :462-470 (clip_grad_norm_ over the whole encoder, then
// optimizer.step()).  The flow's 20 M parameters live in ONE flat fp32 buffer whose gradient the backward fills in the same layout, so
// the update is one elementwise kernel; for the two dense weight families (l.1.weight and c.j.weight: 98 % of the bytes) it also writes
// the half / bfloat16 split planes the tensor-core GEMMs read, straight from the registers, instead of a separate re-pack pass over the
// weights at the start of the next step.  The thin, padded families (l.0 / l.2 weights) are re-packed by mhe_flow_pack_weights pieces.
// The global gradient norm of the reference's clipping spans modules outside this library (the CNN backbone): the caller combines
// mhe_flow_grad_sqnorm() with its own terms and passes the resulting scale.
#include "flow_tc.cuh"

namespace mhe {
namespace tcflow {
using namespace tc;

struct AdamArgs {
    float* p; const float* g; float* m; float* v;
    uint16_t *w1h, *w1b, *cwh, *cwb;            // plane bases (NULL: no planes)
    size_t n, blk, oW1, hh, cw_base, cw_stride, hc, cb_base;
    int nblk, ncw;
    float omb1, b2, omb2, eps, step_size, bc2s, gscale;   // 1 - b1, b2, 1 - b2, eps, lr / (1 - b1^t), sqrt(1 - b2^t): formed in double on the host
};

__device__ __forceinline__ void put4(uint16_t* hi, size_t plane, const float* x, bool f16) {
    uint16_t h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (f16) { h[k] = to16<true>(x[k]); l[k] = to16<true>(x[k] - from16<true>(h[k])); }
        else { h[k] = to16<false>(x[k]); l[k] = to16<false>(x[k] - from16<false>(h[k])); }
    }
    *reinterpret_cast<uint2*>(hi) = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
    *reinterpret_cast<uint2*>(hi + plane) = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
}

// 4 consecutive parameters per thread (every segment starts on a 64-float boundary, so a group never straddles two tensors)
__global__ void __launch_bounds__(256) adam_planes_kernel(AdamArgs a) {
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= a.n) return;
    const float4 g4 = *reinterpret_cast<const float4*>(a.g + i);
    float4 p4 = *reinterpret_cast<float4*>(a.p + i), m4 = *reinterpret_cast<float4*>(a.m + i), v4 = *reinterpret_cast<float4*>(a.v + i);
    float p[4] = {p4.x, p4.y, p4.z, p4.w}, m[4] = {m4.x, m4.y, m4.z, m4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w};
    const float g[4] = {g4.x * a.gscale, g4.y * a.gscale, g4.z * a.gscale, g4.w * a.gscale};
#pragma unroll
    for (int k = 0; k < 4; ++k) {   // torch.optim.Adam (no weight decay, no amsgrad), same operation order
        m[k] = m[k] + a.omb1 * (g[k] - m[k]);                             // exp_avg.lerp_(grad, 1 - beta1)
        v[k] = a.b2 * v[k] + a.omb2 * (g[k] * g[k]);                      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        const float denom = sqrtf(v[k]) / a.bc2s + a.eps;
        p[k] = p[k] - a.step_size * (m[k] / denom);
    }
    *reinterpret_cast<float4*>(a.p + i) = make_float4(p[0], p[1], p[2], p[3]);
    *reinterpret_cast<float4*>(a.m + i) = make_float4(m[0], m[1], m[2], m[3]);
    *reinterpret_cast<float4*>(a.v + i) = make_float4(v[0], v[1], v[2], v[3]);
    if (!a.w1h) return;
    if (i < a.cw_base) {                                                  // coupling blocks: l.1.weight [H][H] of block b
        const size_t b = i / a.blk, o = i - b * a.blk;
        if (o >= a.oW1 && o < a.oW1 + a.hh) {
            const size_t e = o - a.oW1, base = b * 2 * a.hh + e;         // planes [block][2][H][H]
            put4(a.w1h + base, a.hh, p, true);
            put4(a.w1b + base, a.hh, p, false);
        }
    } else if (i < a.cb_base) {                                           // conditioning weights c.j.weight [H][C] of idx
        const size_t idx = (i - a.cw_base) / a.cw_stride, e = (i - a.cw_base) - idx * a.cw_stride;
        if (e < a.hc) {
            const size_t base = idx * 2 * a.hc + e;                       // planes [idx][2][H][C]
            put4(a.cwh + base, a.hc, p, true);
            put4(a.cwb + base, a.hc, p, false);
        }
    }
}

__global__ void __launch_bounds__(256) sqnorm_kernel(const float* __restrict__ x, size_t n, double* __restrict__ out) {
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) acc += (double)x[i] * (double)x[i];
    __shared__ double s[256];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) atomicAdd(out, s[0]);
}

// the reduction step of the peer-memory gradient exchange: acc[i] += sum over the landing slots s != skip of land[s * stride + i]
template <bool kVec>
__global__ void __launch_bounds__(256) sum_shards_kernel(float* __restrict__ acc, const float* __restrict__ land, int n_src, int skip,
                                                         size_t stride, size_t n) {
    const size_t step = (size_t)gridDim.x * blockDim.x;
    if (kVec) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n / 4; i += step) {
            float4 a = reinterpret_cast<float4*>(acc)[i];
            for (int s = 0; s < n_src; ++s) {
                if (s == skip) continue;
                const float4 v = __ldcs(reinterpret_cast<const float4*>(land + (size_t)s * stride) + i);
                a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            }
            reinterpret_cast<float4*>(acc)[i] = a;
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
            float a = acc[i];
            for (int s = 0; s < n_src; ++s) if (s != skip) a += land[(size_t)s * stride + i];
            acc[i] = a;
        }
    }
}

}  // namespace tcflow
}  // namespace mhe

using namespace mhe;

extern "C" {

int mhe_flow_grad_sqnorm(mhe_flow_shape s, const float* dparams, double* out, void* stream) {
    MHE_REQUIRE(valid_shape(s) && dparams && out, "grad_sqnorm: bad args");
    FlowLayout L(s);
    if (cudaMemsetAsync(out, 0, sizeof(double), (cudaStream_t)stream) != cudaSuccess) { set_error("grad_sqnorm: memset failed"); return MHE_ERR_CUDA; }
    tcflow::sqnorm_kernel<<<592, 256, 0, (cudaStream_t)stream>>>(dparams, L.total, out);
    return check_launch("grad sqnorm");
}

int mhe_flow_adam_step(mhe_flow_shape s, float* params, const float* dparams, float* exp_avg, float* exp_avg_sq, void* packed, int step,
                       double lr, double beta1, double beta2, double eps, double grad_scale, void* stream_) {
    MHE_REQUIRE(valid_shape(s) && params && dparams && exp_avg && exp_avg_sq && step >= 1, "adam_step: bad args");
    FlowLayout L(s);
    cudaStream_t stream = (cudaStream_t)stream_;
    tcflow::AdamArgs a{};
    a.p = params; a.g = dparams; a.m = exp_avg; a.v = exp_avg_sq;
    a.n = L.total; a.blk = L.blk; a.oW1 = L.oW1; a.hh = (size_t)L.H * L.H; a.cw_base = L.cw_base; a.cw_stride = L.cw_stride;
    a.hc = (size_t)L.H * L.C; a.cb_base = L.cb_base; a.nblk = L.L * 2; a.ncw = L.L * 4;
    a.omb1 = (float)(1.0 - beta1); a.b2 = (float)beta2; a.omb2 = (float)(1.0 - beta2); a.eps = (float)eps; a.gscale = (float)grad_scale;
    a.step_size = (float)(lr / (1.0 - pow(beta1, (double)step)));
    a.bc2s = (float)sqrt(1.0 - pow(beta2, (double)step));
    if (packed) {
        if (!tcflow::supported(L)) { set_error("adam_step: shape outside the tensor-core path"); return MHE_ERR_UNSUPPORTED; }
        tcflow::Packed P(L, (__nv_bfloat16*)packed);
        a.w1h = (uint16_t*)P.w1; a.w1b = (uint16_t*)P.w1b; a.cwh = (uint16_t*)P.cw; a.cwb = (uint16_t*)P.cwb;
    }
    tcflow::adam_planes_kernel<<<cdiv((int)((L.total + 3) / 4), 256), 256, 0, stream>>>(a);
    MHE_TRY(check_launch("adam step"));
    if (packed) {   // the thin, padded families: re-packed from the updated weights (2 % of the bytes)
        tcflow::Packed P(L, (__nv_bfloat16*)packed);
        for (int f16 = 1; f16 >= 0; --f16) {
            MHE_TRY(tc::split_planes(params + L.oW0, L.D, (long)L.blk, L.H, L.D, nullptr, f16 ? P.w0 : P.w0b, L.H, tcflow::kDp, 2, L.L * 2, f16 != 0, stream));
            MHE_TRY(tc::split_planes(params + L.oW2, L.H, (long)L.blk, L.D, L.H, nullptr, f16 ? P.w2 : P.w2b, tcflow::kDp, L.H, 2, L.L * 2, f16 != 0, stream));
        }
    }
    return MHE_OK;
}

int mhe_sum_shards(float* acc, const float* land, int n_src, int skip, size_t stride, size_t n, void* stream) {
    MHE_REQUIRE(acc && land && n_src >= 1 && n_src <= 64, "sum_shards: bad args");
    if (n == 0) return MHE_OK;
    const bool vec = n % 4 == 0 && stride % 4 == 0 && ((uintptr_t)acc | (uintptr_t)land) % 16 == 0;
    const size_t work = vec ? n / 4 : n;
    const int grid = (int)(work < (size_t)148 * 8 * 256 ? (work + 255) / 256 : 148 * 8);
    if (vec) tcflow::sum_shards_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(acc, land, n_src, skip, stride, n);
    else     tcflow::sum_shards_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(acc, land, n_src, skip, stride, n);
    return check_launch("sum shards");
}

}  // extern "C"
