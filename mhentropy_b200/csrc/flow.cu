// Conditional RealNVP coupling layers (reference hand/flows.py:75-122, 210-227) — exact-fp32 path.
//
// Per layer and per net (s, t):  a0 = lrelu(x_masked W0^T + cp0),  a1 = lrelu(a0 W1^T + cp1),
// o = a1 W2^T + b2 (tanh for s), then the affine coupling.  cp0/cp1 are the conditioning
// projections hoisted to one GEMM per step (mhe_flow_cond_fwd) with the l.0/l.1 biases folded in.
// The backward recomputes a0/a1 from the saved layer input instead of storing them.
#include "gemm_simt.cuh"
#include "flow_tc.cuh"
#include "flow_fused.cuh"

namespace mhe {

// ---- epilogues ------------------------------------------------------------------------------
struct EpiCpLrelu {  // act[batch][m][n] = lrelu(acc + cp[m % B][cp_off + batch*cp_bstride + n])
    float* out; long ld; long strideOut;
    const float* cp; long cp_ld; long cp_off; long cp_bstride; int B;
    __device__ void operator()(int b, int, int m, int n, float acc) const {
        const float c = __ldg(cp + (long)(m % B) * cp_ld + cp_off + (long)b * cp_bstride + n);
        out[(long)b * strideOut + (long)m * ld + n] = lrelu(acc + c);
    }
};
struct EpiOutHead {  // st[batch][m][n] = batch == 0 ? tanh(acc + b2) : acc + b2
    float* out; long ld; long strideOut; const float* bias; long strideBias;
    __device__ void operator()(int b, int, int m, int n, float acc) const {
        float v = acc + __ldg(bias + (long)b * strideBias + n);
        if (b == 0) v = tanhf(v);
        out[(long)b * strideOut + (long)m * ld + n] = v;
    }
};
struct EpiActGrad {  // out = acc * lrelu'(act)
    float* out; const float* act; long ld; long stride;
    __device__ void operator()(int b, int, int m, int n, float acc) const {
        const long i = (long)b * stride + (long)m * ld + n;
        out[i] = acc * lrelu_grad_from_out(act[i]);
    }
};
struct EpiMaskAtomicAdd {  // out[m][n] += mask[n] * acc   (the two nets land on the same gradient)
    float* out; long ld; const float* mask;
    __device__ void operator()(int, int, int m, int n, float acc) const {
        const float w = __ldg(mask + n);
        if (w != 0.f) atomicAdd(out + (long)m * ld + n, w * acc);
    }
};
struct EpiCondFwd {  // cp[m][idx*H + n] = acc + cb[idx][n] + bj[idx][n]
    float* cp; long cp_ld; int H; const float* params; size_t cb_base, cb_stride, blk, ob0, ob1;
    __device__ void operator()(int idx, int, int m, int n, float acc) const {
        const int j = idx & 1;
        const size_t blkoff = (size_t)(idx >> 1) * blk + (j ? ob1 : ob0);
        cp[(long)m * cp_ld + (long)idx * H + n] = acc + __ldg(params + cb_base + (size_t)idx * cb_stride + n) + __ldg(params + blkoff + n);
    }
};
struct EpiAtomicStore {  // C[m][n] += acc (K or batch split)
    float* C; long ldc;
    __device__ void operator()(int, int, int m, int n, float acc) const { atomicAdd(C + (long)m * ldc + n, acc); }
};

// ---- elementwise kernels ----------------------------------------------------------------------
// affine coupling forward. st = [2][R][D] (s then t). direction 0: y = m x + (1-m)(x e^s + t), ld += sum s
// direction 1: y = (1-m)(x - t) e^-s + m x, ld -= sum s
__global__ void coupling_fwd_kernel(const float* __restrict__ x, const float* __restrict__ st, const float* __restrict__ mask,
                                    int R, int D, int direction, float* __restrict__ y, float* __restrict__ logdet) {
    const int r = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    float ssum = 0.f;
    for (int d = lane; d < D; d += 32) {
        const float m = mask[d];
        const float xv = x[(long)r * D + d];
        float out = xv;
        if (m == 0.f) {
            const float s = st[(long)r * D + d];
            const float t = st[(long)R * D + (long)r * D + d];
            out = direction == 0 ? fmaf(xv, expf(s), t) : (xv - t) * expf(-s);
            ssum += s;
        }
        y[(long)r * D + d] = out;
    }
    if (logdet) {
        ssum = warp_sum(ssum);
        if (lane == 0) logdet[r] += direction == 0 ? ssum : -ssum;
    }
}

// coupling backward: from g = dL/dy, gl = dL/dlogdet to
//   dpre[2][R][D]: gradient at the pre-activation of the output heads (s before tanh, t), zero on passive dims
//   gx[R][D]: dL/dx without the path through the nets (added later by the dgrad of layer 0)
__global__ void coupling_bwd_kernel(const float* __restrict__ x, const float* __restrict__ st, const float* __restrict__ mask,
                                    const float* g, const float* __restrict__ gl, float gl_scale, int R, int D, int direction,
                                    float* __restrict__ dpre, float* gx) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)R * D) return;
    const int r = (int)(i / D), d = (int)(i % D);
    const float m = mask[d];
    const float gv = g[i];
    if (m != 0.f) {
        dpre[i] = 0.f;
        dpre[(long)R * D + i] = 0.f;
        gx[i] = gv;
        return;
    }
    const float s = st[i], t = st[(long)R * D + i], xv = x[i];
    const float glv = gl ? gl_scale * gl[r] : 0.f;
    float ds, dt, dx;
    if (direction == 0) {
        const float e = expf(s);
        dx = gv * e;
        dt = gv;
        ds = gv * xv * e + glv;
    } else {
        const float e = expf(-s);
        dx = gv * e;
        dt = -gv * e;
        ds = -gv * (xv - t) * e - glv;
    }
    dpre[i] = ds * (1.f - s * s);  // tanh'
    dpre[(long)R * D + i] = dt;
    gx[i] = dx;
}

// out[z][n] (+)= sum_m in[z][m][n]   (bias gradients)
__global__ void colsum_kernel(const float* __restrict__ in, int M, int N, long strideIn, float* __restrict__ out, long strideOut) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int z = blockIdx.z;
    const int chunks = gridDim.y;
    if (n >= N) return;
    const int per = (M + chunks - 1) / chunks;
    const int mb = blockIdx.y * per, me = min(M, mb + per);
    float acc = 0.f;
    for (int m = mb; m < me; ++m) acc += in[(long)z * strideIn + (long)m * N + n];
    if (me > mb) atomicAdd(out + (long)z * strideOut + n, acc);
}

// dcp[b][off + z*zstride + n] += sum_s in[z][s*B + b][n]   (sum over the hypotheses of an image)
__global__ void hyp_sum_kernel(const float* __restrict__ in, int R, int B, int N, long strideIn,
                               float* __restrict__ dcp, long cp_ld, long off, long zstride) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    const int z = blockIdx.z;
    if (n >= N) return;
    float acc = 0.f;
    for (int r = b; r < R; r += B) acc += in[(long)z * strideIn + (long)r * N + n];
    dcp[(long)b * cp_ld + off + (long)z * zstride + n] += acc;
}

__global__ void std_normal_logp_fwd_kernel(const float* __restrict__ z, const float* __restrict__ logdet, float sign, int R, int D, float* __restrict__ logp) {
    const int r = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) { const float v = z[(long)r * D + d]; ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    if (lane == 0) logp[r] = -0.5f * ss - 0.5f * D * 1.8378770664093453f + (logdet ? sign * logdet[r] : 0.f);
}
__global__ void std_normal_logp_bwd_kernel(const float* __restrict__ z, const float* __restrict__ dlogp, int R, int D, float* __restrict__ dz) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)R * D) return;
    dz[i] = -z[i] * dlogp[i / D];
}

// ---- workspace carving ------------------------------------------------------------------------
struct FlowWs {
    float *a0, *a1, *st, *dh1, *dh0, *dpre, *gx;
    static size_t floats(const FlowLayout& L, int R) {
        return 4 * (size_t)2 * R * L.H + 2 * (size_t)2 * R * L.D + (size_t)R * L.D + 7 * 64;
    }
    FlowWs(float* base, const FlowLayout& L, int R) {
        auto take = [&](size_t n) { float* p = base; base += (n + 63) / 64 * 64; return p; };
        a0 = take((size_t)2 * R * L.H);
        a1 = take((size_t)2 * R * L.H);
        dh1 = take((size_t)2 * R * L.H);
        dh0 = take((size_t)2 * R * L.H);
        st = take((size_t)2 * R * L.D);
        dpre = take((size_t)2 * R * L.D);
        gx = take((size_t)R * L.D);
    }
};

// nets of one layer: x (layer input) -> a0, a1, st
static int layer_nets_fwd(const FlowLayout& L, const float* params, const float* mask_row, const float* cp, const float* x,
                          int R, int B, int layer, FlowWs& ws, cudaStream_t stream) {
    const long cp_ld = (long)L.L * 4 * L.H;
    const float* blk = params + L.block(layer, 0);
    {   // G0: [R][D] x W0^T -> a0, both nets
        GemmArgs g; g.A = x; g.lda = L.D; g.strideA = 0; g.a_kscale = mask_row;
        g.B = blk + L.oW0; g.ldb = L.D; g.strideB = (long)L.blk;
        g.M = R; g.N = L.H; g.K = L.D; g.batches = 2;
        EpiCpLrelu e{ws.a0, L.H, (long)R * L.H, cp, cp_ld, (long)(layer * 4 + 0) * L.H, (long)2 * L.H, B};
        MHE_TRY((launch_sgemm<Major::K, Major::K>(g, e, stream, "flow G0")));
    }
    {   // G1: a0 x W1^T -> a1
        GemmArgs g; g.A = ws.a0; g.lda = L.H; g.strideA = (long)R * L.H;
        g.B = blk + L.oW1; g.ldb = L.H; g.strideB = (long)L.blk;
        g.M = R; g.N = L.H; g.K = L.H; g.batches = 2;
        EpiCpLrelu e{ws.a1, L.H, (long)R * L.H, cp, cp_ld, (long)(layer * 4 + 1) * L.H, (long)2 * L.H, B};
        MHE_TRY((launch_sgemm<Major::K, Major::K>(g, e, stream, "flow G1")));
    }
    {   // G2: a1 x W2^T + b2 -> st (tanh on s)
        GemmArgs g; g.A = ws.a1; g.lda = L.H; g.strideA = (long)R * L.H;
        g.B = blk + L.oW2; g.ldb = L.H; g.strideB = (long)L.blk;
        g.M = R; g.N = L.D; g.K = L.H; g.batches = 2;
        EpiOutHead e{ws.st, L.D, (long)R * L.D, blk + L.ob2, (long)L.blk};
        MHE_TRY((launch_sgemm<Major::K, Major::K>(g, e, stream, "flow G2")));
    }
    return MHE_OK;
}


__global__ void cond_bias_grad_kernel(const float* __restrict__ dcp, int B, long cp_ld, int H, float* __restrict__ dparams,
                                      size_t cb_base, size_t cb_stride, size_t blk, size_t ob0, size_t ob1) {
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= cp_ld) return;
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc += dcp[(long)b * cp_ld + n];
    const int idx = (int)(n / H), h = (int)(n % H);
    dparams[cb_base + (size_t)idx * cb_stride + h] += acc;                          // c.j.bias
    dparams[(size_t)(idx >> 1) * blk + ((idx & 1) ? ob1 : ob0) + h] += acc;         // folded l.j.bias
}

static int cuda_ok(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return MHE_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return MHE_ERR_CUDA;
}

}  // namespace mhe

using namespace mhe;

extern "C" {

size_t mhe_flow_param_floats(mhe_flow_shape s) { return valid_shape(s) ? FlowLayout(s).total : 0; }

size_t mhe_flow_param_offset(mhe_flow_shape s, int layer, int net, int which) {
    if (!valid_shape(s) || layer < 0 || layer >= s.layers || net < 0 || net > 1) return (size_t)-1;
    FlowLayout L(s);
    const size_t b = L.block(layer, net);
    switch (which) {
        case 0: return b + L.oW0;
        case 1: return b + L.ob0;
        case 2: return b + L.oW1;
        case 3: return b + L.ob1;
        case 4: return b + L.oW2;
        case 5: return b + L.ob2;
        case 6: return L.cw(layer, net, 0);
        case 7: return L.cb(layer, net, 0);
        case 8: return L.cw(layer, net, 1);
        case 9: return L.cb(layer, net, 1);
    }
    return (size_t)-1;
}

size_t mhe_flow_cp_floats_per_image(mhe_flow_shape s) { return (size_t)s.layers * 4 * s.hidden; }

size_t mhe_flow_workspace_bytes(mhe_flow_shape s, int R, int tensor_core) {
    if (!valid_shape(s) || R < 0) return 0;
    FlowLayout L(s);
    if (tensor_core) {
        if (!tcflow::supported(L)) return 0;
        size_t n = tcflow::Ws::bytes(L, R);
        if (fused::supported(L, R)) {
            n = n > fused::FWs::bytes(L, R) ? n : fused::FWs::bytes(L, R);
            n = n > fused::BWs::bytes(L, R) ? n : fused::BWs::bytes(L, R);
        }
        return n;
    }
    return FlowWs::floats(L, R) * sizeof(float);
}

size_t mhe_flow_saved_bytes(mhe_flow_shape s, int R, int tensor_core) {
    if (!valid_shape(s) || R < 0) return 0;
    FlowLayout L(s);
    if (tensor_core) {
        if (!tcflow::supported(L)) return 0;
        return fused::supported(L, R) ? fused::FSaved::bytes(L, R) : tcflow::Saved::bytes(L, R);
    }
    return (size_t)(L.L + 1) * R * L.D * sizeof(float);
}

size_t mhe_flow_cond_workspace_bytes(mhe_flow_shape s, int B) {
    if (!valid_shape(s) || B < 0) return 0;
    FlowLayout L(s);
    return tcflow::supported(L) ? tcflow::cond_ws_bytes(L, B) : 0;
}

size_t mhe_flow_packed_bytes(mhe_flow_shape s) {
    if (!valid_shape(s)) return 0;
    FlowLayout L(s);
    return tcflow::supported(L) ? tcflow::Packed::elems(L) * 2 : 0;
}

int mhe_flow_pack_weights(mhe_flow_shape s, const float* params, void* packed, int which, void* stream) {
    MHE_REQUIRE(valid_shape(s) && params && packed && which >= 1 && which <= 63, "pack_weights: bad args");
    FlowLayout L(s);
    if (!tcflow::supported(L)) { set_error("pack_weights: shape outside the tensor-core path (dim <= 64, hidden %% 64 == 0, cond %% 8 == 0)"); return MHE_ERR_UNSUPPORTED; }
    return tcflow::pack_weights(L, params, packed, which, (cudaStream_t)stream);
}

int mhe_flow_zero_bias_grads(mhe_flow_shape s, float* dparams, void* stream) {
    MHE_REQUIRE(valid_shape(s) && dparams, "zero_bias_grads: bad args");
    FlowLayout L(s);
    return tcflow::zero_bias_grads(L, dparams, (cudaStream_t)stream);
}

int mhe_flow_cond_fwd_uses_planes(mhe_flow_shape s, int B) {
    if (!valid_shape(s)) return 0;
    FlowLayout L(s);
    return tcflow::supported(L) && !tcflow::cond_direct_supported(L, B) ? 1 : 0;
}

int mhe_flow_cond_fwd(mhe_flow_shape s, const float* params, const void* packed, const float* feat, int B, float* cp,
                      void* workspace, size_t workspace_bytes, void* stream_) {
    MHE_REQUIRE(valid_shape(s), "cond_fwd: bad shape");
    MHE_REQUIRE(params && feat && cp && B >= 0, "cond_fwd: null pointer or negative B");
    if (B == 0) return MHE_OK;
    FlowLayout L(s);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (packed) {
        if (!tcflow::supported(L)) { set_error("cond_fwd: shape outside the tensor-core path"); return MHE_ERR_UNSUPPORTED; }
        if (!workspace || workspace_bytes < tcflow::cond_ws_bytes(L, B)) { set_error("cond_fwd: workspace too small"); return MHE_ERR_WORKSPACE; }
        return tcflow::cond_fwd(L, params, packed, feat, B, cp, workspace, stream);
    }
    GemmArgs g; g.A = feat; g.lda = L.C; g.strideA = 0;
    g.B = params + L.cw_base; g.ldb = L.C; g.strideB = (long)L.cw_stride;
    g.M = B; g.N = L.H; g.K = L.C; g.batches = L.L * 4;
    EpiCondFwd e{cp, (long)L.L * 4 * L.H, L.H, params, L.cb_base, L.cb_stride, L.blk, L.ob0, L.ob1};
    return launch_sgemm<Major::K, Major::K>(g, e, stream, "cond fwd");
}

int mhe_flow_cond_bwd(mhe_flow_shape s, const float* params, const void* packed, const float* feat, const float* dcp, int B,
                      float* dparams, float* dfeat, void* workspace, size_t workspace_bytes, void* stream_) {
    MHE_REQUIRE(valid_shape(s), "cond_bwd: bad shape");
    MHE_REQUIRE(params && feat && dcp && dparams && B >= 0, "cond_bwd: null pointer or negative B");
    if (B == 0) return MHE_OK;
    FlowLayout L(s);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (packed) {
        if (!tcflow::supported(L)) { set_error("cond_bwd: shape outside the tensor-core path"); return MHE_ERR_UNSUPPORTED; }
        if (!workspace || workspace_bytes < tcflow::cond_ws_bytes(L, B)) { set_error("cond_bwd: workspace too small"); return MHE_ERR_WORKSPACE; }
        return tcflow::cond_bwd(L, params, packed, feat, dcp, B, dparams, dfeat, workspace, stream);
    }
    const long cp_ld = (long)L.L * 4 * L.H;
    {   // dCw[idx] [H][C] += dcp[:, idx, :]^T feat
        GemmArgs g; g.A = dcp; g.lda = cp_ld; g.strideA = L.H;  // A(m=h, k=b) = dcp[b*cp_ld + idx*H + h]
        g.B = feat; g.ldb = L.C; g.strideB = 0;                 // B(k=b, n=c) = feat[b*C + c]
        g.M = L.H; g.N = L.C; g.K = B; g.batches = L.L * 4;
        EpiAccumulate e{dparams + L.cw_base, L.C, (long)L.cw_stride, nullptr, 0};
        MHE_TRY((launch_sgemm<Major::MN, Major::MN>(g, e, stream, "cond wgrad")));
    }
    cond_bias_grad_kernel<<<cdiv((int)cp_ld, 256), 256, 0, stream>>>(dcp, B, cp_ld, L.H, dparams, L.cb_base, L.cb_stride, L.blk, L.ob0, L.ob1);
    MHE_TRY(check_launch("cond bias grad"));
    if (dfeat) {  // dfeat [B][C] = sum_idx dcp[:, idx, :] Cw[idx]
        MHE_TRY(cuda_ok(cudaMemsetAsync(dfeat, 0, (size_t)B * L.C * sizeof(float), stream), "memset dfeat"));
        GemmArgs g; g.A = dcp; g.lda = cp_ld; g.strideA = L.H;                  // A(m=b, k=h) = dcp[b*cp_ld + idx*H + h]
        g.B = params + L.cw_base; g.ldb = L.C; g.strideB = (long)L.cw_stride;  // B(k=h, n=c) = Cw[idx][h*C + c]
        g.M = B; g.N = L.C; g.K = L.H; g.batches = L.L * 4;
        EpiAtomicStore e{dfeat, L.C};
        MHE_TRY((launch_sgemm<Major::K, Major::MN>(g, e, stream, "cond dfeat")));
    }
    return MHE_OK;
}

int mhe_flow_rowcond_supported(mhe_flow_shape s, int R) {
    if (!valid_shape(s) || R <= 0) return 0;
    FlowLayout L(s);
    // below MHE_ROWCOND_MIN_ROWS the cluster-fused kernels on materialised projections are faster (measured: 1,024 rows 0.45 vs 0.53 ms,
    // 2,048 rows 0.68 vs 0.53 ms, 4,096 rows 1.16 vs 0.75 ms; the per-GEMM pass has a ~0.5 ms floor of 60 launches)
    static const int min_rows = [] { const char* e = getenv("MHE_ROWCOND_MIN_ROWS"); return e ? atoi(e) : 1536; }();
    if (!tcflow::rowcond_supported(L)) return 0;
    if (!fused::supported(L, R)) return 1;
    return min_rows > 0 && R >= min_rows ? 1 : 0;
}
size_t mhe_flow_rowcond_workspace_bytes(mhe_flow_shape s, int R) {
    if (!valid_shape(s) || R <= 0) return 0;
    FlowLayout L(s);
    return tcflow::rowcond_ws_bytes(L, R);
}
int mhe_flow_pass_fwd_rowcond(mhe_flow_shape s, const float* params, const void* packed, const float* mask, const float* feat,
                              const float* in, int R, int direction, float* out, float* logdet,
                              void* workspace, size_t workspace_bytes, void* stream_) {
    MHE_REQUIRE(valid_shape(s), "pass_fwd_rowcond: bad shape");
    MHE_REQUIRE(R >= 0 && direction >= 0 && direction <= 1, "pass_fwd_rowcond: bad R/direction");
    if (R == 0) return MHE_OK;
    MHE_REQUIRE(params && packed && mask && feat && in && out && workspace, "pass_fwd_rowcond: null pointer");
    FlowLayout L(s);
    if (!tcflow::rowcond_supported(L)) { set_error("pass_fwd_rowcond: shape outside the tensor-core path"); return MHE_ERR_UNSUPPORTED; }
    if (workspace_bytes < tcflow::rowcond_ws_bytes(L, R)) { set_error("pass_fwd_rowcond: workspace too small"); return MHE_ERR_WORKSPACE; }
    return tcflow::pass_fwd_rowcond(L, params, packed, mask, feat, in, R, direction, out, logdet, workspace, (cudaStream_t)stream_);
}

int mhe_flow_pass_fwd(mhe_flow_shape s, const float* params, const void* packed, const float* mask, const float* cp,
                      const float* in, int R, int B, int direction,
                      float* out, float* logdet, float* saved,
                      void* workspace, size_t workspace_bytes, void* stream_) {
    MHE_REQUIRE(valid_shape(s), "pass_fwd: bad shape");
    MHE_REQUIRE(R >= 0 && B > 0 && direction >= 0 && direction <= 1, "pass_fwd: bad R/B/direction");
    if (R == 0) return MHE_OK;
    MHE_REQUIRE(params && mask && cp && in && out && workspace, "pass_fwd: null pointer");
    FlowLayout L(s);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (packed) {
        if (!tcflow::supported(L)) { set_error("pass_fwd: shape outside the tensor-core path"); return MHE_ERR_UNSUPPORTED; }
        if (workspace_bytes < mhe_flow_workspace_bytes(s, R, 1)) { set_error("pass_fwd: workspace too small"); return MHE_ERR_WORKSPACE; }
        if (fused::supported(L, R))
            return fused::pass_fwd(L, params, packed, mask, cp, in, R, B, direction, out, logdet, saved, workspace, stream);
        return tcflow::pass_fwd(L, params, packed, mask, cp, in, R, B, direction, out, logdet, saved, workspace, stream);
    }
    if (workspace_bytes < mhe_flow_workspace_bytes(s, R, 0)) { set_error("pass_fwd: workspace too small"); return MHE_ERR_WORKSPACE; }
    FlowWs ws((float*)workspace, L, R);
    const size_t row_bytes = (size_t)R * L.D * sizeof(float);
    if (logdet) MHE_TRY(cuda_ok(cudaMemsetAsync(logdet, 0, (size_t)R * sizeof(float), stream), "memset logdet"));
    // layer inputs ping-pong between `out` and ws.gx unless they are being saved
    const float* x = in;
    for (int step = 0; step < L.L; ++step) {
        const int layer = direction == 0 ? step : L.L - 1 - step;
        if (saved) {
            float* slot = saved + (size_t)step * R * L.D;
            if (step == 0) MHE_TRY(cuda_ok(cudaMemcpyAsync(slot, in, row_bytes, cudaMemcpyDeviceToDevice, stream), "save input"));
            x = slot;
        }
        MHE_TRY(layer_nets_fwd(L, params, mask + (size_t)layer * L.D, cp, x, R, B, layer, ws, stream));
        float* y;
        if (saved) y = (step == L.L - 1) ? saved + (size_t)L.L * R * L.D : saved + (size_t)(step + 1) * R * L.D;
        else y = (step == L.L - 1) ? out : ((x == ws.gx) ? ws.dpre : ws.gx);
        coupling_fwd_kernel<<<cdiv(R, 8), 256, 0, stream>>>(x, ws.st, mask + (size_t)layer * L.D, R, L.D, direction, y, logdet);
        MHE_TRY(check_launch("coupling fwd"));
        x = y;
    }
    if (saved) MHE_TRY(cuda_ok(cudaMemcpyAsync(out, saved + (size_t)L.L * R * L.D, row_bytes, cudaMemcpyDeviceToDevice, stream), "copy out"));
    return MHE_OK;
}

int mhe_flow_pass_cond_bwd(mhe_flow_shape s, const float* params, const void* packed, const float* mask, const float* cp,
                           const float* saved, int R, int B, int direction, const float* dout, const float* dlogdet, float dlogdet_scale,
                           float* din, float* dparams, float* dcp, const float* feat, float* dfeat,
                           void* workspace, size_t workspace_bytes, void* cond_workspace, size_t cond_workspace_bytes, void* stream_) {
    MHE_REQUIRE(valid_shape(s) && feat, "pass_cond_bwd: bad shape or null feat");
    FlowLayout L(s);
    if (R > 0 && packed && tcflow::supported(L) && fused::supported(L, R)) {
        MHE_REQUIRE(B > 0 && direction >= 0 && direction <= 1, "pass_cond_bwd: bad B/direction");
        MHE_REQUIRE(params && mask && cp && saved && dout && din && dparams && dcp && workspace && cond_workspace, "pass_cond_bwd: null pointer");
        if (workspace_bytes < mhe_flow_workspace_bytes(s, R, 1)) { set_error("pass_cond_bwd: workspace too small"); return MHE_ERR_WORKSPACE; }
        if (cond_workspace_bytes < tcflow::cond_ws_bytes(L, B)) { set_error("pass_cond_bwd: conditioning workspace too small"); return MHE_ERR_WORKSPACE; }
        fused::CondBwd c{feat, dfeat, cond_workspace};
        return fused::pass_bwd(L, params, packed, mask, saved, R, B, direction, dout, dlogdet, dlogdet_scale, din, dparams, dcp, workspace,
                               (cudaStream_t)stream_, &c);
    }
    // other paths: the two entry points one after the other
    MHE_TRY(mhe_flow_pass_bwd(s, params, packed, mask, cp, saved, R, B, direction, dout, dlogdet, dlogdet_scale, din, dparams, dcp, workspace,
                              workspace_bytes, stream_));
    return mhe_flow_cond_bwd(s, params, packed, feat, dcp, B, dparams, dfeat, cond_workspace, cond_workspace_bytes, stream_);
}

int mhe_flow_cond_wgrad(mhe_flow_shape s, const float* feat, const float* dcp, int Bt, float* dparams, void* workspace, size_t workspace_bytes,
                        void* stream) {
    MHE_REQUIRE(valid_shape(s) && feat && dcp && dparams && workspace && Bt >= 0, "cond_wgrad: bad args");
    if (Bt == 0) return MHE_OK;
    FlowLayout L(s);
    if (!tcflow::supported(L)) { set_error("cond_wgrad: shape outside the tensor-core path"); return MHE_ERR_UNSUPPORTED; }
    if (workspace_bytes < tcflow::cond_ws_bytes(L, Bt)) { set_error("cond_wgrad: workspace too small"); return MHE_ERR_WORKSPACE; }
    return tcflow::cond_wgrad(L, feat, dcp, Bt, dparams, workspace, (cudaStream_t)stream);
}

int mhe_flow_pass_is_fused(mhe_flow_shape s, int R) {
    if (!valid_shape(s) || R <= 0) return 0;
    FlowLayout L(s);
    return tcflow::supported(L) && fused::supported(L, R) ? 1 : 0;
}
int mhe_flow_bwd_chunk_count(mhe_flow_shape s, int R) {
    if (!valid_shape(s)) return 1;
    FlowLayout L(s);
    return tcflow::supported(L) && fused::supported(L, R) ? fused::chunk_count(L.L) : 1;
}
int mhe_flow_bwd_chunk_layers(mhe_flow_shape s, int R, int direction, int chunk, int* first_layer, int* layers) {
    MHE_REQUIRE(valid_shape(s) && first_layer && layers && chunk >= 0 && chunk < mhe_flow_bwd_chunk_count(s, R), "bwd_chunk_layers: bad args");
    FlowLayout L(s);
    if (mhe_flow_bwd_chunk_count(s, R) == 1) { *first_layer = 0; *layers = L.L; return MHE_OK; }
    fused::chunk_layers(L.L, direction, chunk, first_layer, layers);
    return MHE_OK;
}
int mhe_flow_join_chunk(void* stream, int chunk) { return fused::join_chunk((cudaStream_t)stream, chunk); }

int mhe_flow_pass_bwd_prepare(mhe_flow_shape s, const float* mask, const float* saved, int R, int direction, void* workspace,
                              size_t workspace_bytes, void* stream) {
    MHE_REQUIRE(valid_shape(s) && R >= 0 && direction >= 0 && direction <= 1, "pass_bwd_prepare: bad args");
    if (R == 0) return MHE_OK;
    MHE_REQUIRE(mask && saved && workspace, "pass_bwd_prepare: null pointer");
    FlowLayout L(s);
    if (!tcflow::supported(L) || !fused::supported(L, R)) return MHE_ERR_UNSUPPORTED;   // nothing to prepare on the other paths
    if (workspace_bytes < mhe_flow_workspace_bytes(s, R, 1)) { set_error("pass_bwd_prepare: workspace too small"); return MHE_ERR_WORKSPACE; }
    return fused::pass_bwd_prepare(L, mask, saved, R, direction, workspace, (cudaStream_t)stream);
}

int mhe_flow_pass_bwd(mhe_flow_shape s, const float* params, const void* packed, const float* mask, const float* cp,
                      const float* saved, int R, int B, int direction,
                      const float* dout, const float* dlogdet, float dlogdet_scale,
                      float* din, float* dparams, float* dcp,
                      void* workspace, size_t workspace_bytes, void* stream_) {
    MHE_REQUIRE(valid_shape(s), "pass_bwd: bad shape");
    MHE_REQUIRE(R >= 0 && B > 0 && direction >= 0 && direction <= 1, "pass_bwd: bad R/B/direction");
    if (R == 0) return MHE_OK;
    MHE_REQUIRE(params && mask && cp && saved && dout && din && dparams && dcp && workspace, "pass_bwd: null pointer");
    FlowLayout L(s);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (packed) {
        if (!tcflow::supported(L)) { set_error("pass_bwd: shape outside the tensor-core path"); return MHE_ERR_UNSUPPORTED; }
        if (workspace_bytes < mhe_flow_workspace_bytes(s, R, 1)) { set_error("pass_bwd: workspace too small"); return MHE_ERR_WORKSPACE; }
        if (fused::supported(L, R))
            return fused::pass_bwd(L, params, packed, mask, saved, R, B, direction, dout, dlogdet, dlogdet_scale, din, dparams, dcp, workspace, stream);
        return tcflow::pass_bwd(L, params, packed, mask, cp, saved, R, B, direction, dout, dlogdet, dlogdet_scale, din, dparams, dcp, workspace, stream);
    }
    if (workspace_bytes < mhe_flow_workspace_bytes(s, R, 0)) { set_error("pass_bwd: workspace too small"); return MHE_ERR_WORKSPACE; }
    FlowWs ws((float*)workspace, L, R);
    const long cp_ld = (long)L.L * 4 * L.H;
    const long RH = (long)R * L.H, RD = (long)R * L.D;
    const float* g = dout;  // dL/d(layer output)
    for (int step = L.L - 1; step >= 0; --step) {
        const int layer = direction == 0 ? step : L.L - 1 - step;
        const float* x = saved + (size_t)step * R * L.D;
        const float* mrow = mask + (size_t)layer * L.D;
        const float* blk = params + L.block(layer, 0);
        float* dblk = dparams + L.block(layer, 0);
        MHE_TRY(layer_nets_fwd(L, params, mrow, cp, x, R, B, layer, ws, stream));  // recompute a0, a1, st
        // gradient wrt this layer's input is built in din when it is the final result, else in ws.gx
        float* gx = (step == 0) ? din : ws.gx;
        coupling_bwd_kernel<<<cdiv((int)RD, 256), 256, 0, stream>>>(x, ws.st, mrow, g, dlogdet, dlogdet_scale, R, L.D, direction, ws.dpre, gx);
        MHE_TRY(check_launch("coupling bwd"));
        {   // dgrad G2: dh1 = (dpre W2) * lrelu'(a1)
            GemmArgs a; a.A = ws.dpre; a.lda = L.D; a.strideA = RD;
            a.B = blk + L.oW2; a.ldb = L.H; a.strideB = (long)L.blk;
            a.M = R; a.N = L.H; a.K = L.D; a.batches = 2;
            EpiActGrad e{ws.dh1, ws.a1, L.H, RH};
            MHE_TRY((launch_sgemm<Major::K, Major::MN>(a, e, stream, "dgrad G2")));
        }
        {   // dgrad G1: dh0 = (dh1 W1) * lrelu'(a0)
            GemmArgs a; a.A = ws.dh1; a.lda = L.H; a.strideA = RH;
            a.B = blk + L.oW1; a.ldb = L.H; a.strideB = (long)L.blk;
            a.M = R; a.N = L.H; a.K = L.H; a.batches = 2;
            EpiActGrad e{ws.dh0, ws.a0, L.H, RH};
            MHE_TRY((launch_sgemm<Major::K, Major::MN>(a, e, stream, "dgrad G1")));
        }
        {   // dgrad G0: gx += mask * (dh0 W0), both nets
            GemmArgs a; a.A = ws.dh0; a.lda = L.H; a.strideA = RH;
            a.B = blk + L.oW0; a.ldb = L.D; a.strideB = (long)L.blk;
            a.M = R; a.N = L.D; a.K = L.H; a.batches = 2;
            EpiMaskAtomicAdd e{gx, L.D, mrow};
            MHE_TRY((launch_sgemm<Major::K, Major::MN>(a, e, stream, "dgrad G0")));
        }
        // weight gradients (K = R)
        const int ks = (R >= 16384) ? 8 : 1;
        {   // dW2 [D][H] += dpre^T a1
            GemmArgs a; a.A = ws.dpre; a.lda = L.D; a.strideA = RD;
            a.B = ws.a1; a.ldb = L.H; a.strideB = RH;
            a.M = L.D; a.N = L.H; a.K = R; a.batches = 2; a.ksplit = ks;
            EpiAccumulate e{dblk + L.oW2, L.H, (long)L.blk, nullptr, ks > 1};
            MHE_TRY((launch_sgemm<Major::MN, Major::MN>(a, e, stream, "wgrad W2")));
        }
        {   // dW1 [H][H] += dh1^T a0
            GemmArgs a; a.A = ws.dh1; a.lda = L.H; a.strideA = RH;
            a.B = ws.a0; a.ldb = L.H; a.strideB = RH;
            a.M = L.H; a.N = L.H; a.K = R; a.batches = 2; a.ksplit = ks;
            EpiAccumulate e{dblk + L.oW1, L.H, (long)L.blk, nullptr, ks > 1};
            MHE_TRY((launch_sgemm<Major::MN, Major::MN>(a, e, stream, "wgrad W1")));
        }
        {   // dW0 [H][D] += dh0^T (x * mask)
            GemmArgs a; a.A = ws.dh0; a.lda = L.H; a.strideA = RH;
            a.B = x; a.ldb = L.D; a.strideB = 0;
            a.M = L.H; a.N = L.D; a.K = R; a.batches = 2; a.ksplit = ks;
            EpiAccumulate e{dblk + L.oW0, L.D, (long)L.blk, mrow, ks > 1};
            MHE_TRY((launch_sgemm<Major::MN, Major::MN>(a, e, stream, "wgrad W0")));
        }
        {   // db2 += colsum(dpre)
            dim3 grid(cdiv(L.D, 64), min(64, cdiv(R, 64)), 2);
            colsum_kernel<<<grid, 64, 0, stream>>>(ws.dpre, R, L.D, RD, dblk + L.ob2, (long)L.blk);
            MHE_TRY(check_launch("db2"));
        }
        {   // dcp: sum over hypotheses of dh0 (j = 0) and dh1 (j = 1)
            dim3 grid(cdiv(L.H, 128), B, 2);
            hyp_sum_kernel<<<grid, 128, 0, stream>>>(ws.dh0, R, B, L.H, RH, dcp, cp_ld, (long)(layer * 4 + 0) * L.H, (long)2 * L.H);
            MHE_TRY(check_launch("dcp0"));
            hyp_sum_kernel<<<grid, 128, 0, stream>>>(ws.dh1, R, B, L.H, RH, dcp, cp_ld, (long)(layer * 4 + 1) * L.H, (long)2 * L.H);
            MHE_TRY(check_launch("dcp1"));
        }
        g = gx;  // the next coupling_bwd reads g[i] and writes gx[i] from the same thread: in place is safe
    }
    return MHE_OK;
}

int mhe_flow_set_async(int on) {
    fused::set_async_wgrad(on & 1);
    tcflow::set_grads_are_zero((on >> 1) & 1);
    tcflow::set_dfeat_is_zero((on >> 2) & 1);
    fused::set_wgrad_operands_prepared((on >> 3) & 1);
    tcflow::set_skip_cond_wgrad((on >> 4) & 1);
    return MHE_OK;
}
int mhe_flow_join(void* stream) { return fused::join((cudaStream_t)stream); }

int mhe_std_normal_logp_fwd(const float* z, const float* logdet, float logdet_sign, int R, int D, float* logp, void* stream) {
    if (R == 0) return MHE_OK;
    MHE_REQUIRE(z && logp && R >= 0 && D > 0, "std_normal_logp_fwd: bad args");
    if (R == 0) return MHE_OK;
    std_normal_logp_fwd_kernel<<<cdiv(R, 8), 256, 0, (cudaStream_t)stream>>>(z, logdet, logdet_sign, R, D, logp);
    return check_launch("std normal logp fwd");
}

int mhe_std_normal_logp_bwd(const float* z, const float* dlogp, int R, int D, float* dz, void* stream) {
    if (R == 0) return MHE_OK;
    MHE_REQUIRE(z && dlogp && dz && R >= 0 && D > 0, "std_normal_logp_bwd: bad args");
    if (R == 0) return MHE_OK;
    std_normal_logp_bwd_kernel<<<cdiv((int)((long)R * D), 256), 256, 0, (cudaStream_t)stream>>>(z, dlogp, R, D, dz);
    return check_launch("std normal logp bwd");
}

}  // extern "C"
