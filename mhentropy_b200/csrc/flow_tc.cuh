// Declarations of the tensor-core flow path (flow_tc.cu), shared with the C-ABI dispatch in flow.cu.
#pragma once
#include "tc_gemm.cuh"

namespace mhe {
namespace tcflow {

constexpr int kDp = 64;   // flow dimension padded to one 64-wide K block / N tile

inline bool supported(const FlowLayout& L) { return L.D <= kDp && L.H % 64 == 0 && L.C % 8 == 0; }

// packed split weights, once as half planes (forward GEMMs) and once as bfloat16 planes (backward GEMMs; A and B of one MMA
// must share a format): w0 [L*2][2][H][64], w1 [L*2][2][H][H], w2 [L*2][2][64][H], cw [L*4][2][H][C]
struct Packed {
    __nv_bfloat16 *w0, *w1, *w2, *cw;       // half planes
    __nv_bfloat16 *w0b, *w1b, *w2b, *cwb;   // bfloat16 planes
    static size_t elems(const FlowLayout& L) {
        return 2 * ((size_t)L.L * 2 * 2 * ((size_t)L.H * kDp * 2 + (size_t)L.H * L.H) + (size_t)L.L * 4 * 2 * L.H * L.C) + 8 * 512;
    }
    Packed(const FlowLayout& L, __nv_bfloat16* base) {
        auto take = [&](size_t n) { __nv_bfloat16* p = base; base += (n + 511) / 512 * 512; return p; };
        w0 = take((size_t)L.L * 2 * 2 * L.H * kDp); w1 = take((size_t)L.L * 2 * 2 * L.H * L.H);
        w2 = take((size_t)L.L * 2 * 2 * kDp * L.H); cw = take((size_t)L.L * 4 * 2 * L.H * L.C);
        w0b = take((size_t)L.L * 2 * 2 * L.H * kDp); w1b = take((size_t)L.L * 2 * 2 * L.H * L.H);
        w2b = take((size_t)L.L * 2 * 2 * kDp * L.H); cwb = take((size_t)L.L * 4 * 2 * L.H * L.C);
    }
};

// per-pass scratch: activations as planes, small fp32 buffers
struct Ws {
    __nv_bfloat16 *xm, *a0, *a1, *dh0[2], *dh1[2], *dprep[2];   // gradient planes are double-buffered by layer parity:
    __nv_bfloat16 *a0b, *a1b, *xmb;                              // the weight-gradient GEMMs of a layer run on side streams,
    float *st, *dpre, *gx;                                        // on bfloat16 re-planed copies of the saved activations
    static size_t bytes(const FlowLayout& L, int R) {
        const size_t act = (size_t)2 * 2 * R * L.H * 2, small = (size_t)2 * 2 * R * kDp * 2;
        return 8 * act + 5 * small + ((size_t)5 * R * L.D) * 4 + 48 * 1024;
    }
    Ws(void* base_, const FlowLayout& L, int R) {
        uint8_t* base = (uint8_t*)base_;
        auto take = [&](size_t n) { uint8_t* p = base; base += (n + 1023) / 1024 * 1024; return p; };
        const size_t act = (size_t)2 * 2 * R * L.H * 2;
        xm = (__nv_bfloat16*)take((size_t)2 * R * kDp * 2);
        a0 = (__nv_bfloat16*)take(act); a1 = (__nv_bfloat16*)take(act);
        for (int i = 0; i < 2; ++i) {
            dh0[i] = (__nv_bfloat16*)take(act); dh1[i] = (__nv_bfloat16*)take(act);
            dprep[i] = (__nv_bfloat16*)take((size_t)2 * 2 * R * kDp * 2);
        }
        a0b = (__nv_bfloat16*)take(act); a1b = (__nv_bfloat16*)take(act); xmb = (__nv_bfloat16*)take((size_t)2 * R * kDp * 2);
        st = (float*)take((size_t)2 * R * L.D * 4);
        dpre = (float*)take((size_t)2 * R * L.D * 4);
        gx = (float*)take((size_t)R * L.D * 4);
    }
};

// saved-for-backward block of the tensor-core path: layer inputs, head outputs and the hidden activations
//   x  fp32 [(L+1)][R][D] | st fp32 [L][2][R][D] | xm bf16 [L][2][R][64] | a0, a1 bf16 [L][2 nets][2 planes][R][H]
struct Saved {
    float *x_, *st_;
    __nv_bfloat16 *xm_, *a0_, *a1_;
    size_t RD, RDp, RH;
    static size_t bytes(const FlowLayout& L, int R) {
        const size_t RD = (size_t)R * L.D, RDp = (size_t)R * kDp, RH = (size_t)R * L.H;
        return ((size_t)(L.L + 1) * RD + (size_t)L.L * 2 * RD) * 4 + ((size_t)L.L * 2 * RDp + 2 * (size_t)L.L * 4 * RH) * 2 + 8 * 1024;
    }
    Saved(float* base_, const FlowLayout& L, int R) : RD((size_t)R * L.D), RDp((size_t)R * kDp), RH((size_t)R * L.H) {
        uint8_t* base = (uint8_t*)base_;
        auto take = [&](size_t n) { uint8_t* p = base; base += (n + 1023) / 1024 * 1024; return p; };
        x_ = (float*)take((size_t)(L.L + 1) * RD * 4);
        st_ = (float*)take((size_t)L.L * 2 * RD * 4);
        xm_ = (__nv_bfloat16*)take((size_t)L.L * 2 * RDp * 2);
        a0_ = (__nv_bfloat16*)take((size_t)L.L * 4 * RH * 2);
        a1_ = (__nv_bfloat16*)take((size_t)L.L * 4 * RH * 2);
    }
    float* x(int step) const { return x_ + (size_t)step * RD; }
    float* st(int step) const { return st_ + (size_t)step * 2 * RD; }
    __nv_bfloat16* xm(int step) const { return xm_ + (size_t)step * 2 * RDp; }
    __nv_bfloat16* a0(int step) const { return a0_ + (size_t)step * 4 * RH; }
    __nv_bfloat16* a1(int step) const { return a1_ + (size_t)step * 4 * RH; }
};

inline size_t cond_ws_bytes(const FlowLayout& L, int B) {
    return ((size_t)2 * B * L.C + 512 + (size_t)2 * B * L.L * 4 * L.H) * 2 + 4096;
}

// dW[batch] (+)= A[batch] . B[batch]^T for K-major bfloat16 plane operands (A: [M rows][K], B: [N rows][K]); rows of dW have pitch ld.
// transposed != 0 stores element (m, n) at dW[n][m] instead.
// The caller may promise that the weight slots of dparams are zero when the backward entry points run (mhe_flow_set_async bit 1):
// the weight-gradient epilogues then store instead of read-modify-write.
// conditioning projections straight from the fp32 weights (cond_direct.cu): no conditioning planes needed for the forward
bool cond_direct_supported(const FlowLayout& L, int B);
int cond_fwd_direct(const FlowLayout& L, const float* params, const float* feat, int B, float* cp, void* ws, cudaStream_t stream);
int zero_bias_grads(const FlowLayout& L, float* dparams, cudaStream_t stream);
void set_skip_cond_wgrad(int on);
int cond_wgrad(const FlowLayout& L, const float* feat, const float* dcp, int Bt, float* dparams, void* ws, cudaStream_t stream);
int cond_bwd_feat_planes(const FlowLayout& L, const float* feat, int B, void* ws, cudaStream_t stream);
int cond_bwd_dcp_planes(const FlowLayout& L, const float* dcp, int B, void* ws, int l0, int nl, cudaStream_t stream);
int cond_bwd_layers(const FlowLayout& L, const void* packed, const float* dcp, int B, float* dparams, float* dfeat, void* ws, int l0, int nl,
                    cudaStream_t s_bias, cudaStream_t s_wgrad, cudaStream_t s_dfeat);
bool dfeat_is_zero();
void set_grads_are_zero(int on);
void set_dfeat_is_zero(int on);   // mhe_flow_set_async bit 2: dfeat is zero when cond_bwd runs (its memset is skipped)
bool grads_are_zero();

int wgrad_kmajor(const tc::PlaneTensor& A, const tc::PlaneTensor& B, tc::GemmShape g, float* dW, long ld, long batch_stride, int ncols,
                 int transposed, cudaStream_t stream, const char* what);

int pack_weights(const FlowLayout& L, const float* params, void* packed, int which, cudaStream_t stream);
int cond_fwd(const FlowLayout& L, const float* params, const void* packed, const float* feat, int B, float* cp, void* ws, cudaStream_t stream);
int cond_bwd(const FlowLayout& L, const float* params, const void* packed, const float* feat, const float* dcp, int B, float* dparams,
             float* dfeat, void* ws, cudaStream_t stream);
int pass_fwd(const FlowLayout& L, const float* params, const void* packed, const float* mask, const float* cp, const float* in, int R, int B,
             int direction, float* out, float* logdet, float* saved, void* workspace, cudaStream_t stream);
// forward pass with ONE conditioning row per flow row, the projections contracted inside the coupling GEMMs (flow_tc.cu RowCond)
bool rowcond_supported(const FlowLayout& L);
size_t rowcond_ws_bytes(const FlowLayout& L, int R);
int pass_fwd_rowcond(const FlowLayout& L, const float* params, const void* packed, const float* mask, const float* feat, const float* in, int R,
                     int direction, float* out, float* logdet, void* workspace, cudaStream_t stream);
int pass_bwd(const FlowLayout& L, const float* params, const void* packed, const float* mask, const float* cp, const float* saved, int R, int B,
             int direction, const float* dout, const float* dlogdet, float dlogdet_scale, float* din, float* dparams, float* dcp,
             void* workspace, cudaStream_t stream);

}  // namespace tcflow
}  // namespace mhe
