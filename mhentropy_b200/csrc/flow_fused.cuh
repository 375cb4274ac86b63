// Cluster-fused flow passes (flow_fused.cu): ONE kernel runs all L coupling layers of a pass for a tile of rows.
//
// Mapping.  Rows sit on the MMA N dimension (tiles of NT = 64 rows), hidden features on M (128 per CTA).  A cluster
// of 8 CTAs owns one row tile: CTAs 0-3 the s-net, 4-7 the t-net, each a 128-feature slice.  Per layer
//   G0  a0T[slice] = W0[slice] . xm^T            N-split: every CTA holds xm (64 rows x 45) in shared memory
//   G1  a1T[slice] = W1[slice, :] . a0T          a0T of the net gathered from the 4 CTAs through L2 (global scratch, TMA)
//   G2  part[d]    = W2[:, slice] . a1T[slice]   K-split: partial head outputs, summed by every CTA after an exchange
// Activations live transposed ([feature][row], split 16-bit planes), so they are directly the K-major operands of the
// batched weight-gradient GEMMs that follow the backward pass.  Synchronisation inside the cluster uses mbarriers with
// remote (cluster-scope release) arrivals: 2 exchanges per layer, no kernel boundaries, weights TMA-prefetched.
#pragma once
#include "flow_tc.cuh"

namespace mhe {
namespace fused {

constexpr int NT = 64;          // rows per tile (MMA N)
constexpr int FS = 128;         // features per CTA (MMA M)
constexpr int kCluster = 8;     // 2 nets x (512 / 128) slices
constexpr int kDp = 64;         // flow dimension padded

inline int tiles_of(int R) { return (R + NT - 1) / NT; }
inline int padded_rows(int R) { return tiles_of(R) * NT; }
// the fused path covers the production shape; other shapes / very long batches use the per-GEMM path (flow_tc.cu)
bool supported(const FlowLayout& L, int R);

// saved-for-backward block of the fused path
//   x fp32 [(L+1)][R][D] | st fp32 [L][2][R][D] (s after tanh, t) | a0T, a1T half planes [L][2 nets][2 planes][H][Rp]
struct FSaved {
    float *x_, *st_;
    __nv_bfloat16 *a0_, *a1_;
    size_t RD, act;   // act = elements of one layer's [2 nets][2 planes][H][Rp]
    static size_t bytes(const FlowLayout& L, int R) {
        const size_t RD = (size_t)R * L.D, act = (size_t)4 * L.H * padded_rows(R);
        return ((size_t)(L.L + 1) * RD + (size_t)L.L * 2 * RD) * 4 + 2 * (size_t)L.L * act * 2 + 8 * 1024;
    }
    FSaved(void* base_, const FlowLayout& L, int R) : RD((size_t)R * L.D), act((size_t)4 * L.H * padded_rows(R)) {
        uint8_t* base = (uint8_t*)base_;
        auto take = [&](size_t n) { uint8_t* p = base; base += (n + 1023) / 1024 * 1024; return p; };
        x_ = (float*)take((size_t)(L.L + 1) * RD * 4);
        st_ = (float*)take((size_t)L.L * 2 * RD * 4);
        a0_ = (__nv_bfloat16*)take((size_t)L.L * act * 2);
        a1_ = (__nv_bfloat16*)take((size_t)L.L * act * 2);
    }
    float* x(int step) const { return x_ + (size_t)step * RD; }
    float* st(int step) const { return st_ + (size_t)step * 2 * RD; }
};

// per-pass scratch
struct FWs {
    __nv_bfloat16* a0T;   // [2 nets][2 planes][H][Rp]  (forward without saving)
    float* partial;       // [2 parities][tiles][8 CTAs][64][NT]
    static size_t bytes(const FlowLayout& L, int R) {
        return (size_t)4 * L.H * padded_rows(R) * 2 + (size_t)2 * tiles_of(R) * kCluster * kDp * NT * 4 + 4096;
    }
    FWs(void* base_, const FlowLayout& L, int R) {
        uint8_t* base = (uint8_t*)base_;
        auto take = [&](size_t n) { uint8_t* p = base; base += (n + 1023) / 1024 * 1024; return p; };
        a0T = (__nv_bfloat16*)take((size_t)4 * L.H * padded_rows(R) * 2);
        partial = (float*)take((size_t)2 * tiles_of(R) * kCluster * kDp * NT * 4);
    }
};

// backward scratch: gradient planes of every layer (consumed by the batched weight-gradient GEMMs), bfloat16 re-planes of the saved
// activations and of the masked layer inputs, the exchange buffer
struct BWs {
    __nv_bfloat16 *dh1T, *dh0T;   // [L][2 nets][2 planes][H][Rp]   (indexed by layer)
    __nv_bfloat16 *dpreT;         // [L][2][2][64][Rp]
    __nv_bfloat16 *a0b, *a1b;     // [L][2][2][H][Rp]  bfloat16 planes of the saved half activations, re-indexed by layer
    __nv_bfloat16 *xmT;           // [L][2][2][64][Rp] masked layer inputs, transposed, one copy per net
    float* partial;
    static size_t bytes(const FlowLayout& L, int R) {
        const size_t act = (size_t)L.L * 4 * L.H * padded_rows(R) * 2, small = (size_t)L.L * 4 * kDp * padded_rows(R) * 2;
        return 4 * act + 2 * small + (size_t)2 * tiles_of(R) * kCluster * kDp * NT * 4 + 16 * 1024;
    }
    BWs(void* base_, const FlowLayout& L, int R) {
        uint8_t* base = (uint8_t*)base_;
        auto take = [&](size_t n) { uint8_t* p = base; base += (n + 1023) / 1024 * 1024; return p; };
        const size_t act = (size_t)L.L * 4 * L.H * padded_rows(R) * 2, small = (size_t)L.L * 4 * kDp * padded_rows(R) * 2;
        dh1T = (__nv_bfloat16*)take(act); dh0T = (__nv_bfloat16*)take(act);
        a0b = (__nv_bfloat16*)take(act); a1b = (__nv_bfloat16*)take(act);
        dpreT = (__nv_bfloat16*)take(small); xmT = (__nv_bfloat16*)take(small);
        partial = (float*)take((size_t)2 * tiles_of(R) * kCluster * kDp * NT * 4);
    }
};

// optional: the conditioning backward (mhe_flow_cond_bwd) pipelined into the pass, chunk by chunk (feat [B][C], dfeat [B][C] or NULL,
// ws: the conditioning workspace)
struct CondBwd { const float* feat; float* dfeat; void* ws; };
int pass_bwd(const FlowLayout& L, const float* params, const void* packed, const float* mask, const float* saved, int R, int B, int direction,
             const float* dout, const float* dlogdet, float dlogdet_scale, float* din, float* dparams, float* dcp, void* workspace,
             cudaStream_t stream, const CondBwd* cond = nullptr);

// Asynchronous weight gradients: with set_async_wgrad(1) pass_bwd returns with its weight-gradient GEMMs still running on internal
// streams; join(stream) makes `stream` wait for them (mhe_flow_set_async / mhe_flow_join in the C ABI).
void set_async_wgrad(int on);
int chunk_count(int L);
void chunk_layers(int L, int direction, int c, int* l0, int* nl);
int join_chunk(cudaStream_t stream, int c);   // `stream` waits for every gradient of chunk c of the last (async) pass
void set_wgrad_operands_prepared(int on);
int pass_bwd_prepare(const FlowLayout& L, const float* mask, const float* saved, int R, int direction, void* workspace, cudaStream_t stream);
int join(cudaStream_t stream);

int pass_fwd(const FlowLayout& L, const float* params, const void* packed, const float* mask, const float* cp, const float* in, int R, int B,
             int direction, float* out, float* logdet, float* saved, void* workspace, cudaStream_t stream);

}  // namespace fused
}  // namespace mhe
