// Tensor-core path of the conditional RealNVP (reference hand/flows.py:75-122, 210-227): every contraction of
// the coupling MLPs and of the hoisted conditioning runs on tcgen05 through tc_gemm.cuh in bf16x3 split
// precision; activations travel between kernels as split-bf16 planes.  Same maths and same C ABI as the fp32
// path in flow.cu — selected by passing packed weights (mhe_flow_pack_weights) to the entry points.
#include "flow_tc.cuh"

namespace mhe {
namespace tcflow {
using namespace tc;
typedef __nv_bfloat16 bf16;

// ---- plane stores -------------------------------------------------------------------------------
// Forward operands (weights, activations) are half planes, gradients bfloat16 planes (see tc_gemm.cuh).
template <bool F16>
__device__ __forceinline__ void store_planes32(bf16* hi_ptr, bf16* lo_ptr, const float* v) {
    uint32_t h[16], l[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const uint16_t h0 = to16<F16>(v[2 * j]), h1 = to16<F16>(v[2 * j + 1]);
        const uint16_t l0 = to16<F16>(v[2 * j] - from16<F16>(h0)), l1 = to16<F16>(v[2 * j + 1] - from16<F16>(h1));
        h[j] = (uint32_t)h0 | ((uint32_t)h1 << 16);
        l[j] = (uint32_t)l0 | ((uint32_t)l1 << 16);
    }
    uint4* ph = reinterpret_cast<uint4*>(hi_ptr);
    uint4* pl = reinterpret_cast<uint4*>(lo_ptr);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        ph[j] = make_uint4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
        pl[j] = make_uint4(l[4 * j], l[4 * j + 1], l[4 * j + 2], l[4 * j + 3]);
    }
}

// ---- epilogues (one call per thread per 32-column chunk of its accumulator row) -----------------------
struct EpiHiddenPlanes {   // a = lrelu(acc + cp[row % B][...]) -> planes out[batch][plane][row][col]
    static constexpr bool kDirect = true, kStaged = false, kRmw = false, kTile8 = true;
    bf16* out; long ld; long plane_stride; long batch_stride;
    const float* cp; long cp_ld; long cp_off; long cp_bstride; int B;
    struct Pre { float4 a, b; };                         // the row's 8 conditioning terms, loaded ahead of the accumulator
    __device__ void pre8(int b, int row, int col, Pre& p) const {
        const float4* c = reinterpret_cast<const float4*>(cp + (long)(row % B) * cp_ld + cp_off + (long)b * cp_bstride + col);
        p.a = __ldg(c); p.b = __ldg(c + 1);
    }
    typedef const float* RowRef;                         // the row's conditioning terms (the modulo once per tile, not once per chunk)
    __device__ RowRef pre_row(int b, int row) const { return cp + (long)(row % B) * cp_ld + cp_off + (long)b * cp_bstride; }
    __device__ void pre8r(RowRef r, int col, Pre& p) const {
        const float4* c = reinterpret_cast<const float4*>(r + col);
        p.a = __ldg(c); p.b = __ldg(c + 1);
    }
    __device__ void tile8(int b, int, int row, int col, float* v, const Pre& pr, const GemmShape&) const {   // 8 columns of one row (see tc_gemm.cuh)
        const float4 t0 = pr.a, t1 = pr.b;
        v[0] = lrelu(v[0] + t0.x); v[1] = lrelu(v[1] + t0.y); v[2] = lrelu(v[2] + t0.z); v[3] = lrelu(v[3] + t0.w);
        v[4] = lrelu(v[4] + t1.x); v[5] = lrelu(v[5] + t1.y); v[6] = lrelu(v[6] + t1.z); v[7] = lrelu(v[7] + t1.w);
        uint4 hi, lo;
        split8<true>(v, hi, lo);
        bf16* p = out + (long)b * batch_stride + (long)row * ld + col;
        *reinterpret_cast<uint4*>(p) = hi;
        *reinterpret_cast<uint4*>(p + plane_stride) = lo;
    }
    __device__ void operator()(int b, int, int row, int col0, float* v, const GemmShape&) const {
        const float4* c = reinterpret_cast<const float4*>(cp + (long)(row % B) * cp_ld + cp_off + (long)b * cp_bstride + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 t = __ldg(c + j);
            v[4 * j] = lrelu(v[4 * j] + t.x); v[4 * j + 1] = lrelu(v[4 * j + 1] + t.y);
            v[4 * j + 2] = lrelu(v[4 * j + 2] + t.z); v[4 * j + 3] = lrelu(v[4 * j + 3] + t.w);
        }
        bf16* p = out + (long)b * batch_stride + (long)row * ld + col0;
        store_planes32<true>(p, p + plane_stride, v);
    }
    __device__ void elem(int, int, int, int, float, const GemmShape&) const {}
};
__device__ __forceinline__ float fast_tanh(float x) {   // 1 - 2/(e^{2x}+1); abs error ~1e-7, saturates cleanly
    const float e = __expf(2.f * x);
    return 1.f - __fdividef(2.f, e + 1.f);
}
struct EpiOutHead {   // st[batch][row][col] = batch == 0 ? tanh(acc + b2) : acc + b2, col < D
    static constexpr bool kDirect = false, kStaged = true, kRmw = false;
    float* out; int D; long batch_stride; const float* bias; long bias_bstride;
    __device__ void operator()(int, int, int, int, float*, const GemmShape&) const {}
    __device__ void elem(int b, int, int row, int col, float v, const GemmShape&) const {
        if (col < D) {
            float t = v + __ldg(bias + (long)b * bias_bstride + col);
            if (b == 0) t = fast_tanh(t);
            out[(long)b * batch_stride + (long)row * D + col] = t;
        }
    }
};
struct EpiActGradPlanes {   // dh = acc * lrelu'(act) -> planes.  (dcp = sum over the hypotheses of an image: dcp_from_row_planes_kernel)
    static constexpr bool kDirect = true, kStaged = false, kRmw = false, kTile8 = true;
    bf16* out; const bf16* act_hi; long ld; long plane_stride; long batch_stride; long act_plane_stride; long act_batch_stride;
    struct Pre { uint4 a; };                             // hi plane of the 8 saved activations (their signs)
    __device__ void pre8(int b, int row, int col, Pre& p) const {
        p.a = __ldg(reinterpret_cast<const uint4*>(act_hi + (long)b * act_batch_stride + (long)row * ld + col));
    }
    typedef const bf16* RowRef;
    __device__ RowRef pre_row(int b, int row) const { return act_hi + (long)b * act_batch_stride + (long)row * ld; }
    __device__ void pre8r(RowRef r, int col, Pre& p) const { p.a = __ldg(reinterpret_cast<const uint4*>(r + col)); }
    __device__ void tile8(int b, int, int row, int col, float* v, const Pre& pr, const GemmShape&) const {   // 8 columns of one row (see tc_gemm.cuh)
        const uint4 t = pr.a;
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {   // half > 0  <=>  sign bit clear and magnitude non-zero (the hi plane carries the activation's sign)
            const uint32_t e0 = w[i] & 0xFFFFu, e1 = w[i] >> 16;
            v[2 * i] *= ((e0 & 0x8000u) == 0 && (e0 & 0x7FFFu) != 0) ? 1.f : kLeakySlope;
            v[2 * i + 1] *= ((e1 & 0x8000u) == 0 && (e1 & 0x7FFFu) != 0) ? 1.f : kLeakySlope;
        }
        uint4 hi, lo;
        split8<false>(v, hi, lo);
        const long off = (long)b * batch_stride + (long)row * ld + col;
        *reinterpret_cast<uint4*>(out + off) = hi;
        *reinterpret_cast<uint4*>(out + off + plane_stride) = lo;
    }
    __device__ void operator()(int b, int, int row, int col0, float* v, const GemmShape&) const {
        const uint4* a = reinterpret_cast<const uint4*>(act_hi + (long)b * act_batch_stride + (long)row * ld + col0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint4 t = a[j];
            const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                // half > 0  <=>  sign bit clear and magnitude non-zero (the hi plane carries the activation's sign)
                const uint32_t e0 = w[i] & 0xFFFFu, e1 = w[i] >> 16;
                v[8 * j + 2 * i] *= ((e0 & 0x8000u) == 0 && (e0 & 0x7FFFu) != 0) ? 1.f : kLeakySlope;
                v[8 * j + 2 * i + 1] *= ((e1 & 0x8000u) == 0 && (e1 & 0x7FFFu) != 0) ? 1.f : kLeakySlope;
            }
        }
        const long off = (long)b * batch_stride + (long)row * ld + col0;
        store_planes32<false>(out + off, out + off + plane_stride, v);
    }
    __device__ void elem(int, int, int, int, float, const GemmShape&) const {}
};
// dcp[b][cp_off + net * cp_bstride + f] += sum over the rows r = b, b + B, ... of (hi + lo)[net][r][f]   (reference flows.py:107-109
// differentiated: the conditioning of an image is shared by its hypotheses).  dh: bfloat16 planes [2 nets][2 planes][R][H].
// One thread owns 8 consecutive features of one (image, net): 16-byte plane loads, no atomics - each output is written by one thread.
// (The GEMM epilogue used to add every element into dcp with an atomic: 16.7 M atomics per launch at 32,768 rows, half its run time.)
__global__ void __launch_bounds__(256) dcp_from_row_planes_kernel(const bf16* __restrict__ dh, int R, int B, int H, long plane_stride, long batch_stride,
                                                                   float* __restrict__ dcp, long cp_ld, long cp_off, long cp_bstride) {
    const int per_img = H >> 3;                                   // threads per (image, net)
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = (int)(t / per_img), f = (int)(t % per_img) * 8, net = blockIdx.y;
    if (b >= B) return;
    const uint16_t* hi = reinterpret_cast<const uint16_t*>(dh) + (long)net * batch_stride + f;
    const uint16_t* lo = hi + plane_stride;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (long r = b; r < R; r += B) {
        const uint4 h4 = *reinterpret_cast<const uint4*>(hi + r * H), l4 = *reinterpret_cast<const uint4*>(lo + r * H);
        const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
        for (int k = 0; k < 8; ++k)
            acc[k] += from16<false>((uint16_t)(hw[k >> 1] >> ((k & 1) * 16))) + from16<false>((uint16_t)(lw[k >> 1] >> ((k & 1) * 16)));
    }
    float4* o = reinterpret_cast<float4*>(dcp + (long)b * cp_ld + cp_off + (long)net * cp_bstride + f);
    float4 a = o[0], c = o[1];
    a.x += acc[0]; a.y += acc[1]; a.z += acc[2]; a.w += acc[3];
    c.x += acc[4]; c.y += acc[5]; c.z += acc[6]; c.w += acc[7];
    o[0] = a; o[1] = c;
}
struct EpiMaskAtomicAdd {   // gx[row][col] += mask[col] * acc, col < D
    static constexpr bool kDirect = false, kStaged = true, kRmw = false;
    float* out; int D; const float* mask;
    __device__ void operator()(int, int, int, int, float*, const GemmShape&) const {}
    __device__ void elem(int, int, int row, int col, float v, const GemmShape&) const {
        if (col < D) { const float w = __ldg(mask + col); if (w != 0.f) atomicAdd(out + (long)row * D + col, w * v); }
    }
};
struct EpiWgrad {   // dW[batch][row][col] += acc, col < ncols  (staged: lanes along the columns; loads batched before stores)
    static constexpr bool kDirect = false, kStaged = true, kRmw = true;
    float* dW; long ld; long batch_stride; int ncols; int atomic; int overwrite = 0;
    __device__ void operator()(int, int, int, int, float*, const GemmShape&) const {}
    __device__ void elem(int, int, int, int, float, const GemmShape&) const {}
    __device__ float* rmw_ptr(int b, int row0, int col, long& stride) const {
        stride = ld;
        return col < ncols ? dW + (long)b * batch_stride + (long)row0 * ld + col : nullptr;
    }
};
struct EpiWgradVec {   // the same for unsplit contractions into 16-byte aligned rows (ld % 4 == 0, ncols % 8 == 0): 8 columns of a row per
    // lane, 128 contiguous bytes per row and instruction - the short contractions (K = 64 images / rows) are all epilogue
    static constexpr bool kDirect = false, kStaged = false, kRmw = false, kTile8 = true;
    float* dW; long ld; long batch_stride; int overwrite;
    struct Pre { float4 a, b; };
    __device__ void operator()(int, int, int, int, float*, const GemmShape&) const {}
    __device__ void elem(int, int, int, int, float, const GemmShape&) const {}
    __device__ void pre8(int b, int row, int col, Pre& p) const {
        if (overwrite) return;
        const float4* q = reinterpret_cast<const float4*>(dW + (long)b * batch_stride + (long)row * ld + col);
        p.a = q[0]; p.b = q[1];
    }
    __device__ void tile8(int b, int, int row, int col, float* v, const Pre& p, const GemmShape&) const {
        float4* q = reinterpret_cast<float4*>(dW + (long)b * batch_stride + (long)row * ld + col);
        if (overwrite) { q[0] = make_float4(v[0], v[1], v[2], v[3]); q[1] = make_float4(v[4], v[5], v[6], v[7]); }
        else {
            q[0] = make_float4(p.a.x + v[0], p.a.y + v[1], p.a.z + v[2], p.a.w + v[3]);
            q[1] = make_float4(p.b.x + v[4], p.b.y + v[5], p.b.z + v[6], p.b.w + v[7]);
        }
    }
};
struct EpiWgradT {   // element (row, col) goes to dW[batch][col][row]: lanes = rows are already contiguous in memory
    static constexpr bool kDirect = true, kStaged = false, kRmw = false;
    float* dW; long ld; long batch_stride; int ncols; int atomic; int overwrite = 0;
    __device__ void operator()(int b, int, int row, int col0, float* v, const GemmShape&) const {
        float* base = dW + (long)b * batch_stride + row;
        if (overwrite && !atomic) {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (col0 + j < ncols) base[(long)(col0 + j) * ld] = v[j];
            return;
        }
        if (atomic) {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (col0 + j < ncols) atomicAdd(base + (long)(col0 + j) * ld, v[j]);
            return;
        }
        float old[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) old[j] = (col0 + j < ncols) ? base[(long)(col0 + j) * ld] : 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) if (col0 + j < ncols) base[(long)(col0 + j) * ld] = old[j] + v[j];
    }
    __device__ void elem(int, int, int, int, float, const GemmShape&) const {}
};
struct EpiCondFwd {   // cp[row][idx*H + col] = acc + Cb[idx][col] + b_j[col]
    static constexpr bool kDirect = true, kStaged = false, kRmw = false;
    float* cp; long cp_ld; int H; const float* params; size_t cb_base, cb_stride, blk, ob0, ob1;
    __device__ void operator()(int idx, int, int row, int col0, float* v, const GemmShape&) const {
        const float* cb = params + cb_base + (size_t)idx * cb_stride + col0;
        const float* bj = params + (size_t)(idx >> 1) * blk + ((idx & 1) ? ob1 : ob0) + col0;
        float4* o = reinterpret_cast<float4*>(cp + (long)row * cp_ld + (long)idx * H + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            o[j] = make_float4(v[4 * j] + __ldg(cb + 4 * j) + __ldg(bj + 4 * j), v[4 * j + 1] + __ldg(cb + 4 * j + 1) + __ldg(bj + 4 * j + 1),
                               v[4 * j + 2] + __ldg(cb + 4 * j + 2) + __ldg(bj + 4 * j + 2), v[4 * j + 3] + __ldg(cb + 4 * j + 3) + __ldg(bj + 4 * j + 3));
    }
    __device__ void elem(int, int, int, int, float, const GemmShape&) const {}
};
struct EpiAtomicRows {   // C[row][col] += acc (every batch lands on the same output)
    static constexpr bool kDirect = false, kStaged = true, kRmw = false;
    float* C; long ld; int ncols;
    __device__ void operator()(int, int, int, int, float*, const GemmShape&) const {}
    __device__ void elem(int, int, int row, int col, float v, const GemmShape&) const {
        if (col < ncols) atomicAdd(C + (long)row * ld + col, v);
    }
};

// ---- elementwise kernels --------------------------------------------------------------------------
template <bool F16>
__device__ __forceinline__ void put_planes(bf16* hi, long plane_stride, float v) {
    uint16_t* p = reinterpret_cast<uint16_t*>(hi);
    const uint16_t h = to16<F16>(v);
    p[0] = h;
    p[plane_stride] = to16<F16>(v - from16<F16>(h));
}

// coupling forward (see flow.cu) + masked split planes of the result for the next layer's first GEMM
__global__ void coupling_fwd_kernel(const float* __restrict__ x, const float* __restrict__ st, const float* __restrict__ mask,
                                    const float* __restrict__ next_mask, int R, int D, int direction, float* __restrict__ y,
                                    float* __restrict__ logdet, bf16* __restrict__ ym) {
    pdl_launch_dependents();
    pdl_wait();
    const int r = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    float ssum = 0.f;
    for (int d = lane; d < kDp; d += 32) {
        float out = 0.f;
        if (d < D) {
            const float xv = x[(long)r * D + d];
            out = xv;
            if (mask[d] == 0.f) {
                const float s = st[(long)r * D + d], t = st[(long)R * D + (long)r * D + d];
                out = direction == 0 ? fmaf(xv, expf(s), t) : (xv - t) * expf(-s);
                ssum += s;
            }
            y[(long)r * D + d] = out;
        }
        if (ym) put_planes<true>(ym + (long)r * kDp + d, (long)R * kDp, (d < D && next_mask) ? out * next_mask[d] : 0.f);
    }
    if (logdet) {
        ssum = warp_sum(ssum);
        if (lane == 0) logdet[r] += direction == 0 ? ssum : -ssum;
    }
}

// coupling backward (see flow.cu) + split planes of the head gradients [2 nets][2 planes][R][64] + the bias
// gradient db2 += colsum(dpre).  Block = 4 row slots x 64 columns, kCbRows rows per block (few rows per thread: the loop
// is a chain of dependent memory round trips).
constexpr int kCbRows = 8;
__global__ void __launch_bounds__(256) coupling_bwd_kernel(const float* __restrict__ x, const float* __restrict__ st, const float* __restrict__ mask,
                                    const float* g, const float* __restrict__ gl, float gl_scale, int R, int D, int direction,
                                    bf16* __restrict__ dprep, float* gx, float* __restrict__ db2, long db2_stride) {
    __shared__ float red[2][4][kDp];
    pdl_launch_dependents();
    pdl_wait();
    const int d = threadIdx.x & (kDp - 1), slot = threadIdx.x >> 6;
    const long ps = (long)R * kDp;
    const float m = d < D ? mask[d] : 1.f;
    float sds = 0.f, sdt = 0.f;
    for (int rr = slot; rr < kCbRows; rr += 4) {
        const int r = blockIdx.x * kCbRows + rr;
        if (r >= R) break;
        float ds = 0.f, dt = 0.f;
        if (d < D) {
            const long k = (long)r * D + d;
            const float gv = g[k];
            float dx = gv;
            if (m == 0.f) {
                const float s = st[k], t = st[(long)R * D + k], xv = x[k];
                const float glv = gl ? gl_scale * gl[r] : 0.f;
                if (direction == 0) { const float e = expf(s); dx = gv * e; dt = gv; ds = gv * xv * e + glv; }
                else { const float e = expf(-s); dx = gv * e; dt = -gv * e; ds = -gv * (xv - t) * e - glv; }
                ds *= (1.f - s * s);
            }
            gx[k] = dx;
        }
        const long i = (long)r * kDp + d;
        put_planes<false>(dprep + i, ps, ds);
        put_planes<false>(dprep + 2 * ps + i, ps, dt);
        sds += ds; sdt += dt;
    }
    red[0][slot][d] = sds; red[1][slot][d] = sdt;
    __syncthreads();
    if (slot == 0 && d < D && m == 0.f) {
        atomicAdd(db2 + d, red[0][0][d] + red[0][1][d] + red[0][2][d] + red[0][3][d]);
        atomicAdd(db2 + db2_stride + d, red[1][0][d] + red[1][1][d] + red[1][2][d] + red[1][3][d]);
    }
}

// half planes -> bfloat16 planes of the same values (for the backward GEMMs): n elements per plane (a multiple of 8), batches of
// [2 planes][n]; 8 values per thread, 16-byte loads and stores
__global__ void __launch_bounds__(256) replane_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, long n, int batches) {
    pdl_launch_dependents();
    pdl_wait();
    const long i8 = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i8 >= n * batches) return;
    const long b = i8 / n, k = i8 % n;
    const uint16_t* s16 = reinterpret_cast<const uint16_t*>(src) + b * 2 * n + k;
    uint16_t* d16 = reinterpret_cast<uint16_t*>(dst) + b * 2 * n + k;
    const uint4 h4 = *reinterpret_cast<const uint4*>(s16), l4 = *reinterpret_cast<const uint4*>(s16 + n);
    const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
    uint32_t oh[4], ol[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hw[q])), c = __half22float2(*reinterpret_cast<const __half2*>(&lw[q]));
        const float v0 = a.x + c.x, v1 = a.y + c.y;
        const __nv_bfloat162 hh = __floats2bfloat162_rn(v0, v1);
        const float2 hf = __bfloat1622float2(hh);
        const __nv_bfloat162 ll = __floats2bfloat162_rn(v0 - hf.x, v1 - hf.y);
        oh[q] = *reinterpret_cast<const uint32_t*>(&hh);
        ol[q] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    *reinterpret_cast<uint4*>(d16) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
    *reinterpret_cast<uint4*>(d16 + n) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
}

// grid (cp_ld / 256, 8): block row y sums images y, y + 8, ... and adds its share (the slots accumulate anyway)
__global__ void cond_bias_grad_kernel(const float* __restrict__ dcp, int B, long cp_ld, int H, float* __restrict__ dparams,
                                      size_t cb_base, size_t cb_stride, size_t blk, size_t ob0, size_t ob1, long n_begin, long n_end) {
    const long n = n_begin + (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_end) return;
    float acc = 0.f;
    for (int b = blockIdx.y; b < B; b += gridDim.y) acc += dcp[(long)b * cp_ld + n];
    const int idx = (int)(n / H), h = (int)(n % H);
    atomicAdd(dparams + cb_base + (size_t)idx * cb_stride + h, acc);
    atomicAdd(dparams + (size_t)(idx >> 1) * blk + ((idx & 1) ? ob1 : ob0) + h, acc);
}

static int cuda_ok(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return MHE_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return MHE_ERR_CUDA;
}

static PlaneTensor pt(const bf16* base, int cols, int rows, long pitch, long plane_stride, int batches, long batch_stride) {
    PlaneTensor t;
    t.base = base; t.cols = cols; t.rows = rows; t.planes = 2; t.batches = batches;
    t.row_pitch = pitch; t.plane_stride = plane_stride; t.batch_stride = batch_stride;
    return t;
}

static bool bn256_enabled() {
    static const bool on = [] { const char* e = getenv("MHE_TC_BN256"); return e ? atoi(e) != 0 : true; }();
    return on;
}
// BN = 64 while the grid would not fill the chip with 128-wide tiles.  F16: both operands are half planes (forward GEMMs) or both bfloat16 planes (backward GEMMs).
template <bool A_MN, bool B_MN, bool F16, class Epi>
static int gemm(const PlaneTensor& A, const PlaneTensor& B, GemmShape g, const Epi& e, cudaStream_t s, const char* what,
                const PlaneTensor* A2 = nullptr, const PlaneTensor* B2 = nullptr) {
    const long ctas128 = (long)cdiv(g.M, BM) * cdiv(g.N, 128) * g.batches * g.ksplit;
    if (A2) {   // K-concatenated second operand pair (GemmShape::kb2): persistent kernel, wide tiles
        if (bn256_enabled() && g.N % 256 == 0 && ctas128 >= 2 * 148) return launch_tc_gemm<256, A_MN, B_MN, 3, F16>(A, B, g, e, s, what, A2, B2);
        return launch_tc_gemm<128, A_MN, B_MN, 3, F16>(A, B, g, e, s, what, A2, B2);
    }
    // 128 x 256 tiles for the long contractions: the 128 x 128 tile streams 64 KB of operand planes per 64-deep k-block, more than an
    // SM ingests from L2 in the 768 tensor cycles the block's three products take; the wider tile moves 25 % fewer bytes per flop
    const bool bn256 = bn256_enabled();
    if (bn256 && g.N % 256 == 0 && g.K >= 256 && ctas128 >= 2 * 148) return launch_tc_gemm<256, A_MN, B_MN, 3, F16>(A, B, g, e, s, what);
    if (g.N > 64 && ctas128 >= 148) return launch_tc_gemm<128, A_MN, B_MN, 3, F16>(A, B, g, e, s, what);
    return launch_tc_gemm<64, A_MN, B_MN, 3, F16>(A, B, g, e, s, what);
}

// zero the slots of a gradient buffer that ACCUMULATE on the tensor-core path (the biases): with mhe_flow_set_async bit 1 the weight
// slots are stored, not accumulated, so they need no zeroing
__global__ void zero_bias_grads_kernel(float* __restrict__ d, int nblk, size_t blk, size_t ob0, size_t ob1, size_t ob2, int H, int D,
                                       size_t cb_base, size_t cb_stride, int ncb) {
    const int b = blockIdx.x;
    if (b < nblk) {
        float* p = d + (size_t)b * blk;
        for (int i = threadIdx.x; i < H; i += blockDim.x) { p[ob0 + i] = 0.f; p[ob1 + i] = 0.f; }
        for (int i = threadIdx.x; i < D; i += blockDim.x) p[ob2 + i] = 0.f;
    } else if (b - nblk < ncb) {
        float* p = d + cb_base + (size_t)(b - nblk) * cb_stride;
        for (int i = threadIdx.x; i < H; i += blockDim.x) p[i] = 0.f;
    }
}
int zero_bias_grads(const FlowLayout& L, float* dparams, cudaStream_t stream) {
    zero_bias_grads_kernel<<<L.L * 2 + L.L * 4, 256, 0, stream>>>(dparams, L.L * 2, L.blk, L.ob0, L.ob1, L.ob2, L.H, L.D, L.cb_base, L.cb_stride, L.L * 4);
    return check_launch("zero bias grads");
}

static int g_grads_zero = 0, g_dfeat_zero = 0, g_skip_cond_wgrad = 0;
void set_skip_cond_wgrad(int on) { g_skip_cond_wgrad = on; }
void set_grads_are_zero(int on) { g_grads_zero = on; }
bool grads_are_zero() { return g_grads_zero != 0; }
void set_dfeat_is_zero(int on) { g_dfeat_zero = on; }

int wgrad_kmajor(const PlaneTensor& A, const PlaneTensor& B, GemmShape g, float* dW, long ld, long batch_stride, int ncols, int transposed,
                 cudaStream_t stream, const char* what) {
    if (transposed) {
        EpiWgradT e{dW, ld, batch_stride, ncols, g.ksplit > 1, grads_are_zero() && g.ksplit == 1};
        return gemm<false, false, false>(A, B, g, e, stream, what);
    }
    if (g.ksplit == 1 && ld % 4 == 0 && ncols % 8 == 0 && g.N % 8 == 0 && ((uintptr_t)dW & 15) == 0 && batch_stride % 4 == 0) {
        EpiWgradVec e{dW, ld, batch_stride, grads_are_zero() ? 1 : 0};
        return gemm<false, false, false>(A, B, g, e, stream, what);
    }
    EpiWgrad e{dW, ld, batch_stride, ncols, g.ksplit > 1, grads_are_zero() && g.ksplit == 1};
    return gemm<false, false, false>(A, B, g, e, stream, what);
}

// ---- side stream for the weight-gradient GEMMs ---------------------------------------------------------
// dW GEMMs are off the critical path of the backward (nothing downstream in the pass reads them), so they are
// forked onto an internal stream, layer by layer, and joined at the end of the pass.  Fork/join uses events, which
// also makes the branches parallel nodes when the caller captures the pass into a CUDA graph.
constexpr int kMaxDevices = 16;
struct Aux {
    cudaStream_t stream[2] = {nullptr, nullptr};   // [0]: the HxH gradient, [1]: the two thin ones
    cudaEvent_t ready[64] = {}, done[2][64] = {};
    bool ok = false;
};
static Aux& aux_ctx() {   // one context per device: streams and events belong to the device that was current when they were created
    static Aux ctx[kMaxDevices];
    static bool tried_dev[kMaxDevices] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
    Aux& a = ctx[dev];
    bool& tried = tried_dev[dev];
    if (!tried) {
        tried = true;
        bool ok = cudaStreamCreateWithFlags(&a.stream[0], cudaStreamNonBlocking) == cudaSuccess &&
                  cudaStreamCreateWithFlags(&a.stream[1], cudaStreamNonBlocking) == cudaSuccess;
        for (int i = 0; i < 64 && ok; ++i)
            ok = cudaEventCreateWithFlags(&a.ready[i], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&a.done[0][i], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&a.done[1][i], cudaEventDisableTiming) == cudaSuccess;
        a.ok = ok;
    }
    return a;
}

// ---- packed weights -----------------------------------------------------------------------------------
int pack_weights(const FlowLayout& L, const float* params, void* packed, int which, cudaStream_t stream) {
    Packed P(L, (bf16*)packed);
    // W0 [H][D] -> [H][64]; W1 [H][H]; W2 [D][H] -> [64][H]; Cw [H][C]
    // half planes (forward GEMMs): bit 0 = all of them, bit 2 = only the conditioning weights, bit 3 = only the coupling weights
    if (which & (1 | 4))
        MHE_TRY(split_planes(params + L.cw_base, L.C, (long)L.cw_stride, L.H, L.C, nullptr, P.cw, L.H, L.C, 2, L.L * 4, true, stream));
    if (which & (1 | 8)) {
        MHE_TRY(split_planes(params + L.oW0, L.D, (long)L.blk, L.H, L.D, nullptr, P.w0, L.H, kDp, 2, L.L * 2, true, stream));
        MHE_TRY(split_planes(params + L.oW1, L.H, (long)L.blk, L.H, L.H, nullptr, P.w1, L.H, L.H, 2, L.L * 2, true, stream));
        MHE_TRY(split_planes(params + L.oW2, L.H, (long)L.blk, L.D, L.H, nullptr, P.w2, kDp, L.H, 2, L.L * 2, true, stream));
    }
    // bfloat16 planes (backward GEMMs): bit 1 = all of them, bit 4 = only the coupling weights, bit 5 = only the conditioning weights
    if (which & (2 | 16)) {
        MHE_TRY(split_planes(params + L.oW0, L.D, (long)L.blk, L.H, L.D, nullptr, P.w0b, L.H, kDp, 2, L.L * 2, false, stream));
        MHE_TRY(split_planes(params + L.oW1, L.H, (long)L.blk, L.H, L.H, nullptr, P.w1b, L.H, L.H, 2, L.L * 2, false, stream));
        MHE_TRY(split_planes(params + L.oW2, L.H, (long)L.blk, L.D, L.H, nullptr, P.w2b, kDp, L.H, 2, L.L * 2, false, stream));
    }
    if (which & (2 | 32))
        MHE_TRY(split_planes(params + L.cw_base, L.C, (long)L.cw_stride, L.H, L.C, nullptr, P.cwb, L.H, L.C, 2, L.L * 4, false, stream));
    return MHE_OK;
}

// ---- conditioning ---------------------------------------------------------------------------------------
int cond_fwd(const FlowLayout& L, const float* params, const void* packed, const float* feat, int B, float* cp, void* ws_, cudaStream_t stream) {
    if (cond_direct_supported(L, B)) return cond_fwd_direct(L, params, feat, B, cp, ws_, stream);
    Packed P(L, (bf16*)packed);
    bf16* featp = (bf16*)ws_;   // [2][B][C]
    MHE_TRY(split_planes(feat, L.C, 0, B, L.C, nullptr, featp, B, L.C, 2, 1, true, stream));
    PlaneTensor A = pt(featp, L.C, B, L.C, (long)B * L.C, 1, 0);
    PlaneTensor Bt = pt(P.cw, L.C, L.H, L.C, (long)L.H * L.C, L.L * 4, (long)2 * L.H * L.C);
    GemmShape g{B, L.H, L.C, L.L * 4, 1, 0, 1};
    EpiCondFwd e{cp, (long)L.L * 4 * L.H, L.H, params, L.cb_base, L.cb_stride, L.blk, L.ob0, L.ob1};
    return gemm<false, false, true>(A, Bt, g, e, stream, "tc cond fwd");
}

int cond_bwd(const FlowLayout& L, const float* params, const void* packed, const float* feat, const float* dcp, int B,
             float* dparams, float* dfeat, void* ws_, cudaStream_t stream) {
    Packed P(L, (bf16*)packed);
    const long cp_ld = (long)L.L * 4 * L.H;
    bf16* featp = (bf16*)ws_;                                   // [2][B][C]
    bf16* dcpp = featp + (((size_t)2 * B * L.C + 511) / 512) * 512;   // [2][B][cp_ld]
    MHE_TRY(split_planes(feat, L.C, 0, B, L.C, nullptr, featp, B, L.C, 2, 1, false, stream));   // bfloat16: partner of dcp planes
    MHE_TRY(split_planes(dcp, cp_ld, 0, B, (int)cp_ld, nullptr, dcpp, B, (int)cp_ld, 2, 1, false, stream));
    // The weight / bias gradients and the feature gradient only share their inputs: fork the former onto an internal stream.
    Aux& aux = aux_ctx();
    const bool fork = aux.ok && dfeat != nullptr;
    cudaStream_t wstream = fork ? aux.stream[0] : stream;
    if (fork) {
        MHE_TRY(cuda_ok(cudaEventRecord(aux.ready[0], stream), "fork cond wgrad"));
        MHE_TRY(cuda_ok(cudaStreamWaitEvent(wstream, aux.ready[0], 0), "fork cond wgrad"));
    }
    {   // bias gradients: column sums of dcp, independent of both GEMMs -> their own stream, enqueued first (a small kernel behind a
        // many-wave GEMM only starts when that GEMM's last wave has been placed)
        cudaStream_t bstream = fork ? aux.stream[1] : stream;
        if (fork) MHE_TRY(cuda_ok(cudaStreamWaitEvent(bstream, aux.ready[0], 0), "fork cond bias grad"));
        cond_bias_grad_kernel<<<dim3(cdiv((int)cp_ld, 256), 8), 256, 0, bstream>>>(dcp, B, cp_ld, L.H, dparams, L.cb_base, L.cb_stride, L.blk, L.ob0, L.ob1, 0, cp_ld);
        MHE_TRY(check_launch("cond bias grad"));
        if (fork) MHE_TRY(cuda_ok(cudaEventRecord(aux.done[1][0], bstream), "join cond bias grad"));
    }
    if (!g_skip_cond_wgrad) {   // dCw[idx] [H][C] += dcp[:, idx, :]^T feat      (A MN-major: cols = h, rows = b; B MN-major: cols = c, rows = b)
        PlaneTensor A = pt(dcpp, L.H, B, cp_ld, (long)B * cp_ld, L.L * 4, L.H);
        PlaneTensor Bt = pt(featp, L.C, B, L.C, (long)B * L.C, 1, 0);
        GemmShape g{L.H, L.C, B, L.L * 4, 1, 1, 0};
        if (L.C % 8 == 0 && L.cw_stride % 4 == 0 && L.cw_base % 4 == 0) {
            EpiWgradVec e{dparams + L.cw_base, L.C, (long)L.cw_stride, grads_are_zero() ? 1 : 0};
            MHE_TRY((gemm<true, true, false>(A, Bt, g, e, wstream, "tc cond wgrad")));
        } else {
            EpiWgrad e{dparams + L.cw_base, L.C, (long)L.cw_stride, L.C, 0, grads_are_zero()};
            MHE_TRY((gemm<true, true, false>(A, Bt, g, e, wstream, "tc cond wgrad")));
        }
    }
    if (fork) MHE_TRY(cuda_ok(cudaEventRecord(aux.done[0][0], wstream), "join cond wgrad"));
    if (dfeat) {   // dfeat [B][C] = sum_idx dcp[:, idx, :] Cw[idx]   (A K-major over h; B MN-major: cols = c, rows = h)
        // (a memset node between the dcp planes and this GEMM costs a scheduling hop inside a captured graph: the caller may zero
        // dfeat ahead of time instead, mhe_flow_set_async bit 2)
        if (!g_dfeat_zero) MHE_TRY(cuda_ok(cudaMemsetAsync(dfeat, 0, (size_t)B * L.C * sizeof(float), stream), "memset dfeat"));
        PlaneTensor A = pt(dcpp, L.H, B, cp_ld, (long)B * cp_ld, L.L * 4, L.H);
        PlaneTensor Bt = pt(P.cwb, L.C, L.H, L.C, (long)L.H * L.C, L.L * 4, (long)2 * L.H * L.C);
        GemmShape g{B, L.C, L.H, L.L * 4, 1, 1, 1};
        g.kfold = (L.L * 4) % 4 == 0 ? 4 : ((L.L * 4) % 2 == 0 ? 2 : 1);
        if (const char* e = getenv("MHE_DFEAT_KFOLD")) { const int kf = atoi(e); if (kf >= 1 && (L.L * 4) % kf == 0) g.kfold = kf; }   // every batch lands on the same output: contract 4 per CTA, 4x fewer atomics
        EpiAtomicRows e{dfeat, L.C, L.C};
        MHE_TRY((gemm<false, true, false>(A, Bt, g, e, stream, "tc cond dfeat")));
    }
    if (fork) {
        MHE_TRY(cuda_ok(cudaStreamWaitEvent(stream, aux.done[0][0], 0), "join cond wgrad"));
        MHE_TRY(cuda_ok(cudaStreamWaitEvent(stream, aux.done[1][0], 0), "join cond bias grad"));
    }
    return MHE_OK;
}

// The conditioning weight gradient alone, from its two factors: dCw[idx] (+)= dcp[:, idx, :]^T feat over Bt rows.  Data parallelism uses
// it with the factors GATHERED from all ranks (Bt = world x B): the gradient of the 50 MB of conditioning weights has rank <= Bt per
// (layer, net, j), so exchanging dcp (B x L*4*H) and feat (B x C) is far less traffic than all-reducing the dense gradient.
int cond_wgrad(const FlowLayout& L, const float* feat, const float* dcp, int Bt, float* dparams, void* ws_, cudaStream_t stream) {
    const long cp_ld = (long)L.L * 4 * L.H;
    bf16* featp = (bf16*)ws_;
    bf16* dcpp = featp + (((size_t)2 * Bt * L.C + 511) / 512) * 512;
    MHE_TRY(split_planes(feat, L.C, 0, Bt, L.C, nullptr, featp, Bt, L.C, 2, 1, false, stream));
    MHE_TRY(split_planes(dcp, cp_ld, 0, Bt, (int)cp_ld, nullptr, dcpp, Bt, (int)cp_ld, 2, 1, false, stream));
    PlaneTensor A = pt(dcpp, L.H, Bt, cp_ld, (long)Bt * cp_ld, L.L * 4, L.H);
    PlaneTensor Bt_ = pt(featp, L.C, Bt, L.C, (long)Bt * L.C, 1, 0);
    GemmShape g{L.H, L.C, Bt, L.L * 4, 1, 1, 0};
    if (L.C % 8 == 0 && L.cw_stride % 4 == 0 && L.cw_base % 4 == 0) {
        EpiWgradVec e{dparams + L.cw_base, L.C, (long)L.cw_stride, grads_are_zero() ? 1 : 0};
        return gemm<true, true, false>(A, Bt_, g, e, stream, "tc cond wgrad");
    }
    EpiWgrad e{dparams + L.cw_base, L.C, (long)L.cw_stride, L.C, 0, grads_are_zero()};
    return gemm<true, true, false>(A, Bt_, g, e, stream, "tc cond wgrad");
}

// ---- conditioning backward by layer range (the chunked fused pass pipelines it behind each chunk's dcp sums) ----------------
// feat planes (bfloat16: partner of the dcp planes), once per pass
int cond_bwd_feat_planes(const FlowLayout& L, const float* feat, int B, void* ws_, cudaStream_t stream) {
    return split_planes(feat, L.C, 0, B, L.C, nullptr, (bf16*)ws_, B, L.C, 2, 1, false, stream);
}
// dcp planes of layers [l0, l0 + nl): dense [2][B][nl*4*H] at their own offset of the workspace
int cond_bwd_dcp_planes(const FlowLayout& L, const float* dcp, int B, void* ws_, int l0, int nl, cudaStream_t stream) {
    const long cp_ld = (long)L.L * 4 * L.H;
    bf16* dcpp = (bf16*)ws_ + (((size_t)2 * B * L.C + 511) / 512) * 512 + (size_t)2 * B * l0 * 4 * L.H;
    return split_planes(dcp + (size_t)l0 * 4 * L.H, cp_ld, 0, B, nl * 4 * L.H, nullptr, dcpp, B, nl * 4 * L.H, 2, 1, false, stream);
}
// the three independent pieces for layers [l0, l0 + nl), each on the stream given (no fork / join here)
int cond_bwd_layers(const FlowLayout& L, const void* packed, const float* dcp, int B, float* dparams, float* dfeat, void* ws_, int l0, int nl,
                    cudaStream_t s_bias, cudaStream_t s_wgrad, cudaStream_t s_dfeat) {
    Packed P(L, (bf16*)packed);
    const long cp_ld = (long)L.L * 4 * L.H, ld_c = (long)nl * 4 * L.H;
    bf16* featp = (bf16*)ws_;
    bf16* dcpp = featp + (((size_t)2 * B * L.C + 511) / 512) * 512 + (size_t)2 * B * l0 * 4 * L.H;
    const int nb = nl * 4;
    cond_bias_grad_kernel<<<dim3(cdiv((int)ld_c, 256), 8), 256, 0, s_bias>>>(dcp, B, cp_ld, L.H, dparams, L.cb_base, L.cb_stride, L.blk, L.ob0, L.ob1,
                                                                              (long)l0 * 4 * L.H, (long)(l0 + nl) * 4 * L.H);
    MHE_TRY(check_launch("cond bias grad"));
    {   // dCw[idx] [H][C] (+)= dcp[:, idx, :]^T feat
        PlaneTensor A = pt(dcpp, L.H, B, ld_c, (long)B * ld_c, nb, L.H);
        PlaneTensor Bt = pt(featp, L.C, B, L.C, (long)B * L.C, 1, 0);
        GemmShape g{L.H, L.C, B, nb, 1, 1, 0};
        if (L.C % 8 == 0 && L.cw_stride % 4 == 0 && L.cw_base % 4 == 0) {
            EpiWgradVec e{dparams + L.cw_base + (size_t)l0 * 4 * L.cw_stride, L.C, (long)L.cw_stride, grads_are_zero() ? 1 : 0};
            MHE_TRY((gemm<true, true, false>(A, Bt, g, e, s_wgrad, "tc cond wgrad")));
        } else {
            EpiWgrad e{dparams + L.cw_base + (size_t)l0 * 4 * L.cw_stride, L.C, (long)L.cw_stride, L.C, 0, grads_are_zero()};
            MHE_TRY((gemm<true, true, false>(A, Bt, g, e, s_wgrad, "tc cond wgrad")));
        }
    }
    if (dfeat) {   // dfeat [B][C] += sum_idx dcp[:, idx, :] Cw[idx]   (dfeat zeroed by the caller of the pass)
        PlaneTensor A = pt(dcpp, L.H, B, ld_c, (long)B * ld_c, nb, L.H);
        PlaneTensor Bt = pt(P.cwb + (size_t)l0 * 4 * 2 * L.H * L.C, L.C, L.H, L.C, (long)L.H * L.C, nb, (long)2 * L.H * L.C);
        GemmShape g{B, L.C, L.H, nb, 1, 1, 1};
        g.kfold = nb % 4 == 0 ? 4 : (nb % 2 == 0 ? 2 : 1);
        EpiAtomicRows e{dfeat, L.C, L.C};
        MHE_TRY((gemm<false, true, false>(A, Bt, g, e, s_dfeat, "tc cond dfeat")));
    }
    return MHE_OK;
}
bool dfeat_is_zero() { return g_dfeat_zero != 0; }

// ---- coupling layers ------------------------------------------------------------------------------------
struct LayerBufs {   // where one layer's activations live (workspace, or the saved-for-backward block)
    bf16 *xm, *a0, *a1;
    float* st;
};

// One conditioning row per flow row (log_prob / sample with per-row features, reference flows.py:107-109 with cond (R, C)): the
// projections c.j(cond) are NOT materialised - an R x L*4*H fp32 tensor written and read back, 1.6 GB at 16,384 rows - but contracted
// inside the coupling GEMMs: h_j = lrelu([a | cond] [W_j | Cw_j]^T + (b_j + Cb_j)), the second operand pair of the GEMM (GemmShape::kb2).
struct RowCond {
    const bf16* featp;     // half planes of the conditioning rows [2][R][C]
    const float* bias;     // [L*4][H]: Cb[idx] + b_j of the owning net, idx = layer*4 + net*2 + j
};

static int layer_nets_fwd(const FlowLayout& L, const float* params, const Packed& P, const float* cp, int R, int B, int layer,
                          const LayerBufs& bf, cudaStream_t stream, const RowCond* rc = nullptr) {
    const long cp_ld = (long)L.L * 4 * L.H;
    const long RH = (long)R * L.H, RD = (long)R * kDp;
    const long HC = (long)L.H * L.C;
    for (int j = 0; j < 2; ++j) {   // G0: xm [R][64] x W0^T -> a0;  G1: a0 x W1^T -> a1
        PlaneTensor A = j == 0 ? pt(bf.xm, kDp, R, kDp, RD, 1, 0) : pt(bf.a0, L.H, R, L.H, RH, 2, 2 * RH);
        PlaneTensor Bt = j == 0 ? pt(P.w0 + (size_t)layer * 2 * 2 * L.H * kDp, kDp, L.H, kDp, (long)L.H * kDp, 2, (long)2 * L.H * kDp)
                                : pt(P.w1 + (size_t)layer * 2 * 2 * L.H * L.H, L.H, L.H, L.H, (long)L.H * L.H, 2, (long)2 * L.H * L.H);
        GemmShape g{R, L.H, j == 0 ? kDp : L.H, 2, 1, j, 1};
        bf16* out = j == 0 ? bf.a0 : bf.a1;
        if (rc) {
            PlaneTensor A2 = pt(rc->featp, L.C, R, L.C, (long)R * L.C, 1, 0);                                   // shared by both nets
            PlaneTensor B2 = pt(P.cw + (size_t)(layer * 4 + j) * 2 * HC, L.C, L.H, L.C, HC, 2, 4 * HC);       // nets are 2 idx apart
            g.kb2 = L.C / 64; g.a2_batch_mul = 0; g.b2_batch_mul = 1;
            EpiHiddenPlanes e{out, L.H, RH, 2 * RH, rc->bias, 0, (long)(layer * 4 + j) * L.H, (long)2 * L.H, 1};
            MHE_TRY((gemm<false, false, true>(A, Bt, g, e, stream, j == 0 ? "tc flow G0 rowcond" : "tc flow G1 rowcond", &A2, &B2)));
        } else {
            EpiHiddenPlanes e{out, L.H, RH, 2 * RH, cp, cp_ld, (long)(layer * 4 + j) * L.H, (long)2 * L.H, B};
            MHE_TRY((gemm<false, false, true>(A, Bt, g, e, stream, j == 0 ? "tc flow G0" : "tc flow G1")));
        }
    }
    {   // G2: a1 x W2^T + b2 -> st
        PlaneTensor A = pt(bf.a1, L.H, R, L.H, RH, 2, 2 * RH);
        PlaneTensor Bt = pt(P.w2 + (size_t)layer * 2 * 2 * kDp * L.H, L.H, kDp, L.H, (long)kDp * L.H, 2, (long)2 * kDp * L.H);
        GemmShape g{R, kDp, L.H, 2, 1, 1, 1};
        EpiOutHead e{bf.st, L.D, (long)R * L.D, params + L.block(layer, 0) + L.ob2, (long)L.blk};
        MHE_TRY((gemm<false, false, true>(A, Bt, g, e, stream, "tc flow G2")));
    }
    return MHE_OK;
}

static int pass_fwd_impl(const FlowLayout& L, const float* params, const void* packed, const float* mask, const float* cp, const float* in, int R,
                         int B, int direction, float* out, float* logdet, float* saved, void* workspace, cudaStream_t stream, const RowCond* rc) {
    Packed P(L, (bf16*)packed);
    Ws ws(workspace, L, R);
    Saved S(saved, L, R);
    const size_t row_bytes = (size_t)R * L.D * sizeof(float);
    if (logdet) MHE_TRY(cuda_ok(cudaMemsetAsync(logdet, 0, (size_t)R * sizeof(float), stream), "memset logdet"));
    const int first = direction == 0 ? 0 : L.L - 1;
    if (saved) MHE_TRY(cuda_ok(cudaMemcpyAsync(S.x(0), in, row_bytes, cudaMemcpyDeviceToDevice, stream), "save input"));
    MHE_TRY(split_planes(in, L.D, 0, R, L.D, mask + (size_t)first * L.D, saved ? S.xm(0) : ws.xm, R, kDp, 2, 1, true, stream));
    const float* x = saved ? S.x(0) : in;
    for (int step = 0; step < L.L; ++step) {
        const int layer = direction == 0 ? step : L.L - 1 - step;
        const bool last = step == L.L - 1;
        LayerBufs bf = saved ? LayerBufs{S.xm(step), S.a0(step), S.a1(step), S.st(step)} : LayerBufs{ws.xm, ws.a0, ws.a1, ws.st};
        MHE_TRY(layer_nets_fwd(L, params, P, cp, R, B, layer, bf, stream, rc));
        float* y;
        if (saved) y = S.x(step + 1);
        else y = last ? out : ((x == ws.gx) ? ws.dpre : ws.gx);
        bf16* ym = last ? nullptr : (saved ? S.xm(step + 1) : ws.xm);
        const int next_layer = direction == 0 ? layer + 1 : layer - 1;
        MHE_TRY(cuda_ok(launch_chain(coupling_fwd_kernel, dim3(cdiv(R, 8)), dim3(256), 0, stream, x, (const float*)bf.st, mask + (size_t)layer * L.D,
                                     last ? (const float*)nullptr : mask + (size_t)next_layer * L.D, R, L.D, direction, y, logdet, ym), "tc coupling fwd"));
        MHE_TRY(check_launch("tc coupling fwd"));
        x = y;
    }
    if (saved) MHE_TRY(cuda_ok(cudaMemcpyAsync(out, S.x(L.L), row_bytes, cudaMemcpyDeviceToDevice, stream), "copy out"));
    return MHE_OK;
}
int pass_fwd(const FlowLayout& L, const float* params, const void* packed, const float* mask, const float* cp, const float* in, int R, int B,
             int direction, float* out, float* logdet, float* saved, void* workspace, cudaStream_t stream) {
    return pass_fwd_impl(L, params, packed, mask, cp, in, R, B, direction, out, logdet, saved, workspace, stream, nullptr);
}

// ---- per-row conditioning folded into the coupling GEMMs (no saved state: inference / scoring) ----------------
__global__ void cond_bias_kernel(const float* __restrict__ params, float* __restrict__ bias, int H, int n, size_t cb_base, size_t cb_stride,
                                 size_t blk, size_t ob0, size_t ob1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int idx = i / H, col = i - idx * H;
    bias[i] = params[cb_base + (size_t)idx * cb_stride + col] + params[(size_t)(idx >> 1) * blk + ((idx & 1) ? ob1 : ob0) + col];
}
bool rowcond_supported(const FlowLayout& L) { return supported(L) && L.C % 64 == 0 && L.C >= 64; }
static size_t up1k(size_t n) { return (n + 1023) / 1024 * 1024; }
size_t rowcond_ws_bytes(const FlowLayout& L, int R) {
    return up1k(Ws::bytes(L, R)) + up1k((size_t)2 * R * L.C * 2) + up1k((size_t)L.L * 4 * L.H * 4) + 1024;
}
int pass_fwd_rowcond(const FlowLayout& L, const float* params, const void* packed, const float* mask, const float* feat, const float* in, int R,
                     int direction, float* out, float* logdet, void* workspace, cudaStream_t stream) {
    if (R <= 0) return MHE_OK;
    uint8_t* extra = (uint8_t*)workspace + up1k(Ws::bytes(L, R));
    bf16* featp = (bf16*)extra;
    float* bias = (float*)(extra + up1k((size_t)2 * R * L.C * 2));
    MHE_TRY(split_planes(feat, L.C, 0, R, L.C, nullptr, featp, R, L.C, 2, 1, true, stream));
    const int n = L.L * 4 * L.H;
    cond_bias_kernel<<<cdiv(n, 256), 256, 0, stream>>>(params, bias, L.H, n, L.cb_base, L.cb_stride, L.blk, L.ob0, L.ob1);
    MHE_TRY(check_launch("cond bias"));
    RowCond rc{featp, bias};
    return pass_fwd_impl(L, params, packed, mask, nullptr, in, R, 1, direction, out, logdet, nullptr, workspace, stream, &rc);
}

// The backward reads the hidden activations the forward saved (no recomputation on this path).
int pass_bwd(const FlowLayout& L, const float* params, const void* packed, const float* mask, const float* cp, const float* saved, int R, int B,
             int direction, const float* dout, const float* dlogdet, float dlogdet_scale, float* din, float* dparams, float* dcp,
             void* workspace, cudaStream_t stream) {
    Packed P(L, (bf16*)packed);
    Ws ws(workspace, L, R);
    Saved S(const_cast<float*>(saved), L, R);
    const long cp_ld = (long)L.L * 4 * L.H;
    const long RH = (long)R * L.H, RD = (long)R * kDp;
    const float* g = dout;
    // wgrad contracts over the rows: split K once it is long.  8 splits give the H x H gradient 256 tiles of 128 x 128 (both nets), enough
    // for the 128-wide tile: with 4 splits the launcher fell back to 64-wide tiles, whose 48 KB of operands per 384 tensor cycles is more
    // than an SM ingests (measured 151 us per layer at 32,768 rows against 82 us for the same contraction in the forward GEMM)
    const int ks = R >= 8192 ? 8 : 1;
    // the two thin gradients (dW0, dW2: 512 x 64 outputs per net = 8 tiles in all) need a deeper split to occupy the chip: as many
    // splits as divide the k-blocks evenly, up to 16 (128 CTAs)
    int ks_thin = ks;
    if (R >= 2048) { const int kb = cdiv(R, BK); for (int c = 16; c >= 1; c >>= 1) if (kb % c == 0 && kb / c >= 4) { ks_thin = c; break; } }
    Aux& aux = aux_ctx();
    const bool fork = aux.ok && L.L <= 64;
    cudaStream_t wstream = fork ? aux.stream[0] : stream, wstream2 = fork ? aux.stream[1] : stream;
    for (int step = L.L - 1; step >= 0; --step) {
        const int layer = direction == 0 ? step : L.L - 1 - step;
        const int pb = step & 1;
        const float* mrow = mask + (size_t)layer * L.D;
        float* dblk = dparams + L.block(layer, 0);
        float* gx = (step == 0) ? din : ws.gx;
        // this parity's gradient planes were last read by the side-stream GEMMs of step + 2
        if (fork && step + 2 < L.L) {
            MHE_TRY(cuda_ok(cudaStreamWaitEvent(stream, aux.done[0][step + 2], 0), "wait wgrad"));
            MHE_TRY(cuda_ok(cudaStreamWaitEvent(stream, aux.done[1][step + 2], 0), "wait wgrad"));
        }
        MHE_TRY(cuda_ok(launch_chain(coupling_bwd_kernel, dim3(cdiv(R, kCbRows)), dim3(256), 0, stream, (const float*)S.x(step), (const float*)S.st(step), mrow,
                                     g, dlogdet, dlogdet_scale, R, L.D, direction, ws.dprep[pb], gx, dblk + L.ob2, (long)L.blk), "tc coupling bwd"));
        MHE_TRY(check_launch("tc coupling bwd"));
        PlaneTensor dpreK = pt(ws.dprep[pb], kDp, R, kDp, RD, 2, 2 * RD);
        PlaneTensor w0 = pt(P.w0b + (size_t)layer * 2 * 2 * L.H * kDp, kDp, L.H, kDp, (long)L.H * kDp, 2, (long)2 * L.H * kDp);
        PlaneTensor w1 = pt(P.w1b + (size_t)layer * 2 * 2 * L.H * L.H, L.H, L.H, L.H, (long)L.H * L.H, 2, (long)2 * L.H * L.H);
        PlaneTensor w2 = pt(P.w2b + (size_t)layer * 2 * 2 * kDp * L.H, L.H, kDp, L.H, (long)kDp * L.H, 2, (long)2 * kDp * L.H);
        PlaneTensor a0 = pt(ws.a0b, L.H, R, L.H, RH, 2, 2 * RH), a1 = pt(ws.a1b, L.H, R, L.H, RH, 2, 2 * RH);
        PlaneTensor dh0 = pt(ws.dh0[pb], L.H, R, L.H, RH, 2, 2 * RH), dh1 = pt(ws.dh1[pb], L.H, R, L.H, RH, 2, 2 * RH);
        PlaneTensor xm = pt(ws.xmb, kDp, R, kDp, RD, 1, 0);
        {   // dgrad G2: dh1 = (dpre W2) * lrelu'(a1);  W2 planes [64][H] read MN-major (cols = h);  dcp1 += sum_s dh1
            GemmShape s{R, L.H, kDp, 2, 1, 1, 1};
            EpiActGradPlanes e{ws.dh1[pb], S.a1(step), L.H, RH, 2 * RH, RH, 2 * RH};
            MHE_TRY((gemm<false, true, false>(dpreK, w2, s, e, stream, "tc dgrad G2")));
        }
        {   // dgrad G1: dh0 = (dh1 W1) * lrelu'(a0);  dcp0 += sum_s dh0
            GemmShape s{R, L.H, L.H, 2, 1, 1, 1};
            EpiActGradPlanes e{ws.dh0[pb], S.a0(step), L.H, RH, 2 * RH, RH, 2 * RH};
            MHE_TRY((gemm<false, true, false>(dh1, w1, s, e, stream, "tc dgrad G1")));
        }
        if (fork) {
            MHE_TRY(cuda_ok(cudaEventRecord(aux.ready[step], stream), "fork wgrad"));
            MHE_TRY(cuda_ok(cudaStreamWaitEvent(wstream, aux.ready[step], 0), "fork wgrad"));
            MHE_TRY(cuda_ok(cudaStreamWaitEvent(wstream2, aux.ready[step], 0), "fork wgrad"));
        }
        {   // dgrad G0: gx += mask * (dh0 W0), both nets;  W0 planes [H][64] read MN-major (cols = d)
            GemmShape s{R, kDp, L.H, 2, 1, 1, 1};
            EpiMaskAtomicAdd e{gx, L.D, mrow};
            MHE_TRY((gemm<false, true, false>(dh0, w0, s, e, stream, "tc dgrad G0")));
        }
        {   // dcp of this layer: sums over the hypotheses of dh0 (j = 0) and dh1 (j = 1), off the critical path (side stream; the
            // conditioning backward that reads dcp runs after the pass has joined its side streams)
            dim3 grid(cdiv(B * (L.H / 8), 256), 2);
            dcp_from_row_planes_kernel<<<grid, 256, 0, wstream2>>>(ws.dh1[pb], R, B, L.H, RH, 2 * RH, dcp, cp_ld, (long)(layer * 4 + 1) * L.H, (long)2 * L.H);
            MHE_TRY(check_launch("dcp from dh1 planes"));
            dcp_from_row_planes_kernel<<<grid, 256, 0, wstream2>>>(ws.dh0[pb], R, B, L.H, RH, 2 * RH, dcp, cp_ld, (long)(layer * 4 + 0) * L.H, (long)2 * L.H);
            MHE_TRY(check_launch("dcp from dh0 planes"));
        }
        // the saved activations are half planes; the side streams re-plane them to bfloat16 for their GEMMs
        replane_kernel<<<cdiv((int)(2 * RH / 8), 256), 256, 0, wstream>>>(S.a0(step), ws.a0b, RH, 2);
        MHE_TRY(check_launch("replane a0"));
        replane_kernel<<<cdiv((int)(2 * RH / 8), 256), 256, 0, wstream2>>>(S.a1(step), ws.a1b, RH, 2);
        MHE_TRY(check_launch("replane a1"));
        replane_kernel<<<cdiv((int)(RD / 8), 256), 256, 0, wstream2>>>(S.xm(step), ws.xmb, RD, 1);
        MHE_TRY(check_launch("replane xm"));
        {   // dW1 [out][in] += dh1^T a0
            GemmShape s{L.H, L.H, R, 2, ks, 1, 1};
            EpiWgrad e{dblk + L.oW1, L.H, (long)L.blk, L.H, ks > 1};
            MHE_TRY((gemm<true, true, false>(dh1, a0, s, e, wstream, "tc wgrad W1")));
        }
        {   // dW0 [out][d] += dh0^T xm
            GemmShape s{L.H, kDp, R, 2, ks_thin, 1, 0};
            EpiWgrad e{dblk + L.oW0, L.D, (long)L.blk, L.D, ks_thin > 1};
            MHE_TRY((gemm<true, true, false>(dh0, xm, s, e, wstream2, "tc wgrad W0")));
        }
        {   // dW2 [d][h] += dpre^T a1, computed as (a1^T dpre)[h][d] and stored transposed
            GemmShape s{L.H, kDp, R, 2, ks_thin, 1, 1};
            EpiWgradT e{dblk + L.oW2, L.H, (long)L.blk, L.D, ks_thin > 1};
            MHE_TRY((gemm<true, true, false>(a1, dpreK, s, e, wstream2, "tc wgrad W2")));
        }
        if (fork) {
            MHE_TRY(cuda_ok(cudaEventRecord(aux.done[0][step], wstream), "join wgrad"));
            MHE_TRY(cuda_ok(cudaEventRecord(aux.done[1][step], wstream2), "join wgrad"));
        }
        g = gx;
    }
    if (fork) {   // join: everything the side stream did belongs to this pass
        MHE_TRY(cuda_ok(cudaStreamWaitEvent(stream, aux.done[0][0], 0), "join wgrad"));
        MHE_TRY(cuda_ok(cudaStreamWaitEvent(stream, aux.done[1][0], 0), "join wgrad"));
    }
    return MHE_OK;
}

}  // namespace tcflow
}  // namespace mhe
