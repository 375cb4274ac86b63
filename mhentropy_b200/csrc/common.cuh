// Shared host/device helpers for the mhentropy_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <atomic>
#include "../../include/mhentropy_b200.h"

#if defined(__CUDACC__)
#define MHE_HD __host__ __device__ __forceinline__
#else
#define MHE_HD inline
#endif

namespace mhe {

constexpr float kLeakySlope = 0.01f;  // F.leaky_relu default, reference flows.py:117
constexpr int kSegAlign = 64;         // floats; every parameter segment starts on a 256-byte boundary

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

inline int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return MHE_ERR_CUDA;
    }
    return MHE_OK;
}

// Optional timing probe (bench.py's roofline leg): CUDA events around every launch whose label contains
// the configured substring, recorded on the launching stream.
struct ProbeScope {
    bool on;
    cudaStream_t stream;
    ProbeScope(const char* what, cudaStream_t s);
    ~ProbeScope();
};

#define MHE_TRY(expr)                      \
    do {                                   \
        int _st = (expr);                  \
        if (_st != MHE_OK) return _st;     \
    } while (0)

#define MHE_REQUIRE(cond, ...)             \
    do {                                   \
        if (!(cond)) {                     \
            ::mhe::set_error(__VA_ARGS__); \
            return MHE_ERR_INVALID_ARG;    \
        }                                  \
    } while (0)

inline size_t pad_seg(size_t n) { return (n + kSegAlign - 1) / kSegAlign * kSegAlign; }
inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// flat parameter layout (see include/mhentropy_b200.h)
struct FlowLayout {
    int D, H, C, L;
    int max_split;                             // largest per-layer count of transformed / conditioning dims (see mhe_flow_shape)
    size_t oW0, ob0, oW1, ob1, oW2, ob2, blk;  // offsets inside a (layer, net) block
    size_t cw_base, cw_stride, cb_base, cb_stride, total;
    explicit FlowLayout(mhe_flow_shape s) : D(s.dim), H(s.hidden), C(s.cond), L(s.layers), max_split(s.max_split > 0 ? s.max_split : (s.dim + 1) / 2) {
        oW0 = 0;
        ob0 = oW0 + pad_seg((size_t)H * D);
        oW1 = ob0 + pad_seg(H);
        ob1 = oW1 + pad_seg((size_t)H * H);
        oW2 = ob1 + pad_seg(H);
        ob2 = oW2 + pad_seg((size_t)D * H);
        blk = ob2 + pad_seg(D);
        cw_base = (size_t)L * 2 * blk;
        cw_stride = pad_seg((size_t)H * C);
        cb_base = cw_base + (size_t)L * 4 * cw_stride;
        cb_stride = pad_seg(H);
        total = cb_base + (size_t)L * 4 * cb_stride;
    }
    size_t block(int layer, int net) const { return (size_t)(layer * 2 + net) * blk; }
    size_t cw(int layer, int net, int j) const { return cw_base + (size_t)((layer * 2 + net) * 2 + j) * cw_stride; }
    size_t cb(int layer, int net, int j) const { return cb_base + (size_t)((layer * 2 + net) * 2 + j) * cb_stride; }
};

inline bool valid_shape(mhe_flow_shape s) {
    return s.dim >= 2 && s.dim <= 64 && s.hidden >= 1 && s.cond >= 1 && s.layers >= 1 && s.layers <= 64 && s.max_split >= 0 && s.max_split <= s.dim;
}

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------------
// Kernels of the dependent launch chain are launched with programmatic stream serialization: the next kernel's
// CTAs may start while the previous kernel drains, run their private prologue (barrier init, TMEM allocation,
// descriptor prefetch) and then block in pdl_wait() until the previous kernel has completed and flushed.
// Rule: nothing before pdl_wait() reads or writes global memory.  MHE_PDL=0 in the environment disables it.
bool pdl_enabled();

#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

#if defined(__CUDACC__)
__device__ __forceinline__ float lrelu(float v) { return v > 0.f ? v : kLeakySlope * v; }
// derivative of leaky_relu expressed through its OUTPUT a = lrelu(h): a > 0 <=> h > 0 (slope > 0)
__device__ __forceinline__ float lrelu_grad_from_out(float a) { return a > 0.f ? 1.f : kLeakySlope; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

}  // namespace mhe
