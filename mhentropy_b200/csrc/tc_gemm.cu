// Host side of the tensor-core GEMM: TMA tensor-map construction (driver entry point, no libcuda link),
// a small tensor-map cache, the fp32 -> split-bf16 plane conversion and a raw GEMM entry point for tests.
#include <cudaTypedefs.h>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>
#include "tc_gemm.cuh"

#ifndef MHE_TC_PERSISTENT_DEFAULT
#define MHE_TC_PERSISTENT_DEFAULT "1"
#endif

namespace mhe {
namespace tc {

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

int make_tensor_map(CUtensorMap* map, const PlaneTensor& t, int box_rows) {
    auto encode = get_encode();
    if (!encode) { set_error("cuTensorMapEncodeTiled entry point not available"); return MHE_ERR_CUDA; }
    if (((uintptr_t)t.base & 15) || (t.row_pitch % 8) || (t.plane_stride % 8) || (t.batch_stride % 8)) {
        set_error("tensor map: base/strides must be 16-byte aligned");
        return MHE_ERR_INVALID_ARG;
    }
    const long plane_stride = t.plane_stride > 0 ? t.plane_stride : (long)t.rows * t.row_pitch;
    const long batch_stride = t.batch_stride > 0 ? t.batch_stride : plane_stride * t.planes;
    cuuint64_t gdim[4] = {(cuuint64_t)t.cols, (cuuint64_t)t.rows, (cuuint64_t)t.planes, (cuuint64_t)t.batches};
    cuuint64_t gstr[3] = {(cuuint64_t)t.row_pitch * 2, (cuuint64_t)plane_stride * 2, (cuuint64_t)batch_stride * 2};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)t.base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) cols=%d rows=%d planes=%d batches=%d pitch=%ld box_rows=%d", (int)r, t.cols, t.rows,
                  t.planes, t.batches, t.row_pitch, box_rows);
        return MHE_ERR_CUDA;
    }
    return MHE_OK;
}

struct MapKey {
    const void* base; int cols, rows, planes, batches, box_rows; long row_pitch, plane_stride, batch_stride;
    bool operator==(const MapKey& o) const {
        return base == o.base && cols == o.cols && rows == o.rows && planes == o.planes && batches == o.batches && box_rows == o.box_rows &&
               row_pitch == o.row_pitch && plane_stride == o.plane_stride && batch_stride == o.batch_stride;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = std::hash<const void*>()(k.base);
        auto mix = [&](size_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
        mix(k.cols); mix(k.rows); mix(k.planes); mix(k.batches); mix(k.box_rows); mix((size_t)k.row_pitch); mix((size_t)k.plane_stride); mix((size_t)k.batch_stride);
        return h;
    }
};

// fp32 [batches][rows][cols] (row_pitch / batch_stride in elements) as a 3-D map with SWIZZLE_128B boxes of 32 floats x box_rows: the
// staging format of kernels that convert fp32 weights to split planes on the fly (cond_direct.cu)
const CUtensorMap* cached_map_f32(const float* base, int cols, int rows, int batches, long row_pitch, long batch_stride, int box_rows, int* status) {
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap*, MapKeyHash> cache;
    MapKey key{base, cols, rows, -4, batches, box_rows, row_pitch, 0, batch_stride};
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *status = MHE_OK; return it->second; }
    auto encode = get_encode();
    if (!encode) { set_error("cuTensorMapEncodeTiled entry point not available"); *status = MHE_ERR_CUDA; return nullptr; }
    if (((uintptr_t)base & 15) || (row_pitch % 4) || (batch_stride % 4)) { set_error("fp32 tensor map: base/strides must be 16-byte aligned"); *status = MHE_ERR_INVALID_ARG; return nullptr; }
    CUtensorMap* map = new CUtensorMap;
    cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batches};
    cuuint64_t gstr[2] = {(cuuint64_t)row_pitch * 4, (cuuint64_t)batch_stride * 4};
    cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (fp32) failed (%d) cols=%d rows=%d batches=%d", (int)r, cols, rows, batches);
        delete map;
        *status = MHE_ERR_CUDA;
        return nullptr;
    }
    cache.emplace(key, map);
    *status = MHE_OK;
    return map;
}

// MHE_TC_PERSISTENT: unset / "0" = one tile per CTA everywhere; "1" = persistent kernel everywhere; otherwise a comma-separated list of
// launch-label substrings that take the persistent kernel
bool tc_persistent_enabled(const char* what) {
    static std::string spec;
    static bool parsed = false;
    if (!parsed) { parsed = true; if (const char* e = getenv("MHE_TC_PERSISTENT")) spec = e; else spec = MHE_TC_PERSISTENT_DEFAULT; }
    if (spec.empty() || spec == "0") return false;
    if (spec == "1") return true;
    size_t pos = 0;
    while (pos < spec.size()) {
        size_t end = spec.find(',', pos);
        if (end == std::string::npos) end = spec.size();
        const std::string item = spec.substr(pos, end - pos);
        if (!item.empty() && what && strstr(what, item.c_str())) return true;
        pos = end + 1;
    }
    return false;
}

int stages_for(const char* what, int requested, int max_stages) {
    static std::vector<std::pair<std::string, int>> overrides;
    static bool parsed = false;
    if (!parsed) {
        parsed = true;
        if (const char* e = getenv("MHE_TC_STAGES")) {
            std::string s(e);
            size_t pos = 0;
            while (pos < s.size()) {
                size_t end = s.find(',', pos);
                if (end == std::string::npos) end = s.size();
                const std::string item = s.substr(pos, end - pos);
                const size_t eq = item.find('=');
                if (eq != std::string::npos) overrides.emplace_back(item.substr(0, eq), atoi(item.c_str() + eq + 1));
                pos = end + 1;
            }
        }
    }
    int n = requested;
    for (const auto& o : overrides)
        if (what && strstr(what, o.first.c_str())) n = o.second;
    if (n <= 0 || n > max_stages) n = max_stages;
    return n;
}

// Tensor maps depend only on (pointer, geometry); they are pure descriptors, so caching them is safe even
// when a buffer is freed and a new one is later allocated at the same address with the same geometry.
const CUtensorMap* cached_map(const PlaneTensor& t, int box_rows, int* status) {
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap*, MapKeyHash> cache;
    MapKey key{t.base, t.cols, t.rows, t.planes, t.batches, box_rows, t.row_pitch, t.plane_stride, t.batch_stride};
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *status = MHE_OK; return it->second; }
    if (cache.size() > 65536) {   // unbounded pointer churn: start over
        for (auto& kv : cache) free(kv.second);
        cache.clear();
    }
    CUtensorMap* m = nullptr;
    if (posix_memalign((void**)&m, 64, sizeof(CUtensorMap)) != 0) { set_error("tensor map alloc failed"); *status = MHE_ERR_CUDA; return nullptr; }
    *status = make_tensor_map(m, t, box_rows);
    if (*status != MHE_OK) { free(m); return nullptr; }
    cache.emplace(key, m);
    return m;
}

// fp32 [batches][rows][cols] (src_ld pitch) -> 16-bit planes [batches][planes][rows_p][cols_p], zero padded,
// optional per-column multiplier (the coupling mask).
template <bool F16>
__global__ void split_planes_kernel(const float* __restrict__ src, long src_ld, long src_batch, int rows, int cols, const float* __restrict__ colscale,
                                    uint16_t* __restrict__ dst, int rows_p, int cols_p, int planes) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long per = (long)rows_p * cols_p;
    if (idx >= per) return;
    const int b = blockIdx.y;
    const int r = (int)(idx / cols_p), c = (int)(idx % cols_p);
    float v = 0.f;
    if (r < rows && c < cols) {
        v = src[(long)b * src_batch + (long)r * src_ld + c];
        if (colscale) v *= colscale[c];
    }
    const uint16_t hi = to16<F16>(v);
    uint16_t* d = dst + (long)b * planes * per + idx;
    d[0] = hi;
    if (planes > 1) d[per] = to16<F16>(v - from16<F16>(hi));
}

// padded / strided sources, 8 outputs per thread (16-byte stores into both planes; cols_p % 8 == 0): scalar loads, L1-cached
template <bool F16>
__global__ void __launch_bounds__(256) split_planes_pad8_kernel(const float* __restrict__ src, long src_ld, long src_batch, int rows, int cols,
                                                                const float* __restrict__ colscale, uint16_t* __restrict__ dst, int rows_p, int cols_p) {
    const long i8 = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    const long per = (long)rows_p * cols_p;
    if (i8 >= per) return;
    const int b = blockIdx.y;
    const int r = (int)(i8 / cols_p), c0 = (int)(i8 % cols_p);
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int c = c0 + 2 * k + e;
            float x = 0.f;
            if (r < rows && c < cols) {
                x = __ldg(src + (long)b * src_batch + (long)r * src_ld + c);
                if (colscale) x *= __ldg(colscale + c);
            }
            v[e] = x;
        }
        const uint16_t h0 = to16<F16>(v[0]), h1 = to16<F16>(v[1]);
        const uint16_t l0 = to16<F16>(v[0] - from16<F16>(h0)), l1 = to16<F16>(v[1] - from16<F16>(h1));
        hw[k] = (uint32_t)h0 | ((uint32_t)h1 << 16);
        lw[k] = (uint32_t)l0 | ((uint32_t)l1 << 16);
    }
    uint16_t* d = dst + (long)b * 2 * per + i8;
    *reinterpret_cast<uint4*>(d) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(d + per) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

// dense, unpadded, 4 elements per thread (weights): src [batches][n4*4] -> hi / lo planes
template <bool F16>
__global__ void split_planes_vec4_kernel(const float4* __restrict__ src, long src_batch4, long n4, uint16_t* __restrict__ dst) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const int b = blockIdx.y;
    const float4 v = __ldg(src + (long)b * src_batch4 + i);
    const uint16_t h0 = to16<F16>(v.x), h1 = to16<F16>(v.y), h2 = to16<F16>(v.z), h3 = to16<F16>(v.w);
    const uint16_t l0 = to16<F16>(v.x - from16<F16>(h0)), l1 = to16<F16>(v.y - from16<F16>(h1)), l2 = to16<F16>(v.z - from16<F16>(h2)),
                   l3 = to16<F16>(v.w - from16<F16>(h3));
    uint2 hh, ll;
    hh.x = (uint32_t)h0 | ((uint32_t)h1 << 16); hh.y = (uint32_t)h2 | ((uint32_t)h3 << 16);
    ll.x = (uint32_t)l0 | ((uint32_t)l1 << 16); ll.y = (uint32_t)l2 | ((uint32_t)l3 << 16);
    uint2* d = reinterpret_cast<uint2*>(dst + (long)b * 2 * n4 * 4) + i;
    d[0] = hh;
    d[n4] = ll;
}

int split_planes(const float* src, long src_ld, long src_batch, int rows, int cols, const float* colscale, __nv_bfloat16* dst_, int rows_p,
                 int cols_p, int planes, int batches, bool f16, cudaStream_t stream) {
    if (rows_p <= 0 || cols_p <= 0 || batches <= 0) return MHE_OK;
    uint16_t* dst = reinterpret_cast<uint16_t*>(dst_);
    const long n = (long)rows * cols;
    if (planes == 2 && !colscale && rows == rows_p && cols == cols_p && src_ld == cols && n % 4 == 0 && src_batch % 4 == 0 &&
        ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 7) == 0) {
        dim3 grid(cdiv((int)(n / 4), 256), batches);
        if (f16) split_planes_vec4_kernel<true><<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(src), src_batch / 4, n / 4, dst);
        else split_planes_vec4_kernel<false><<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(src), src_batch / 4, n / 4, dst);
        return check_launch("split planes vec4");
    }
    if (planes == 2 && cols_p % 8 == 0 && ((uintptr_t)dst & 15) == 0) {
        dim3 grid8(cdiv((int)((long)rows_p * cols_p / 8), 256), batches);
        if (f16) split_planes_pad8_kernel<true><<<grid8, 256, 0, stream>>>(src, src_ld, src_batch, rows, cols, colscale, dst, rows_p, cols_p);
        else split_planes_pad8_kernel<false><<<grid8, 256, 0, stream>>>(src, src_ld, src_batch, rows, cols, colscale, dst, rows_p, cols_p);
        return check_launch("split planes pad8");
    }
    dim3 grid(cdiv((int)((long)rows_p * cols_p), 256), batches);
    if (f16) split_planes_kernel<true><<<grid, 256, 0, stream>>>(src, src_ld, src_batch, rows, cols, colscale, dst, rows_p, cols_p, planes);
    else split_planes_kernel<false><<<grid, 256, 0, stream>>>(src, src_ld, src_batch, rows, cols, colscale, dst, rows_p, cols_p, planes);
    return check_launch("split planes");
}

struct EpiRawStore {   // C[batch][row][col] = acc  (+= when accumulate; atomic when K is split)
    static constexpr bool kDirect = true, kStaged = false, kRmw = false;
    float* C; long ldc; long strideC; int accumulate; int atomic; int null_epi = 0;
    __device__ void operator()(int b, int, int row, int col0, float* v, const GemmShape& g) const {
        float* p = C + (long)b * strideC + (long)row * ldc + col0;
        if (null_epi) {   // MHE_RAW_NULL_EPI=1 (tools/bench_tc_gemm.py): the main loop alone - nothing is stored unless a value is NaN
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) s += v[j];
            if (s != s) p[0] = s;
            return;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            if (col0 + j < g.N) {
                if (atomic) atomicAdd(p + j, v[j]);
                else if (accumulate) p[j] += v[j];
                else p[j] = v[j];
            }
        }
    }
    __device__ void elem(int, int, int, int, float, const GemmShape&) const {}
};

}  // namespace tc
}  // namespace mhe

using namespace mhe;
using namespace mhe::tc;

extern "C" {

int mhe_split_planes(const float* src, int rows, int cols, void* dst, int rows_p, int cols_p, int planes, int batches, int f16, void* stream) {
    MHE_REQUIRE(src && dst && rows >= 0 && cols >= 0 && rows_p >= rows && cols_p >= cols && cols_p % 8 == 0 && (planes == 1 || planes == 2),
                "split_planes: bad args");
    return split_planes(src, cols, (long)rows * cols, rows, cols, nullptr, (__nv_bfloat16*)dst, rows_p, cols_p, planes, batches, f16 != 0, (cudaStream_t)stream);
}

// Raw tensor-core GEMM for tests: A, B are plane tensors [batches][planes][rows][cols] (dense), C fp32 [batches][M][N].
// a_mn / b_mn select MN-major operands (A stored [K][M], B stored [K][N]); otherwise A is [M][K], B is [N][K].
int mhe_tc_gemm_raw(const void* A, const void* B, float* C, int M, int N, int K, int batches, int planes, int a_mn, int b_mn, int bn,
                    int ksplit, int f16, void* stream_) {
    MHE_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0 && batches > 0 && (planes == 1 || planes == 2) && (bn == 64 || bn == 128) && ksplit >= 1,
                "tc_gemm_raw: bad args");
    cudaStream_t stream = (cudaStream_t)stream_;
    PlaneTensor ta, tb;
    ta.base = (const __nv_bfloat16*)A; ta.planes = planes; ta.batches = batches;
    tb.base = (const __nv_bfloat16*)B; tb.planes = planes; tb.batches = batches;
    if (a_mn) { ta.cols = M; ta.rows = K; } else { ta.cols = K; ta.rows = M; }
    if (b_mn) { tb.cols = N; tb.rows = K; } else { tb.cols = K; tb.rows = N; }
    ta.row_pitch = ta.cols; tb.row_pitch = tb.cols;
    MHE_REQUIRE(ta.cols % 8 == 0 && tb.cols % 8 == 0, "tc_gemm_raw: contiguous extents must be multiples of 8");
    GemmShape g{M, N, K, batches, ksplit, 1, 1};
    static const int null_epi = [] { const char* e = getenv("MHE_RAW_NULL_EPI"); return e ? atoi(e) : 0; }();
    EpiRawStore e{C, N, (long)M * N, 0, ksplit > 1, null_epi};
    if (ksplit > 1 && cudaMemsetAsync(C, 0, (size_t)batches * M * N * sizeof(float), stream) != cudaSuccess) return MHE_ERR_CUDA;
#define MHE_RAW(BN_, AMN_, BMN_, NP_) do { if (f16) return launch_tc_gemm<BN_, AMN_, BMN_, NP_, true>(ta, tb, g, e, stream, "tc raw"); \
                                             return launch_tc_gemm<BN_, AMN_, BMN_, NP_, false>(ta, tb, g, e, stream, "tc raw"); } while (0)
#define MHE_RAW_NP(BN_, AMN_, BMN_) do { if (planes == 1) MHE_RAW(BN_, AMN_, BMN_, 1); else MHE_RAW(BN_, AMN_, BMN_, 3); } while (0)
#define MHE_RAW_MAJ(BN_) do { if (!a_mn && !b_mn) MHE_RAW_NP(BN_, false, false); else if (!a_mn && b_mn) MHE_RAW_NP(BN_, false, true); \
                              else if (a_mn && !b_mn) MHE_RAW_NP(BN_, true, false); else MHE_RAW_NP(BN_, true, true); } while (0)
    if (bn == 64) MHE_RAW_MAJ(64); else MHE_RAW_MAJ(128);
#undef MHE_RAW
#undef MHE_RAW_NP
#undef MHE_RAW_MAJ
    return MHE_OK;
}

}  // extern "C"
