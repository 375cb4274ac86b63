// Correctness probe: one k-block (K = 64) of tcgen05.mma kind::f16, M = 128, A K-major SW128, B MN-major SW128 given as
// two 64-column groups LBO bytes apart.  Checks D = A.B for N = 64 (group 0 only) and N = 128 (both groups).
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__host__ __device__ constexpr uint32_t instr_desc(int m, int n, bool a_mn, bool b_mn) {
    return (1u << 4) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__host__ __device__ inline float aval(int m, int k) { return ((m * 3 + k * 5) % 11 - 5) * 0.125f; }
__host__ __device__ inline float bval(int g, int k, int n) { return ((k * 7 + n * 3 + g * 2) % 13 - 6) * 0.0625f; }

// mode 0: N = 64 on group 0;  mode 1: N = 128 over both groups;  lbo: byte distance between the groups
__global__ void __launch_bounds__(128, 1) probe(int mode, int lbo, float* out, int boff) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base = smem_raw + (smem0 - smem_u32(smem_raw));
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // A: K-major [128 rows][64 k] SW128 at +0 (16 KB).  B groups at +32768 and +32768 + lbo, each [64 k-rows][64 n] MN-major SW128.
    for (int i = threadIdx.x; i < 128 * 64; i += 128) {
        const int m = i / 64, k = i % 64;
        const uint32_t off = (m / 8) * 1024 + (m % 8) * 128 + (((k / 8) ^ (m % 8)) << 4) + (k % 8) * 2;
        *reinterpret_cast<__half*>(base + off) = __float2half(aval(m, k));
        *reinterpret_cast<__half*>(base + 16384 + off) = __float2half(0.5f * aval(m, k));
    }
    for (int g = 0; g < 2; ++g)
        for (int i = threadIdx.x; i < 64 * 64; i += 128) {
            const int k = i / 64, n = i % 64;
            const uint32_t off = boff + g * lbo + k * 128 + (((n / 8) ^ (k % 8)) << 4) + (n % 8) * 2;
            *reinterpret_cast<__half*>(base + off) = __float2half(bval(g, k, n));
        }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = mode ? instr_desc(128, 128, false, true) : instr_desc(128, 64, false, true);
        const uint32_t kHi = (uint32_t)((1024u >> 4) | (1u << 14) | (2u << 29));
        const uint32_t a0 = (1u << 16) | (smem0 >> 4), b0 = (((uint32_t)lbo >> 4) << 16) | ((smem0 + boff) >> 4);
        if (mode == 2) {
            const uint32_t i128 = instr_desc(128, 128, false, true), i64 = instr_desc(128, 64, false, true);
            for (int ks = 0; ks < 4; ++ks) {
                umma(tmem, a0 + ks * 2u, b0 + ks * (2048u >> 4), kHi, i128, ks > 0);
                umma(tmem, a0 + (16384u >> 4) + ks * 2u, b0 + ks * (2048u >> 4), kHi, i64, 1u);
            }
        } else
        for (int ks = 0; ks < 4; ++ks) umma(tmem, a0 + ks * 2u, b0 + ks * (2048u >> 4), kHi, idesc, ks > 0);
        commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = threadIdx.x >> 5;
    for (int c0 = 0; c0 < 128; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int jj = 0; jj < 32; ++jj) out[threadIdx.x * 128 + c0 + jj] = v[jj];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

int main() {
    float* d; cudaMalloc(&d, 128 * 128 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    static float h[128 * 128];
    const int cfg[][3] = {{1, 8192, 32768}, {2, 8192, 32768}, {2, 8192, 131072}, {2, 8192, 180224}, {2, 8192, 196608}, {2, 16384, 180224}};
    for (auto& c : cfg) {
        cudaMemset(d, 0, sizeof(h));
        probe<<<1, 128, 220 * 1024>>>(c[0], c[1], d, c[2]);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double maxerr[2] = {0, 0};
        int nan = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < (c[0] ? 128 : 64); ++n) {
                double ref = 0;
                for (int k = 0; k < 64; ++k) ref += (double)aval(m, k) * bval(n / 64, k, n % 64) * ((c[0] == 2 && n < 64) ? 1.5 : 1.0);
                const float got = h[m * 128 + n];
                if (got != got) { ++nan; continue; }
                maxerr[n / 64] = fmax(maxerr[n / 64], fabs(got - ref));
            }
        printf("boff %d mode N=%d lbo=%d: %s  max err group0 %.3g group1 %.3g  nan %d   sample D[1][0]=%g D[1][64]=%g\n", c[2], c[0] ? 128 : 64, c[1], cudaGetErrorString(e),
               maxerr[0], maxerr[1], nan, h[128], h[128 + 64]);
    }
    return 0;
}
