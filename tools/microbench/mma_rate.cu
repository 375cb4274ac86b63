// Microbenchmark: issue rate of tcgen05.mma (kind::f16, M=128, cta_group::1) for several N and operand majors,
// operands resident in shared memory (contents irrelevant).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__host__ __device__ constexpr uint32_t instr_desc(int m, int n, bool a_mn, bool b_mn) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate) : "memory");
}

template <int M, int N, bool A_MN, bool B_MN, int NACC>
__global__ void __launch_bounds__(128, 1) bench(int iters, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // zero the operands (finite values)
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (smem0 - smem_u32(smem_raw)))[i] = 0x3c003c00u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = instr_desc(M, N, A_MN, B_MN);
        constexpr uint32_t kHi = (uint32_t)((1024u >> 4) | (1u << 14) | (2u << 29));
        constexpr uint32_t kLoA = A_MN ? ((8192u >> 4) << 16) : (1u << 16);
        constexpr uint32_t kLoB = B_MN ? ((8192u >> 4) << 16) : (1u << 16);
        constexpr uint32_t kStepA = A_MN ? (2048u >> 4) : 2u, kStepB = B_MN ? (2048u >> 4) : 2u;
        const uint32_t a0 = kLoA | (smem0 >> 4), b0 = kLoB | ((smem0 + 65536) >> 4);
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma(tmem + (it % NACC) * N, a0 + ks * kStepA, b0 + ks * kStepB, kHi, idesc, 1);
        }
        commit(smem_u32(&bar));
        long long t1 = clock64();
        mbar_wait(smem_u32(&bar), 0);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int M, int N, bool A_MN, bool B_MN, int NACC>
void run(const char* name) {
    long long* d; cudaMalloc(&d, 16);
    auto k = bench<M, N, A_MN, B_MN, NACC>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 512;
    for (int rep = 0; rep < 2; ++rep) { k<<<1, 128, 200 * 1024>>>(iters, d); cudaDeviceSynchronize(); }
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    printf("%-44s issue %7.1f cyc/mma   complete %7.1f cyc/mma  (%s)\n", name, (double)h[0] / (iters * 4), (double)h[1] / (iters * 4), cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    run<128, 64, false, false, 1>("M128 N64  A K-major  B K-major  1 acc");
    run<128, 64, false, true, 1>("M128 N64  A K-major  B MN-major 1 acc");
    run<128, 64, true, true, 1>("M128 N64  A MN-major B MN-major 1 acc");
    run<128, 64, false, true, 4>("M128 N64  A K-major  B MN-major 4 acc");
    run<128, 128, false, false, 1>("M128 N128 A K-major  B K-major  1 acc");
    run<128, 128, false, true, 1>("M128 N128 A K-major  B MN-major 1 acc");
    run<128, 256, false, false, 1>("M128 N256 A K-major  B K-major  1 acc");
    run<128, 256, false, true, 1>("M128 N256 A K-major  B MN-major 1 acc");
    run<128, 32, false, true, 1>("M128 N32  A K-major  B MN-major 1 acc");
    run<64, 64, false, true, 1>("M64  N64  A K-major  B MN-major 1 acc");
    run<64, 128, false, true, 1>("M64  N128 A K-major  B MN-major 1 acc");
    return 0;
}
