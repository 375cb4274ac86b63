// Microbenchmark: cluster-internal bulk copies shared::cta -> shared::cluster (cp.async.bulk, mbarrier complete_tx) and plain
// st.shared::cluster stores.  Cluster of 8; every CTA sends `bytes` to each of `fanout` peers, `iters` times.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t cta) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(cta)); return r; }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(256, 1) k(int bytes, int fanout, int iters, int mode, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const uint32_t src = smem_u32(smem), dst = smem_u32(smem) + 96 * 1024;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    for (int i = threadIdx.x; i < 48 * 1024; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = i;
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (mode == 0) {
            if (threadIdx.x == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes * fanout) : "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                for (int f = 1; f <= fanout; ++f) {
                    const uint32_t peer = (rank + f) & 7;
                    // my slot in the peer: (f-1) * bytes
                    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(mapa(dst + (f - 1) * bytes, peer)), "r"(src), "r"(bytes), "r"(mapa(smem_u32(&bar), peer)) : "memory");
                }
                mbar_wait(smem_u32(&bar), it & 1);
            }
            asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        } else {
            for (int f = 1; f <= fanout; ++f) {
                const uint32_t peer = (rank + f) & 7;
                const uint32_t base = mapa(dst + (f - 1) * bytes, peer);
                for (int o = threadIdx.x * 16; o < bytes; o += 256 * 16) {
                    uint4 v = *reinterpret_cast<uint4*>(smem + o);
                    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + o), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                }
            }
            asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
int main() {
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int cfg[][3] = {{32768, 3, 0}, {32768, 1, 0}, {8192, 7, 0}, {32768, 3, 1}, {8192, 7, 1}};
    for (auto& c : cfg)
        for (int grid : {8, 80}) {
            const int iters = 20;
            k<<<grid, 256, 200 * 1024>>>(c[0], c[1], iters, c[2], d);
            cudaError_t e = cudaDeviceSynchronize();
            long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            printf("%s bytes %6d fanout %d grid %3d: %8.0f cyc/iter  -> %6.1f B/clk out per CTA (%s)\n", c[2] ? "st.cluster" : "bulk copy ", c[0], c[1], grid,
                   (double)h / iters, (double)c[0] * c[1] * iters / h, cudaGetErrorString(e));
        }
    return 0;
}
