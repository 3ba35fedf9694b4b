// Development probe: tcgen05.mma issue/execute rate for the shapes the decoder uses (M=128, K=16, fp16 -> fp32).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
template <int N, int MODE /*0 SS, 1 TS*/, int NACC /*distinct accumulators cycled*/>
__global__ void __launch_bounds__(128) rate_kernel(long long* out, int iters) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar; __shared__ uint32_t tm;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (128 + N) * 128 * 2 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tm)), "n"(512));
                     asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t ad = make_desc(smem_u32(smem), 128 * 16, 128), bd = make_desc(smem_u32(smem) + 128 * 128 * 2, N * 16, 128);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const uint32_t d = tm + ((it * 8 + ks) % NACC) * 128;
                const uint64_t b = bd + ((ks * 2 * N * 16) >> 4);
                if (MODE == 0) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" :: "r"(d), "l"(ad + ((ks * 2 * 128 * 16) >> 4)), "l"(b), "r"(idesc) : "memory");
                else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" :: "r"(d), "r"(tm + 384 + ks * 8), "l"(b), "r"(idesc) : "memory");
            }
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "n"(512));
}
template <int N, int MODE, int NACC>
static int run(const char* name, int grid) {
    long long* d; CK(cudaMalloc(&d, 16));
    const int iters = 2000; const size_t smem = (128 + N) * 128 * 2;
    CK(cudaFuncSetAttribute(rate_kernel<N, MODE, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rate_kernel<N, MODE, NACC><<<grid, 128, smem>>>(d, iters);
    CK(cudaDeviceSynchronize());
    long long h[2]; CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
    printf("%-34s grid=%3d  issue %.1f cyc/MMA   complete %.1f cyc/MMA\n", name, grid, (double)h[0] / (iters * 8), (double)h[1] / (iters * 8));
    cudaFree(d); return 0;
}
int main() {
    for (int grid : {1, 148}) {
        run<128, 0, 1>("N=128 SS same accumulator", grid);
        run<128, 1, 1>("N=128 TS same accumulator", grid);
        run<128, 1, 2>("N=128 TS 2 accumulators", grid);
        run<96, 1, 1>("N=96  TS same accumulator", grid);
        run<64, 1, 1>("N=64  TS same accumulator", grid);
        run<256, 1, 1>("N=256 TS same accumulator", grid);
        run<256, 0, 1>("N=256 SS same accumulator", grid);
    }
    return 0;
}
