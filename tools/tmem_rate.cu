// Development probe: tcgen05.ld / tcgen05.st throughput per SM vs number of warps, alone and concurrently with MMAs.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
#define LD32(addr, v) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15," \
    "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), \
      "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), \
      "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), \
      "=r"(v[30]), "=r"(v[31]) : "r"(addr))
#define ST16(addr, v) asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
    :: "r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), \
       "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory")

// warps [0, nw): TMEM traffic (op 0: ld x32 ; 1: st x16 ; 2: ld x32 + 2 st x16).  warp 16: n_mma MMAs (N=128, TS).
__global__ void __launch_bounds__(544) tmem_kernel(long long* out, int iters, int nw, int op, int n_mma) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 65536 / 4; i += 544) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (warp == 16) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tm)), "n"(512));
                      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 16) {
        if (n_mma > 0 && lane == 0) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint64_t bd = make_desc(smem_u32(smem) + 32768, 128 * 16, 128);
            const long long t0 = clock64();
            for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                                 :: "r"(tm + 256), "r"(tm + 384 + ks * 8), "l"(bd + ((ks * 2 * 128 * 16) >> 4)), "r"(idesc) : "memory");
            }
            const long long t1 = clock64();
            if (blockIdx.x == 0) out[1] = t1 - t0;
        }
    } else if (warp < nw) {
        const uint32_t addr = tm + ((uint32_t)((warp & 3) * 32) << 16) + ((warp >> 2) & 3) * 32;
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = lane + j;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (op == 0 || op == 2) { LD32(addr, v); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
            if (op == 1 || op == 2) { ST16(addr + 128, v); if (op == 2) ST16(addr + 144, v); asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
        uint32_t acc = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) acc += v[j];
        if (acc == 0xdeadbeef) out[5] = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 16) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "n"(512));
}
int main() {
    long long* d; CK(cudaMalloc(&d, 64));
    CK(cudaFuncSetAttribute(tmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    const int iters = 4000;
    const char* names[3] = {"ld x32 (4 KB/warp-iter)", "st x16 (2 KB/warp-iter)", "ld x32 + 2 st x16"};
    for (int op = 0; op < 3; ++op)
        for (int nw : {1, 4, 8, 16})
            for (int mma : {0, 1}) {
                CK(cudaMemset(d, 0, 64));
                const int n_mma = mma ? iters * 16 : 0;
                tmem_kernel<<<148, 544, 65536>>>(d, iters, nw, op, n_mma);
                CK(cudaDeviceSynchronize());
                long long h[2]; CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
                const double cyc = (double)h[0] / iters;
                const double bytes = (op == 0 ? 4096.0 : op == 1 ? 2048.0 : 4096.0) * nw;   // ld bytes (op2: ld part)
                printf("%-26s warps=%2d mma=%d : %7.1f cyc/iter/warp  -> %6.1f B/clk/SM (ld or st bytes)", names[op], nw, mma, cyc, bytes / cyc);
                if (mma) printf("   MMA %.1f cyc each", (double)h[1] / n_mma);
                printf("\n");
            }
    return 0;
}
