// Hardware probe (development tool, not product code): validates on a B200 the tcgen05 building blocks the decoder
// kernel relies on -- no-swizzle K-major UMMA shared-memory descriptors, the kind::f16 instruction descriptor,
// A-operand-from-TMEM (packed half2 written with tcgen05.st.32x32b), tcgen05.ld of the fp32 accumulator,
// tcgen05.commit -> mbarrier.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tc_probe tc_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <cmath>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;          // descriptor version (Blackwell)
    return d;                        // layout_type = 0 (no swizzle), base_offset = 0
}

__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    return (1u << 4) /*D=f32*/ | (0u << 7) /*A=f16*/ | (0u << 10) /*B=f16*/ | (0u << 15) /*A K-major*/ | (0u << 16) /*B K-major*/ |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}

// variant: 0 = (LBO = k-chunk stride, SBO = 8-row-group stride); 1 = swapped.   mode: 0 = A from smem (SS), 1 = A from TMEM (TS)
template <int N, int K>
__global__ void __launch_bounds__(128) probe_kernel(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D,
                                                    int variant, int mode) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __half* sA = reinterpret_cast<__half*>(smem);                       // [K/8][128][8]
    __half* sB = reinterpret_cast<__half*>(smem + 128 * K * 2);         // [K/8][N][8]
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;

    // no-swizzle K-major: element (r, k) at (k/8)*(rows*16) + r*16 + (k%8)*2 bytes
    for (int i = tid; i < 128 * K; i += 128) { const int r = i / K, k = i % K; sA[(k / 8) * 128 * 8 + r * 8 + (k % 8)] = A[i]; }
    for (int i = tid; i < N * K; i += 128) { const int r = i / K, k = i % K; sB[(k / 8) * N * 8 + r * 8 + (k % 8)] = B[i]; }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic-proxy smem writes -> visible to the MMA (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t acc = tmem;                 // columns [0, N)
    const uint32_t a_t = tmem + 256;           // columns [256, 256 + K/2): A operand, 2 halves per column

    if (mode == 1) {
        // each thread owns TMEM lane = its row; write K/2 packed columns, 16 columns per tcgen05.st
        const int row = tid;
        for (int c0 = 0; c0 < K / 2; c0 += 8) {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const __half lo = A[row * K + 2 * (c0 + j)], hi = A[row * K + 2 * (c0 + j) + 1];
                v[j] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
            }
            const uint32_t addr = a_t + ((uint32_t)(warp * 32) << 16) + c0;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                         :: "r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = make_idesc(128, N);
        const uint32_t a_kchunk = 128 * 16, b_kchunk = N * 16, rowgrp = 128;
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint32_t a_addr = smem_u32(sA) + ks * 2 * a_kchunk, b_addr = smem_u32(sB) + ks * 2 * b_kchunk;
            const uint64_t ad = variant == 0 ? make_desc(a_addr, a_kchunk, rowgrp) : make_desc(a_addr, rowgrp, a_kchunk);
            const uint64_t bd = variant == 0 ? make_desc(b_addr, b_kchunk, rowgrp) : make_desc(b_addr, rowgrp, b_kchunk);
            if (mode == 0) mma_ss(acc, ad, bd, idesc, ks > 0);
            else mma_ts(acc, a_t + ks * 8, bd, idesc, ks > 0);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        const int row = tid;
        for (int c0 = 0; c0 < N; c0 += 8) {
            uint32_t v[8];
            const uint32_t addr = acc + ((uint32_t)(warp * 32) << 16) + c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(addr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 8; ++j) D[row * N + c0 + j] = __uint_as_float(v[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(512));
}

template <int N, int K>
static void run(const char* name) {
    std::vector<__half> hA(128 * K), hB(N * K);
    std::vector<float> ref(128 * N), out(128 * N);
    srand(1);
    for (auto& v : hA) v = __float2half((rand() % 2001 - 1000) / 1000.0f);
    for (auto& v : hB) v = __float2half((rand() % 2001 - 1000) / 1000.0f);
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
        double a = 0; for (int k = 0; k < K; ++k) a += (double)__half2float(hA[m * K + k]) * __half2float(hB[n * K + k]);
        ref[m * N + n] = (float)a;
    }
    __half *dA, *dB; float* dD;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, out.size() * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    const size_t smem = (128 + N) * K * 2;
    CK(cudaFuncSetAttribute(probe_kernel<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int mode = 0; mode < 2; ++mode) for (int variant = 0; variant < 1; ++variant) {
        CK(cudaMemset(dD, 0, out.size() * 4));
        probe_kernel<N, K><<<1, 128, smem>>>(dA, dB, dD, variant, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s mode=%d variant=%d: CUDA error %s\n", name, mode, variant, cudaGetErrorString(e)); exit(2); }
        CK(cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0; for (size_t i = 0; i < out.size(); ++i) maxerr = fmax(maxerr, fabs((double)out[i] - ref[i]));
        printf("%s  mode=%s  desc-variant=%d  max|err|=%.3e  %s\n", name, mode ? "TS(A in TMEM)" : "SS(A in smem)", variant, maxerr,
               maxerr < 1e-3 ? "OK" : "MISMATCH");
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

int main() {
    run<128, 32>("N=128 K=32 ");
    run<128, 128>("N=128 K=128");
    run<96, 128>("N=96  K=128");
    run<32, 64>("N=32  K=64 ");
    return 0;
}
