// Hardware probe: B operand used TRANSPOSED through an MN-major descriptor over the SAME shared-memory slab the forward pass
// uses (backward GEMM  G[128 x Nout] * W[Nout][Kin] -> [128 x Kin]).  Slab: element (n, k) of W at (k/8)*(Nout*16) + n*16 + (k%8)*2.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <cmath>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
// A: G [128][NOUT] (from TMEM), W [NOUT][KIN] in forward slab layout; D [128][KIN]
template <int NOUT, int KIN>
__global__ void __launch_bounds__(128) probe(const __half* __restrict__ G, const __half* __restrict__ W, float* __restrict__ D, int variant) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __half* sW = reinterpret_cast<__half*>(smem);
    __shared__ uint64_t bar; __shared__ uint32_t tm;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < NOUT * KIN; i += 128) { const int n = i / KIN, k = i % KIN; sW[(k / 8) * NOUT * 8 + n * 8 + (k % 8)] = W[i]; }
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tm)), "n"(512));
                     asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
    if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t acc = tm, a_t = tm + 256;
    for (int c0 = 0; c0 < NOUT / 2; c0 += 8) {
        uint32_t v[8];
        for (int j = 0; j < 8; ++j) v[j] = (uint32_t)__half_as_ushort(G[tid * NOUT + 2 * (c0 + j)]) | ((uint32_t)__half_as_ushort(G[tid * NOUT + 2 * (c0 + j) + 1]) << 16);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     :: "r"(a_t + ((uint32_t)(warp * 32) << 16) + c0), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // D = f32, A = B = f16, A K-major, B MN-major (bit 16), N = KIN, M = 128
        const uint32_t idesc = (1u << 4) | (1u << 16) | ((uint32_t)(KIN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t mn_group = NOUT * 16, k_group = 128;      // stride between 8-wide k_in chunks / between 8-row n_out groups
        for (int ks = 0; ks < NOUT / 16; ++ks) {
            const uint32_t addr = smem_u32(sW) + ks * 2 * k_group;
            const uint64_t bd = variant == 0 ? make_desc(addr, k_group, mn_group) : make_desc(addr, mn_group, k_group);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                         :: "r"(acc), "r"(a_t + ks * 8), "l"(bd), "r"(idesc), "r"((uint32_t)(ks > 0)) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < KIN; c0 += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(acc + ((uint32_t)(warp * 32) << 16) + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) D[tid * KIN + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "n"(512));
}
template <int NOUT, int KIN>
static void run(const char* name, int variant) {
    std::vector<__half> hG(128 * NOUT), hW(NOUT * KIN); std::vector<float> ref(128 * KIN), out(128 * KIN);
    srand(2);
    for (auto& v : hG) v = __float2half((rand() % 2001 - 1000) / 1000.0f);
    for (auto& v : hW) v = __float2half((rand() % 2001 - 1000) / 1000.0f);
    for (int m = 0; m < 128; ++m) for (int k = 0; k < KIN; ++k) { double a = 0; for (int n = 0; n < NOUT; ++n) a += (double)__half2float(hG[m * NOUT + n]) * __half2float(hW[n * KIN + k]); ref[m * KIN + k] = (float)a; }
    __half *dG, *dW; float* dD;
    CK(cudaMalloc(&dG, hG.size() * 2)); CK(cudaMalloc(&dW, hW.size() * 2)); CK(cudaMalloc(&dD, out.size() * 4));
    CK(cudaMemcpy(dG, hG.data(), hG.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, out.size() * 4));
    const size_t smem = NOUT * KIN * 2;
    CK(cudaFuncSetAttribute(probe<NOUT, KIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe<NOUT, KIN><<<1, 128, smem>>>(dG, dW, dD, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s variant=%d: CUDA error %s\n", name, variant, cudaGetErrorString(e)); exit(2); }
    CK(cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0; for (size_t i = 0; i < out.size(); ++i) maxerr = fmax(maxerr, fabs((double)out[i] - ref[i]));
    printf("%s variant=%d (0: LBO=k-group 128B, SBO=mn-group; 1: swapped)  max|err|=%.3e %s\n", name, variant, maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
}
int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    run<128, 128>("Nout=128 Kin=128", variant);
    run<96, 128>("Nout=96  Kin=128", variant);
    run<128, 32>("Nout=128 Kin=32 ", variant);
    return 0;
}
