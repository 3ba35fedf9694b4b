// Development probe: cost of a warp-cooperative row gather (116 B rows) as a function of the shared-memory carve-out.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
template <int MODE>   // 0: __ldg coalesced rows, 8 in flight ; 1: same with ld.cg ; 2: thread-per-row (32 rows x 29 loads)
__global__ void gather_kernel(const float* __restrict__ table, const int* __restrict__ rows, int n_rows, float* out, long long* cyc) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0.f;
    const long long t0 = clock64();
    const int per_warp = n_rows / (gridDim.x * (blockDim.x / 32));
    const int base = (blockIdx.x * (blockDim.x / 32) + warp) * per_warp;
    if (MODE < 2) {
        for (int i = 0; i < per_warp; i += 8) {
            float v[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int src = rows[base + i + r];
                const float* p = table + (int64_t)src * 29 + (lane < 29 ? lane : 28);
                v[r] = MODE == 0 ? __ldg(p) : __ldcg(p);
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) acc += v[r];
        }
    } else {
        for (int i = 0; i < per_warp; i += 32) {
            const int src = rows[base + i + lane];
            const float* p = table + (int64_t)src * 29;
            float v[29];
#pragma unroll
            for (int j = 0; j < 29; ++j) v[j] = __ldg(p + j);
#pragma unroll
            for (int j = 0; j < 29; ++j) acc += v[j];
        }
    }
    const long long t1 = clock64();
    if (acc == 123.456f) out[0] = acc + sm[0];
    if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = (t1 - t0);
    if (blockIdx.x == 0 && threadIdx.x == 0) cyc[1] = per_warp;
}
template <int MODE>
static int run(const char* name, size_t smem, int warps, const float* table, const int* rows, int n_rows, float* out, long long* cyc) {
    CK(cudaFuncSetAttribute(gather_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gather_kernel<MODE><<<148, warps * 32, smem>>>(table, rows, n_rows, out, cyc);
    CK(cudaDeviceSynchronize());
    gather_kernel<MODE><<<148, warps * 32, smem>>>(table, rows, n_rows, out, cyc);
    CK(cudaDeviceSynchronize());
    long long h[2]; CK(cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost));
    printf("%-34s smem=%6zu B warps/CTA=%d : %.1f cycles per row per warp\n", name, smem, warps, (double)h[0] / h[1]);
    return 0;
}
int main() {
    const int n_table = 23000, n_rows = 148 * 2 * 4096;
    float* table; int* rows; float* out; long long* cyc;
    CK(cudaMalloc(&table, n_table * 29 * 4)); CK(cudaMalloc(&rows, n_rows * 4)); CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&cyc, 16));
    CK(cudaMemset(table, 0, n_table * 29 * 4));
    int* h = (int*)malloc(n_rows * 4); srand(1);
    for (int i = 0; i < n_rows; ++i) h[i] = rand() % n_table;
    CK(cudaMemcpy(rows, h, n_rows * 4, cudaMemcpyHostToDevice));
    for (size_t smem : {(size_t)0, (size_t)100 * 1024, (size_t)200 * 1024, (size_t)227 * 1024}) {
        run<0>("coalesced rows __ldg", smem, 2, table, rows, n_rows, out, cyc);
        run<1>("coalesced rows ld.cg", smem, 2, table, rows, n_rows, out, cyc);
        run<2>("thread-per-row __ldg", smem, 2, table, rows, n_rows, out, cyc);
    }
    return 0;
}
