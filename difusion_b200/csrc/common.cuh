// Shared device helpers of libdifusion_b200 (sm_100a).  No torch, no thrust: plain CUDA runtime.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/difusion_b200.h"

#define DIF_L 29                 // latent dim
#define DIF_SUM_STRIDE 32         // row stride of the per-slot encoder sums (16-byte aligned rows: float4 reductions)
#define DIF_NUM_SMS 148          // B200

namespace dif {

extern thread_local char g_last_error[256];

// Optional measurement hooks (bench.py): one-shot CUDA-event brackets around a named kernel, and a launch counter.
struct ProfHook { cudaEvent_t start, stop; };
extern thread_local ProfHook g_prof[DIF_PROF_COUNT];
extern thread_local uint64_t g_launches;
inline void prof_begin(int which, cudaStream_t st) { if (g_prof[which].start) cudaEventRecord(g_prof[which].start, st); }
inline void prof_end(int which, cudaStream_t st) {
    if (g_prof[which].stop) { cudaEventRecord(g_prof[which].stop, st); g_prof[which].start = nullptr; g_prof[which].stop = nullptr; }
}
#define DIF_COUNT_LAUNCH(n) (dif::g_launches += (n))

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what, cudaGetErrorString(e));
        return DIF_E_LAUNCH;
    }
    return DIF_OK;
}

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------------------
// The per-frame path is a chain of short kernels; with the programmatic-stream-serialization attribute a kernel may be
// scheduled while its predecessor is still running and blocks at pdl_wait() until the predecessor has completed and its
// writes are visible, which hides the launch latency (and, for the tensor-core kernels, the weight-image load) behind the
// predecessor.  Rules kept by every kernel launched through launch_pdl: pdl_wait() is executed unconditionally by every
// thread before the first access to memory another kernel of the stream may have written, and nothing is read before it
// except kernel parameters and the prepared (constant) network images.  DIF_PDL=0 turns the attribute off (A/B timing).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Carves a caller-provided workspace into aligned sub-buffers.
struct Carver {
    char* base; size_t off;
    explicit Carver(void* p) : base((char*)p), off(0) {}
    template <class T> T* take(size_t count) { T* r = (T*)(base + off); off += align_up(count * sizeof(T)); return r; }
};

// Map geometry passed by value to kernels.
struct Grid {
    int nx, ny, nz;
    float bx, by, bz, vs;
    float inv_vs; int mul;        // scalar_division_mode 1: multiply by the fp32 reciprocal (torch on CUDA tensors)
    __host__ __device__ int64_t cells() const { return (int64_t)nx * ny * nz; }
};

inline Grid make_grid(const dif_map_view* m) {
    Grid g; g.nx = m->nx; g.ny = m->ny; g.nz = m->nz;
    g.bx = m->bound_min[0]; g.by = m->bound_min[1]; g.bz = m->bound_min[2]; g.vs = m->voxel_size;
    g.inv_vs = 1.0f / m->voxel_size; g.mul = m->scalar_division_mode == 1;
    return g;
}

// (p - bound_min) / voxel_size exactly as the reference computes it: one fp32 subtract, then one fp32 true division (torch on
// CPU tensors) or one fp32 multiplication by the rounded reciprocal (torch on CUDA tensors, dif_map_view.scalar_division_mode)
// (system/map.py:366-367, :565).  Explicit _rn intrinsics stop nvcc from contracting or replacing the operations.
__device__ __forceinline__ float3 normalize_point(const Grid& g, float x, float y, float z) {
    const float dx = __fsub_rn(x, g.bx), dy = __fsub_rn(y, g.by), dz = __fsub_rn(z, g.bz);
    if (g.mul) return make_float3(__fmul_rn(dx, g.inv_vs), __fmul_rn(dy, g.inv_vs), __fmul_rn(dz, g.inv_vs));
    return make_float3(__fdiv_rn(dx, g.vs), __fdiv_rn(dy, g.vs), __fdiv_rn(dz, g.vs));
}

// linear id = z + nz*y + nz*ny*x  (map.py:287-292)
__device__ __forceinline__ int lin_id(const Grid& g, int ix, int iy, int iz) { return iz + g.nz * (iy + g.ny * ix); }

__device__ __forceinline__ bool in_grid(const Grid& g, int ix, int iy, int iz) {
    return (unsigned)ix < (unsigned)g.nx && (unsigned)iy < (unsigned)g.ny && (unsigned)iz < (unsigned)g.nz;
}

// ---- hash-sharded map: ownership --------------------------------------------------------------------------------------------
// owner(cell) = splitmix64(id of the cell's super-block) % world; a super-block is (2^k)^3 cells.  Hashing blocks instead of
// single cells keeps a PLIVox and (almost all of) its 26 neighbours on one rank, so only the rows on a block's surface ever
// have to travel (the blend of mc_interp_kernel.cu:103-181 reads the 3x3x3 neighbourhood).
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
struct Shard { int rank, world, k; };          // k = log2 of the super-block edge in cells
__host__ __device__ __forceinline__ int shard_block_owner(int bx, int by, int bz, int nbx, int nby, int nbz, int world) {
    (void)nbx;
    return (int)(mix64((uint64_t)bz + (uint64_t)nbz * ((uint64_t)by + (uint64_t)nby * (uint64_t)bx)) % (uint64_t)world);
}
__host__ __device__ __forceinline__ int shard_owner_xyz(int nx, int ny, int nz, int ix, int iy, int iz, int k, int world) {
    const int nbx = ((nx - 1) >> k) + 1, nby = ((ny - 1) >> k) + 1, nbz = ((nz - 1) >> k) + 1;
    return shard_block_owner(ix >> k, iy >> k, iz >> k, nbx, nby, nbz, world);
}
__host__ __device__ __forceinline__ int shard_owner_lin(int nx, int ny, int nz, int64_t lin, int k, int world) {
    const int iz = (int)(lin % nz), iy = (int)((lin / nz) % ny), ix = (int)(lin / ((int64_t)nz * ny));
    return shard_owner_xyz(nx, ny, nz, ix, iy, iz, k, world);
}
// Bit r set: rank r owns the cell or one of its 26 neighbours' super-blocks, i.e. rank r keeps this cell's latent row (as owner
// or in its halo).  Only cells on the surface of their super-block have neighbours in other blocks (world <= 32).
__host__ __device__ __forceinline__ uint32_t shard_holder_mask(int nx, int ny, int nz, int64_t lin, int k, int world) {
    const int iz = (int)(lin % nz), iy = (int)((lin / nz) % ny), ix = (int)(lin / ((int64_t)nz * ny));
    const int nbx = ((nx - 1) >> k) + 1, nby = ((ny - 1) >> k) + 1, nbz = ((nz - 1) >> k) + 1;
    const int bx = ix >> k, by = iy >> k, bz = iz >> k, e = (1 << k) - 1;
    // neighbour-block offsets per axis: -1 if the cell sits on the low face of its block, +1 on the high face (grid-clamped)
    const int x0 = ((ix & e) == 0 && ix > 0) ? -1 : 0, x1 = ((ix & e) == e && ix < nx - 1) ? 1 : 0;
    const int y0 = ((iy & e) == 0 && iy > 0) ? -1 : 0, y1 = ((iy & e) == e && iy < ny - 1) ? 1 : 0;
    const int z0 = ((iz & e) == 0 && iz > 0) ? -1 : 0, z1 = ((iz & e) == e && iz < nz - 1) ? 1 : 0;
    uint32_t m = 0;
    for (int dx = x0; dx <= x1; ++dx)
        for (int dy = y0; dy <= y1; ++dy)
            for (int dz = z0; dz <= z1; ++dz) m |= 1u << shard_block_owner(bx + dx, by + dy, bz + dz, nbx, nby, nbz, world);
    return m;
}

// One latent row (29 floats) into x[0..28] (x[29..31] are left for the caller).  Rows padded to 32 floats are 16-byte aligned:
// 8 vector loads; a lane gathering its own row touches one 128-byte line instead of issuing 29 scalar requests.
__device__ __forceinline__ void load_latent_row(const float* __restrict__ table, int64_t row, int stride, float (&x)[32]) {
    const float* lp = table + row * stride;
    if (stride == 32) {
        const float4* q = reinterpret_cast<const float4*>(lp);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float4 v = __ldg(q + j); x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w; }
    } else {
#pragma unroll
        for (int j = 0; j < DIF_L; ++j) x[j] = __ldg(lp + j);
    }
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace dif
