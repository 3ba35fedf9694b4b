// Device-side mesh cache merge.
//   replaces the host part of reference system/map.py:698-714: D2H of every new triangle, `_get_valid_idx` (:20-26, a numba
//   searchsorted loop over all cached triangles), three np.concatenate, plus the voxel->world transform of :698.
// Semantics kept exactly: a cached triangle survives unless its PLIVox id appears among the ids of the NEW triangles (a
// re-meshed PLIVox that produced no triangle keeps its old ones, as in the reference); order = surviving cached triangles in
// their old order, then the new triangles in MC output order.
//
// HBM-bound streaming: every cached row (56 B) is read once and written at most once.  Membership is a byte flag per grid cell
// (self-cleaning: the flags set from the new ids are cleared again from the same list, so a merge costs O(triangles), never
// O(grid)); the stable compaction is count -> single-CTA scan of the per-chunk counts -> scatter with a block scan.
#include "common.cuh"

namespace dif {

constexpr int MCACHE_THREADS = 256;
constexpr int MCACHE_CHUNK = 1024;              // cached triangles per CTA (4 per thread)

__global__ void cache_flag_kernel(const int64_t* __restrict__ ids, int64_t n, uint8_t* __restrict__ flag, uint8_t v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[ids[i]] = v;
}

__device__ __forceinline__ int block_sum(int v, int* s_warp) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
    for (int w = 0; w < MCACHE_THREADS / 32; ++w) t += s_warp[w];
    return t;
}

__global__ void __launch_bounds__(MCACHE_THREADS) cache_count_kernel(const int64_t* __restrict__ cache_id, int64_t n,
                                                                      const uint8_t* __restrict__ flag, int32_t* __restrict__ chunk_count) {
    __shared__ int s_warp[MCACHE_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * MCACHE_CHUNK;
    int keep = 0;
#pragma unroll
    for (int k = 0; k < MCACHE_CHUNK / MCACHE_THREADS; ++k) {
        const int64_t i = base + k * MCACHE_THREADS + threadIdx.x;
        if (i < n) keep += flag[cache_id[i]] == 0;
    }
    const int t = block_sum(keep, s_warp);
    if (threadIdx.x == 0) chunk_count[blockIdx.x] = t;
}

// exclusive scan of the per-chunk counts by one CTA (n_chunks = n_cache / 1024: a few thousand); total -> *kept_out, and
// *total_out = kept + n_new (the size of the merged cache, read back by the host)
__global__ void __launch_bounds__(1024) cache_scan_kernel(int32_t* __restrict__ chunk_count, int64_t n_chunks, int64_t n_new,
                                                          int64_t* __restrict__ totals /*[2]*/) {
    __shared__ int s_warp[32];
    __shared__ int64_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t c0 = 0; c0 < n_chunks; c0 += 1024) {
        const int64_t c = c0 + threadIdx.x;
        const int v = c < n_chunks ? chunk_count[c] : 0;
        int incl = v;
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += u; }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        int wpre = 0, tot = 0;
        for (int w = 0; w < 32; ++w) { const int t = s_warp[w]; if (w < (int)(threadIdx.x >> 5)) wpre += t; tot += t; }
        const int64_t carry = s_carry;
        if (c < n_chunks) chunk_count[c] = (int32_t)(carry + wpre + incl - v);        // exclusive offset (< 2^31 triangles)
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) { totals[0] = s_carry; totals[1] = s_carry + n_new; }
}

__global__ void __launch_bounds__(MCACHE_THREADS) cache_scatter_kernel(const float* __restrict__ c_tri, const int64_t* __restrict__ c_id,
                                                                        const float* __restrict__ c_std, int64_t n, const uint8_t* __restrict__ flag,
                                                                        const int32_t* __restrict__ chunk_off, float* __restrict__ o_tri,
                                                                        int64_t* __restrict__ o_id, float* __restrict__ o_std) {
    __shared__ int s_warp[MCACHE_THREADS / 32];
    __shared__ int s_src[MCACHE_CHUNK];                 // kept rows of this chunk, in order (chunk-local source index)
    __shared__ int s_run;
    const int64_t base = (int64_t)blockIdx.x * MCACHE_CHUNK;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
#pragma unroll 1
    for (int k = 0; k < MCACHE_CHUNK / MCACHE_THREADS; ++k) {        // stable: sub-chunks in order, block scan inside each
        const int li = k * MCACHE_THREADS + threadIdx.x;
        const int64_t i = base + li;
        const int keep = (i < n) ? (flag[c_id[i]] == 0) : 0;
        int incl = keep;
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += u; }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        int wpre = 0, tot = 0;
        for (int w = 0; w < MCACHE_THREADS / 32; ++w) { const int t = s_warp[w]; if (w < (int)(threadIdx.x >> 5)) wpre += t; tot += t; }
        const int run = s_run;
        if (keep) s_src[run + wpre + incl - 1] = li;
        __syncthreads();
        if (threadIdx.x == 0) s_run = run + tot;
        __syncthreads();
    }
    const int kept = s_run;
    const int64_t dst = chunk_off[blockIdx.x];
    // rows are copied as flat float streams: consecutive threads write consecutive words of the compacted output
    for (int e = threadIdx.x; e < kept * 9; e += MCACHE_THREADS) o_tri[dst * 9 + e] = c_tri[(base + s_src[e / 9]) * 9 + e % 9];
    for (int e = threadIdx.x; e < kept * 3; e += MCACHE_THREADS) o_std[dst * 3 + e] = c_std[(base + s_src[e / 3]) * 3 + e % 3];
    for (int e = threadIdx.x; e < kept; e += MCACHE_THREADS) o_id[dst + e] = c_id[base + s_src[e]];
}

// new triangles: voxel units -> world (map.py:698: vertices * voxel_size + bound_min, two separately rounded torch ops) and append
__global__ void cache_append_kernel(const float* __restrict__ n_tri, const int64_t* __restrict__ n_id, const float* __restrict__ n_std, int64_t n_new,
                                    float vs, float bx, float by, float bz, const int64_t* __restrict__ totals,
                                    float* __restrict__ o_tri, int64_t* __restrict__ o_id, float* __restrict__ o_std) {
    const int64_t dst = totals[0];
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n_new * 9) {
        const int a = (int)(e % 3);
        o_tri[dst * 9 + e] = __fadd_rn(__fmul_rn(n_tri[e], vs), a == 0 ? bx : (a == 1 ? by : bz));
    }
    if (e < n_new * 3) o_std[dst * 3 + e] = n_std[e];
    if (e < n_new) o_id[dst + e] = n_id[e];
}

}  // namespace dif

using namespace dif;

extern "C" {

size_t dif_mesh_cache_scratch_bytes(int64_t n_cells, int64_t n_cache) {
    const int64_t chunks = (n_cache + MCACHE_CHUNK - 1) / MCACHE_CHUNK;
    return align_up((size_t)n_cells) + align_up((size_t)(chunks + 1) * sizeof(int32_t)) + 256;
}

int dif_mesh_cache_merge(const float* cache_tri, const int64_t* cache_id, const float* cache_std, int64_t n_cache,
                         const float* new_tri, const int64_t* new_id, const float* new_std, int64_t n_new,
                         float voxel_size, const float* bound_min, int64_t n_cells,
                         float* out_tri, int64_t* out_id, float* out_std, int64_t* totals_dev,
                         void* persist, size_t persist_bytes, void* stream) {
    if (n_cache < 0 || n_new < 0 || n_cells <= 0 || !totals_dev || !persist || !bound_min) return DIF_E_INVALID;
    if (n_cache > 0 && (!cache_tri || !cache_id || !cache_std)) return DIF_E_INVALID;
    if (n_new > 0 && (!new_tri || !new_id || !new_std)) return DIF_E_INVALID;
    if (n_cache + n_new > 0 && (!out_tri || !out_id || !out_std)) return DIF_E_INVALID;
    if (n_cache + n_new >= (int64_t(1) << 31) / 9) return DIF_E_INVALID;
    if (persist_bytes < dif_mesh_cache_scratch_bytes(n_cells, n_cache)) return DIF_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    Carver c(persist);
    uint8_t* flag = c.take<uint8_t>(n_cells);                       // zero on entry (caller zero-fills once), zero on exit
    const int64_t chunks = (n_cache + MCACHE_CHUNK - 1) / MCACHE_CHUNK;
    int32_t* chunk_count = c.take<int32_t>(chunks + 1);
    const unsigned g_new = (unsigned)((n_new + 255) / 256);
    if (n_new) { cache_flag_kernel<<<g_new, 256, 0, st>>>(new_id, n_new, flag, 1); DIF_COUNT_LAUNCH(1); }
    if (n_cache) {
        cache_count_kernel<<<(unsigned)chunks, MCACHE_THREADS, 0, st>>>(cache_id, n_cache, flag, chunk_count);
        DIF_COUNT_LAUNCH(1);
    }
    cache_scan_kernel<<<1, 1024, 0, st>>>(chunk_count, chunks, n_new, totals_dev);
    DIF_COUNT_LAUNCH(1);
    if (n_cache) {
        cache_scatter_kernel<<<(unsigned)chunks, MCACHE_THREADS, 0, st>>>(cache_tri, cache_id, cache_std, n_cache, flag, chunk_count,
                                                                          out_tri, out_id, out_std);
        DIF_COUNT_LAUNCH(1);
    }
    if (n_new) {
        cache_append_kernel<<<(unsigned)((n_new * 9 + 255) / 256), 256, 0, st>>>(new_tri, new_id, new_std, n_new, voxel_size, bound_min[0],
                                                                                 bound_min[1], bound_min[2], totals_dev, out_tri, out_id, out_std);
        cache_flag_kernel<<<g_new, 256, 0, st>>>(new_id, n_new, flag, 0);
        DIF_COUNT_LAUNCH(2);
    }
    return check_launch("dif_mesh_cache_merge");
}

}  // extern "C"
