// Frame pre-processing in front of the fusion path: depth unprojection and the 2 cm box filter.
//   replaces reference system/ext/imgproc/imgproc.cu:5-44 (unproject_depth) and system/tracker.py:13-23 (point_box_filter:
//   min/max reductions, floor, a sort-based torch.unique with inverse, two torch_scatter.scatter_mean, one host sync for n_xyz).
//
// Box filter without a sort: the cells a frame touches are marked in a bitmap laid over the frame's bounding box (1 bit per
// 2 cm cell: ~2 MB for a 6 x 4 x 5 m view), an ordered popcount scan of the bitmap gives every occupied cell its rank in
// ascending key order - exactly the order torch.unique(sorted=True) produces - and points are accumulated into their rank with
// float atomics.  Everything is sized on the device; the caller reads back one count.  HBM-bound: 24 B in per point, bitmap
// traffic, 24 B out per occupied cell.
#include "common.cuh"

namespace dif {

// ------------------------------------------------------------------------------------------------ unproject_depth
// imgproc.cu:5-24: no FMA contraction in the reference's PTX (sub, div.rn, mul).  Invalid pixels: the reference writes only
// x = NaN and leaves y, z uninitialised; here all three are NaN.
__global__ void unproject_depth_kernel(const float* __restrict__ depth, int h, int w, float fx, float fy, float cx, float cy, float* __restrict__ pc) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y * blockDim.y + threadIdx.y;
    if (u >= w || v >= h) return;
    const int64_t px = (int64_t)v * w + u;
    const float d = depth[px];
    float x, y, z;
    if (d == d) {
        x = __fmul_rn(__fdiv_rn(__fsub_rn((float)(unsigned)u, cx), fx), d);
        y = __fmul_rn(__fdiv_rn(__fsub_rn((float)(unsigned)v, cy), fy), d);
        z = d;
    } else {
        x = y = z = __int_as_float(0x7fc00000);
    }
    pc[3 * px] = x; pc[3 * px + 1] = y; pc[3 * px + 2] = z;
}

// ------------------------------------------------------------------------------------------------ point_box_filter
struct BoxState {                       // lives in the scratch buffer
    float mn[3], mx[3];                 // min / max of the points (tracker.py:15-16 before the half-voxel margin)
    int n[3];                           // n_x, n_y, n_z (:18)
    int overflow;                       // bounding box needs more bitmap words than the scratch holds
    unsigned long long n_words;
    int n_cells;                        // occupied cells = output rows
};

__device__ __forceinline__ void atomic_min_f(float* a, float v) {
    int old = __float_as_int(*a);
    while (v < __int_as_float(old)) { const int assumed = old; old = atomicCAS((int*)a, assumed, __float_as_int(v)); if (old == assumed) break; }
}
__device__ __forceinline__ void atomic_max_f(float* a, float v) {
    int old = __float_as_int(*a);
    while (v > __int_as_float(old)) { const int assumed = old; old = atomicCAS((int*)a, assumed, __float_as_int(v)); if (old == assumed) break; }
}

__global__ void box_init_kernel(BoxState* s) {
    for (int k = 0; k < 3; ++k) { s->mn[k] = __int_as_float(0x7f800000); s->mx[k] = __int_as_float(0xff800000); }
    s->overflow = 0; s->n_cells = 0; s->n_words = 0;
}

__global__ void box_minmax_kernel(const float* __restrict__ p, int n, BoxState* s) {
    float mn[3] = {__int_as_float(0x7f800000), __int_as_float(0x7f800000), __int_as_float(0x7f800000)};
    float mx[3] = {__int_as_float(0xff800000), __int_as_float(0xff800000), __int_as_float(0xff800000)};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
#pragma unroll
        for (int k = 0; k < 3; ++k) { const float v = p[3 * i + k]; mn[k] = fminf(mn[k], v); mx[k] = fmaxf(mx[k], v); }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        for (int o = 16; o; o >>= 1) { mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o)); mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o)); }
        if ((threadIdx.x & 31) == 0) { atomic_min_f(&s->mn[k], mn[k]); atomic_max_f(&s->mx[k], mx[k]); }
    }
}

// tracker.py:15-18 in torch's fp32 arithmetic: bound = min -/+ voxel*0.5 (python double product, rounded to fp32 when it meets the
// tensor); n = floor((max_bound - min_bound) / voxel) + 16.  The division by the python scalar is a TRUE fp32 division here, as
// torch evaluates it on CPU tensors - the oracle's arithmetic (the CUDA build of torch multiplies by the fp32 reciprocal, which
// can move a point that sits exactly on a cell face; see DESIGN.md "rounding of scalar divisions").
__global__ void box_dims_kernel(BoxState* s, float voxel, float half_voxel, unsigned long long max_words) {
    unsigned long long cells = 1;
    for (int k = 0; k < 3; ++k) {
        const float lo = __fsub_rn(s->mn[k], half_voxel), hi = __fadd_rn(s->mx[k], half_voxel);
        s->mn[k] = lo; s->mx[k] = hi;
        const long long nk = (long long)floorf(__fdiv_rn(__fsub_rn(hi, lo), voxel)) + 16;
        s->n[k] = (int)nk;
        cells *= (unsigned long long)(nk > 0 ? nk : 1);
    }
    const unsigned long long words = (cells + 31) / 32;
    s->n_words = words;
    if (words > max_words) { s->overflow = 1; s->n_words = 0; }
}

__device__ __forceinline__ long long box_key(const BoxState* s, const float* p, float voxel) {
    const long long x = (long long)floorf(__fdiv_rn(__fsub_rn(p[0], s->mn[0]), voxel));
    const long long y = (long long)floorf(__fdiv_rn(__fsub_rn(p[1], s->mn[1]), voxel));
    const long long z = (long long)floorf(__fdiv_rn(__fsub_rn(p[2], s->mn[2]), voxel));
    return x + y * s->n[0] + z * (long long)s->n[0] * s->n[1];                        // tracker.py:19
}

__global__ void box_mark_kernel(const float* __restrict__ p, int n, const BoxState* __restrict__ s, float voxel, uint32_t* __restrict__ bitmap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || s->overflow) return;
    const long long key = box_key(s, p + 3 * i, voxel);
    atomicOr(bitmap + (key >> 5), 1u << (key & 31));
}

// ordered scan of the per-word popcounts: word_rank[w] = number of occupied cells in words < w.  One CTA per 4096 words writes its
// sum, a single CTA scans the sums, the third pass adds them back.
constexpr int BOX_SCAN_T = 256, BOX_SCAN_W = 4096;

__global__ void __launch_bounds__(BOX_SCAN_T) box_scan1_kernel(const uint32_t* __restrict__ bitmap, const BoxState* __restrict__ s,
                                                               uint32_t* __restrict__ word_rank, uint32_t* __restrict__ chunk_sum) {
    __shared__ uint32_t s_warp[BOX_SCAN_T / 32];
    __shared__ uint32_t s_run;
    const unsigned long long n_words = s->n_words;
    const unsigned long long base = (unsigned long long)blockIdx.x * BOX_SCAN_W;
    if (base >= n_words) return;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (int k = 0; k < BOX_SCAN_W / BOX_SCAN_T; ++k) {
        const unsigned long long w = base + (unsigned long long)k * BOX_SCAN_T + threadIdx.x;
        const uint32_t c = w < n_words ? __popc(bitmap[w]) : 0;
        uint32_t incl = c;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += u; }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint32_t wpre = 0, tot = 0;
        for (int q = 0; q < BOX_SCAN_T / 32; ++q) { const uint32_t t = s_warp[q]; if (q < (int)(threadIdx.x >> 5)) wpre += t; tot += t; }
        const uint32_t run = s_run;
        if (w < n_words) word_rank[w] = run + wpre + incl - c;
        __syncthreads();
        if (threadIdx.x == 0) s_run = run + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) chunk_sum[blockIdx.x] = s_run;
}

__global__ void __launch_bounds__(1024) box_scan2_kernel(uint32_t* __restrict__ chunk_sum, BoxState* __restrict__ s, int32_t* __restrict__ n_out) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const unsigned long long n_chunks = (s->n_words + BOX_SCAN_W - 1) / BOX_SCAN_W;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (unsigned long long c0 = 0; c0 < n_chunks; c0 += 1024) {
        const unsigned long long c = c0 + threadIdx.x;
        const uint32_t v = c < n_chunks ? chunk_sum[c] : 0;
        uint32_t incl = v;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += u; }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint32_t wpre = 0, tot = 0;
        for (int q = 0; q < 32; ++q) { const uint32_t t = s_warp[q]; if (q < (int)(threadIdx.x >> 5)) wpre += t; tot += t; }
        const uint32_t carry = s_carry;
        if (c < n_chunks) chunk_sum[c] = carry + wpre + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) { s->n_cells = (int)s_carry; n_out[0] = s->overflow ? -1 : (int)s_carry; }
}

// rank of the point's cell -> accumulate.  sums: [n_points rows max][8] = (x, y, z, nx, ny, nz, count, -), zero-filled by the caller side
__global__ void box_accumulate_kernel(const float* __restrict__ p, const float* __restrict__ nr, int n, const BoxState* __restrict__ s, float voxel,
                                      const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ word_rank,
                                      const uint32_t* __restrict__ chunk_sum, float* __restrict__ sums) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || s->overflow) return;
    const long long key = box_key(s, p + 3 * i, voxel);
    const long long w = key >> 5;
    const uint32_t rank = chunk_sum[w / BOX_SCAN_W] + word_rank[w] + __popc(bitmap[w] & ((1u << (key & 31)) - 1u));
    float* d = sums + (size_t)rank * 8;
    atomicAdd(d + 0, p[3 * i]); atomicAdd(d + 1, p[3 * i + 1]); atomicAdd(d + 2, p[3 * i + 2]);
    atomicAdd(d + 3, nr[3 * i]); atomicAdd(d + 4, nr[3 * i + 1]); atomicAdd(d + 5, nr[3 * i + 2]);
    atomicAdd(d + 6, 1.0f);
}

// mean = sum / count (torch_scatter.scatter_mean: sum, then division by the clamped count); clears the bitmap words and the sum
// rows this frame touched, so the scratch is all-zero again when the call returns (self-cleaning)
__global__ void box_finalize_kernel(const BoxState* __restrict__ s, float* __restrict__ sums, float* __restrict__ out_p, float* __restrict__ out_n) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= s->n_cells) return;
    float* d = sums + (size_t)r * 8;
    const float c = fmaxf(d[6], 1.0f);
#pragma unroll
    for (int k = 0; k < 3; ++k) { out_p[3 * r + k] = __fdiv_rn(d[k], c); out_n[3 * r + k] = __fdiv_rn(d[3 + k], c); }
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] = 0.f;
}

__global__ void box_unmark_kernel(const float* __restrict__ p, int n, const BoxState* __restrict__ s, float voxel, uint32_t* __restrict__ bitmap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || s->overflow) return;
    const long long key = box_key(s, p + 3 * i, voxel);
    bitmap[key >> 5] = 0u;
}

constexpr unsigned long long BOX_MAX_WORDS = 1ull << 24;           // 2^29 cells of 2 cm = a 16 m cube: 64 MB bitmap + 64 MB ranks

}  // namespace dif

using namespace dif;

extern "C" {

int dif_unproject_depth(const float* depth, int h, int w, float fx, float fy, float cx, float cy, float* pc_out, void* stream) {
    if (!depth || !pc_out || h <= 0 || w <= 0) return DIF_E_INVALID;
    const dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8);
    unproject_depth_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(depth, h, w, fx, fy, cx, cy, pc_out);
    DIF_COUNT_LAUNCH(1);
    return check_launch("unproject_depth_kernel");
}

size_t dif_box_filter_scratch_bytes(int64_t max_points, int64_t max_cells) {
    const size_t words = (size_t)((max_cells + 31) / 32);
    return 256 + align_up(words * 4) * 2 + align_up((words / BOX_SCAN_W + 2) * 4) + align_up((size_t)max_points * 8 * sizeof(float)) + 256;
}

int dif_point_box_filter(const float* points, const float* normals, int64_t n, float voxel_size, int64_t max_cells,
                         float* out_points, float* out_normals, int32_t* n_out_dev, void* scratch, size_t scratch_bytes, void* stream) {
    if (n < 0 || n >= (int64_t(1) << 31) || !n_out_dev || !scratch || !(voxel_size > 0.f) || max_cells <= 0) return DIF_E_INVALID;
    if (n > 0 && (!points || !normals || !out_points || !out_normals)) return DIF_E_INVALID;
    if (scratch_bytes < dif_box_filter_scratch_bytes(n, max_cells)) return DIF_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned long long words = (unsigned long long)((max_cells + 31) / 32);
    if (words > BOX_MAX_WORDS) return DIF_E_INVALID;
    Carver c(scratch);                                   // bitmap and sums are zero on entry (caller zero-fills once) and on exit
    BoxState* s = c.take<BoxState>(1);
    uint32_t* bitmap = c.take<uint32_t>(words);
    uint32_t* word_rank = c.take<uint32_t>(words);
    uint32_t* chunk_sum = c.take<uint32_t>(words / BOX_SCAN_W + 2);
    float* sums = c.take<float>((size_t)n * 8);
    if (n == 0) { cudaMemsetAsync(n_out_dev, 0, sizeof(int32_t), st); return check_launch("dif_point_box_filter"); }
    const int ni = (int)n;
    const unsigned gp = (unsigned)((n + 255) / 256);
    // voxel * 0.5 is a python double product in the reference; it is rounded to fp32 when it is combined with the fp32 tensor
    const float half_voxel = (float)((double)voxel_size * 0.5);
    box_init_kernel<<<1, 1, 0, st>>>(s);
    box_minmax_kernel<<<DIF_NUM_SMS, 256, 0, st>>>(points, ni, s);
    box_dims_kernel<<<1, 1, 0, st>>>(s, voxel_size, half_voxel, words);
    box_mark_kernel<<<gp, 256, 0, st>>>(points, ni, s, voxel_size, bitmap);
    const unsigned chunks = (unsigned)((words + BOX_SCAN_W - 1) / BOX_SCAN_W);
    box_scan1_kernel<<<chunks, BOX_SCAN_T, 0, st>>>(bitmap, s, word_rank, chunk_sum);
    box_scan2_kernel<<<1, 1024, 0, st>>>(chunk_sum, s, n_out_dev);
    box_accumulate_kernel<<<gp, 256, 0, st>>>(points, normals, ni, s, voxel_size, bitmap, word_rank, chunk_sum, sums);
    box_finalize_kernel<<<gp, 256, 0, st>>>(s, sums, out_points, out_normals);
    box_unmark_kernel<<<gp, 256, 0, st>>>(points, ni, s, voxel_size, bitmap);
    DIF_COUNT_LAUNCH(9);
    return check_launch("dif_point_box_filter");
}

}  // extern "C"
