// Photometric (RGB odometry) term of the tracker.
//   replaces reference system/ext/imgproc/photometric.cu:3-138 (gradient_xy_kernel, evaluate_fJ) and the torch reduction
//   chain of system/tracker.py:131-172 (compute_rgb_Hg: boolean-mask compaction, robust weights, einsum, 3 host syncs).
//
// Three entry points share ONE per-pixel evaluator (so the op-table results and the fused result agree bit for bit):
//   dif_gradient_xy   - Sobel gradients, NaN border                                  (ext op `gradient_xy`)
//   dif_rgb_odometry  - per-pixel residual f (NaN = invalid) and 6-vector J           (ext op `rgb_odometry`)
//   dif_rgb_linearize - f, J, robust weight and the 6x6 normal equations in one launch, one 44-double result
//                       (what compute_rgb_Hg returns: H, g, energy; plus the valid-pixel count)
//
// HBM-bound streaming: 5 image planes in (28 B/pixel incl. the gradient pair), nothing out in the fused form (the reference
// writes f (4 B) + J (24 B) per pixel and re-reads them several times through torch).  Warp-shuffle + per-CTA partials + last-CTA
// ordered sum, as in icp.cu; deterministic.
//
// Rounding: explicit round-to-nearest intrinsics, with fmaf exactly where nvcc AND ptxas (-fmad=true) fuse the reference source
// (read off the PTX and the SASS of the reference extension built by oracle/build_ref.py; note that ptxas fuses the three
// cross-product rows although the PTX still has separate mul/sub).  oracle/imgproc_oracle.c is the same arithmetic on the CPU.
#include "common.cuh"

namespace dif {

struct PhotoArgs {
    const float* prev_i; const float* prev_d; const float* cur_i; const float* cur_d; const float* dIdxy;   // [h][w], dIdxy [h][w][2]
    int h, w;
    float fx, fy, cx, cy;               // intr (tracker.py:143 passes the level-0 intrinsics at every level)
    float k[9];                         // K R K^-1, row major
    float kt[3];                        // K t
    float min_grad_scale, max_depth_delta;
};

// photometric.cu:24-78.  Returns validity; f and (if want_J) J[6] are the reference's f_val / J_val (J before the sign flip of
// tracker.py:157).
__device__ __forceinline__ bool eval_pixel(const PhotoArgs& a, int v, int u, bool want_J, float& f, float* J) {
    const int64_t px = (int64_t)v * a.w + u;
    const float dIx = a.dIdxy[2 * px], dIy = a.dIdxy[2 * px + 1];
    const float m2 = __fmaf_rn(dIx, dIx, __fmul_rn(dIy, dIy));
    if (m2 < a.min_grad_scale || m2 != m2) return false;
    const float d1 = a.cur_d[px];
    if (d1 != d1) return false;
    const float uf = (float)(unsigned)u, vf = (float)(unsigned)v;
    const float wd = __fmaf_rn(__fadd_rn(a.k[8], __fmaf_rn(a.k[6], uf, __fmul_rn(a.k[7], vf))), d1, a.kt[2]);
    const float xn = __fmaf_rn(__fadd_rn(a.k[2], __fmaf_rn(a.k[0], uf, __fmul_rn(a.k[1], vf))), d1, a.kt[0]);
    const float yn = __fmaf_rn(__fadd_rn(a.k[5], __fmaf_rn(a.k[3], uf, __fmul_rn(a.k[4], vf))), d1, a.kt[1]);
    const int u0 = __float2int_rn(__fdiv_rn(xn, wd));
    const int v0 = __float2int_rn(__fdiv_rn(yn, wd));
    if (!(u0 >= 0 && u0 < a.w && v0 >= 0 && v0 < a.h)) return false;
    const int64_t p0x = (int64_t)v0 * a.w + u0;
    const float d0 = a.prev_d[p0x];
    if (!(d0 == d0 && fabsf(__fsub_rn(wd, d0)) <= a.max_depth_delta && d0 > 0.0f)) return false;
    f = __fsub_rn(a.cur_i[px], a.prev_i[p0x]);
    if (want_J) {
        const float Gx = __fdiv_rn(__fmul_rn(__fsub_rn((float)(unsigned)u0, a.cx), d0), a.fx);
        const float Gy = __fdiv_rn(__fmul_rn(__fsub_rn((float)(unsigned)v0, a.cy), d0), a.fy);
        const float Gz = d0;
        const float p0 = __fdiv_rn(__fmul_rn(a.fx, dIx), Gz);
        const float p1 = __fdiv_rn(__fmul_rn(a.fy, dIy), Gz);
        const float p2 = __fdiv_rn(-__fmaf_rn(p0, Gx, __fmul_rn(p1, Gy)), Gz);
        J[0] = p0; J[1] = p1; J[2] = p2;
        J[3] = __fmaf_rn(Gy, p2, -__fmul_rn(Gz, p1));
        J[4] = __fmaf_rn(Gz, p0, -__fmul_rn(Gx, p2));
        J[5] = __fmaf_rn(p1, Gx, -__fmul_rn(p0, Gy));
    }
    return true;
}

// photometric.cu:3-22
__global__ void gradient_xy_kernel(const float* __restrict__ I, int h, int w, float* __restrict__ out) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y * blockDim.y + threadIdx.y;
    if (u >= w || v >= h) return;
    float gx, gy;
    if (v < 1 || v > h - 2 || u < 1 || u > w - 2) {
        gx = gy = __int_as_float(0x7fc00000);
    } else {
        const float* r0 = I + (int64_t)(v - 1) * w + u; const float* r1 = r0 + w; const float* r2 = r1 + w;
        const float a = r0[-1], b = r0[0], c = r0[1], d = r1[-1], e = r1[1], f = r2[-1], g = r2[0], hh = r2[1];
        gx = __fmul_rn(__fadd_rn(__fmaf_rn(__fsub_rn(e, d), 2.0f, __fsub_rn(c, a)), __fsub_rn(hh, f)), 0.125f);
        gy = __fmul_rn(__fadd_rn(__fmaf_rn(__fsub_rn(g, b), 2.0f, __fsub_rn(f, a)), __fsub_rn(hh, c)), 0.125f);
    }
    reinterpret_cast<float2*>(out)[(int64_t)v * w + u] = make_float2(gx, gy);
}

__global__ void rgb_odometry_kernel(PhotoArgs a, float* __restrict__ f_out, float* __restrict__ J_out) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y * blockDim.y + threadIdx.y;
    if (u >= a.w || v >= a.h) return;
    float f = 0.f, J[6];
    const bool ok = eval_pixel(a, v, u, J_out != nullptr, f, J);
    const int64_t px = (int64_t)v * a.w + u;
    f_out[px] = ok ? f : __int_as_float(0x7fc00000);
    if (ok && J_out) {
#pragma unroll
        for (int j = 0; j < 6; ++j) J_out[6 * px + j] = J[j];        // invalid pixels stay untouched, like the reference's torch::empty
    }
}

constexpr int PH_THREADS = 256;
constexpr int PH_VALS = 32;            // 21 upper-tri H + 6 g + E + M, padded

__device__ __forceinline__ double warp_sum_d(double v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// robust: 0 none, 1 huber, 2 tukey (tracker.py:58-71)
// kkt_dev (nullable): K R K^-1 [9] and K t [3] read from device memory instead of the launch arguments - the device-driven
// Gauss-Newton loop (gn.cu) rewrites them after every pose update, so the host never touches the pose between iterations.
__global__ void __launch_bounds__(PH_THREADS) rgb_linearize_kernel(PhotoArgs a, const float* __restrict__ kkt_dev, int robust, float robust_k,
                                                                   float weight, int want_grad,
                                                                   double* __restrict__ partials, unsigned int* __restrict__ done_counter,
                                                                   double* __restrict__ out) {
    __shared__ double s_part[PH_THREADS / 32][PH_VALS];
    __shared__ bool is_last;
    if (kkt_dev) {
#pragma unroll
        for (int i = 0; i < 9; ++i) a.k[i] = __ldcg(kkt_dev + i);
#pragma unroll
        for (int i = 0; i < 3; ++i) a.kt[i] = __ldcg(kkt_dev + 9 + i);
    }
    float acc[29];
#pragma unroll
    for (int j = 0; j < 29; ++j) acc[j] = 0.f;
    const int64_t n_px = (int64_t)a.h * a.w;
    for (int64_t px = (int64_t)blockIdx.x * PH_THREADS + threadIdx.x; px < n_px; px += (int64_t)gridDim.x * PH_THREADS) {
        const int v = (int)(px / a.w), u = (int)(px % a.w);
        float f = 0.f, Jr[6];
        if (!eval_pixel(a, v, u, want_grad != 0, f, Jr)) continue;
        float w = 1.f;
        if (robust == 1) { const float af = fabsf(f); if (af > robust_k) w = robust_k / af; }
        else if (robust == 2) { const float q = f / robust_k; const float t = 1.f - q * q; w = fabsf(f) <= robust_k ? t * t : 0.f; }
        const float wf = f * w;
        acc[27] += f * wf;                                           // sum_error (:166)
        acc[28] += 1.f;
        if (want_grad) {
            float J[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) J[j] = -Jr[j];                // "the derivative computed is actually for -xi" (:157)
            int k = 0;
#pragma unroll
            for (int p = 0; p < 6; ++p)
#pragma unroll
                for (int q = p; q < 6; ++q) acc[k++] += (J[p] * w) * J[q];    // H = sum JW^T J (:168)
#pragma unroll
            for (int p = 0; p < 6; ++p) acc[21 + p] += J[p] * wf;              // g = sum J Wf   (:169)
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 29; ++j) { const double s = warp_sum_d((double)acc[j]); if (lane == 0) s_part[warp][j] = s; }
    __syncthreads();
    if (threadIdx.x < 29) {
        double t = 0.0;
        for (int w8 = 0; w8 < PH_THREADS / 32; ++w8) t += s_part[w8][threadIdx.x];
        partials[(size_t)blockIdx.x * PH_VALS + threadIdx.x] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(done_counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (is_last && threadIdx.x < 32) {
        __threadfence();
        double tot = 0.0;
        if (threadIdx.x < 29) for (unsigned b = 0; b < gridDim.x; ++b) tot += __ldcg(partials + (size_t)b * PH_VALS + threadIdx.x);
        const double M = __shfl_sync(0xffffffffu, tot, 28);
        const double scale = M > 0.0 ? (double)weight / M : 0.0;     // error_scale = 1/M * weight (:165)
        if (threadIdx.x < 21) {
            int p = 0, rem = threadIdx.x;
            while (rem >= 6 - p) { rem -= 6 - p; ++p; }
            const int q = p + rem;
            out[p * 6 + q] = tot * scale; out[q * 6 + p] = tot * scale;
        } else if (threadIdx.x < 27) out[36 + threadIdx.x - 21] = tot * scale;
        else if (threadIdx.x == 27) out[42] = tot * scale;
        else if (threadIdx.x == 28) out[43] = M;
        if (threadIdx.x == 0) *done_counter = 0u;
    }
}

static int fill_args(PhotoArgs& a, const float* prev_i, const float* prev_d, const float* cur_i, const float* cur_d, const float* dIdxy, int h, int w,
                     const float* intr, const float* krkinv, const float* kt, float min_grad_scale, float max_depth_delta) {
    if (!prev_i || !prev_d || !cur_i || !cur_d || !dIdxy || !intr || !krkinv || !kt || h <= 0 || w <= 0) return DIF_E_INVALID;
    a.prev_i = prev_i; a.prev_d = prev_d; a.cur_i = cur_i; a.cur_d = cur_d; a.dIdxy = dIdxy; a.h = h; a.w = w;
    a.fx = intr[0]; a.fy = intr[1]; a.cx = intr[2]; a.cy = intr[3];
    for (int i = 0; i < 9; ++i) a.k[i] = krkinv[i];
    for (int i = 0; i < 3; ++i) a.kt[i] = kt[i];
    a.min_grad_scale = min_grad_scale; a.max_depth_delta = max_depth_delta;
    return DIF_OK;
}

size_t rgb_scratch_bytes() { return align_up((size_t)DIF_NUM_SMS * 4 * PH_VALS * sizeof(double)) + 256; }

// krkinv/kt: host values, or (kkt_dev != NULL) 12 floats in device memory read by the kernel
int rgb_launch(const float* prev_i, const float* prev_d, const float* cur_i, const float* cur_d, const float* dIdxy, int h, int w,
               const float* intr, const float* krkinv, const float* kt, const float* kkt_dev, float min_grad_scale, float max_depth_delta,
               int robust_kind, float robust_k, float weight, int want_grad, void* scratch, size_t scratch_bytes, double* out_dev, cudaStream_t st) {
    static const float zero12[12] = {};
    PhotoArgs a;
    const int rc = fill_args(a, prev_i, prev_d, cur_i, cur_d, dIdxy, h, w, intr, kkt_dev ? zero12 : krkinv, kkt_dev ? zero12 : kt,
                             min_grad_scale, max_depth_delta);
    if (rc || !scratch || !out_dev || robust_kind < 0 || robust_kind > 2) return DIF_E_INVALID;
    if (scratch_bytes < rgb_scratch_bytes()) return DIF_E_WORKSPACE;
    Carver c(scratch);                                               // zero-filled once by the caller; the counter is left zeroed
    double* partials = c.take<double>((size_t)DIF_NUM_SMS * 4 * PH_VALS);
    unsigned int* counter = c.take<unsigned int>(1);
    const int64_t n_px = (int64_t)h * w;
    int64_t grid = (n_px + PH_THREADS * 4 - 1) / (PH_THREADS * 4);   // >= 4 pixels per thread
    if (grid > DIF_NUM_SMS * 4) grid = DIF_NUM_SMS * 4;
    if (grid < 1) grid = 1;
    rgb_linearize_kernel<<<(unsigned)grid, PH_THREADS, 0, st>>>(a, kkt_dev, robust_kind, robust_k, weight, want_grad, partials, counter, out_dev);
    DIF_COUNT_LAUNCH(1);
    return check_launch("rgb_linearize_kernel");
}

}  // namespace dif

using namespace dif;

extern "C" {

int dif_gradient_xy(const float* intensity, int h, int w, float* out_grad, void* stream) {
    if (!intensity || !out_grad || h <= 0 || w <= 0) return DIF_E_INVALID;
    const dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8);
    gradient_xy_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(intensity, h, w, out_grad);
    DIF_COUNT_LAUNCH(1);
    return check_launch("gradient_xy_kernel");
}

int dif_rgb_odometry(const float* prev_intensity, const float* prev_depth, const float* cur_intensity, const float* cur_depth,
                     const float* cur_dIdxy, int h, int w, const float* intr, const float* krkinv, const float* kt,
                     float min_grad_scale, float max_depth_delta, float* f_out, float* J_out, void* stream) {
    PhotoArgs a;
    const int rc = fill_args(a, prev_intensity, prev_depth, cur_intensity, cur_depth, cur_dIdxy, h, w, intr, krkinv, kt, min_grad_scale, max_depth_delta);
    if (rc || !f_out) return DIF_E_INVALID;
    const dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8);
    rgb_odometry_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(a, f_out, J_out);
    DIF_COUNT_LAUNCH(1);
    return check_launch("rgb_odometry_kernel");
}

size_t dif_rgb_scratch_bytes(void) { return rgb_scratch_bytes(); }

int dif_rgb_linearize(const float* prev_intensity, const float* prev_depth, const float* cur_intensity, const float* cur_depth,
                      const float* cur_dIdxy, int h, int w, const float* intr, const float* krkinv, const float* kt,
                      float min_grad_scale, float max_depth_delta, int robust_kind, float robust_k, float weight, int want_grad,
                      void* scratch, size_t scratch_bytes, double* out_dev, void* stream) {
    return rgb_launch(prev_intensity, prev_depth, cur_intensity, cur_depth, cur_dIdxy, h, w, intr, krkinv, kt, nullptr, min_grad_scale,
                      max_depth_delta, robust_kind, robust_k, weight, want_grad, scratch, scratch_bytes, out_dev, (cudaStream_t)stream);
}

}  // extern "C"
