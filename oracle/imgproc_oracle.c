/* ORACLE (test infrastructure, not product code): scalar CPU restatement of the reference's image-side kernels.
 *
 *   reference: pytorch/system/ext/imgproc/photometric.cu
 *       gradient_xy_kernel :3-22   -> dif_oracle_gradient_xy()
 *       evaluate_fJ        :24-78  -> dif_oracle_rgb_odometry()
 *   reference: pytorch/system/ext/imgproc/imgproc.cu
 *       unproject_depth_kernel :5-24 -> dif_oracle_unproject_depth()
 *
 * Parity status: PINNED.  tests/golden/make_golden_gpu.py executes the unmodified reference extension (oracle/_ref/imgproc,
 * built from /root/reference by oracle/build_ref.py) on a B200 and stores inputs + outputs in tests/golden/ref_ext_photo.npz /
 * ref_ext_unproject.npz; tests/test_oracle_imgproc.py checks this file against them bit for bit.
 *
 * Rounding: -ffp-contract=off; fmaf() exactly where nvcc + ptxas (default -fmad=true) emit FFMA for the reference source
 * (read off the PTX and SASS of the reference build; ptxas additionally fuses the three cross-product rows of J).
 * Only tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke() may load this.
 */
#include <math.h>
#include <stdint.h>

/* photometric.cu:3-22 */
void dif_oracle_gradient_xy(const float* I, int h, int w, float* out /*[h][w][2]*/) {
    for (int v = 0; v < h; ++v) for (int u = 0; u < w; ++u) {
        float* o = out + 2 * ((int64_t)v * w + u);
        if (v < 1 || v > h - 2 || u < 1 || u > w - 2) { o[0] = o[1] = NAN; continue; }
#define PX(dv, du) I[(int64_t)(v + (dv)) * w + (u + (du))]
        const float u_d1 = PX(-1, 1) - PX(-1, -1), u_d2 = PX(0, 1) - PX(0, -1), u_d3 = PX(1, 1) - PX(1, -1);
        o[0] = (fmaf(u_d2, 2.0f, u_d1) + u_d3) * 0.125f;
        const float v_d1 = PX(1, -1) - PX(-1, -1), v_d2 = PX(1, 0) - PX(-1, 0), v_d3 = PX(1, 1) - PX(-1, 1);
        o[1] = (fmaf(v_d2, 2.0f, v_d1) + v_d3) * 0.125f;
#undef PX
    }
}

/* photometric.cu:24-78.  f_out NaN = rejected; J_out (may be NULL) rows of rejected pixels are left untouched. */
void dif_oracle_rgb_odometry(const float* prev_i, const float* prev_d, const float* cur_i, const float* cur_d, const float* dIdxy,
                             int h, int w, const float* intr, const float* k, const float* kt,
                             float min_grad_scale, float max_depth_delta, float* f_out, float* J_out) {
    const float fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3];
    for (int v = 0; v < h; ++v) for (int u = 0; u < w; ++u) {
        const int64_t px = (int64_t)v * w + u;
        f_out[px] = NAN;
        const float dIx = dIdxy[2 * px], dIy = dIdxy[2 * px + 1];
        const float m2 = fmaf(dIx, dIx, dIy * dIy);
        if (m2 < min_grad_scale || isnan(m2)) continue;
        const float d1 = cur_d[px];
        if (isnan(d1)) continue;
        const float uf = (float)(uint32_t)u, vf = (float)(uint32_t)v;
        const float wd = fmaf(k[8] + fmaf(k[6], uf, k[7] * vf), d1, kt[2]);
        const float xn = fmaf(k[2] + fmaf(k[0], uf, k[1] * vf), d1, kt[0]);
        const float yn = fmaf(k[5] + fmaf(k[3], uf, k[4] * vf), d1, kt[1]);
        const float qx = xn / wd, qy = yn / wd;
        /* __float2int_rn: round half to even, saturating; NaN -> 0 */
        const int u0 = isnan(qx) ? 0 : (qx >= 2147483648.0f ? 2147483647 : (qx <= -2147483648.0f ? (-2147483647 - 1) : (int)nearbyintf(qx)));
        const int v0 = isnan(qy) ? 0 : (qy >= 2147483648.0f ? 2147483647 : (qy <= -2147483648.0f ? (-2147483647 - 1) : (int)nearbyintf(qy)));
        if (!(u0 >= 0 && u0 < w && v0 >= 0 && v0 < h)) continue;
        const int64_t p0x = (int64_t)v0 * w + u0;
        const float d0 = prev_d[p0x];
        if (!(!isnan(d0) && fabsf(wd - d0) <= max_depth_delta && d0 > 0.0f)) continue;
        f_out[px] = cur_i[px] - prev_i[p0x];
        if (J_out) {
            const float Gx = (((float)(uint32_t)u0 - cx) * d0) / fx;
            const float Gy = (((float)(uint32_t)v0 - cy) * d0) / fy;
            const float Gz = d0;
            const float p0 = (fx * dIx) / Gz;
            const float p1 = (fy * dIy) / Gz;
            const float p2 = -fmaf(p0, Gx, p1 * Gy) / Gz;
            float* J = J_out + 6 * px;
            J[0] = p0; J[1] = p1; J[2] = p2;
            J[3] = fmaf(Gy, p2, -(Gz * p1));
            J[4] = fmaf(Gz, p0, -(Gx * p2));
            J[5] = fmaf(p1, Gx, -(p0 * Gy));
        }
    }
}

/* imgproc.cu:5-24: pc[v][u] = ((u - cx) / fx * d, (v - cy) / fy * d, d); NaN depth -> x = NaN (y, z untouched) */
void dif_oracle_unproject_depth(const float* depth, int h, int w, float fx, float fy, float cx, float cy, float* pc /*[h][w][3]*/) {
    for (int v = 0; v < h; ++v) for (int u = 0; u < w; ++u) {
        const int64_t px = (int64_t)v * w + u;
        const float d = depth[px];
        if (!isnan(d)) {
            pc[3 * px + 0] = (((float)(uint32_t)u - cx) / fx) * d;
            pc[3 * px + 1] = (((float)(uint32_t)v - cy) / fy) * d;
            pc[3 * px + 2] = d;
        } else {
            pc[3 * px + 0] = NAN;
        }
    }
}
