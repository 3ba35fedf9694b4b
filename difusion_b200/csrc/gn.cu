// Device-driven Gauss-Newton pose refinement: the loop of reference system/tracker.py:220-283 (gauss_newton) with the pose, the
// energy test, the 6x6 solve and the SE(3) update kept ON THE DEVICE.
//
// The reference (and the Python mirror's host loop) reads H, g and the energy of every term back to the host, sums them, solves
// and rebuilds the pose in numpy: per iteration 3 host syncs per term plus ~100 us of Python between kernels.  Here one iteration
// is: the term kernels (dif_icp_linearize / dif_rgb_linearize, reading the current pose from a device block) followed by
// gn_update_kernel (one thread, fp64): sum the terms, compare the energy with the previous iterate (tracker.py:263-268), solve
// H xi = -g (partial-pivot elimination), delta <- exp(xi) . delta (utils/motion_util.py:205-229,277-278), refresh the fp32 pose
// blocks the term kernels read, and post {sequence, continue/break, delta, energy} to a pinned host mailbox.  The host thread (inside
// this one C call) only spins on that mailbox word to decide whether to enqueue the next iteration: no stream synchronisation, no
// readback copies, no Python between iterations.
#include "common.cuh"
#include "icp_args.cuh"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

namespace dif {
int icp_launch(const dif_map_view* map, const void* decoder_prepared, const float* obs_xyz, int obs_stride, int64_t n, const float* pose_host,
               const dif_frame_params* frame_dev, float huber_k, int want_grad, void* scratch, size_t scratch_sz, double* out_dev, cudaStream_t st);
int rgb_launch(const float* prev_i, const float* prev_d, const float* cur_i, const float* cur_d, const float* dIdxy, int h, int w,
               const float* intr, const float* krkinv, const float* kt, const float* kkt_dev, float min_grad_scale, float max_depth_delta,
               int robust_kind, float robust_k, float weight, int want_grad, void* scratch, size_t scratch_bytes, double* out_dev, cudaStream_t st);
size_t rgb_scratch_bytes();

// device state of one Gauss-Newton run
struct GnState {
    double delta[12];            // current delta pose: R[9] row major, t[3]
    double last_delta[12];       // the iterate last_energy belongs to (tracker.py:267)
    double last_energy;
    double pad;
    dif_frame_params frame;      // what the ICP kernel reads: n_points + fp32 (R_last, t_last, R_delta, t_delta)
    float kkt[12];               // what the photometric kernel reads: K R_delta K^-1 [9], K t_delta [3]
};

struct GnCalib { double K[9], Kinv[9]; };

__device__ inline void gn_publish_pose(GnState* s, const GnCalib& c) {
    for (int i = 0; i < 9; ++i) s->frame.pose[12 + i] = (float)s->delta[i];
    for (int i = 0; i < 3; ++i) s->frame.pose[21 + i] = (float)s->delta[9 + i];
    double kr[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { double a = 0; for (int k = 0; k < 3; ++k) a += c.K[3 * i + k] * s->delta[3 * k + j]; kr[3 * i + j] = a; }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { double a = 0; for (int k = 0; k < 3; ++k) a += kr[3 * i + k] * c.Kinv[3 * k + j]; s->kkt[3 * i + j] = (float)a; }
    for (int i = 0; i < 3; ++i) { double a = 0; for (int k = 0; k < 3; ++k) a += c.K[3 * i + k] * s->delta[9 + k]; s->kkt[9 + i] = (float)a; }
}

struct GnInit { double last[12], delta[12]; int n_points; };

__global__ void gn_init_kernel(GnState* s, GnInit in, GnCalib c) {
    if (threadIdx.x != 0) return;
    for (int i = 0; i < 12; ++i) { s->delta[i] = in.delta[i]; s->last_delta[i] = in.delta[i]; }
    s->last_energy = INFINITY;
    s->frame.n_points = in.n_points; s->frame.seq = 0; s->frame.reserved[0] = s->frame.reserved[1] = 0;
    for (int i = 0; i < 12; ++i) s->frame.pose[i] = (float)in.last[i];
    gn_publish_pose(s, c);
}

// H xi = -g by Gaussian elimination with partial pivoting (what numpy.linalg.solve / LAPACK gesv does); false if singular
__host__ __device__ inline bool gn_solve6(const double* H, const double* g, double* xi) {
    double A[6][7];
    for (int i = 0; i < 6; ++i) { for (int j = 0; j < 6; ++j) A[i][j] = H[6 * i + j]; A[i][6] = -g[i]; }
    for (int c = 0; c < 6; ++c) {
        int p = c; double best = fabs(A[c][c]);
        for (int r = c + 1; r < 6; ++r) if (fabs(A[r][c]) > best) { best = fabs(A[r][c]); p = r; }
        if (!(best > 0.0)) return false;
        if (p != c) for (int j = c; j < 7; ++j) { const double t = A[c][j]; A[c][j] = A[p][j]; A[p][j] = t; }
        for (int r = c + 1; r < 6; ++r) {
            const double f = A[r][c] / A[c][c];
            for (int j = c; j < 7; ++j) A[r][j] -= f * A[c][j];
        }
    }
    for (int i = 5; i >= 0; --i) {
        double a = A[i][6];
        for (int j = i + 1; j < 6; ++j) a -= A[i][j] * xi[j];
        xi[i] = a / A[i][i];
    }
    return true;
}

// delta <- from_twist(xi) . delta   (motion_util.py:205-229: R = exp(phi), t = J_l(phi) rho; :277-278 composition)
__host__ __device__ inline void gn_apply_twist(const double* xi, double* d) {
    const double rho[3] = {xi[0], xi[1], xi[2]}, phi[3] = {xi[3], xi[4], xi[5]};
    const double angle = sqrt(phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2]);
    double R[9], J[9];
    if (angle <= 1e-8) {                                       // np.isclose(angle, 0.)
        const double W[9] = {0, -phi[2], phi[1], phi[2], 0, -phi[0], -phi[1], phi[0], 0};
        for (int i = 0; i < 9; ++i) { R[i] = W[i] + (i % 4 == 0 ? 1.0 : 0.0); J[i] = 0.5 * W[i] + (i % 4 == 0 ? 1.0 : 0.0); }
    } else {
        const double a[3] = {phi[0] / angle, phi[1] / angle, phi[2] / angle};
        const double s = sin(angle), c = cos(angle);
        const double W[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double w2 = 0; for (int k = 0; k < 3; ++k) w2 += W[3 * i + k] * W[3 * k + j];
                const double I = i == j ? 1.0 : 0.0;
                R[3 * i + j] = I + s * W[3 * i + j] + (1.0 - c) * w2;
                J[3 * i + j] = (s / angle) * I + (1.0 - s / angle) * a[i] * a[j] + ((1.0 - c) / angle) * W[3 * i + j];
            }
    }
    double t[3], Rn[9], tn[3];
    for (int i = 0; i < 3; ++i) t[i] = J[3 * i] * rho[0] + J[3 * i + 1] * rho[1] + J[3 * i + 2] * rho[2];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) { double v = 0; for (int k = 0; k < 3; ++k) v += R[3 * i + k] * d[3 * k + j]; Rn[3 * i + j] = v; }
        tn[i] = R[3 * i] * d[9] + R[3 * i + 1] * d[10] + R[3 * i + 2] * d[11] + t[i];
    }
    for (int i = 0; i < 9; ++i) d[i] = Rn[i];
    for (int i = 0; i < 3; ++i) d[9 + i] = tn[i];
}

struct GnTermOuts { const double* out[DIF_GN_MAX_TERMS]; int kind[DIF_GN_MAX_TERMS]; int n; };

// mailbox (pinned host memory, 64-bit words): [0] = seq << 8 | status, [1] = first failing term kind + 1, [2..13] = delta, [14] = energy
__global__ void gn_update_kernel(GnState* s, GnCalib c, GnTermOuts t, int is_first, int is_final, unsigned long long seq,
                                 volatile unsigned long long* mailbox) {
    if (threadIdx.x != 0) return;
    double E = 0.0, H[36], g[6];
    for (int i = 0; i < 36; ++i) H[i] = 0.0;
    for (int i = 0; i < 6; ++i) g[i] = 0.0;
    int status = DIF_GN_CONTINUE, bad = 0;
    for (int k = 0; k < t.n; ++k) {
        const double* o = t.out[k];
        if (!(__ldcg(o + 43) > 0.0)) { if (!bad) bad = t.kind[k] + 1; continue; }      // empty valid set (utility.py:84-85 / tracker.py:165)
        E += __ldcg(o + 42);
        if (!is_final) { for (int i = 0; i < 36; ++i) H[i] += __ldcg(o + i); for (int i = 0; i < 6; ++i) g[i] += __ldcg(o + 36 + i); }
    }
    const double last_E = is_first ? INFINITY : s->last_energy;
    if (bad) status = DIF_GN_EMPTY;
    else if (E > last_E) {                                     // tracker.py:263-265
        for (int i = 0; i < 12; ++i) s->delta[i] = s->last_delta[i];
        gn_publish_pose(s, c);
        status = DIF_GN_BREAK;
    } else {
        for (int i = 0; i < 12; ++i) s->last_delta[i] = s->delta[i];
        s->last_energy = E;
        if (!is_final) {                                       // :270-272
            double xi[6];
            if (gn_solve6(H, g, xi)) { gn_apply_twist(xi, s->delta); gn_publish_pose(s, c); }
            else status = DIF_GN_SINGULAR;
        }
    }
    for (int i = 0; i < 12; ++i) mailbox[2 + i] = (unsigned long long)__double_as_longlong(s->delta[i]);
    mailbox[14] = (unsigned long long)__double_as_longlong(E);
    mailbox[1] = (unsigned long long)bad;
    __threadfence_system();
    mailbox[0] = (seq << 8) | (unsigned long long)status;
    __threadfence_system();
}

}  // namespace dif

using namespace dif;

extern "C" {

// The update step of one iteration (solve H xi = -g, delta <- exp(xi) . delta) evaluated on the HOST with the very functions
// gn_update_kernel runs on the device: lets the CPU test-suite pin the fp64 algebra against numpy without a GPU.
int dif_debug_gn_step(const double* H, const double* g, double* delta_inout) {
    if (!H || !g || !delta_inout) return DIF_E_INVALID;
    double xi[6];
    if (!gn_solve6(H, g, xi)) return DIF_GN_SINGULAR;
    gn_apply_twist(xi, delta_inout);
    return DIF_OK;
}

size_t dif_gn_scratch_bytes(int64_t n_obs) {
    return align_up(sizeof(GnState)) + align_up((size_t)DIF_GN_MAX_TERMS * 44 * sizeof(double))
         + align_up(dif_icp_scratch_bytes(n_obs)) + align_up(rgb_scratch_bytes()) + 256;
}

int dif_gauss_newton(const dif_map_view* map, const void* decoder_prepared, const dif_gn_problem* p, void* scratch, size_t scratch_bytes,
                     void* mailbox_host, dif_gn_result* result, void* stream) {
    if (!p || !scratch || !mailbox_host || !result || p->n_groups < 0 || p->n_groups > DIF_GN_MAX_GROUPS) return DIF_E_INVALID;
    if (scratch_bytes < dif_gn_scratch_bytes(p->n_obs)) return DIF_E_WORKSPACE;
    bool any_sdf = false;
    for (int gi = 0; gi < p->n_groups; ++gi) {
        const dif_gn_group& G = p->group[gi];
        if (G.n_terms < 1 || G.n_terms > DIF_GN_MAX_TERMS || G.n_iters < 0) return DIF_E_INVALID;
        for (int k = 0; k < G.n_terms; ++k) {
            if (G.kind[k] == DIF_GN_TERM_SDF) any_sdf = true;
            else if (G.kind[k] == DIF_GN_TERM_RGB) { if (G.level[k] < 0 || G.level[k] >= p->n_levels || p->n_levels > DIF_GN_MAX_LEVELS) return DIF_E_INVALID; }
            else return DIF_E_INVALID;
        }
    }
    if (any_sdf && (!map || !decoder_prepared || !p->obs_xyz || p->n_obs <= 0)) return DIF_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    Carver c(scratch);                                               // zero-filled once by the caller (the term kernels' counters)
    GnState* state = c.take<GnState>(1);
    double* outs = c.take<double>((size_t)DIF_GN_MAX_TERMS * 44);
    void* icp_scratch = c.take<char>(dif_icp_scratch_bytes(p->n_obs));
    void* rgb_scratch = c.take<char>(rgb_scratch_bytes());
    volatile unsigned long long* mb = (volatile unsigned long long*)mailbox_host;

    GnCalib cal; GnInit in;
    for (int i = 0; i < 9; ++i) { cal.K[i] = p->K[i]; cal.Kinv[i] = p->Kinv[i]; }
    for (int i = 0; i < 12; ++i) { in.last[i] = p->last_pose[i]; in.delta[i] = p->init_delta[i]; }
    in.n_points = (int)p->n_obs;
    mb[0] = 0ull;
    gn_init_kernel<<<1, 32, 0, st>>>(state, in, cal);
    DIF_COUNT_LAUNCH(1);
    int rc = check_launch("gn_init_kernel");
    if (rc) return rc;

    result->n_sdf = result->n_rgb = 0; result->last_iter = 0; result->status = DIF_GN_CONTINUE; result->n_iterations = 0;
    for (int i = 0; i < 12; ++i) result->delta[i] = p->init_delta[i];
    result->energy = 0.0;
    unsigned long long seq = 0;
    const char* trace_env = getenv("DIF_GN_TRACE");
    const bool trace = trace_env && trace_env[0] == '1';
    for (int gi = 0; gi < p->n_groups; ++gi) {
        const dif_gn_group& G = p->group[gi];
        for (int it = 0; it <= G.n_iters; ++it) {                   // iterations 0..n-1 with gradients, then the energy-only pass (-1)
            const int is_final = it == G.n_iters;
            GnTermOuts to; to.n = G.n_terms;
            for (int k = 0; k < DIF_GN_MAX_TERMS; ++k) { to.out[k] = nullptr; to.kind[k] = 0; }
            for (int k = 0; k < G.n_terms; ++k) {
                double* out = outs + 44 * k;
                to.out[k] = out; to.kind[k] = G.kind[k];
                if (G.kind[k] == DIF_GN_TERM_SDF) {
                    rc = icp_launch(map, decoder_prepared, p->obs_xyz, 3, p->n_obs, nullptr, &state->frame, p->huber_k, !is_final,
                                    icp_scratch, dif_icp_scratch_bytes(p->n_obs), out, st);
                    ++result->n_sdf;
                } else {
                    const dif_gn_level& Lv = p->level[G.level[k]];
                    rc = rgb_launch(Lv.prev_i, Lv.prev_d, Lv.cur_i, Lv.cur_d, Lv.cur_grad, Lv.h, Lv.w, p->intr, nullptr, nullptr, state->kkt,
                                    p->min_grad_scale, p->max_depth_delta, p->rgb_robust, p->rgb_robust_k, p->rgb_weight, !is_final,
                                    rgb_scratch, rgb_scratch_bytes(), out, st);
                    ++result->n_rgb;
                }
                if (rc) return rc;
            }
            ++seq;
            gn_update_kernel<<<1, 32, 0, st>>>(state, cal, to, it == 0, is_final, seq, mb);
            DIF_COUNT_LAUNCH(1);
            rc = check_launch("gn_update_kernel");
            if (rc) return rc;
            // wait for this iteration's verdict: spin on the mailbox word the update kernel posts (no stream synchronisation)
            unsigned long long w = 0;
            timespec t_start; clock_gettime(CLOCK_MONOTONIC, &t_start);
            for (unsigned long long spins = 0;; ++spins) {
                w = mb[0];
                if ((w >> 8) == seq) break;
                if ((spins & 0xfffff) == 0xfffff) {                 // every ~1 M polls: has the stream died, drained without posting, or hung?
                    const cudaError_t q = cudaStreamQuery(st);
                    if (q != cudaErrorNotReady) {
                        w = mb[0];
                        if ((w >> 8) == seq) break;
                        snprintf(g_last_error, sizeof(g_last_error), "dif_gauss_newton: the stream %s before iteration %llu posted its verdict",
                                 q == cudaSuccess ? "drained" : cudaGetErrorString(q), seq);
                        if (q != cudaSuccess) { (void)cudaGetLastError(); }
                        return DIF_E_LAUNCH;
                    }
                    timespec now; clock_gettime(CLOCK_MONOTONIC, &now);
                    if ((now.tv_sec - t_start.tv_sec) > 30) {        // one iteration is tens of microseconds
                        snprintf(g_last_error, sizeof(g_last_error), "dif_gauss_newton: no verdict for iteration %llu after 30 s", seq);
                        return DIF_E_LAUNCH;
                    }
                }
            }
            __atomic_thread_fence(__ATOMIC_ACQUIRE);
            const int status = (int)(w & 0xff);
            if (trace) {
                double e, d9, d10, d11; unsigned long long u = mb[14]; memcpy(&e, &u, 8);
                u = mb[11]; memcpy(&d9, &u, 8); u = mb[12]; memcpy(&d10, &u, 8); u = mb[13]; memcpy(&d11, &u, 8);
                fprintf(stderr, "[dif_gauss_newton] group %d iter %d%s: energy %.12g status %d  t_delta %.9g %.9g %.9g\n", gi, it, is_final ? " (final)" : "",
                        e, status, d9, d10, d11);
            }
            ++result->n_iterations;
            result->last_iter = is_final ? -1 : it;
            result->status = status;
            if (status != DIF_GN_CONTINUE) break;
        }
        if (result->status == DIF_GN_EMPTY || result->status == DIF_GN_SINGULAR) break;
    }
    if (seq > 0) {
        for (int i = 0; i < 12; ++i) { const unsigned long long u = mb[2 + i]; double d; memcpy(&d, &u, 8); result->delta[i] = d; }
        const unsigned long long u = mb[14]; memcpy(&result->energy, &u, 8);
        result->empty_term = (int)mb[1];
    } else result->empty_term = 0;
    return DIF_OK;
}

}  // extern "C"
