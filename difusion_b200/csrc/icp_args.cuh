// Arguments shared by the SIMT and the tensor-core ICP linearisation kernels (reference system/tracker.py:174-218).
#pragma once
#include "common.cuh"

namespace dif {

struct MapRO { const int64_t* indexer; const float* latent; const float* obs; Grid g; float ignore_th; int lat_stride;
               const int32_t* row_of; };     // row_of: sharded map, slot -> local latent row (-1: not stored on this rank); NULL = identity
struct Pose { float Rc[9], tc[3], Rd[9], td[3], Rl[9]; };      // composite (last*delta), delta, last rotation

// composite pose exactly as the host path forms it: fp64 products of the fp32 inputs, rounded once (tracker.py:181)
__host__ __device__ inline void compose_pose(const float* pose24, Pose& p) {
    const float *Rl = pose24, *tl = pose24 + 9, *Rd = pose24 + 12, *td = pose24 + 21;
    for (int i = 0; i < 3; ++i) {
        double t = tl[i];
        for (int j = 0; j < 3; ++j) {
            double a = 0;
            for (int k = 0; k < 3; ++k) a += (double)Rl[3 * i + k] * Rd[3 * k + j];
            p.Rc[3 * i + j] = (float)a;
            t += (double)Rl[3 * i + j] * td[j];
        }
        p.tc[i] = (float)t;
    }
    for (int i = 0; i < 9; ++i) { p.Rd[i] = Rd[i]; p.Rl[i] = Rl[i]; }
    for (int i = 0; i < 3; ++i) p.td[i] = td[i];
}

// Point count and pose of a launch: host values, or (frame != NULL) read from the device block after the dependency wait.
struct IcpFrame { Pose pose; int n; };
__device__ inline void icp_resolve_frame(const dif_frame_params* frame, const Pose& host_pose, int host_n, IcpFrame& out) {
    if (frame) {
        compose_pose(frame->pose, out.pose);
        const int f = frame->n_points;
        out.n = f < 0 ? 0 : (f < host_n ? f : host_n);
    } else { out.pose = host_pose; out.n = host_n; }
}

struct IcpTcArgs {
    MapRO m; const float* obs; int obs_stride; int n; const dif_frame_params* frame; Pose pose; float huber_k; int want_grad;
    double* partials;             // [grid][32] per-CTA fp64 sums: 21 upper-triangular H, 6 g, energy, M (every row fully written)
    unsigned int* done_counter;   // zero on entry, left zero on exit
    double* out;                  // [44]
    unsigned long long* ll;       // [grid][32][2] epoch-tagged words of the per-CTA rows (icp_tc2_kernel)
    unsigned int* epoch;          // epoch of the last completed launch on this scratch
};

int launch_icp_tc(const void* decoder_prepared, const IcpTcArgs& a, cudaStream_t st);

}  // namespace dif
