// Arguments shared by the SIMT and the tensor-core ICP linearisation kernels (reference system/tracker.py:174-218).
#pragma once
#include "common.cuh"

namespace dif {

struct MapRO { const int64_t* indexer; const float* latent; const float* obs; Grid g; float ignore_th; };
struct Pose { float Rc[9], tc[3], Rd[9], td[3], Rl[9]; };      // composite (last*delta), delta, last rotation

struct IcpTcArgs {
    MapRO m; const float* obs; int n; Pose pose; float huber_k; int want_grad;
    double* partials;             // [grid][32] per-CTA fp64 sums: 21 upper-triangular H, 6 g, energy, M (every row fully written)
    unsigned int* done_counter;   // zero on entry, left zero on exit
    double* out;                  // [44]
};

int launch_icp_tc(const void* decoder_prepared, const IcpTcArgs& a, cudaStream_t st);

}  // namespace dif
