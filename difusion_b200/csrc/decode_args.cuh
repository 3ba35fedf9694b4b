// Sample-source description shared by the SIMT and the tensor-core decoder kernels.
#pragma once
#include "common.cuh"

namespace dif {

// Where a tile's samples come from.  mode 0: explicit arrays (forward_model replacement).  mode 1: dense lattice, sample
// s -> PLIVox s / n3, lattice point s % n3 (map.py:644-653: the reference materialises B*l^3 x 32 inputs).  mode 2: the same
// lattice addressed through a compacted list of global lattice indices whose length lives on the device (map.py:667-679).
struct DecodeArgs {
    const float* P; const float* latent; int lat_stride; const int32_t* rows; const float* xyz; int64_t n;
    const int32_t* out_index; float sdf_sign; float* sdf; float* std; float* grad; int grad_head;
    int mode; int lat_n; float lat_step, lat_a; const uint32_t* list; const int32_t* n_dev;
    const int32_t* row_map;      // lattice modes on a sharded map: block slot -> local latent row (NULL = identity)
};

__device__ __forceinline__ float lattice_coord(const DecodeArgs& a, int i) {
    // get_samples(): idx * vsize + a, then - 0.5 into network coordinates (utility.py:143-147, map.py:645-646)
    return __fsub_rn(__fadd_rn(__fmul_rn((float)i, a.lat_step), a.lat_a), 0.5f);
}

// sample s of a launch -> (latent row or -1, output index, lattice point index)
__device__ __forceinline__ void decode_sample_source(const DecodeArgs& a, int64_t sidx, int64_t n_total, int n3,
                                                      int64_t& row, int64_t& out, int& li) {
    row = -1; out = sidx; li = 0;
    if (sidx >= n_total) return;
    if (a.mode == 0) {
        row = a.rows ? (int64_t)a.rows[sidx] : sidx;
        out = a.out_index ? (int64_t)a.out_index[sidx] : sidx;
    } else {
        const int64_t g = a.mode == 1 ? sidx : (int64_t)a.list[sidx];
        row = a.rows[g / n3]; li = (int)(g % n3); out = g;
        if (a.row_map && row >= 0) row = a.row_map[row];
    }
}

int launch_decode_tc(const void* prepared, DecodeArgs a, int64_t n_max, cudaStream_t st);
int set_tc_timing_buffer(unsigned long long* dev_buf);

}  // namespace dif
