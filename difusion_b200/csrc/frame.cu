// dif_frame: one frame of the SLAM loop in one call - tracker linearisation of the frame against the map built so far
// (reference system/tracker.py:174-218 via main.py:79) followed by integrate_keyframe (system/map.py:340-452 via main.py:88).
// Every per-frame value (point count, poses) is read from a device block, so the launch sequence is identical for every
// frame and the whole call can be captured once into a CUDA graph and replayed.
#include "common.cuh"

namespace dif {
int icp_launch(const dif_map_view* map, const void* decoder_prepared, const float* obs_xyz, int obs_stride, int64_t n, const float* pose_host,
               const dif_frame_params* frame_dev, float huber_k, int want_grad, void* scratch, size_t scratch_sz, double* out_dev, cudaStream_t st);
int integrate_launch(const dif_map_view* map, const void* encoder_prepared, const float* xyz, const float* normal, int stride, int64_t n,
                     const dif_frame_params* frame, uint8_t* unq_mask, void* persist, size_t persist_sz, void* scratch, size_t scratch_sz,
                     int32_t* stats_dev, cudaStream_t st);
}  // namespace dif

using namespace dif;

extern "C" int dif_frame(const dif_map_view* map, const void* encoder_prepared, const void* decoder_prepared, const float* points,
                         int64_t max_points, const dif_frame_params* frame_dev, float huber_k, int flags, uint8_t* unq_mask,
                         void* persist, size_t persist_sz, void* scratch, size_t scratch_sz, void* icp_scratch, size_t icp_scratch_sz,
                         void* result_dev, void* stream) {
    if (!map || !points || !frame_dev || !result_dev || max_points <= 0 || !(flags & (DIF_FRAME_TRACK | DIF_FRAME_INTEGRATE))) return DIF_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    double* icp_out = (double*)result_dev;
    int32_t* stats = (int32_t*)((char*)result_dev + 44 * sizeof(double));
    const int S = DIF_FRAME_POINT_FLOATS;
    if (flags & DIF_FRAME_TRACK) {
        const int rc = icp_launch(map, decoder_prepared, points, S, max_points, nullptr, frame_dev, huber_k, 1, icp_scratch, icp_scratch_sz, icp_out, st);
        if (rc) return rc;
    }
    if (flags & DIF_FRAME_INTEGRATE) {
        const int rc = integrate_launch(map, encoder_prepared, points + 3, points + 6, S, max_points, frame_dev, unq_mask, persist, persist_sz,
                                        scratch, scratch_sz, stats, st);
        if (rc) return rc;
    }
    return DIF_OK;
}
