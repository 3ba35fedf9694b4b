// dif_frame: one frame of the SLAM loop in one call - tracker linearisation of the frame against the map built so far
// (reference system/tracker.py:174-218 via main.py:79) followed by integrate_keyframe (system/map.py:340-452 via main.py:88).
// Every per-frame value (point count, poses) is read from a device block, so the launch sequence is identical for every
// frame and the whole call can be captured once into a CUDA graph and replayed.
#include "common.cuh"
#include <stdlib.h>

namespace dif {
int icp_launch(const dif_map_view* map, const void* decoder_prepared, const float* obs_xyz, int obs_stride, int64_t n, const float* pose_host,
               const dif_frame_params* frame_dev, float huber_k, int want_grad, void* scratch, size_t scratch_sz, double* out_dev, cudaStream_t st);
int integrate_launch(const dif_map_view* map, const void* encoder_prepared, const float* xyz, const float* normal, int stride, int64_t n,
                     const dif_frame_params* frame, uint8_t* unq_mask, void* persist, size_t persist_sz, void* scratch, size_t scratch_sz,
                     int32_t* stats_dev, cudaStream_t st, cudaEvent_t readers_done);

// Side stream + fork/join events of dif_frame, one set per device (created on first use, never destroyed).
struct FrameSide { cudaStream_t stream; cudaEvent_t fork, join; bool ok; };
static FrameSide* frame_side() {
    static FrameSide side[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    FrameSide& s = side[dev];
    if (!s.ok) {
        if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        s.ok = true;
    }
    return &s;
}
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
    // The tracker term only READS the map; of the integrate chain only its last kernel (fuse_kernel) writes what the term reads
    // (latents, observation counts) - cells allocated by this frame have observation count 0 and are ignored by the term whether it
    // sees them or not.  So the linearisation (one persistent CTA per SM, latency-bound) runs on a side stream BESIDE the index
    // kernels (voxelise / prune / allocate / gather: small blocks that share the SMs with it), and fuse_kernel joins.  The fork and
    // the join are events, so the same code is captured into a CUDA graph as two branches.  DIF_FRAME_OVERLAP=0: one stream.
    const char* ov = getenv("DIF_FRAME_OVERLAP");
    FrameSide* side = (flags & DIF_FRAME_TRACK) && (flags & DIF_FRAME_INTEGRATE) && !(ov && ov[0] == '0') ? frame_side() : nullptr;
    cudaEvent_t join = nullptr;
    if (flags & DIF_FRAME_TRACK) {
        cudaStream_t ist = st;
        if (side) { cudaEventRecord(side->fork, st); cudaStreamWaitEvent(side->stream, side->fork, 0); ist = side->stream; }
        const int rc = icp_launch(map, decoder_prepared, points, S, max_points, nullptr, frame_dev, huber_k, 1, icp_scratch, icp_scratch_sz, icp_out, ist);
        if (side) { cudaEventRecord(side->join, side->stream); join = side->join; }
        if (rc) { if (join) cudaStreamWaitEvent(st, join, 0); return rc; }
    }
    if (flags & DIF_FRAME_INTEGRATE) {
        const int rc = integrate_launch(map, encoder_prepared, points + 3, points + 6, S, max_points, frame_dev, unq_mask, persist, persist_sz,
                                        scratch, scratch_sz, stats, st, join);
        if (rc) return rc;
    }
    return DIF_OK;
}
