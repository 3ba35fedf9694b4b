/* difusion_b200 - C ABI of the B200-native DI-Fusion hot path (libdifusion_b200.so, sm_100a).
 *
 * This is the drop-in boundary for the reference's per-frame fusion path.  The reference
 * (huangjh-pub/di-fusion, paths below relative to its pytorch/ directory) implements that path as
 * chains of torch ops inside system/map.py + system/tracker.py plus two pybind11 torch extensions
 * (system/ext/__init__.py:15-44).  Every entry point here cites the reference interface it replaces.
 *
 * Conventions
 *   - plain C: raw DEVICE pointers + sizes + a CUDA stream handle passed as void* (cudaStream_t);
 *     no torch types, no allocation, no host synchronisation inside any call (the reference ext ops
 *     allocate their outputs and sync: mc_interp_kernel.cu:344-369, indexing.cu:76,96-97);
 *   - every function returns 0 on success or a negative DIF_E_* code and is asynchronous on `stream`;
 *   - scalars that the reference reads back with .item() are written to small device arrays that the
 *     caller may copy when (and if) it needs them;
 *   - re-entrant: no global mutable state; two host threads may drive two streams on disjoint buffers.
 */
#ifndef DIFUSION_B200_H
#define DIFUSION_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIF_ABI_VERSION 3
#define DIF_LATENT_DIM 29                 /* ckpt/default/hyper.json:34 "code_length" */

enum {
    DIF_OK = 0,
    DIF_E_INVALID = -1,                   /* bad argument (null pointer, size, unsupported resolution)   */
    DIF_E_WORKSPACE = -2,                 /* workspace smaller than dif_*_workspace_bytes() says        */
    DIF_E_LAUNCH = -3                     /* CUDA launch error; text via dif_last_error()                */
};

/* ---- folded network weights -------------------------------------------------------------------------
 * fp32 blobs, prepared once on the host (difusion_b200/weights.py) from the reference checkpoint:
 *   decoder: weight-norm folded W_k = g_k v_k/||v_k||  (network/di_decoder.py:36-40)
 *   encoder: eval BatchNorm folded into the 1x1 convs  (utils/pt_util.py:37-116,     SURVEY A.6)
 * Layout (row-major [out][in], then bias):
 *   decoder  W0[128][32] b0[128] W1[128][128] b1[128] W2[96][128] b2[96] W3[128][128] b3[128]
 *            w4[128] b4[1] wu[128] bu[1]                         -> DIF_DECODER_BLOB_FLOATS
 *            (input column order: latent 0..28, x, y, z  - network/utility.py:80;
 *             layer-3 input order: h2 (96) first, then the 32 inputs - di_decoder.py:61-62)
 *   encoder  W0[32][6] b0[32] W1[64][32] b1[64] W2[256][64] b2[256] W3[29][256] b3[29]
 *                                                                -> DIF_ENCODER_BLOB_FLOATS
 * dif_prepare_* expands a blob into the device-side form the kernels read (transposed fp32 copies,
 * fp16 hi/lo split tiles in UMMA shared-memory layout). */
#define DIF_DECODER_BLOB_FLOATS (128*32+128 + 128*128+128 + 96*128+96 + 128*128+128 + 128+1 + 128+1)
#define DIF_ENCODER_BLOB_FLOATS (32*6+32 + 64*32+64 + 256*64+256 + 29*256+29)

size_t dif_decoder_prepared_bytes(void);
size_t dif_encoder_prepared_bytes(void);
int dif_prepare_decoder(const float* blob_dev, void* prepared_dev, void* stream);
int dif_prepare_encoder(const float* blob_dev, void* prepared_dev, void* stream);

/* ---- map state --------------------------------------------------------------------------------------
 * Device view of the reference's DenseIndexedMap.cold_vars (system/map.py:199-211).  The arrays are owned
 * by the caller (torch tensors in the Python mirror) and stay readable as the reference's public
 * attributes `indexer`, `latent_vecs`, `latent_vecs_pos`, `voxel_obs_count`. */
typedef struct dif_map_view {
    int64_t* indexer;            /* [nx*ny*nz]  -1 = empty, else slot; linear id = z + nz*y + nz*ny*x (map.py:287-292) */
    float*   latent_vecs;        /* [capacity][29] */
    int64_t* latent_vecs_pos;    /* [capacity]  slot -> linear id, -1 = unused */
    float*   voxel_obs_count;    /* [capacity]  integer-valued fp32 */
    uint8_t* slot_dirty;         /* [capacity]  1 = latent changed since the last mesh extraction (map.py:303-308) */
    int32_t* n_occupied;         /* device scalar (the reference keeps a host int, map.py:200) */
    int64_t  capacity;
    int32_t  nx, ny, nz;
    float    bound_min[3];
    float    voxel_size;
    int32_t  prune_min_vox_obs;  /* fusion-lr-kt.yaml:32 */
    float    ignore_count_th;    /* :33 */
    float    encoder_count_th;   /* :34 */
    /* -- hash-sharded map (new: the reference is single-GPU).  shard_world <= 1: unsharded.  Integer state is
     * replicated (every rank runs the same index kernels on the same frame); the encoder MLP + latent fusion of a PLIVox run
     * only on its owner.  After dif_integrate, xchg_slots[0 .. stats[DIF_STAT_N_XCHG]) lists the slots whose latent rows this
     * rank owns and changed: dif_shard_pack sends the boundary ones to the ranks that keep them in their halo. */
    int32_t  shard_rank, shard_world;
    int32_t* xchg_slots;         /* [>= min(8*n_points, capacity)] or NULL */
    /* floats between consecutive rows of latent_vecs: 29 = the reference's packed rows, 32 = rows padded to 128 bytes (16-byte
     * aligned: the gather kernels then fetch a row with 8 vector loads instead of 29 scalar ones; the Python mirror stores the
     * table this way and exposes the reference's (capacity, 29) tensor as a strided view).  0 is read as 29. */
    int32_t  latent_stride;
    /* -- sharded storage (see "hash-sharded map" below).  row_of_slot == NULL: latent row of slot s is row s (unsharded map). */
    int32_t  shard_block_log2;   /* super-block edge = 2^k cells (ownership and halo granularity); 0 = per-cell ownership */
    int32_t* row_of_slot;        /* [capacity] slot -> row of latent_vecs on THIS rank, -1 = not stored here (caller initialises to -1) */
    int32_t* n_rows;             /* device scalar: rows of latent_vecs in use on this rank */
    int64_t  row_capacity;       /* rows latent_vecs can hold on this rank */
    /* (p - bound_min) / voxel_size (map.py:366-367, :565): torch evaluates `tensor / python_float` as an IEEE division on CPU
     * tensors and as a multiplication by the rounded fp32 reciprocal on CUDA tensors (its cpu-scalar fast path); a point within an
     * ulp of a PLIVox face lands in different cells under the two.  0 = true division (the reference on CPU tensors: BASELINE
     * configs[0], the committed fixtures); 1 = multiply by 1.f / voxel_size (bit-identical to the reference run on CUDA tensors). */
    int32_t  scalar_division_mode;
    int32_t  reserved_;
} dif_map_view;

/* ---- integrate_keyframe  (system/map.py:340-452; SURVEY rows a-2 .. a-6) -----------------------------
 * One call = voxelise + prune + allocate (ascending linear id) + encoder-target focus + 8-offset gather +
 * encoder MLP + per-PLIVox sum + running-mean fusion.  `persist` is scratch that must be zero-filled ONCE by
 * the caller (and whenever it is re-allocated); the call leaves it zero-filled again.  `scratch` needs no init.
 * unq_mask (nullable) receives the per-point prune mask the reference returns (map.py:375,519).
 * stats_dev[DIF_STAT_COUNT] is written on the device. */
enum {
    DIF_STAT_N_KEPT = 0,         /* points surviving the prune                                    */
    DIF_STAT_N_NEW = 1,          /* slots allocated by this call                                  */
    DIF_STAT_N_SAMPLES = 2,      /* encoder samples gathered (map.py:434)                         */
    DIF_STAT_N_UPDATED = 3,      /* PLIVoxes fused (len(surface_blatent_mapping), map.py:437)     */
    DIF_STAT_N_OCCUPIED = 4,     /* n_occupied after the call                                     */
    DIF_STAT_FLAGS = 5,          /* bit0: a point fell outside the grid (dropped); bit1: capacity exhausted; bit2: row_capacity exhausted (sharded) */
    DIF_STAT_N_FOCUSED = 6,      /* points passing the focus mask (map.py:389-397)                */
    DIF_STAT_N_XCHG = 7,         /* sharded map: owned PLIVoxes fused by this call (length of xchg_slots) */
    DIF_STAT_N_ROWS = 9,         /* sharded map: latent rows in use on this rank after the call (== n_occupied when unsharded) */
    DIF_STAT_SEQ = 8,            /* frame->seq echoed back (0 without a frame block): tells a lagging reader which frame the counters belong to */
    DIF_STAT_COUNT = 12
};

/* ---- per-frame parameter block in DEVICE memory (new: SURVEY 7 step 7, "CUDA-graph the per-frame pipeline") -----------
 * The reference passes the point count and the poses as host values (tensor shapes, Isometry objects), which bakes them into
 * every launch.  With this block the SAME launch sequence serves every frame, so it can be captured once into a CUDA graph and
 * replayed: the caller writes the block (it may sit in front of the frame's points so that ONE H2D copy moves both) and the
 * kernels read the actual point count and the poses from it.  `n` / `max_points` arguments then only bound the launch grids. */
typedef struct dif_frame_params {
    int32_t n_points;            /* valid rows this frame (clamped to the n / max_points argument of the call) */
    int32_t seq;                 /* caller's frame number, echoed to stats[DIF_STAT_SEQ] */
    int32_t reserved[2];
    float   pose[24];            /* R_last[9], t_last[3], R_delta[9], t_delta[3], row-major (the pose_host layout of dif_icp_linearize) */
} dif_frame_params;              /* 112 bytes; callers that pack it in front of the points pad it to DIF_FRAME_HEADER_FLOATS floats */
#define DIF_FRAME_HEADER_FLOATS 32
#define DIF_FRAME_POINT_FLOATS 9          /* dif_frame point row: camera xyz, world xyz, world normal */
size_t dif_integrate_persist_bytes(int64_t n_cells, int64_t capacity);
size_t dif_integrate_scratch_bytes(int64_t max_points);
int dif_integrate(const dif_map_view* map, const void* encoder_prepared,
                  const float* xyz /*[n][3] world*/, const float* normal /*[n][3] world*/, int64_t n,
                  const dif_frame_params* frame_dev /* NULL: n is exact; else n bounds frame_dev->n_points (device) */,
                  uint8_t* unq_mask /*[n] or NULL*/, void* persist, size_t persist_bytes,
                  void* scratch, size_t scratch_bytes, int32_t* stats_dev, void* stream);

/* ---- network evaluation  (network/utility.py:61-126 forward_model; di_decoder.py:55-86; di_encoder.py:26-30) --
 * dif_decode: sdf/std (and optionally d sdf/d xyz, d std/d xyz) for n samples.  Sample i reads latent row
 *   rows ? rows[i] : i   of `latent` (row stride latent_stride floats); rows[i] < 0 marks a padding sample (outputs 0).
 *   out_index (nullable) scatters result i to position out_index[i]; sdf_sign multiplies sdf (map.py:687). */
int dif_decode(const void* decoder_prepared, const float* latent, int latent_stride /* floats per row: 29, or 32 (16-byte aligned rows) */,
               const int32_t* rows, const float* xyz /*[n][3]*/, int64_t n, const int32_t* out_index, float sdf_sign,
               float* sdf, float* std, float* dsdf_dxyz /*[n][3] or NULL*/, float* dstd_dxyz /*[n][3] or NULL*/, void* stream);
int dif_encode(const void* encoder_prepared, const float* xyzn /*[n][6]*/, int64_t n, float* latent_out /*[n][29]*/, void* stream);

/* ---- get_sdf lookup  (system/map.py:565-575; SURVEY a-8) ---------------------------------------------
 * slot_out[i] = latent slot or -1 (empty cell, outside the grid, or obs_count <= ignore_count_th);
 * rel_out[i]  = (xyz-bound_min)/voxel_size - cell - 0.5   (network coordinates);   n_valid_dev: device scalar. */
int dif_map_query(const dif_map_view* map, const float* xyz, int64_t n, int32_t* slot_out, float* rel_out,
                  int32_t* n_valid_dev, void* stream);

/* ---- compute_sdf_Hg  (system/tracker.py:174-218; SURVEY a-9) ------------------------------------------
 * Fused: transform obs by last*delta, lookup, decoder fwd (+bwd wrt xyz), r = sdf/std, J = [G R_last^T, q x .],
 * robust kernel (tracker.py:58-71: huber_k > 0 Huber(k); huber_k < 0 Tukey(-huber_k); 0 none), normal equations.  pose = {R_last[9], t_last[3], R_delta[9], t_delta[3]}
 * row-major fp32 (host memory, copied at call time).  out_dev[44] (fp64): H[36] row-major, g[6], energy, M (valid count);
 * already divided by M as the reference does.  want_grad = 0 reproduces no_grad=True (only energy and M are written).
 * `scratch` must be zero-filled ONCE by the caller; every call leaves it zero-filled again.
 * frame_dev != NULL: the point count (bounded by n) and the poses are read from the device block instead of n / pose_host
 * (pose_host may then be NULL); the composite pose last*delta is formed on the device with the same fp64 arithmetic. */
size_t dif_icp_scratch_bytes(int64_t n);
int dif_icp_linearize(const dif_map_view* map, const void* decoder_prepared, const float* obs_xyz /*[n][3] camera frame*/,
                      int64_t n, const float* pose_host /*[24]*/, const dif_frame_params* frame_dev, float huber_k, int want_grad,
                      void* scratch, size_t scratch_bytes, double* out_dev /*[44]*/, void* stream);

/* ---- one frame, one call  (main.py:71-94: tracker linearisation against the map so far, then integrate_keyframe) -----------
 * points: [max_points][DIF_FRAME_POINT_FLOATS] rows = camera-frame xyz (tracker.last_processed_pc), world xyz and world normal
 * (what main.py:88 passes to integrate_keyframe); frame_dev->n_points of them are valid.  Equivalent to dif_icp_linearize (if
 * flags & DIF_FRAME_TRACK) followed by dif_integrate (if flags & DIF_FRAME_INTEGRATE) on the same stream, with every per-frame
 * value read from device memory - the whole call is CUDA-graph capturable and replayable for any frame that fits max_points.
 * result_dev: [44] doubles (H, g, energy, M as dif_icp_linearize) followed by DIF_STAT_COUNT int32 counters (as dif_integrate),
 * DIF_FRAME_RESULT_BYTES in all, so that one D2H copy returns everything the host loop reads. */
enum { DIF_FRAME_TRACK = 1, DIF_FRAME_INTEGRATE = 2 };
#define DIF_FRAME_RESULT_BYTES (44 * 8 + DIF_STAT_COUNT * 4)
int dif_frame(const dif_map_view* map, const void* encoder_prepared, const void* decoder_prepared,
              const float* points, int64_t max_points, const dif_frame_params* frame_dev, float huber_k, int flags,
              uint8_t* unq_mask /*[max_points] or NULL*/, void* persist, size_t persist_bytes, void* scratch, size_t scratch_bytes,
              void* icp_scratch, size_t icp_scratch_bytes, void* result_dev, void* stream);

/* ---- mesh extraction  (system/map.py:624-691; SURVEY a-10, a-11) ---------------------------------------
 * dif_mesh_select: which PLIVoxes to decode.  updated_slots == NULL means "all occupied" (no_cache=True, map.py:615).
 *   focused_ids_out[K]   = latent_vecs_pos[updated]                      (map.py:627)
 *   block_slots_out[B]   = slots of (focused U allocated 6-neighbours), ascending linear id, obs_count > ignore_count_th (:628-631)
 *   mapping_out[capacity]= slot -> batch index or -1                     (:633-635; the reference sizes it max_slot+1)
 *   counts_dev[2]        = {K, B}
 * dif_mesh_decode: cube_sdf/std [B][2r][2r][2r]; fast != 0: r^3 low pass, trilinear x2 (align_corners), re-evaluate
 *   |sdf| < 0.05 (map.py:655-682); sdf is stored negated (:687).  counts_dev[2] = {n_low, n_high}.
 * dif_marching_cubes: system.ext.marching_cubes_interp (mc.cpp:3-16, mc_interp_kernel.cu:202-382).  Triangles in voxel
 *   units; *count_dev = number produced (may exceed max_tri: the excess is dropped, as in the reference :308,375-379). */
size_t dif_mesh_select_scratch_bytes(int64_t n_cells, int64_t capacity);
int dif_mesh_select(const dif_map_view* map, const int32_t* updated_slots, int64_t n_updated,
                    int64_t* focused_ids_out, int32_t* block_slots_out, int32_t* mapping_out, int32_t* counts_dev,
                    void* persist, size_t persist_bytes, void* stream);
size_t dif_mesh_decode_scratch_bytes(int64_t n_blocks, int r);
int dif_mesh_decode(const dif_map_view* map, const void* decoder_prepared, const int32_t* block_slots, int64_t n_blocks,
                    int r, int fast, float* cube_sdf, float* cube_std, void* scratch, size_t scratch_bytes,
                    int32_t* counts_dev, void* stream);
int dif_marching_cubes(const int64_t* indexer, int nx, int ny, int nz, const int64_t* valid_blocks, int64_t n_valid,
                       const int32_t* vec_batch_mapping, int64_t mapping_len, const float* cube_sdf, const float* cube_std,
                       int r, float max_std, float* tri /*[max_tri][3][3]*/, int64_t* tri_flatten_id /*[max_tri]*/,
                       float* tri_std /*[max_tri][3]*/, int64_t max_tri, int32_t* count_dev, void* stream);

/* ---- device-side mesh cache merge  (replaces the host code of system/map.py:698-714 + _get_valid_idx :20-26) --
 * out = [cached triangles whose PLIVox id does not occur in new_id, in their old order] ++ [new triangles, voxel units -> world:
 * v * voxel_size + bound_min (map.py:698)].  Rows: tri [.][3][3] f32, id [.] i64 (linear PLIVox id), std [.][3] f32.
 * out_* must hold n_cache + n_new rows and may not alias the inputs.  totals_dev[2] (int64) = {kept, kept + n_new}.
 * persist: dif_mesh_cache_scratch_bytes(n_cells, n_cache) bytes, zero-filled ONCE by the caller; left zeroed (self-cleaning), so
 * it can be reused while n_cache does not outgrow it. */
size_t dif_mesh_cache_scratch_bytes(int64_t n_cells, int64_t n_cache);
int dif_mesh_cache_merge(const float* cache_tri, const int64_t* cache_id, const float* cache_std, int64_t n_cache,
                         const float* new_tri, const int64_t* new_id, const float* new_std, int64_t n_new,
                         float voxel_size, const float* bound_min /*[3], host*/, int64_t n_cells,
                         float* out_tri, int64_t* out_id, float* out_std, int64_t* totals_dev,
                         void* persist, size_t persist_bytes, void* stream);

/* ---- frame pre-processing  (system/tracker.py:88-117, :13-23; system/ext/imgproc/imgproc.cu:5-44) -------------
 * dif_unproject_depth : ext op unproject_depth: pc[v][u] = ((u - cx) / fx * d, (v - cy) / fy * d, d); NaN depth -> NaN point.
 * dif_point_box_filter: tracker.point_box_filter (:13-23): mean point / mean normal per voxel_size cell of the frame's bounding box,
 *   rows in ascending cell-key order (== torch.unique's order).  out_* hold up to n rows; *n_out_dev = rows written, or -1 when
 *   the bounding box has more than max_cells cells.  scratch: dif_box_filter_scratch_bytes(max n, max_cells) bytes, zero-filled
 *   ONCE by the caller and left zeroed by every call.
 * dif_remove_radius_outlier: ext op remove_radius_outlier (pcproc.cu:98-105,172-196): mask[i] = the nb_points-th nearest point
 *   (i itself included) is closer than radius.  pc: rows of `stride` floats (the reference passes (N,4) rows), xyz first.
 * dif_estimate_normals: ext op estimate_normals (pcproc.cu:107-170,198-220): PCA normal of the <= max_nn - 1 nearest neighbours
 *   within radius (max_nn <= 32), NaN if fewer than 5, flipped towards cam_xyz (host float[3]).
 *   Both replace the reference's per-call kd-tree (cuda_kdtree.cu) by a uniform grid of cell edge = radius over the frame's
 *   bounding box; *status_dev = 1 when that box has more than max_cells cells (outputs are then undefined).
 *   scratch: dif_knn_scratch_bytes(max n, max_cells), zero-filled ONCE by the caller, left zeroed by every call. */
size_t dif_knn_scratch_bytes(int64_t max_points, int64_t max_cells);
int dif_remove_radius_outlier(const float* pc, int stride, int64_t n, int nb_points, float radius, int64_t max_cells,
                              uint8_t* mask_out, int32_t* status_dev, void* scratch, size_t scratch_bytes, void* stream);
int dif_estimate_normals(const float* pc, int stride, int64_t n, int max_nn, float radius, const float* cam_xyz, int64_t max_cells,
                         float* normal_out /*[n][3]*/, int32_t* status_dev, void* scratch, size_t scratch_bytes, void* stream);
int dif_unproject_depth(const float* depth, int h, int w, float fx, float fy, float cx, float cy, float* pc_out /*[h][w][3]*/, void* stream);
size_t dif_box_filter_scratch_bytes(int64_t max_points, int64_t max_cells);
int dif_point_box_filter(const float* points /*[n][3]*/, const float* normals /*[n][3]*/, int64_t n, float voxel_size, int64_t max_cells,
                         float* out_points, float* out_normals, int32_t* n_out_dev, void* scratch, size_t scratch_bytes, void* stream);

/* ---- photometric term  (system/ext/imgproc/photometric.cu:3-138, system/tracker.py:131-172) -------------------
 * Images are row-major [h][w] f32 (NaN = invalid depth); gradients [h][w][2]; intr = {fx, fy, cx, cy}; krkinv = K R K^-1 (row
 * major, 9), kt = K t (3) - host arrays, exactly the lists the reference passes to rgb_odometry (tracker.py:139-146).
 * dif_gradient_xy : ext op gradient_xy (photometric.cu:3-22,80-93): Sobel/8, NaN on the one-pixel border.
 * dif_rgb_odometry: ext op rgb_odometry (photometric.cu:24-78,95-138): f_out [h][w] (NaN = rejected pixel); J_out [h][w][6] or NULL
 *                   (compute_J = false); J rows of rejected pixels are left untouched, as in the reference (torch::empty).
 * dif_rgb_linearize: compute_rgb_Hg (tracker.py:131-172) in one launch: out_dev[44] doubles = H (36, row major), g (6), energy,
 *                   M (valid pixels); J is negated as in :157, error_scale = weight / M (:165).  robust_kind 0 none / 1 huber /
 *                   2 tukey (:58-71).  want_grad = 0 fills only energy and M.  scratch: dif_rgb_scratch_bytes(), zero-filled once. */
int dif_gradient_xy(const float* intensity, int h, int w, float* out_grad, void* stream);
int dif_rgb_odometry(const float* prev_intensity, const float* prev_depth, const float* cur_intensity, const float* cur_depth,
                     const float* cur_dIdxy, int h, int w, const float* intr, const float* krkinv, const float* kt,
                     float min_grad_scale, float max_depth_delta, float* f_out, float* J_out, void* stream);
size_t dif_rgb_scratch_bytes(void);
int dif_rgb_linearize(const float* prev_intensity, const float* prev_depth, const float* cur_intensity, const float* cur_depth,
                      const float* cur_dIdxy, int h, int w, const float* intr, const float* krkinv, const float* kt,
                      float min_grad_scale, float max_depth_delta, int robust_kind, float robust_k, float weight, int want_grad,
                      void* scratch, size_t scratch_bytes, double* out_dev, void* stream);

/* ---- Gauss-Newton pose refinement  (system/tracker.py:220-283 gauss_newton; :174-218, :131-172 for the two terms) -----------
 * The whole loop in one call: for every group of `iter_config`, iterations 0..n-1 (terms with gradients) and the closing
 * energy-only pass; after each iteration the energy test (:263-268), the 6x6 solve and delta <- exp(xi) . delta
 * (utils/motion_util.py:205-229,277-278) run on the device in fp64, and the term kernels read the pose from device memory.
 * The call blocks until the last iteration's verdict has arrived in `mailbox_host` (>= 128 bytes of pinned host memory that the
 * device can write: cudaHostAlloc / torch pin_memory under UVA); it never synchronises the stream.
 *   terms: DIF_GN_TERM_SDF = compute_sdf_Hg on obs_xyz (camera frame) against `map`; DIF_GN_TERM_RGB = compute_rgb_Hg at pyramid
 *   `level` (the reference passes the level-0 intrinsics at every level, tracker.py:135-146; K and K^-1 are taken as given).
 *   result: delta = refined delta pose (R row major [9], t [3]); last_iter = the reference's loop variable `i_iter` on exit
 *   (:276: -1 after a completed group, else the iteration that raised the energy); status of the last executed iteration:
 *   DIF_GN_EMPTY = a term had no valid sample (the reference asserts / divides by zero there; empty_term = 1 sdf, 2 rgb),
 *   DIF_GN_SINGULAR = H was singular (numpy.linalg.solve raises).  scratch: dif_gn_scratch_bytes(n_obs), zero-filled once. */
enum { DIF_GN_TERM_SDF = 0, DIF_GN_TERM_RGB = 1 };
enum { DIF_GN_CONTINUE = 1, DIF_GN_BREAK = 2, DIF_GN_EMPTY = 3, DIF_GN_SINGULAR = 4 };
#define DIF_GN_MAX_TERMS 4
#define DIF_GN_MAX_GROUPS 8
#define DIF_GN_MAX_LEVELS 4
typedef struct dif_gn_level { const float *prev_i, *prev_d, *cur_i, *cur_d, *cur_grad; int32_t h, w; } dif_gn_level;
typedef struct dif_gn_group { int32_t n_iters, n_terms; int32_t kind[DIF_GN_MAX_TERMS]; int32_t level[DIF_GN_MAX_TERMS]; } dif_gn_group;
typedef struct dif_gn_problem {
    const float* obs_xyz;        /* [n_obs][3] camera frame (tracker.last_processed_pc[0]); sdf terms only */
    int64_t n_obs;
    float   huber_k;             /* sdf robust kernel (fusion-lr-kt.yaml:47): > 0 Huber(k), < 0 Tukey(-k), 0 none */
    int32_t n_levels;
    dif_gn_level level[DIF_GN_MAX_LEVELS];
    float   intr[4];             /* fx, fy, cx, cy */
    double  K[9], Kinv[9];       /* calib.to_K() and its inverse, row major */
    float   min_grad_scale, max_depth_delta;
    int32_t rgb_robust;          /* 0 none, 1 huber, 2 tukey */
    float   rgb_robust_k, rgb_weight;
    int32_t n_groups;
    dif_gn_group group[DIF_GN_MAX_GROUPS];
    double  last_pose[12];       /* R [9] row major, t [3]: all_pd_pose[-1] */
    double  init_delta[12];      /* last_pose^-1 . init_pose */
} dif_gn_problem;
typedef struct dif_gn_result {
    double  delta[12];
    double  energy;              /* energy of the last evaluated iterate */
    int32_t last_iter, status, empty_term, n_iterations, n_sdf, n_rgb;
} dif_gn_result;
/* host evaluation of one update step with the device loop's own functions (H [36], g [6], delta [12] in/out): DIF_OK, or
 * DIF_GN_SINGULAR; test hook, no GPU needed */
int dif_debug_gn_step(const double* H, const double* g, double* delta_inout);
size_t dif_gn_scratch_bytes(int64_t n_obs);
int dif_gauss_newton(const dif_map_view* map, const void* decoder_prepared, const dif_gn_problem* problem, void* scratch, size_t scratch_bytes,
                     void* mailbox_host, dif_gn_result* result, void* stream);

/* ---- latent optimisation  (system/map.py:80-117 OptimizeProcess.do_optimize; disabled in the shipped loop) ------
 * One backward pass of the refinement loss  sum_i -log N(clamp(gt_sdf_i, +-0.2); clamp(sdf_i, +-0.2), std_i) / n_div  through the
 * decoder into the unique latent rows: sample i decodes latent_u[inv[i]] at rel_xyz[i]; grad_u[inv[i]][:] += d loss / d latent
 * (grad_u zero-filled by the caller, map.py:102 optimizer.zero_grad); *loss_out += loss (nullable).  n_div = the reference's
 * n_samples (map.py:86: all gathered samples, also when forward_model splits them into chunks).  Exact fp32. */
int dif_latent_grad(const void* decoder_prepared, const float* latent_u /*[U][29]*/, const int64_t* inv /*[n]*/, const float* rel_xyz /*[n][3]*/,
                    const float* gt_sdf /*[n]*/, int64_t n, int64_t n_div, float* grad_u /*[U][29]*/, double* loss_out, void* stream);

/* ---- groupby_sum  (system/ext/indexing/indexing.cu:59-109; indexing.cpp:4) -----------------------------
 * sum[indices[i]][:] += values[i][:];  count[indices[i]] += L  (the reference bumps the count once per column, :70).
 * sum/count must be zero-filled by the caller (the reference allocates zeros, :96-97). */
int dif_groupby_sum(const float* values /*[n][L]*/, const int64_t* indices /*[n]*/, int64_t n, int32_t L, int64_t C,
                    float* sum /*[C][L]*/, int32_t* count /*[C]*/, void* stream);

/* ---- measurement hooks (bench.py; not part of the reference's interface) -------------------------------------
 * dif_profile_hook: bracket the NEXT launch of the named kernel on this host thread with the two CUDA events
 *   (cudaEvent_t handles, recorded on the launch stream); one-shot, NULL/NULL disarms.
 * dif_launch_count: kernels launched by this host thread since the last reset. */
enum { DIF_PROF_ENCODE = 0, DIF_PROF_ICP = 1, DIF_PROF_DECODE = 2, DIF_PROF_MC = 3,
       DIF_PROF_INDEX = 4 /* voxelize .. gather of dif_integrate */, DIF_PROF_FUSE = 5, DIF_PROF_COUNT = 6 };
int dif_profile_hook(int which, void* start_event, void* stop_event);
uint64_t dif_launch_count(int reset);
/* dif_debug_tc_timing: dev_buf = uint64[148*20*8] or NULL; when set, the tensor-core decoder records per-warp phase cycles
 * (development aid used by tools/tc_timing.py). */
int dif_debug_tc_timing(void* dev_buf);

/* ---- hash-sharded map (new; the reference is single-GPU.  SURVEY 8e, BASELINE configs[4]) -----------------------------------------
 * Ownership: owner(cell) = splitmix64(id of the cell's super-block) % shard_world, super-block = (2^shard_block_log2)^3 cells
 * (default 16^3).  Integer state (indexer, slot numbering, latent_vecs_pos, voxel_obs_count) is replicated: every rank runs the same
 * index kernels on the same frame.  The floating-point payload is sharded: a rank stores latent rows only for the PLIVoxes it owns
 * plus a one-cell halo around its super-blocks (the 26-neighbourhood the marching-cubes blend of an owned PLIVox reads,
 * mc_interp_kernel.cu:103-181), addressed through row_of_slot; the encoder MLP and the fusion of a PLIVox run on its owner only.
 * Per frame ONE exchange: the owner of a row that another rank keeps in its halo (a BOUNDARY row: the PLIVox touches a face, edge
 * or corner of its super-block) sends it to exactly those ranks.
 *   dif_shard_pack   : map->xchg_slots[0 .. *n_xchg_dev) (owned rows fused this frame, written by dif_integrate) -> send_buf =
 *                      [world][1 + cap_rows][32] floats; segment d is what rank d receives.  Segment row 0 = header: word 0 = rows for
 *                      d (may exceed cap_rows: the excess is dropped and reported), word 1 = the largest per-destination count of
 *                      this sender; row 1+i: slot (int32 bits), 29 latents, 2 pad words.
 *   one all-to-all   : equal splits of dif_shard_xchg_bytes(cap_rows, 1) bytes (torch.distributed.all_to_all_single).
 *   dif_shard_unpack : recv_buf [world][1 + cap_rows][32] -> the halo rows of this rank.  *overflow_dev = max over senders of header
 *                      word 1 if it exceeds cap_rows - the same value on every rank (each receives every sender's header), so the
 *                      host's lazy recovery (grow + re-publish) is a collective decision without a collective.
 * No host synchronisation anywhere; sizes never leave the device.
 *   dif_shard_select_points: the tracker's points whose PLIVox this rank owns, compacted (camera frame, order not preserved) into
 *                      out_obs, with their count and the poses in *out_frame_dev: dif_icp_linearize(out_obs, n, NULL, out_frame_dev)
 *                      then linearises this rank's share; the caller all-reduces the 44 doubles. */
size_t dif_shard_xchg_bytes(int64_t cap_rows, int world);
int dif_shard_pack(const dif_map_view* map, const int32_t* n_xchg_dev, int64_t cap_rows, float* send_buf, void* stream);
int dif_shard_unpack(const dif_map_view* map, const float* recv_buf /*[world][1 + cap_rows][32]*/, int64_t cap_rows,
                     int32_t* overflow_dev, void* stream);
int dif_shard_select_points(const dif_map_view* map, const float* obs_xyz /*[n][3] camera frame*/, int64_t n, const float* pose_host /*[24]*/,
                            float* out_obs /*[n][3]*/, dif_frame_params* out_frame_dev, void* stream);

/* owner rank of a PLIVox (host mirror of the device function): linear id -> super-block -> splitmix64 % world. */
int dif_shard_owner(int64_t linear_id, int nx, int ny, int nz, int block_log2, int world);

int dif_abi_version(void);
const char* dif_last_error(void);        /* thread-local text of the last DIF_E_LAUNCH */

#ifdef __cplusplus
}
#endif
#endif
