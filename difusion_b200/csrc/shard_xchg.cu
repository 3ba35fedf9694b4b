// Hash-sharded map (BASELINE configs[4]): the per-frame exchange of BOUNDARY latent rows and the tracker's point
// selection.  The reference is single-GPU; there is nothing to cite except the coupling the design has to respect: the blend of
// ext/marching_cubes/mc_interp_kernel.cu:103-181 reads the 3x3x3 neighbourhood of a PLIVox, so a rank keeps a one-cell halo around
// the super-blocks it owns and the owner of a halo row sends it to exactly the ranks that keep it.
//
// Protocol (one all-to-all with equal splits per frame, no host synchronisation):
//   send_buf [world][1 + cap_rows][32] floats, segment d = what rank d receives
//     row 0      header: word 0 = rows this sender has for d (may exceed cap_rows: excess dropped, reported), word 1 = the largest
//                per-destination count of this sender (every receiver sees it, so all ranks derive the same overflow value)
//     row 1 + i  word 0 = PLIVox slot (int32 bits; slots are global because the integer state is replicated), words 1..29 = latent
//                row, words 30..31 = 0                                                              (128-byte rows)
// HBM-bound copies of 128-byte rows: one warp per row, lane = word.
#include "common.cuh"
#include "icp_args.cuh"

namespace dif {

constexpr int XROW = 32;

__global__ void shard_zero_headers_kernel(float* __restrict__ send, int world, int64_t seg_words) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < world * XROW) send[(int64_t)(i / XROW) * seg_words + (i % XROW)] = 0.f;
}

__global__ void shard_pack_kernel(const float* __restrict__ latent, int lat_stride, const int32_t* __restrict__ row_of,
                                  const int64_t* __restrict__ pos, Grid g, Shard sh, const int32_t* __restrict__ xchg_slots,
                                  const int32_t* __restrict__ n_xchg, int64_t cap_rows, float* __restrict__ send) {
    const int count = *n_xchg;
    const int lane = threadIdx.x & 31;
    const int64_t seg_words = (cap_rows + 1) * XROW;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < count; w += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const int slot = xchg_slots[w];
        const int64_t lin = pos[slot];
        uint32_t dests = shard_holder_mask(g.nx, g.ny, g.nz, lin, sh.k, sh.world) & ~(1u << sh.rank);     // boundary row: kept by other ranks
        if (!dests) continue;
        const int64_t r = row_of ? (int64_t)row_of[slot] : (int64_t)slot;
        float v = 0.f;
        if (lane == 0) v = __int_as_float(slot);
        else if (lane <= DIF_L && r >= 0) v = latent[r * lat_stride + (lane - 1)];
        while (dests) {
            const int d = __ffs(dests) - 1; dests &= dests - 1;
            float* seg = send + (int64_t)d * seg_words;
            int at = 0;
            if (lane == 0) at = atomicAdd(reinterpret_cast<int*>(seg), 1);
            at = __shfl_sync(0xffffffffu, at, 0);
            if (at < cap_rows) seg[(int64_t)(at + 1) * XROW + lane] = v;
        }
    }
}

// word 1 of every header = the sender's largest per-destination count
__global__ void shard_finish_headers_kernel(float* __restrict__ send, int world, int64_t seg_words) {
    int mx = 0;
    for (int d = 0; d < world; ++d) mx = max(mx, __float_as_int(send[(int64_t)d * seg_words]));
    if (threadIdx.x < world) send[(int64_t)threadIdx.x * seg_words + 1] = __int_as_float(mx);
}

__global__ void shard_unpack_kernel(float* __restrict__ latent, int lat_stride, const int32_t* __restrict__ row_of, int64_t capacity,
                                    const float* __restrict__ recv, int my_rank, int64_t cap_rows, int32_t* __restrict__ overflow) {
    const int src = blockIdx.y;
    const int64_t seg_words = (cap_rows + 1) * XROW;
    const float* seg = recv + (int64_t)src * seg_words;
    const int count = __float_as_int(seg[0]), sender_max = __float_as_int(seg[1]);
    if (sender_max > cap_rows && blockIdx.x == 0 && threadIdx.x == 0) atomicMax(overflow, sender_max);   // identical on every rank
    if (src == my_rank) return;
    const int64_t rows = count < cap_rows ? count : cap_rows;
    const int lane = threadIdx.x & 31;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < rows; w += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const float v = seg[(w + 1) * XROW + lane];
        const int slot = __float_as_int(__shfl_sync(0xffffffffu, v, 0));
        if (slot < 0 || slot >= capacity) continue;
        const int64_t r = row_of ? (int64_t)row_of[slot] : (int64_t)slot;
        if (r >= 0 && lane >= 1 && lane <= DIF_L) latent[r * lat_stride + (lane - 1)] = v;
    }
}

// tracker points owned by this rank, compacted (warp-aggregated; order not preserved) + the frame block for dif_icp_linearize
__global__ void shard_select_points_kernel(const int64_t* __restrict__ indexer, Grid g, Shard sh, const float* __restrict__ obs, int n,
                                           Pose pose, float* __restrict__ out_obs, dif_frame_params* __restrict__ out_frame) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool mine = false;
    float ox = 0.f, oy = 0.f, oz = 0.f;
    if (i < n) {
        ox = obs[3 * i]; oy = obs[3 * i + 1]; oz = obs[3 * i + 2];
        const float wx = fmaf(oz, pose.Rc[2], fmaf(oy, pose.Rc[1], ox * pose.Rc[0])) + pose.tc[0];      // same arithmetic as the ICP kernels
        const float wy = fmaf(oz, pose.Rc[5], fmaf(oy, pose.Rc[4], ox * pose.Rc[3])) + pose.tc[1];
        const float wz = fmaf(oz, pose.Rc[8], fmaf(oy, pose.Rc[7], ox * pose.Rc[6])) + pose.tc[2];
        const float3 p = normalize_point(g, wx, wy, wz);
        const int ix = (int)ceilf(p.x) - 1, iy = (int)ceilf(p.y) - 1, iz = (int)ceilf(p.z) - 1;
        if (p.x == p.x && p.y == p.y && p.z == p.z && in_grid(g, ix, iy, iz))
            mine = shard_owner_xyz(g.nx, g.ny, g.nz, ix, iy, iz, sh.k, sh.world) == sh.rank;
    }
    const unsigned b = __ballot_sync(0xffffffffu, mine);
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0 && b) base = atomicAdd(&out_frame->n_points, __popc(b));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (mine) {
        const int at = base + __popc(b & ((1u << lane) - 1u));
        out_obs[3 * at] = ox; out_obs[3 * at + 1] = oy; out_obs[3 * at + 2] = oz;
    }
}

struct Pose24 { float v[24]; };
__global__ void shard_begin_select_kernel(dif_frame_params* __restrict__ out_frame, Pose24 p) {
    if (threadIdx.x == 0) { out_frame->n_points = 0; out_frame->seq = 0; }
    if (threadIdx.x < 24) out_frame->pose[threadIdx.x] = p.v[threadIdx.x];
}

}  // namespace dif

using namespace dif;

static Shard shard_of(const dif_map_view* m) { Shard s; s.rank = m->shard_rank; s.world = m->shard_world > 1 ? m->shard_world : 1; s.k = m->shard_block_log2; return s; }

extern "C" {

size_t dif_shard_xchg_bytes(int64_t cap_rows, int world) { return (size_t)(world > 0 ? world : 1) * (size_t)(cap_rows + 1) * XROW * sizeof(float); }

int dif_shard_pack(const dif_map_view* map, const int32_t* n_xchg_dev, int64_t cap_rows, float* send_buf, void* stream) {
    if (!map || !map->xchg_slots || !n_xchg_dev || cap_rows <= 0 || !send_buf || map->shard_world < 2 || map->shard_world > 32) return DIF_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const int world = map->shard_world;
    const int64_t seg_words = (cap_rows + 1) * XROW;
    shard_zero_headers_kernel<<<(world * XROW + 255) / 256, 256, 0, st>>>(send_buf, world, seg_words);
    shard_pack_kernel<<<DIF_NUM_SMS * 4, 256, 0, st>>>(map->latent_vecs, map->latent_stride > 0 ? map->latent_stride : DIF_L, map->row_of_slot,
                                                       map->latent_vecs_pos, make_grid(map), shard_of(map), map->xchg_slots, n_xchg_dev, cap_rows, send_buf);
    shard_finish_headers_kernel<<<1, 32, 0, st>>>(send_buf, world, seg_words);
    DIF_COUNT_LAUNCH(3);
    return check_launch("shard_pack_kernel");
}

int dif_shard_unpack(const dif_map_view* map, const float* recv_buf, int64_t cap_rows, int32_t* overflow_dev, void* stream) {
    if (!map || !recv_buf || cap_rows <= 0 || !overflow_dev || map->shard_world < 2 || map->shard_world > 32) return DIF_E_INVALID;
    const dim3 grid(64, (unsigned)map->shard_world);
    shard_unpack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(map->latent_vecs, map->latent_stride > 0 ? map->latent_stride : DIF_L, map->row_of_slot,
                                                                map->capacity, recv_buf, map->shard_rank, cap_rows, overflow_dev);
    DIF_COUNT_LAUNCH(1);
    return check_launch("shard_unpack_kernel");
}

int dif_shard_select_points(const dif_map_view* map, const float* obs_xyz, int64_t n, const float* pose_host, float* out_obs,
                            dif_frame_params* out_frame_dev, void* stream) {
    if (!map || !pose_host || !out_obs || !out_frame_dev || n < 0 || n >= (int64_t(1) << 31) || (n > 0 && !obs_xyz)) return DIF_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    Pose p;
    compose_pose(pose_host, p);
    Pose24 p24;                                             // the poses travel as a kernel argument: the frame block is written on the device
    for (int i = 0; i < 24; ++i) p24.v[i] = pose_host[i];
    shard_begin_select_kernel<<<1, 32, 0, st>>>(out_frame_dev, p24);
    if (n > 0)
        shard_select_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(map->indexer, make_grid(map), shard_of(map), obs_xyz, (int)n, p,
                                                                                out_obs, out_frame_dev);
    DIF_COUNT_LAUNCH(2);
    return check_launch("shard_select_points_kernel");
}

int dif_shard_owner(int64_t linear_id, int nx, int ny, int nz, int block_log2, int world) {
    return world > 1 ? shard_owner_lin(nx, ny, nz, linear_id, block_log2, world) : 0;
}

}  // extern "C"
