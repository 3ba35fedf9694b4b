// Fused point-to-implicit ICP linearisation: pose transform -> PLIVox lookup -> decoder forward (+ backward wrt xyz) ->
// residual / Jacobian / Huber -> 6x6 normal equations, one launch, one 44-double result.
//   replaces reference system/tracker.py:174-218 (compute_sdf_Hg) + system/map.py:559-579 (get_sdf) + the autograd
//   backward of network/di_decoder.py.  SURVEY rows a-8, a-9, A.8, A.9.
// The reference runs ~15 torch launches + 2 boolean-mask syncs for the lookup, cuBLAS forward, an autograd backward that
// also builds parameter gradients, and three host syncs (.cpu() x2, .item()) per Gauss-Newton iteration.
#include "mlp_simt.cuh"
#include "icp_args.cuh"
#include <stdlib.h>

namespace dif {


constexpr int ICP_VALS = 32;     // 21 upper-tri H + 6 g + E + M, padded

__global__ void __launch_bounds__(MLP_THREADS) icp_linearize_kernel(
        MapRO m, const float* __restrict__ P, const float* __restrict__ obs, int obs_stride, int n_host, const dif_frame_params* __restrict__ frame,
        Pose pose_host, float huber_k, int want_grad, float* __restrict__ partials, unsigned int* __restrict__ done_counter,
        double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DecoderSmem& s = *reinterpret_cast<DecoderSmem*>(smem_raw);
    __shared__ float q_s[3 * MLP_T];
    __shared__ float r_s[MLP_T];
    __shared__ int slot_s[MLP_T];
    __shared__ bool is_last;
    __shared__ IcpFrame fr_s;
    if (threadIdx.x == 0) icp_resolve_frame(frame, pose_host, n_host, fr_s);
    __syncthreads();
    const Pose& pose = fr_s.pose;
    const int n = fr_s.n;
    float acc[29];                               // meaningful on lane 0 of warp 0 only
#pragma unroll
    for (int j = 0; j < 29; ++j) acc[j] = 0.f;

    const int n_tiles = (n + MLP_T - 1) / MLP_T;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int base = tile * MLP_T;
        if (threadIdx.x < MLP_T) {
            const int t = threadIdx.x, i = base + t;
            int slot = -1; float rx = 0.f, ry = 0.f, rz = 0.f;
            if (i < n) {
                const float* op = obs + (int64_t)obs_stride * i;
                const float x = op[0], y = op[1], z = op[2];
                // cur = (last . delta) @ obs  (tracker.py:181, motion_util.py:322-327)
                const float wx = fmaf(z, pose.Rc[2], fmaf(y, pose.Rc[1], x * pose.Rc[0])) + pose.tc[0];
                const float wy = fmaf(z, pose.Rc[5], fmaf(y, pose.Rc[4], x * pose.Rc[3])) + pose.tc[1];
                const float wz = fmaf(z, pose.Rc[8], fmaf(y, pose.Rc[7], x * pose.Rc[6])) + pose.tc[2];
                q_s[t] = fmaf(z, pose.Rd[2], fmaf(y, pose.Rd[1], x * pose.Rd[0])) + pose.td[0];             // delta @ obs (:196)
                q_s[MLP_T + t] = fmaf(z, pose.Rd[5], fmaf(y, pose.Rd[4], x * pose.Rd[3])) + pose.td[1];
                q_s[2 * MLP_T + t] = fmaf(z, pose.Rd[8], fmaf(y, pose.Rd[7], x * pose.Rd[6])) + pose.td[2];
                const float3 p = normalize_point(m.g, wx, wy, wz);
                const int ix = (int)ceilf(p.x) - 1, iy = (int)ceilf(p.y) - 1, iz = (int)ceilf(p.z) - 1;
                if (p.x == p.x && p.y == p.y && p.z == p.z && in_grid(m.g, ix, iy, iz)) {
                    const int64_t sl = m.indexer[lin_id(m.g, ix, iy, iz)];
                    if (sl >= 0 && m.obs[sl] > m.ignore_th) slot = m.row_of ? m.row_of[sl] : (int)sl;       // (latent ROW from here on)
                }
                rx = __fsub_rn(__fsub_rn(p.x, (float)ix), 0.5f); ry = __fsub_rn(__fsub_rn(p.y, (float)iy), 0.5f); rz = __fsub_rn(__fsub_rn(p.z, (float)iz), 0.5f);
            }
            slot_s[t] = slot;
            s.cat[(96 + 29) * MLP_TP + t] = slot >= 0 ? rx : 0.f;
            s.cat[(96 + 30) * MLP_TP + t] = slot >= 0 ? ry : 0.f;
            s.cat[(96 + 31) * MLP_TP + t] = slot >= 0 ? rz : 0.f;
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < MLP_T * 32; idx += MLP_THREADS) {
            const int t = idx / 32, j = idx % 32;
            if (j < DIF_L) { const int sl = slot_s[t]; s.cat[(96 + j) * MLP_TP + t] = sl >= 0 ? __ldg(m.latent + (int64_t)sl * m.lat_stride + j) : 0.f; }
        }
        __syncthreads();
        decoder_forward_tile(P, s);
        if (threadIdx.x < MLP_T) {
            const int t = threadIdx.x;
            const float sdf = tanhf(s.pre[t]);
            const float sd = 0.05f + 0.5f * softplus_ref(s.pre[MLP_T + t]);
            const float inv = 1.f / sd;
            r_s[t] = sdf / sd;                                   // tracker.py:186
            s.seed[t] = (1.f - sdf * sdf) * inv;                 // d r / d pre_sdf   (std detached)
        }
        __syncthreads();
        if (want_grad) decoder_backward_tile(P, s, 0);
        if (threadIdx.x < MLP_T) {
            const int t = threadIdx.x;
            const bool valid = slot_s[t] >= 0;
            float v[29];
#pragma unroll
            for (int j = 0; j < 29; ++j) v[j] = 0.f;
            if (valid) {
                const float r = r_s[t];
                float w = 1.f;
                if (huber_k > 0.f) { const float ar = fabsf(r); if (ar > huber_k) w = huber_k / ar; }       // tracker.py:59-65
                else if (huber_k < 0.f) { const float q = r / -huber_k, t1 = 1.f - q * q; w = fabsf(r) <= -huber_k ? t1 * t1 : 0.f; }   // Tukey (:66-69)
                v[27] = r * (r * w);                             // energy term  (:210)
                v[28] = 1.f;
                if (want_grad) {
                    // G = d r / d p_world = gx / voxel_size ; A = G @ R_last^T as coded (:197-198) ; B = q x A (:199)
                    const float gx = s.gx[t] / m.g.vs, gy = s.gx[MLP_T + t] / m.g.vs, gz = s.gx[2 * MLP_T + t] / m.g.vs;
                    float J[6];
                    J[0] = gx * pose.Rl[0] + gy * pose.Rl[1] + gz * pose.Rl[2];
                    J[1] = gx * pose.Rl[3] + gy * pose.Rl[4] + gz * pose.Rl[5];
                    J[2] = gx * pose.Rl[6] + gy * pose.Rl[7] + gz * pose.Rl[8];
                    const float qx = q_s[t], qy = q_s[MLP_T + t], qz = q_s[2 * MLP_T + t];
                    J[3] = qy * J[2] - qz * J[1];
                    J[4] = qz * J[0] - qx * J[2];
                    J[5] = qx * J[1] - qy * J[0];
                    int k = 0;
#pragma unroll
                    for (int a = 0; a < 6; ++a)
#pragma unroll
                        for (int b = a; b < 6; ++b) v[k++] = (w * J[a]) * J[b];       // H = sum (wJ)^T J  (:215)
#pragma unroll
                    for (int a = 0; a < 6; ++a) v[21 + a] = J[a] * (r * w);           // g = sum J (w r)    (:216)
                }
            }
#pragma unroll
            for (int j = 0; j < 29; ++j) acc[j] += warp_sum(v[j]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < 29; ++j) partials[blockIdx.x * ICP_VALS + j] = acc[j];
        __threadfence();
        is_last = atomicAdd(done_counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last && threadIdx.x < 32) {
        __threadfence();
        double tot = 0.0;
        if (threadIdx.x < 29) for (unsigned b = 0; b < gridDim.x; ++b) tot += (double)__ldcg(partials + b * ICP_VALS + threadIdx.x);
        const double M = __shfl_sync(0xffffffffu, tot, 28);
        const double scale = M > 0.0 ? 1.0 / M : 0.0;              // error_scale = 1/M (:209)
        if (threadIdx.x < 21) {
            int a = 0, rem = threadIdx.x;
            while (rem >= 6 - a) { rem -= 6 - a; ++a; }
            const int b = a + rem;
            out[a * 6 + b] = tot * scale; out[b * 6 + a] = tot * scale;
        } else if (threadIdx.x < 27) out[36 + threadIdx.x - 21] = tot * scale;
        else if (threadIdx.x == 27) out[42] = tot * scale;
        else if (threadIdx.x == 28) out[43] = M;
        if (threadIdx.x == 0) *done_counter = 0u;
    }
}

}  // namespace dif

extern "C" size_t dif_icp_scratch_bytes(int64_t n);

namespace dif {
int icp_launch(const dif_map_view* map, const void* decoder_prepared, const float* obs_xyz, int obs_stride, int64_t n, const float* pose_host,
               const dif_frame_params* frame_dev, float huber_k, int want_grad, void* scratch, size_t scratch_sz, double* out_dev, cudaStream_t st) {
    if (!map || !decoder_prepared || (!pose_host && !frame_dev) || !scratch || !out_dev || n < 0 || n >= (int64_t(1) << 31) || (n > 0 && !obs_xyz))
        return DIF_E_INVALID;
    if (scratch_sz < dif_icp_scratch_bytes(n)) return DIF_E_WORKSPACE;
    MapRO m{map->indexer, map->latent_vecs, map->voxel_obs_count, make_grid(map), map->ignore_count_th, map->latent_stride > 0 ? map->latent_stride : DIF_L,
            map->shard_world > 1 ? map->row_of_slot : nullptr};
    Pose p = {};
    if (pose_host) compose_pose(pose_host, p);
    Carver c(scratch);
    float* partials = c.take<float>((size_t)DIF_NUM_SMS * 3 * ICP_VALS);
    unsigned int* counter = c.take<unsigned int>(1);
    unsigned long long* ll = c.take<unsigned long long>((size_t)DIF_NUM_SMS * 64);
    unsigned int* epoch = c.take<unsigned int>(1);
    // `scratch` is zero-filled once by the caller; both kernels leave the counter zeroed again when they finish and write
    // every partial row they later read, so no memset is needed per call.
    // tensor-core path (decoder forward + backward on tcgen05) for frames worth of points; DIF_ICP_PATH=simt forces fp32 SIMT
    const char* path_env = getenv("DIF_ICP_PATH");                     // read per call so tests can compare both paths
    const bool force_simt = path_env && path_env[0] == 's';
    if (!force_simt && n >= 2048) {
        static_assert((size_t)DIF_NUM_SMS * 32 * sizeof(double) <= (size_t)DIF_NUM_SMS * 3 * ICP_VALS * sizeof(float), "partials region");
        IcpTcArgs a{m, obs_xyz, obs_stride, (int)n, frame_dev, p, huber_k, want_grad, reinterpret_cast<double*>(partials), counter, out_dev, ll, epoch};
        return launch_icp_tc(decoder_prepared, a, st);
    }
    const int64_t n_tiles = (n + MLP_T - 1) / MLP_T;
    int grid = (int)(n_tiles < DIF_NUM_SMS * 3 ? n_tiles : DIF_NUM_SMS * 3);
    if (grid < 1) grid = 1;
    const size_t smem = sizeof(DecoderSmem);
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(icp_linearize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_set = true; }
    prof_begin(DIF_PROF_ICP, st);
    icp_linearize_kernel<<<grid, MLP_THREADS, smem, st>>>(m, (const float*)decoder_prepared, obs_xyz, obs_stride, (int)n, frame_dev, p, huber_k,
                                                          want_grad, partials, counter, out_dev);
    prof_end(DIF_PROF_ICP, st);
    DIF_COUNT_LAUNCH(1);
    return check_launch("icp_linearize_kernel");
}
}  // namespace dif

using namespace dif;

extern "C" {

size_t dif_icp_scratch_bytes(int64_t n) {
    (void)n;
    return align_up((size_t)DIF_NUM_SMS * 3 * ICP_VALS * sizeof(float)) + 256 + align_up((size_t)DIF_NUM_SMS * 64 * sizeof(unsigned long long)) + 256;
}

int dif_icp_linearize(const dif_map_view* map, const void* decoder_prepared, const float* obs_xyz, int64_t n, const float* pose_host,
                      const dif_frame_params* frame_dev, float huber_k, int want_grad, void* scratch, size_t scratch_sz, double* out_dev, void* stream) {
    return icp_launch(map, decoder_prepared, obs_xyz, 3, n, pose_host, frame_dev, huber_k, want_grad, scratch, scratch_sz, out_dev, (cudaStream_t)stream);
}

}  // extern "C"
