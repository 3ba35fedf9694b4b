// Mesh extraction: per-PLIVox lattice decode (low pass -> trilinear x2 -> select -> high pass) and sparse marching cubes
// with cross-PLIVox std-weighted blending; plus groupby_sum.
//   replaces reference system/map.py:640-687 (do_meshing decode), system/ext/marching_cubes/mc_interp_kernel.cu:7-382,
//   system/ext/indexing/indexing.cu:59-109.  SURVEY rows a-10, a-11, a-6, A.10, A.11.
#include "mlp_simt.cuh"
#include "../../include/dif_mc_tables.h"

namespace dif {

int launch_decode_lattice(const float* P, const float* latent, const int32_t* block_slots, int64_t n_blocks, int lat_n, float lat_step,
                          float lat_a, const uint32_t* list, const int32_t* n_dev, int64_t n_max, float sdf_sign, float* sdf, float* std,
                          cudaStream_t st);

// ------------------------------------------------------------------------------------------------ trilinear x2 + select
// torch.nn.functional.interpolate(mode='trilinear', align_corners=True), l^3 -> (2l)^3 per PLIVox (map.py:658-665), then the
// |sdf| < 0.05 test (:667).  Writes the NEGATED sdf (:687) and std into the cubes and appends the global lattice index of
// every selected sample to `list` (warp-aggregated).
__global__ void upsample_select_kernel(const float* __restrict__ low_sdf, const float* __restrict__ low_std, int64_t n_blocks, int l,
                                       float* __restrict__ cube_sdf, float* __restrict__ cube_std, uint32_t* __restrict__ list,
                                       int32_t* __restrict__ n_sel) {
    const int h = 2 * l, h3 = h * h * h, l3 = l * l * l;
    const int64_t total = n_blocks * h3;
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool sel = false;
    if (g < total) {
        const int64_t b = g / h3; const int i = (int)(g % h3);
        const int X = i / (h * h), Y = (i / h) % h, Z = i % h;
        const float scale = (float)(l - 1) / (float)(h - 1);                  // area_pixel_compute_scale, align_corners
        int i0[3], i1[3]; float w0[3], w1[3];
        const int D[3] = {X, Y, Z};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float src = __fmul_rn(scale, (float)D[a]);
            i0[a] = (int)src; i1[a] = i0[a] + (i0[a] < l - 1 ? 1 : 0);
            w1[a] = __fsub_rn(src, (float)i0[a]); w0[a] = __fsub_rn(1.f, w1[a]);
        }
        const float* ps = low_sdf + b * l3; const float* pd = low_std + b * l3;
        float out[2];
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            const float* p = which ? pd : ps;
            float acc_d[2];
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int xo = (dx ? i1[0] : i0[0]) * l * l;
                float acc_h[2];
#pragma unroll
                for (int dy = 0; dy < 2; ++dy) {
                    const int yo = xo + (dy ? i1[1] : i0[1]) * l;
                    acc_h[dy] = __fadd_rn(__fmul_rn(w0[2], p[yo + i0[2]]), __fmul_rn(w1[2], p[yo + i1[2]]));
                }
                acc_d[dx] = __fadd_rn(__fmul_rn(w0[1], acc_h[0]), __fmul_rn(w1[1], acc_h[1]));
            }
            out[which] = __fadd_rn(__fmul_rn(w0[0], acc_d[0]), __fmul_rn(w1[0], acc_d[1]));
        }
        cube_sdf[g] = -out[0];
        cube_std[g] = out[1];
        sel = fabsf(out[0]) < 0.05f;
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, sel);
    if (ballot) {
        const int lane = threadIdx.x & 31;
        int base = 0;
        if (lane == 0) base = atomicAdd(n_sel, __popc(ballot));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (sel) list[base + __popc(ballot & ((1u << lane) - 1u))] = (uint32_t)g;
    }
}

__global__ void write_counts_kernel(int32_t* counts, int32_t n_low, const int32_t* n_sel) { counts[0] = n_low; counts[1] = n_sel ? *n_sel : 0; }

// ------------------------------------------------------------------------------------------------ marching cubes
// One CTA per focused PLIVox.  Stage 1: the 27 neighbour batch indices.  Stage 2: the (r+1)^3 blended corner values
// (each computed once instead of up to 8x as in the reference's per-sub-cube threads).  Stage 3: one thread per
// sub-cube: case lookup, edge vertices, block-scan of triangle counts, ONE atomicAdd per CTA to reserve output, emit.
// Float arithmetic is written with explicit round-to-nearest intrinsics and explicit fmaf in exactly the places where nvcc
// (-fmad=true) contracts the reference source (read off the reference's PTX; see oracle/mc_oracle.c header), so case indices
// and vertices are bit-identical to the reference extension and to the scalar restatement in oracle/mc_oracle.c.
constexpr int MC_THREADS = 128;
constexpr int MC_MAX_R = 8;

struct McArgs {
    const int64_t* indexer; int nx, ny, nz; const int64_t* valid_blocks; int64_t n_valid; const int32_t* mapping; int64_t mapping_len;
    const float* cube_sdf; const float* cube_std; int r; float max_std; float* tri; int64_t* tri_id; float* tri_std; int64_t max_tri;
    int32_t* count;
};

__device__ __forceinline__ float2 blended_corner(const McArgs& a, const int* nb /*[27]*/, int px, int py, int pz) {
    const int r = a.r, n = 2 * r;
    const int rbound = (r - 1) / 2, rstart = r / 2;
    const float rmid = r / 2.0f, rf = (float)r;
    const int pos[3] = {px, py, pz};
    float w_m[3], w_p[3]; int b_m[3], b_p[3], a_m[3], a_p[3], own_is_p[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (pos[k] <= rbound) {                      // contributors: previous PLIVox (far half of its cube) and own
            b_m[k] = -1; b_p[k] = 0; a_m[k] = pos[k] + rstart + r; a_p[k] = pos[k] + rstart;
            w_p[k] = __fadd_rn((float)pos[k], rmid); w_m[k] = __fsub_rn(rmid, (float)pos[k]); own_is_p[k] = 1;
        } else {                                     // own and next PLIVox
            b_m[k] = 0; b_p[k] = 1; a_m[k] = pos[k] + rstart; a_p[k] = pos[k] + rstart - r;
            w_p[k] = __fsub_rn((float)pos[k], rmid); w_m[k] = __fsub_rn(__fadd_rn(rmid, rf), (float)pos[k]); own_is_p[k] = 0;
        }
        w_m[k] = __fdiv_rn(w_m[k], rf); w_p[k] = __fdiv_rn(w_p[k], rf);
    }
    const int own = own_is_p[0] * 4 + own_is_p[1] * 2 + own_is_p[2];
    float s1 = 0.f, s2 = 0.f, s4 = 0.f;
    const float qnan = __int_as_float(0x7fc00000);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int xp = (k >> 2) & 1, yp = (k >> 1) & 1, zp = k & 1;
        const int bx = xp ? b_p[0] : b_m[0], by = yp ? b_p[1] : b_m[1], bz = zp ? b_p[2] : b_m[2];
        const int batch = nb[(bx + 1) * 9 + (by + 1) * 3 + (bz + 1)];
        float sdf = qnan, sd = qnan;
        if (batch >= 0) {
            const int64_t off = (((int64_t)batch * n + (xp ? a_p[0] : a_m[0])) * n + (yp ? a_p[1] : a_m[1])) * n + (zp ? a_p[2] : a_m[2]);
            sdf = __ldg(a.cube_sdf + off); sd = __ldg(a.cube_std + off);
        }
        const float w = __fmul_rn(__fmul_rn(xp ? w_p[0] : w_m[0], yp ? w_p[1] : w_m[1]), zp ? w_p[2] : w_m[2]);
        if (sdf == sdf) {
            s1 = __fmaf_rn(__fmul_rn(sdf, w), sd, s1);
            s2 = __fmaf_rn(w, sd, s2);
            s4 = __fadd_rn(s4, w);
        } else if (own == k) {
            return make_float2(qnan, qnan);
        }
    }
    return make_float2(__fdiv_rn(s1, s2), __fdiv_rn(s2, s4));
}

__device__ __forceinline__ float4 edge_vertex(float3 p1, float3 p2, float std1, float std2, float v1, float v2) {
    if (fabsf(__fsub_rn(0.0f, v1)) < 1.0e-5f) return make_float4(p1.x, p1.y, p1.z, std1);
    if (fabsf(__fsub_rn(0.0f, v2)) < 1.0e-5f) return make_float4(p2.x, p2.y, p2.z, std2);
    if (fabsf(__fsub_rn(v1, v2)) < 1.0e-5f) return make_float4(p1.x, p1.y, p1.z, std1);
    const float w2 = __fdiv_rn(__fsub_rn(0.0f, v1), __fsub_rn(v2, v1));
    const float w1 = __fsub_rn(1.f, w2);
    return make_float4(__fmaf_rn(p2.x, w2, __fmul_rn(p1.x, w1)), __fmaf_rn(p2.y, w2, __fmul_rn(p1.y, w1)),
                       __fmaf_rn(p2.z, w2, __fmul_rn(p1.z, w1)), __fmaf_rn(std2, w2, __fmul_rn(std1, w1)));
}

__global__ void __launch_bounds__(MC_THREADS) marching_cubes_kernel(McArgs a) {
    __shared__ int nb[27];
    __shared__ float c_sdf[(MC_MAX_R + 1) * (MC_MAX_R + 1) * (MC_MAX_R + 1)];
    __shared__ float c_std[(MC_MAX_R + 1) * (MC_MAX_R + 1) * (MC_MAX_R + 1)];
    __shared__ int warp_tot[MC_THREADS / 32];
    __shared__ int block_base;
    const int r = a.r, r1 = r + 1, r3 = r * r * r, nc = r1 * r1 * r1;
    const float sbs = __fdiv_rn(1.0f, (float)r);
    const int dx8[8] = {0, 1, 1, 0, 0, 1, 1, 0}, dy8[8] = {0, 0, 1, 1, 0, 0, 1, 1}, dz8[8] = {0, 0, 0, 0, 1, 1, 1, 1};
    const int e_a[12] = {0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3}, e_b[12] = {1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7};

    for (int64_t blk = blockIdx.x; blk < a.n_valid; blk += gridDim.x) {
        const int64_t id = a.valid_blocks[blk];
        const int bx = (int)((id / ((int64_t)a.ny * a.nz)) % a.nx), by = (int)((id / a.nz) % a.ny), bz = (int)(id % a.nz);
        if (threadIdx.x < 27) {
            const int ox = threadIdx.x / 9 - 1, oy = (threadIdx.x / 3) % 3 - 1, oz = threadIdx.x % 3 - 1;
            const int x = bx + ox, y = by + oy, z = bz + oz;
            int batch = -1;
            if ((unsigned)x < (unsigned)a.nx && (unsigned)y < (unsigned)a.ny && (unsigned)z < (unsigned)a.nz) {
                const int64_t slot = a.indexer[((int64_t)x * a.ny + y) * a.nz + z];
                if (slot != -1 && slot < a.mapping_len) batch = a.mapping[slot];
            }
            nb[threadIdx.x] = batch;
        }
        __syncthreads();
        for (int c = threadIdx.x; c < nc; c += MC_THREADS) {
            const float2 v = blended_corner(a, nb, c / (r1 * r1), (c / r1) % r1, c % r1);
            c_sdf[c] = v.x; c_std[c] = v.y;
        }
        __syncthreads();
        for (int sub0 = 0; sub0 < r3; sub0 += MC_THREADS) {
            const int sub = sub0 + threadIdx.x;
            int n_tri = 0, type = 0;
            float4 vert[12];
            int rx = 0, ry = 0, rz = 0;
            if (sub < r3) {
                rx = sub / (r * r); ry = (sub / r) % r; rz = sub % r;
                float v[8], sd[8]; bool bad = false;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int c = ((rx + dx8[i]) * r1 + ry + dy8[i]) * r1 + rz + dz8[i];
                    v[i] = c_sdf[c]; sd[i] = c_std[c];
                    bad |= !(v[i] == v[i]);
                    if (v[i] < 0.f) type |= 1 << i;
                }
                const int emask = bad ? 0 : dif_mc_edge_mask[type];
                if (emask) {
#pragma unroll
                    for (int e = 0; e < 12; ++e) {
                        if (emask & (1 << e)) {
                            const int p = e_a[e], q = e_b[e];
                            const float3 p1 = make_float3(__fmaf_rn((float)(rx + dx8[p]), sbs, (float)bx),
                                                          __fmaf_rn((float)(ry + dy8[p]), sbs, (float)by),
                                                          __fmaf_rn((float)(rz + dz8[p]), sbs, (float)bz));
                            const float3 p2 = make_float3(__fmaf_rn((float)(rx + dx8[q]), sbs, (float)bx),
                                                          __fmaf_rn((float)(ry + dy8[q]), sbs, (float)by),
                                                          __fmaf_rn((float)(rz + dz8[q]), sbs, (float)bz));
                            vert[e] = edge_vertex(p1, p2, sd[p], sd[q], v[p], v[q]);
                        }
                    }
                    const int nt = dif_mc_tri_count[type];
                    for (int t = 0; t < nt; ++t) {
                        const float s0 = vert[dif_mc_tri_edges[type][3 * t]].w, s1 = vert[dif_mc_tri_edges[type][3 * t + 1]].w,
                                    s2 = vert[dif_mc_tri_edges[type][3 * t + 2]].w;
                        if (!(s0 > a.max_std || s1 > a.max_std || s2 > a.max_std)) ++n_tri;
                    }
                } else {
                    type = 0;
                }
            }
            // block exclusive scan of n_tri, one reservation per CTA
            int incl = n_tri;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += u; }
            if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
            __syncthreads();
            if (threadIdx.x == 0) {
                int tot = 0;
                for (int w = 0; w < MC_THREADS / 32; ++w) { const int t = warp_tot[w]; warp_tot[w] = tot; tot += t; }
                block_base = tot ? atomicAdd(a.count, tot) : 0;
            }
            __syncthreads();
            int64_t out = (int64_t)block_base + warp_tot[threadIdx.x >> 5] + incl - n_tri;
            if (n_tri) {
                const int nt = dif_mc_tri_count[type];
                for (int t = 0; t < nt; ++t) {
                    const float4 v0 = vert[dif_mc_tri_edges[type][3 * t]], v1 = vert[dif_mc_tri_edges[type][3 * t + 1]],
                                 v2 = vert[dif_mc_tri_edges[type][3 * t + 2]];
                    if (v0.w > a.max_std || v1.w > a.max_std || v2.w > a.max_std) continue;
                    if (out < a.max_tri) {
                        float* o = a.tri + out * 9;
                        o[0] = v0.x; o[1] = v0.y; o[2] = v0.z; o[3] = v1.x; o[4] = v1.y; o[5] = v1.z; o[6] = v2.x; o[7] = v2.y; o[8] = v2.z;
                        float* os = a.tri_std + out * 3;
                        os[0] = v0.w; os[1] = v1.w; os[2] = v2.w;
                        a.tri_id[out] = id;
                    }
                    ++out;
                }
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------------ groupby_sum
__global__ void groupby_sum_kernel(const float* __restrict__ values, const int64_t* __restrict__ indices, int64_t n, int L, int64_t C,
                                   float* __restrict__ sum, int32_t* __restrict__ count) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * L) return;
    const int64_t i = e / L; const int l = (int)(e % L);
    const int64_t g = indices[i];
    if (g < 0 || g >= C) return;
    atomicAdd(sum + g * L + l, values[e]);
    if (l == 0) atomicAdd(count + g, L);
}

}  // namespace dif

using namespace dif;

extern "C" {

size_t dif_mesh_decode_scratch_bytes(int64_t n_blocks, int r) {
    const int64_t l3 = (int64_t)r * r * r, h3 = 8 * l3;
    return 2 * align_up((size_t)n_blocks * l3 * 4) + align_up((size_t)n_blocks * h3 * 4) + 256;
}

int dif_mesh_decode(const dif_map_view* map, const void* decoder_prepared, const int32_t* block_slots, int64_t n_blocks, int r, int fast,
                    float* cube_sdf, float* cube_std, void* scratch, size_t scratch_sz, int32_t* counts_dev, void* stream) {
    if (!map || !decoder_prepared || n_blocks < 0 || r < 1 || r > MC_MAX_R || !counts_dev) return DIF_E_INVALID;
    if (n_blocks > 0 && (!block_slots || !cube_sdf || !cube_std || !scratch)) return DIF_E_INVALID;
    if (n_blocks * 8 * r * r * r >= (int64_t(1) << 32)) return DIF_E_INVALID;
    if (scratch_sz < dif_mesh_decode_scratch_bytes(n_blocks, r)) return DIF_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const float* P = (const float*)decoder_prepared;
    // sample lattice of map.py:640-646: [a, b] in voxel units, spacing (b-a)/(n-1)
    const double sa = -(r / 2) * (1.0 / r), sb = 1.0 + ((r - 1) / 2) * (1.0 / r);
    const int hr = 2 * r, lr = fast ? r : hr;
    const int64_t l3 = (int64_t)lr * lr * lr, h3 = (int64_t)hr * hr * hr;
    if (n_blocks == 0) { write_counts_kernel<<<1, 1, 0, st>>>(counts_dev, 0, nullptr); return check_launch("dif_mesh_decode"); }
    int rc;
    if (!fast || lr < 2) {
        const float step = (float)((sb - sa) / (hr - 1));
        rc = launch_decode_lattice(P, map->latent_vecs, block_slots, n_blocks, hr, step, (float)sa, nullptr, nullptr, n_blocks * h3, -1.f,
                                   cube_sdf, cube_std, st);
        if (rc) return rc;
        write_counts_kernel<<<1, 1, 0, st>>>(counts_dev, (int32_t)(n_blocks * h3), nullptr);
        return check_launch("dif_mesh_decode");
    }
    Carver c(scratch);
    float* low_sdf = c.take<float>(n_blocks * l3);
    float* low_std = c.take<float>(n_blocks * l3);
    uint32_t* list = c.take<uint32_t>(n_blocks * h3);
    int32_t* n_sel = c.take<int32_t>(1);
    cudaMemsetAsync(n_sel, 0, sizeof(int32_t), st);
    const float step_l = (float)((sb - sa) / (lr - 1)), step_h = (float)((sb - sa) / (hr - 1));
    rc = launch_decode_lattice(P, map->latent_vecs, block_slots, n_blocks, lr, step_l, (float)sa, nullptr, nullptr, n_blocks * l3, 1.f,
                               low_sdf, low_std, st);
    if (rc) return rc;
    const int64_t total = n_blocks * h3;
    upsample_select_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(low_sdf, low_std, n_blocks, lr, cube_sdf, cube_std, list, n_sel);
    DIF_COUNT_LAUNCH(2);
    rc = launch_decode_lattice(P, map->latent_vecs, block_slots, n_blocks, hr, step_h, (float)sa, list, n_sel, total, -1.f, cube_sdf, cube_std, st);
    if (rc) return rc;
    write_counts_kernel<<<1, 1, 0, st>>>(counts_dev, (int32_t)(n_blocks * l3), n_sel);
    return check_launch("dif_mesh_decode");
}

int dif_marching_cubes(const int64_t* indexer, int nx, int ny, int nz, const int64_t* valid_blocks, int64_t n_valid,
                       const int32_t* vec_batch_mapping, int64_t mapping_len, const float* cube_sdf, const float* cube_std, int r,
                       float max_std, float* tri, int64_t* tri_flatten_id, float* tri_std, int64_t max_tri, int32_t* count_dev, void* stream) {
    if (!indexer || !count_dev || n_valid < 0 || r < 1 || r > MC_MAX_R || max_tri <= 0) return DIF_E_INVALID;
    if (n_valid > 0 && (!valid_blocks || !vec_batch_mapping || !cube_sdf || !cube_std || !tri || !tri_flatten_id || !tri_std)) return DIF_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(count_dev, 0, sizeof(int32_t), st);
    if (n_valid == 0) return check_launch("dif_marching_cubes");
    McArgs a{indexer, nx, ny, nz, valid_blocks, n_valid, vec_batch_mapping, mapping_len, cube_sdf, cube_std, r, max_std,
             tri, tri_flatten_id, tri_std, max_tri, count_dev};
    const int64_t cap = (int64_t)DIF_NUM_SMS * 16;
    prof_begin(DIF_PROF_MC, st);
    marching_cubes_kernel<<<(unsigned)(n_valid < cap ? n_valid : cap), MC_THREADS, 0, st>>>(a);
    prof_end(DIF_PROF_MC, st);
    DIF_COUNT_LAUNCH(1);
    return check_launch("marching_cubes_kernel");
}

int dif_groupby_sum(const float* values, const int64_t* indices, int64_t n, int32_t L, int64_t C, float* sum, int32_t* count, void* stream) {
    if (n < 0 || L <= 0 || C < 0 || (n > 0 && (!values || !indices || !sum || !count))) return DIF_E_INVALID;
    if (n == 0) return DIF_OK;
    const int64_t total = n * L;
    groupby_sum_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(values, indices, n, L, C, sum, count);
    DIF_COUNT_LAUNCH(1);
    return check_launch("groupby_sum_kernel");
}

}  // extern "C"
