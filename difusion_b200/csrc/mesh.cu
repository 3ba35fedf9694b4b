// Mesh extraction: per-PLIVox lattice decode (low pass -> trilinear x2 -> select -> high pass) and sparse marching cubes
// with cross-PLIVox std-weighted blending; plus groupby_sum.
//   replaces reference system/map.py:640-687 (do_meshing decode), system/ext/marching_cubes/mc_interp_kernel.cu:7-382,
//   system/ext/indexing/indexing.cu:59-109.  SURVEY rows a-10, a-11, a-6, A.10, A.11.
#include "mlp_simt.cuh"
#include "../../include/dif_mc_tables.h"

namespace dif {

int launch_decode_lattice(const float* P, const float* latent, int lat_stride, const int32_t* row_map, const int32_t* block_slots, int64_t n_blocks,
                          int lat_n, float lat_step, float lat_a, const uint32_t* list, const int32_t* n_dev, int64_t n_max, float sdf_sign, float* sdf, float* std,
                          cudaStream_t st);

// ------------------------------------------------------------------------------------------------ trilinear x2 + select
// torch.nn.functional.interpolate(mode='trilinear', align_corners=True), l^3 -> (2l)^3 per PLIVox (map.py:658-665), then the
// |sdf| < 0.05 test (:667).  Writes the NEGATED sdf (:687) and std into the cubes and appends the global lattice index of
// every selected sample to `list`.
// One CTA per PLIVox (grid-stride): the low cube is staged in shared memory with coalesced loads, the per-axis source index /
// weight tables (2l entries) are computed once per CTA, outputs are written as contiguous float streams, and the selected
// indices are collected in shared memory so that the global cursor sees ONE atomicAdd per PLIVox.
constexpr int UP_THREADS = 256;

template <int L>
__global__ void __launch_bounds__(UP_THREADS) upsample_select_kernel(const float* __restrict__ low_sdf, const float* __restrict__ low_std,
                                                                     int64_t n_blocks, float* __restrict__ cube_sdf, float* __restrict__ cube_std,
                                                                     uint32_t* __restrict__ list, int32_t* __restrict__ n_sel) {
    constexpr int H = 2 * L, H3 = H * H * H, L3 = L * L * L;
    __shared__ float s_sdf[L3], s_std[L3];
    __shared__ int t_i0[H], t_i1[H];
    __shared__ float t_w0[H], t_w1[H];
    __shared__ float z_sdf[L * L * H], z_std[L * L * H], y_sdf[L * H * H], y_std[L * H * H];
    __shared__ uint16_t s_list[H3];
    __shared__ int s_n, s_base;
    const int tid = threadIdx.x;
    if (tid < H) {
        const float scale = (float)(L - 1) / (float)(H - 1);                  // area_pixel_compute_scale, align_corners
        const float src = __fmul_rn(scale, (float)tid);
        const int i0 = (int)src;
        t_i0[tid] = i0; t_i1[tid] = i0 + (i0 < L - 1 ? 1 : 0);
        const float w1 = __fsub_rn(src, (float)i0);
        t_w1[tid] = w1; t_w0[tid] = __fsub_rn(1.f, w1);
    }
    for (int64_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        if (tid == 0) s_n = 0;
        for (int i = tid; i < L3; i += UP_THREADS) { s_sdf[i] = low_sdf[b * L3 + i]; s_std[i] = low_std[b * L3 + i]; }
        __syncthreads();
        // separable evaluation, same operations in the same order as the direct 8-tap form (z, then y, then x), each partial
        // result computed once instead of once per output that shares it: 5.25 instead of 21 flops per output and field
        for (int j = tid; j < L * L * H; j += UP_THREADS) {                    // along z: [x][y][Z]
            const int xy = j / H, Z = j % H;
            const int i0 = xy * L + t_i0[Z], i1 = xy * L + t_i1[Z];
            const float w0 = t_w0[Z], w1 = t_w1[Z];
            z_sdf[j] = __fadd_rn(__fmul_rn(w0, s_sdf[i0]), __fmul_rn(w1, s_sdf[i1]));
            z_std[j] = __fadd_rn(__fmul_rn(w0, s_std[i0]), __fmul_rn(w1, s_std[i1]));
        }
        __syncthreads();
        for (int j = tid; j < L * H * H; j += UP_THREADS) {                    // along y: [x][Y][Z]
            const int x = j / (H * H), Y = (j / H) % H, Z = j % H;
            const int i0 = (x * L + t_i0[Y]) * H + Z, i1 = (x * L + t_i1[Y]) * H + Z;
            const float w0 = t_w0[Y], w1 = t_w1[Y];
            y_sdf[j] = __fadd_rn(__fmul_rn(w0, z_sdf[i0]), __fmul_rn(w1, z_sdf[i1]));
            y_std[j] = __fadd_rn(__fmul_rn(w0, z_std[i0]), __fmul_rn(w1, z_std[i1]));
        }
        __syncthreads();
        for (int i = tid; i < H3; i += UP_THREADS) {                           // along x: [X][Y][Z], coalesced stores
            const int X = i / (H * H), yz = i % (H * H);
            const int i0 = t_i0[X] * (H * H) + yz, i1 = t_i1[X] * (H * H) + yz;
            const float w0 = t_w0[X], w1 = t_w1[X];
            const float o_sdf = __fadd_rn(__fmul_rn(w0, y_sdf[i0]), __fmul_rn(w1, y_sdf[i1]));
            const float o_std = __fadd_rn(__fmul_rn(w0, y_std[i0]), __fmul_rn(w1, y_std[i1]));
            cube_sdf[b * H3 + i] = -o_sdf;
            cube_std[b * H3 + i] = o_std;
            const bool sel = fabsf(o_sdf) < 0.05f;
            const unsigned ballot = __ballot_sync(__activemask(), sel);
            if (sel) {
                const unsigned lane = tid & 31, leader = __ffs(ballot) - 1;
                int base = 0;
                if (lane == leader) base = atomicAdd(&s_n, __popc(ballot));
                base = __shfl_sync(ballot, base, leader);
                s_list[base + __popc(ballot & ((1u << lane) - 1u))] = (uint16_t)i;
            }
        }
        __syncthreads();
        const int n = s_n;
        if (tid == 0 && n) s_base = atomicAdd(n_sel, n);
        __syncthreads();
        if (n) {
            const int base = s_base;
            const uint32_t g0 = (uint32_t)(b * H3);
            for (int i = tid; i < n; i += UP_THREADS) list[base + i] = g0 + s_list[i];
        }
        __syncthreads();
    }
}

__global__ void write_counts_kernel(int32_t* counts, int32_t n_low, const int32_t* n_sel) { counts[0] = n_low; counts[1] = n_sel ? *n_sel : 0; }

// ------------------------------------------------------------------------------------------------ marching cubes
// One CTA per focused PLIVox (grid-stride), compiled per sub-voxel resolution R.
//   stage 0 (once per CTA): case tables -> shared memory (a divergent __constant__ index serialises), per-axis blend tables
//            (weights need an IEEE division each: computed R+1 times per CTA instead of 6 times per corner).
//   stage 1: the 27 neighbour batch indices.
//   stage 2: the (R+1)^3 blended corner values, each computed once (the reference recomputes a corner in up to 8 sub-cube
//            threads).  All 16 cube loads of a corner are issued before the first is consumed (no early exit between them),
//            so a thread has 16 independent L2/HBM requests in flight instead of a chain of 8 round trips.
//   stage 3: (a) one thread per sub-cube: case lookup, block scan of the candidate triangle counts -> work list; (b) one thread
//            per candidate triangle: its three edge vertices, the max_std filter, compaction, ONE atomicAdd per chunk to reserve
//            output; triangles are staged in shared memory and written out as contiguous float streams (full sectors).
// Float arithmetic is written with explicit round-to-nearest intrinsics and explicit fmaf in exactly the places where nvcc
// (-fmad=true) contracts the reference source (read off the reference's PTX; see oracle/mc_oracle.c header), so case indices
// and vertices are bit-identical to the reference extension and to the scalar restatement in oracle/mc_oracle.c.
#ifndef DIF_MC_PREFETCH
#define DIF_MC_PREFETCH 1
#endif
constexpr int MC_THREADS = 128;
constexpr int MC_MAX_R = 8;
constexpr int MC_STAGE_TRIS = MC_THREADS;      // triangles staged per chunk: one candidate triangle per thread

struct McArgs {
    const int64_t* indexer; int nx, ny, nz; const int64_t* valid_blocks; int64_t n_valid; const int32_t* mapping; int64_t mapping_len;
    const float* cube_sdf; const float* cube_std; int r; float max_std; float* tri; int64_t* tri_id; float* tri_std; int64_t max_tri;
    int32_t* count;
};

__device__ __forceinline__ float4 edge_vertex(float3 p1, float3 p2, float std1, float std2, float v1, float v2) {
    if (fabsf(__fsub_rn(0.0f, v1)) < 1.0e-5f) return make_float4(p1.x, p1.y, p1.z, std1);
    if (fabsf(__fsub_rn(0.0f, v2)) < 1.0e-5f) return make_float4(p2.x, p2.y, p2.z, std2);
    if (fabsf(__fsub_rn(v1, v2)) < 1.0e-5f) return make_float4(p1.x, p1.y, p1.z, std1);
    const float w2 = __fdiv_rn(__fsub_rn(0.0f, v1), __fsub_rn(v2, v1));
    const float w1 = __fsub_rn(1.f, w2);
    return make_float4(__fmaf_rn(p2.x, w2, __fmul_rn(p1.x, w1)), __fmaf_rn(p2.y, w2, __fmul_rn(p1.y, w1)),
                       __fmaf_rn(p2.z, w2, __fmul_rn(p1.z, w1)), __fmaf_rn(std2, w2, __fmul_rn(std1, w1)));
}

template <int R>
__global__ void __launch_bounds__(MC_THREADS) marching_cubes_kernel(McArgs a) {
    constexpr int R1 = R + 1, NC = R1 * R1 * R1, R3 = R * R * R, N = 2 * R, N3 = N * N * N;
    constexpr int RBOUND = (R - 1) / 2, RSTART = R / 2;
    __shared__ int nb2[2][27];                              // double buffered: the next PLIVox's neighbours are fetched one round ahead
    __shared__ int s_b2[2][3];
    __shared__ int64_t s_id2[2];
    __shared__ float c_sdf[NC], c_std[NC];
    __shared__ float t_wm[R1], t_wp[R1];                    // per-axis blend weights of corner position p (mc_interp_kernel.cu:47-58)
    __shared__ int t_om[R1], t_op[R1];                      // per-axis cube coordinate read from the "minus" / "plus" contributor
    __shared__ uint8_t s_ntri[256];
    __shared__ int8_t s_tri[256][16];
    __shared__ float st_tri[MC_STAGE_TRIS * 9];
    __shared__ float st_std[MC_STAGE_TRIS * 3];
    __shared__ uint32_t s_work[MC_THREADS * 5];             // candidate triangles of one sub-cube pass: sub | t << 9 | case << 12
    __shared__ int warp_tot[MC_THREADS / 32];
    __shared__ int block_base;
    const int tid = threadIdx.x;
    const float sbs = __fdiv_rn(1.0f, (float)R);
    const float qnan = __int_as_float(0x7fc00000);
    const int dx8[8] = {0, 1, 1, 0, 0, 1, 1, 0}, dy8[8] = {0, 0, 1, 1, 0, 0, 1, 1}, dz8[8] = {0, 0, 0, 0, 1, 1, 1, 1};

    for (int i = tid; i < 256; i += MC_THREADS) s_ntri[i] = dif_mc_tri_count[i];
    for (int i = tid; i < 256 * 16; i += MC_THREADS) s_tri[i >> 4][i & 15] = dif_mc_tri_edges[i >> 4][i & 15];
    if (tid < R1) {
        const float rmid = R / 2.0f, rf = (float)R, pf = (float)tid;
        float wm, wp;
        if (tid <= RBOUND) {                                // contributors: previous PLIVox (far half of its cube) and own
            wp = __fadd_rn(pf, rmid); wm = __fsub_rn(rmid, pf);
            t_om[tid] = tid + RSTART + R; t_op[tid] = tid + RSTART;
        } else {                                            // own and next PLIVox
            wp = __fsub_rn(pf, rmid); wm = __fsub_rn(__fadd_rn(rmid, rf), pf);
            t_om[tid] = tid + RSTART; t_op[tid] = tid + RSTART - R;
        }
        t_wm[tid] = __fdiv_rn(wm, rf); t_wp[tid] = __fdiv_rn(wp, rf);
    }

    // neighbour fetch of one PLIVox: valid_blocks -> indexer -> mapping is a chain of three dependent global loads.  It is issued
    // by the 27 threads [NB_T0, NB_T0 + 27) of the CTA's last warp, which has no corner left in the second corner pass, for the
    // PLIVox of the NEXT round while the other warps finish the current one.
    constexpr int NB_T0 = MC_THREADS - 32;
    auto fetch_neighbours = [&](int64_t blk, int buf) {
        const int t = tid - NB_T0;
        if (t < 0 || t >= 27 || blk >= a.n_valid) return;
        const int64_t id = a.valid_blocks[blk];
        const int bx = (int)((id / ((int64_t)a.ny * a.nz)) % a.nx), by = (int)((id / a.nz) % a.ny), bz = (int)(id % a.nz);
        if (t == 0) { s_b2[buf][0] = bx; s_b2[buf][1] = by; s_b2[buf][2] = bz; s_id2[buf] = id; }
        const int x = bx + t / 9 - 1, y = by + (t / 3) % 3 - 1, z = bz + t % 3 - 1;
        int batch = -1;
        if ((unsigned)x < (unsigned)a.nx && (unsigned)y < (unsigned)a.ny && (unsigned)z < (unsigned)a.nz) {
            const int64_t slot = a.indexer[((int64_t)x * a.ny + y) * a.nz + z];
            if (slot != -1 && slot < a.mapping_len) batch = a.mapping[slot];
        }
        nb2[buf][t] = batch;
#if DIF_MC_PREFETCH
        // The PLIVox's OWN cube pair is pulled into L2 as two dense bulk prefetches (UBLKPF) one round ahead of its use: the gathers
        // below touch 12-24 bytes per 40-byte row, and when they are the first touch DRAM is read in sparse 64-byte pieces (ncu,
        // round 1: 2.4x the cube bytes from DRAM, 9 of 32 bytes used per sector); neighbours then find whole cubes in L2.
        if (t == 13 && batch >= 0) {
            constexpr uint32_t CB = (uint32_t)(N3 * sizeof(float)) & ~15u;
            const int64_t off = ((int64_t)batch * N3) & ~(int64_t)3;                       // 16-byte aligned start
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(a.cube_sdf + off), "r"(CB) : "memory");
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(a.cube_std + off), "r"(CB) : "memory");
        }
#endif
    };
#ifndef DIF_MC_ORDER
#define DIF_MC_ORDER 0                  // 0: grid-stride (CTAs sweep the sorted block list in lockstep windows)  1: one contiguous run per CTA
#endif
#if DIF_MC_ORDER == 1
    const int64_t per_cta = (a.n_valid + gridDim.x - 1) / gridDim.x;
    const int64_t blk_begin = blockIdx.x * per_cta, blk_end = blk_begin + per_cta < a.n_valid ? blk_begin + per_cta : a.n_valid, blk_step = 1;
#else
    const int64_t blk_begin = blockIdx.x, blk_end = a.n_valid, blk_step = gridDim.x;
#endif
    if (blk_begin < blk_end) fetch_neighbours(blk_begin, 0);
    int buf = 0;
    for (int64_t blk = blk_begin; blk < blk_end; blk += blk_step, buf ^= 1) {
        __syncthreads();                                    // nb2[buf] is complete (and, first round, the one-time tables)
        const int* nb = nb2[buf];
        const int64_t id = s_id2[buf];
        const bool own_ok = nb[13] >= 0;                    // a missing own cube makes every corner NaN: nothing to emit
        if (own_ok) {
            for (int c = tid; c < NC; c += MC_THREADS) {
                const int p[3] = {c / (R1 * R1), (c / R1) % R1, c % R1};
                const int lo[3] = {p[0] <= RBOUND, p[1] <= RBOUND, p[2] <= RBOUND};     // 1: (previous, own)  0: (own, next)
                const int nb0 = (1 - lo[0]) * 9 + (1 - lo[1]) * 3 + (1 - lo[2]);           // neighbour index of the "minus" contributors
                const int own = lo[0] * 4 + lo[1] * 2 + lo[2];                             // which k is the own cube
                const float wx[2] = {t_wm[p[0]], t_wp[p[0]]}, wy[2] = {t_wm[p[1]], t_wp[p[1]]}, wz[2] = {t_wm[p[2]], t_wp[p[2]]};
                const int ox[2] = {t_om[p[0]] * N * N, t_op[p[0]] * N * N}, oy[2] = {t_om[p[1]] * N, t_op[p[1]] * N}, oz[2] = {t_om[p[2]], t_op[p[2]]};
                float sv[8], dv[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {               // order mmm, mmp, mpm, mpp, pmm, pmp, ppm, ppp
                    const int xp = (k >> 2) & 1, yp = (k >> 1) & 1, zp = k & 1;
                    const int batch = nb[nb0 + xp * 9 + yp * 3 + zp];
                    sv[k] = qnan; dv[k] = qnan;
                    if (batch >= 0) {
                        const int64_t off = (int64_t)batch * N3 + (ox[xp] + oy[yp] + oz[zp]);
                        sv[k] = __ldg(a.cube_sdf + off); dv[k] = __ldg(a.cube_std + off);
                    }
                }
                float s1 = 0.f, s2 = 0.f, s4 = 0.f;
                bool own_missing = false;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int xp = (k >> 2) & 1, yp = (k >> 1) & 1, zp = k & 1;
                    const float w = __fmul_rn(__fmul_rn(wx[xp], wy[yp]), wz[zp]);
                    if (sv[k] == sv[k]) {
                        s1 = __fmaf_rn(__fmul_rn(sv[k], w), dv[k], s1);
                        s2 = __fmaf_rn(w, dv[k], s2);
                        s4 = __fadd_rn(s4, w);
                    } else if (own == k) {
                        own_missing = true;
                    }
                }
                c_sdf[c] = own_missing ? qnan : __fdiv_rn(s1, s2);
                c_std[c] = own_missing ? qnan : __fdiv_rn(s2, s4);
            }
        }
        if (blk + blk_step < blk_end) fetch_neighbours(blk + blk_step, buf ^ 1);   // other buffer: nobody reads it before the barrier at the loop top
        if (!own_ok) continue;                              // uniform per CTA
        __syncthreads();
        const int bx = s_b2[buf][0], by = s_b2[buf][1], bz = s_b2[buf][2];
        for (int sub0 = 0; sub0 < R3; sub0 += MC_THREADS) {
            // ---- 3a: one thread per sub-cube: case index and candidate triangle count; exclusive block scan -> work list of
            //      (sub-cube, triangle, case) items.  Only ~10 % of the sub-cubes of a surface PLIVox cross the surface: generating
            //      the triangles per sub-cube ran 12 predicated edge evaluations per warp for 3-4 active lanes.
            const int sub = sub0 + tid;
            int nt = 0, type = 0;
            if (sub < R3) {
                const int rx = sub / (R * R), ry = (sub / R) % R, rz = sub % R;
                bool bad = false;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float v = c_sdf[((rx + dx8[i]) * R1 + ry + dy8[i]) * R1 + rz + dz8[i]];
                    bad |= !(v == v);
                    if (v < 0.f) type |= 1 << i;
                }
                nt = bad ? 0 : s_ntri[type];
            }
            int incl = nt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += u; }
            if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
            __syncthreads();
            int n_items = 0, wpre = 0;
#pragma unroll
            for (int w = 0; w < MC_THREADS / 32; ++w) { const int t = warp_tot[w]; if (w < (tid >> 5)) wpre += t; n_items += t; }
            for (int t = 0; t < nt; ++t) s_work[wpre + incl - nt + t] = (uint32_t)sub | ((uint32_t)t << 9) | ((uint32_t)type << 12);
            __syncthreads();                                         // work list complete (warp_tot is free again after this barrier)
            // ---- 3b: one thread per candidate triangle: its three edge vertices, the max_std filter, compaction, ONE reservation
            //      per chunk; triangles are staged in shared memory and written out as contiguous float streams (full sectors)
            for (int i0 = 0; i0 < n_items; i0 += MC_THREADS) {
                const int i = i0 + tid;
                bool keep = false;
                float4 vv[3];
                if (i < n_items) {
                    const uint32_t wk = s_work[i];
                    const int sb = wk & 511, t = (wk >> 9) & 7, ty = wk >> 12;
                    const int rx = sb / (R * R), ry = (sb / R) % R, rz = sb % R;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int e = s_tri[ty][3 * t + k];
                        const int p = e < 8 ? e : e - 8, q = e < 4 ? ((e + 1) & 3) : (e < 8 ? 4 + ((e + 1) & 3) : e - 4);      // edge e joins corners p, q
                        const int px = rx + (((p + 1) >> 1) & 1), py = ry + ((p >> 1) & 1), pz = rz + (p >> 2);
                        const int qx = rx + (((q + 1) >> 1) & 1), qy = ry + ((q >> 1) & 1), qz = rz + (q >> 2);
                        const int cp = (px * R1 + py) * R1 + pz, cq = (qx * R1 + qy) * R1 + qz;
                        const float3 p1 = make_float3(__fmaf_rn((float)px, sbs, (float)bx), __fmaf_rn((float)py, sbs, (float)by), __fmaf_rn((float)pz, sbs, (float)bz));
                        const float3 p2 = make_float3(__fmaf_rn((float)qx, sbs, (float)bx), __fmaf_rn((float)qy, sbs, (float)by), __fmaf_rn((float)qz, sbs, (float)bz));
                        vv[k] = edge_vertex(p1, p2, c_std[cp], c_std[cq], c_sdf[cp], c_sdf[cq]);
                    }
                    keep = !(vv[0].w > a.max_std || vv[1].w > a.max_std || vv[2].w > a.max_std);
                }
                const unsigned ballot = __ballot_sync(0xffffffffu, keep);
                if ((tid & 31) == 0) warp_tot[tid >> 5] = __popc(ballot);
                __syncthreads();
                int tot = 0, pre = 0;
#pragma unroll
                for (int w = 0; w < MC_THREADS / 32; ++w) { const int t = warp_tot[w]; if (w < (tid >> 5)) pre += t; tot += t; }
                if (tid == 0 && tot) block_base = atomicAdd(a.count, tot);
                if (keep) {
                    const int local = pre + __popc(ballot & ((1u << (tid & 31)) - 1u));
                    float* o = st_tri + local * 9;
                    o[0] = vv[0].x; o[1] = vv[0].y; o[2] = vv[0].z; o[3] = vv[1].x; o[4] = vv[1].y; o[5] = vv[1].z; o[6] = vv[2].x; o[7] = vv[2].y; o[8] = vv[2].z;
                    float* os = st_std + local * 3;
                    os[0] = vv[0].w; os[1] = vv[1].w; os[2] = vv[2].w;
                }
                __syncthreads();                                     // staging complete, block_base visible
                if (tot) {
                    const int64_t base = block_base;
                    const int64_t room = a.max_tri - base;           // triangles past max_tri are counted, not written
                    const int n_out = room <= 0 ? 0 : (room < tot ? (int)room : tot);
                    float* gt = a.tri + base * 9; float* gs = a.tri_std + base * 3; int64_t* gi = a.tri_id + base;
                    for (int j = tid; j < n_out * 9; j += MC_THREADS) gt[j] = st_tri[j];
                    for (int j = tid; j < n_out * 3; j += MC_THREADS) gs[j] = st_std[j];
                    for (int j = tid; j < n_out; j += MC_THREADS) gi[j] = id;
                }
                __syncthreads();                                     // warp_tot / staging / block_base / (last chunk) the work list are reused
            }
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
        rc = launch_decode_lattice(P, map->latent_vecs, map->latent_stride > 0 ? map->latent_stride : DIF_L, map->shard_world > 1 ? map->row_of_slot : nullptr, block_slots, n_blocks, hr, step, (float)sa, nullptr, nullptr, n_blocks * h3, -1.f,
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
    rc = launch_decode_lattice(P, map->latent_vecs, map->latent_stride > 0 ? map->latent_stride : DIF_L, map->shard_world > 1 ? map->row_of_slot : nullptr, block_slots, n_blocks, lr, step_l, (float)sa, nullptr, nullptr, n_blocks * l3, 1.f,
                               low_sdf, low_std, st);
    if (rc) return rc;
    const int64_t total = n_blocks * h3;
    {
        const int64_t cap = (int64_t)DIF_NUM_SMS * 16;
        const unsigned grid = (unsigned)(n_blocks < cap ? n_blocks : cap);
        switch (lr) {
#define DIF_UP_CASE(L) case L: upsample_select_kernel<L><<<grid, UP_THREADS, 0, st>>>(low_sdf, low_std, n_blocks, cube_sdf, cube_std, list, n_sel); break;
            DIF_UP_CASE(2) DIF_UP_CASE(3) DIF_UP_CASE(4) DIF_UP_CASE(5) DIF_UP_CASE(6) DIF_UP_CASE(7) DIF_UP_CASE(8)
#undef DIF_UP_CASE
        }
    }
    DIF_COUNT_LAUNCH(2);
    rc = launch_decode_lattice(P, map->latent_vecs, map->latent_stride > 0 ? map->latent_stride : DIF_L, map->shard_world > 1 ? map->row_of_slot : nullptr, block_slots, n_blocks, hr, step_h, (float)sa, list, n_sel, total, -1.f, cube_sdf, cube_std, st);
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
#ifndef DIF_MC_CTAS_PER_SM
#define DIF_MC_CTAS_PER_SM 32
#endif
    const int64_t cap = (int64_t)DIF_NUM_SMS * DIF_MC_CTAS_PER_SM;
    prof_begin(DIF_PROF_MC, st);
    const unsigned grid = (unsigned)(n_valid < cap ? n_valid : cap);
    switch (r) {
#define DIF_MC_CASE(R) case R: marching_cubes_kernel<R><<<grid, MC_THREADS, 0, st>>>(a); break;
        DIF_MC_CASE(1) DIF_MC_CASE(2) DIF_MC_CASE(3) DIF_MC_CASE(4) DIF_MC_CASE(5) DIF_MC_CASE(6) DIF_MC_CASE(7) DIF_MC_CASE(8)
#undef DIF_MC_CASE
    }
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
