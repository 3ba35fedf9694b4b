// Tensor-core point-to-implicit ICP linearisation: pose transform -> PLIVox lookup -> decoder FORWARD and BACKWARD (wrt xyz)
// on tcgen05 -> residual / Jacobian / Huber -> 6x6 normal equations, one launch.  Included by decode_tc.cu (same translation
// unit: shares the constant-memory head weights).
//   replaces reference system/tracker.py:174-218 (compute_sdf_Hg) + system/map.py:559-579 (get_sdf) + autograd backward.
//
// Same skeleton as decode_tc_kernel (20 warps: 16 epilogue, MMA issuer, 2 gather producers, two 128-sample tiles in flight,
// weights resident in shared memory, activations chained through TMEM) with 8 GEMM stages per tile instead of 4:
//
//   F0..F3  forward layers (as in decode_tc_kernel); every epilogue thread keeps the ReLU masks of its columns (6 registers)
//   B3      g3 = d r/d pre3 = seed * w4 * relu'(h3)         D = g3 * W3   -> [g2' (96) | d/d latent (29) | d/d xyz skip (3)]
//   B2      g2 = g2' * relu'(h2)                             D = g2 * W2   (K = 96)
//   B1      g1 = D * relu'(h1)                               D = g1 * W1
//   B0      g0 = D * relu'(h0)                               D = g0 * W0   (N = 32; columns 29..31 = d/d xyz direct path)
//
// The backward GEMMs read the SAME shared-memory weight slabs through MN-major descriptors (instruction-descriptor bit 16:
// the slab's 8x16-byte core matrices are valid transposed core matrices; validated by tools/tc_probe_mn.cu), so no transposed
// copy of the weights is needed.  Gradients are split into fp16 hi/lo like activations (3 MMA passes).
#pragma once

namespace dif {
namespace tc {

constexpr uint32_t OFF_AUX = OFF_BAR + 96 + 16;          // per slot: validity byte of the tile's 128 samples
constexpr uint32_t OFF_FRAME = OFF_AUX + 2 * TILE;         // IcpFrame: pose + point count of the launch (host values or the device block)
constexpr uint32_t ICP_SMEM_B = OFF_FRAME + 144;
static_assert(sizeof(IcpFrame) <= 144, "frame block");
static_assert(ICP_SMEM_B <= 232448, "shared memory budget");

__device__ __forceinline__ constexpr uint32_t idesc_f16_bt(int N) { return idesc_f16(N) | (1u << 16); }      // B operand MN-major

// TS-only stage: D (+)= A_tmem * B, three passes hi*hi, lo*hi, hi*lo; w_step16 = descriptor increment per K=16 step
__device__ __forceinline__ void issue_stage(uint32_t idesc, uint32_t w_step16, int ksteps, uint32_t acc, uint32_t a_hi_t, uint32_t a_lo_t,
                                            uint64_t w_hi_d, uint64_t w_lo_d) {
    uint32_t accumulate = 0;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a_t = pass == 1 ? a_lo_t : a_hi_t;
        uint64_t w = pass == 2 ? w_lo_d : w_hi_d;
#pragma unroll 4
        for (int ks = 0; ks < ksteps; ++ks) { mma_ts(acc, a_t + ks * 8, w, idesc, accumulate); accumulate = 1; w += w_step16; }
    }
}

// forward conversion of 16 accumulator columns with ReLU-mask capture
__device__ __forceinline__ uint32_t convert_fwd16(const uint32_t* v, const float* b, uint32_t a_hi, uint32_t a_lo) {
    uint32_t hi[8], lo[8], mask = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float2 bb = *reinterpret_cast<const float2*>(b + 2 * j);
        const float f0 = __uint_as_float(v[2 * j]) + bb.x, f1 = __uint_as_float(v[2 * j + 1]) + bb.y;
        mask |= (f0 > 0.f ? 1u : 0u) << (2 * j);
        mask |= (f1 > 0.f ? 1u : 0u) << (2 * j + 1);
        split_pair(fmaxf(f0, 0.f), fmaxf(f1, 0.f), hi[j], lo[j]);
    }
    tmem_st8(a_hi, hi); tmem_st8(a_lo, lo);
    return mask;
}
// backward conversion of 16 gradient columns: g = relu'(h) ? D : 0
__device__ __forceinline__ void convert_bwd16(const uint32_t* v, uint32_t mask, uint32_t a_hi, uint32_t a_lo) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
        split_pair((mask >> (2 * j)) & 1u ? __uint_as_float(v[2 * j]) : 0.f, (mask >> (2 * j + 1)) & 1u ? __uint_as_float(v[2 * j + 1]) : 0.f, hi[j], lo[j]);
    tmem_st8(a_hi, hi); tmem_st8(a_lo, lo);
}

// producer: one observation per lane -> world point -> PLIVox lookup (map.py:565-575) -> latent row + rel xyz
__device__ __forceinline__ void icp_gather_row(const IcpTcArgs& a, const IcpFrame& fr, int64_t i, float (&x)[32], bool& valid) {
    valid = false;
    int64_t slot = 0;
    float rx = 0.f, ry = 0.f, rz = 0.f;
    if (i < fr.n) {
        const float* op = a.obs + (int64_t)a.obs_stride * i;
        const float ox = __ldg(op), oy = __ldg(op + 1), oz = __ldg(op + 2);
        // cur = (last . delta) @ obs  (tracker.py:181, motion_util.py:322-327) -- same arithmetic as icp_linearize_kernel
        const float wx = fmaf(oz, fr.pose.Rc[2], fmaf(oy, fr.pose.Rc[1], ox * fr.pose.Rc[0])) + fr.pose.tc[0];
        const float wy = fmaf(oz, fr.pose.Rc[5], fmaf(oy, fr.pose.Rc[4], ox * fr.pose.Rc[3])) + fr.pose.tc[1];
        const float wz = fmaf(oz, fr.pose.Rc[8], fmaf(oy, fr.pose.Rc[7], ox * fr.pose.Rc[6])) + fr.pose.tc[2];
        const float3 p = normalize_point(a.m.g, wx, wy, wz);
        const int ix = (int)ceilf(p.x) - 1, iy = (int)ceilf(p.y) - 1, iz = (int)ceilf(p.z) - 1;
        if (p.x == p.x && p.y == p.y && p.z == p.z && in_grid(a.m.g, ix, iy, iz)) {
            const int64_t sl = a.m.indexer[lin_id(a.m.g, ix, iy, iz)];
            if (sl >= 0 && a.m.obs[sl] > a.m.ignore_th) {
                slot = a.m.row_of ? (int64_t)a.m.row_of[sl] : sl;                 // latent ROW (sharded map: -1 = not on this rank)
                valid = slot >= 0;
                if (!valid) slot = 0;
            }
        }
        rx = __fsub_rn(__fsub_rn(p.x, (float)ix), 0.5f); ry = __fsub_rn(__fsub_rn(p.y, (float)iy), 0.5f); rz = __fsub_rn(__fsub_rn(p.z, (float)iz), 0.5f);
    }
    load_latent_row(a.m.latent, slot, a.m.lat_stride, x);
    x[29] = rx; x[30] = ry; x[31] = rz;
}

__global__ void __launch_bounds__(THREADS, 1) icp_tc_kernel(const unsigned char* __restrict__ image, const float* __restrict__ P, IcpTcArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = sbase + OFF_BAR;
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 96);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_stages = a.want_grad ? 8 : 4;
    IcpFrame& fr = *reinterpret_cast<IcpFrame*>(smem + OFF_FRAME);

    if (threadIdx.x == 0) {
        mbar_init(bar0 + 8 * BAR_W, 1);
        mbar_init(bar0 + 8 * BAR_X0, 1); mbar_init(bar0 + 8 * BAR_X1, 1);
        mbar_init(bar0 + 8 * BAR_XF0, 8); mbar_init(bar0 + 8 * BAR_XF1, 8);
        mbar_init(bar0 + 8 * BAR_E0, 8); mbar_init(bar0 + 8 * BAR_E1, 8);
        mbar_init(bar0 + 8 * BAR_ACC0, 1); mbar_init(bar0 + 8 * BAR_ACC1, 1);
        mbar_init(bar0 + 8 * BAR_A0, 8); mbar_init(bar0 + 8 * BAR_A1, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_ptr_s)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_ptr_s, 0);

    // The weight image is constant data: its load is issued BEFORE the programmatic-dependent-launch wait, so it (and the
    // barrier / TMEM set-up above) overlaps with the tail of the previous kernel of the stream.
    if (warp == MMA_WARP && lane == 0) {
        mbar_expect_tx(bar0 + 8 * BAR_W, IMAGE_B);
        constexpr uint32_t CH = 32768;
        for (uint32_t off = 0; off < IMAGE_B; off += CH) bulk_g2s(sbase + off, image + off, (IMAGE_B - off) < CH ? (IMAGE_B - off) : CH, bar0 + 8 * BAR_W);
    }
    pdl_wait(); pdl_launch_dependents();
    if (threadIdx.x == 0) icp_resolve_frame(a.frame, a.pose, a.n, fr);     // (a device-side frame block is only read after the dependency wait)
    __syncthreads();
    const int64_t n_tiles = ((int64_t)fr.n + TILE - 1) / TILE;

    if (warp == MMA_WARP) {
        // ===================================================== MMA issuer (warp-uniform, instructions elected)
        __syncwarp();
        mbar_wait(bar0 + 8 * BAR_W, 0);
        // forward (K-major) and backward (MN-major, LBO = 128 B between 8-row groups, SBO = slab chunk stride) descriptors
        const uint64_t f0h = smem_desc(sbase + OFF_W0, 2048, 128), f0l = smem_desc(sbase + PLANE_B + OFF_W0, 2048, 128);
        const uint64_t f1h = smem_desc(sbase + OFF_W1, 2048, 128), f1l = smem_desc(sbase + PLANE_B + OFF_W1, 2048, 128);
        const uint64_t f2h = smem_desc(sbase + OFF_W2, 1536, 128), f2l = smem_desc(sbase + PLANE_B + OFF_W2, 1536, 128);
        const uint64_t f3h = smem_desc(sbase + OFF_W3, 2048, 128), f3l = smem_desc(sbase + PLANE_B + OFF_W3, 2048, 128);
        const uint64_t b0h = smem_desc(sbase + OFF_W0, 128, 2048), b0l = smem_desc(sbase + PLANE_B + OFF_W0, 128, 2048);
        const uint64_t b1h = smem_desc(sbase + OFF_W1, 128, 2048), b1l = smem_desc(sbase + PLANE_B + OFF_W1, 128, 2048);
        const uint64_t b2h = smem_desc(sbase + OFF_W2, 128, 1536), b2l = smem_desc(sbase + PLANE_B + OFF_W2, 128, 1536);
        const uint64_t b3h = smem_desc(sbase + OFF_W3, 128, 2048), b3l = smem_desc(sbase + PLANE_B + OFF_W3, 128, 2048);
        uint32_t ph_x = 0, ph_a = 0, ph_e = 0;
        for (int64_t it = 0;; ++it) {
            const int64_t t0 = blockIdx.x + (int64_t)gridDim.x * (2 * it), t1 = t0 + gridDim.x;
            const bool live[2] = {t0 < n_tiles, t1 < n_tiles};
            if (!live[0]) break;
#pragma unroll 1
            for (int stage = 0; stage < n_stages; ++stage) {
#pragma unroll 1
                for (int s = 0; s < 2; ++s) {
                    if (!live[s]) continue;
                    const uint32_t acc = tmem + s * 256, a_hi = acc + 128, a_lo = acc + 192;
                    if (stage == 0) {
                        mbar_wait_tight(bar0 + 8 * (BAR_X0 + s), (ph_x >> s) & 1); ph_x ^= 1u << s;
                        if (it > 0) { mbar_wait_tight(bar0 + 8 * (BAR_E0 + s), (ph_e >> s) & 1); ph_e ^= 1u << s; }
                    } else {
                        mbar_wait_tight(bar0 + 8 * (BAR_A0 + s), (ph_a >> s) & 1); ph_a ^= 1u << s;
                    }
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t xh = smem_desc(sbase + OFF_X + s * 2 * X_PLANE_B, X_CHUNK_B, 128);
                        const uint64_t xl = smem_desc(sbase + OFF_X + s * 2 * X_PLANE_B + X_PLANE_B, X_CHUNK_B, 128);
                        switch (stage) {
                            case 0: issue_layer(idesc_f16(128), 128 * 16, 0, 2, acc, 0, 0, xh, xl, f0h, f0l); break;
                            case 1: issue_stage(idesc_f16(128), (2 * 2048) >> 4, 8, acc, a_hi, a_lo, f1h, f1l); break;
                            case 2: issue_stage(idesc_f16(96), (2 * 1536) >> 4, 8, acc, a_hi, a_lo, f2h, f2l); break;
                            case 3: issue_stage(idesc_f16(128), (2 * 2048) >> 4, 8, acc, a_hi, a_lo, f3h, f3l); break;
                            case 4: issue_stage(idesc_f16_bt(128), 256 >> 4, 8, acc, a_hi, a_lo, b3h, b3l); break;      // g3 * W3
                            case 5: issue_stage(idesc_f16_bt(128), 256 >> 4, 6, acc, a_hi, a_lo, b2h, b2l); break;      // g2 * W2 (K = 96)
                            case 6: issue_stage(idesc_f16_bt(128), 256 >> 4, 8, acc, a_hi, a_lo, b1h, b1l); break;      // g1 * W1
                            default: issue_stage(idesc_f16_bt(32), 256 >> 4, 8, acc, a_hi, a_lo, b0h, b0l); break;      // g0 * W0 (N = 32)
                        }
                        mma_commit(bar0 + 8 * (BAR_ACC0 + s));
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= PRODUCER_WARP0) {
        // ===================================================== gather producer of slot s (one observation per lane)
        const int s = warp - PRODUCER_WARP0;
        if (s < 2) {
            unsigned char* x_hi_p = smem + OFF_X + s * 2 * X_PLANE_B;
            unsigned char* aux = smem + OFF_AUX + s * TILE;
            float xa[32], xb[32];
            bool va = false, vb = false;
            uint32_t ph_xf = 0;
            int64_t tile = blockIdx.x + (int64_t)gridDim.x * s;
            if (tile < n_tiles) icp_gather_row(a, fr, tile * TILE + lane, xa, va);
            for (int64_t it = 0; tile < n_tiles; ++it) {
                if (it > 0) { mbar_wait(bar0 + 8 * (BAR_XF0 + s), ph_xf); ph_xf ^= 1; }
                icp_gather_row(a, fr, tile * TILE + 32 + lane, xb, vb);
                gather_store_row(xa, va, lane, x_hi_p); aux[lane] = va;
                icp_gather_row(a, fr, tile * TILE + 64 + lane, xa, va);
                gather_store_row(xb, vb, 32 + lane, x_hi_p); aux[32 + lane] = vb;
                icp_gather_row(a, fr, tile * TILE + 96 + lane, xb, vb);
                gather_store_row(xa, va, 64 + lane, x_hi_p); aux[64 + lane] = va;
                const int64_t next = tile + 2 * (int64_t)gridDim.x;
                if (next < n_tiles) icp_gather_row(a, fr, next * TILE + lane, xa, va);
                gather_store_row(xb, vb, 96 + lane, x_hi_p); aux[96 + lane] = vb;
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar0 + 8 * (BAR_X0 + s));
                tile = next;
            }
        }
    } else {
        // ===================================================== epilogue warps of slot s
        const int s = warp >> 3;
        const int quad = warp & 3, half = (warp >> 2) & 1;
        const int row = quad * 32 + lane;
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        const uint32_t acc = tmem + s * 256 + lane_base, a_hi = acc + 128, a_lo = acc + 192;
        const float* bias = reinterpret_cast<const float*>(smem + OFF_BIAS);
        const unsigned char* aux = smem + OFF_AUX + s * TILE;
        const int head_slot = *reinterpret_cast<const int*>(image + IMAGE_B);
        const float* w4c = c_head_w[head_slot][0];
        const float* wuc = c_head_w[head_slot][1];
        const float b4 = __ldg(P + DecW::b4), bu = __ldg(P + DecW::bu);
        uint32_t ph_acc = 0;
        double lane_sum = 0.0;                          // lane j of a half-1 warp: running sum of value j over this warp's tiles
        mbar_wait(bar0 + 8 * BAR_W, 0);
        for (int64_t it = 0;; ++it) {
            const int64_t tile = blockIdx.x + (int64_t)gridDim.x * (2 * it + s);
            if (tile >= n_tiles) break;
            uint64_t m0 = 0, m1 = 0, m2 = 0;          // ReLU masks of this thread's columns of h0, h1, h2
            bool valid = false;
            // ---- forward hidden layers 0..2
#pragma unroll 1
            for (int layer = 0; layer < 3; ++layer) {
                const int hw = layer == 2 ? 48 : 64;
                const int c_base = half * hw;
                const float* b = bias + layer * 128 + c_base;
                mbar_wait(bar0 + 8 * (BAR_ACC0 + s), ph_acc); ph_acc ^= 1;
                tc_fence_after();
                if (layer == 0) valid = aux[row] != 0;
                if (layer == 2) {      // park the inputs for the skip connection (see decode_tc_kernel) and release the x tile
                    const unsigned char* xp = smem + OFF_X + s * 2 * X_PLANE_B + half * X_PLANE_B + row * 16;
                    uint32_t xv[16];
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const uint4 q = *reinterpret_cast<const uint4*>(xp + cc * X_CHUNK_B);
                        xv[4 * cc] = q.x; xv[4 * cc + 1] = q.y; xv[4 * cc + 2] = q.z; xv[4 * cc + 3] = q.w;
                    }
                    tmem_st16((half ? a_lo : a_hi) + 48, xv);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar0 + 8 * (BAR_XF0 + s));
                }
                uint64_t mm = 0;
#pragma unroll 1
                for (int c = 0; c < hw; c += 32) {
                    uint32_t v0[16], v1[16];
                    const bool two = c + 16 < hw;
                    tmem_ld16_nowait(acc + c_base + c, v0);
                    if (two) tmem_ld16_nowait(acc + c_base + c + 16, v1);
                    tmem_ld_wait();
                    mm |= (uint64_t)convert_fwd16(v0, b + c, a_hi + (c_base + c) / 2, a_lo + (c_base + c) / 2) << c;
                    if (two) mm |= (uint64_t)convert_fwd16(v1, b + c + 16, a_hi + (c_base + c) / 2 + 8, a_lo + (c_base + c) / 2 + 8) << (c + 16);
                }
                if (layer == 0) m0 = mm; else if (layer == 1) m1 = mm; else m2 = mm;
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar0 + 8 * (BAR_A0 + s));
            }
            // ---- layer 3: both heads by every thread (r and the backward seed need sdf AND std of the row)
            mbar_wait(bar0 + 8 * (BAR_ACC0 + s), ph_acc); ph_acc ^= 1;
            tc_fence_after();
            float p_sdf = 0.f, p_std = 0.f;
#pragma unroll 1
            for (int c0 = 0; c0 < 128; c0 += 32) {
                uint32_t v0[16], v1[16];
                tmem_ld16_nowait(acc + c0, v0);
                tmem_ld16_nowait(acc + c0 + 16, v1);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float h0 = fmaxf(__uint_as_float(v0[j]) + bias[352 + c0 + j], 0.f), h1 = fmaxf(__uint_as_float(v1[j]) + bias[352 + c0 + 16 + j], 0.f);
                    p_sdf = fmaf(w4c[c0 + j], h0, p_sdf); p_sdf = fmaf(w4c[c0 + 16 + j], h1, p_sdf);
                    p_std = fmaf(wuc[c0 + j], h0, p_std); p_std = fmaf(wuc[c0 + 16 + j], h1, p_std);
                }
            }
            const float sdf = tanhf(p_sdf + b4);
            const float sd = 0.05f + 0.5f * softplus_ref(p_std + bu);
            const float r = sdf / sd;                                  // tracker.py:186
            const float seed = (1.f - sdf * sdf) / sd;                 // d r / d pre_sdf (std detached)
            float gxs0 = 0.f, gxs1 = 0.f, gxs2 = 0.f;                  // skip-path d r / d xyz (half-1 threads)
            if (a.want_grad) {
                // g3 for this thread's 64 columns: re-read them, relu'(h3) * seed * w4 -> A operand
                const int c_base = half * 64;
#pragma unroll 1
                for (int c = 0; c < 64; c += 32) {
                    uint32_t v0[16], v1[16];
                    tmem_ld16_nowait(acc + c_base + c, v0);
                    tmem_ld16_nowait(acc + c_base + c + 16, v1);
                    tmem_ld_wait();
                    uint32_t hi0[8], lo0[8], hi1[8], lo1[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int k0 = c_base + c + 2 * j, k1 = k0 + 16;
                        const float ga = __uint_as_float(v0[2 * j]) + bias[352 + k0] > 0.f ? seed * w4c[k0] : 0.f;
                        const float gb = __uint_as_float(v0[2 * j + 1]) + bias[352 + k0 + 1] > 0.f ? seed * w4c[k0 + 1] : 0.f;
                        const float gc = __uint_as_float(v1[2 * j]) + bias[352 + k1] > 0.f ? seed * w4c[k1] : 0.f;
                        const float gd = __uint_as_float(v1[2 * j + 1]) + bias[352 + k1 + 1] > 0.f ? seed * w4c[k1 + 1] : 0.f;
                        split_pair(ga, gb, hi0[j], lo0[j]); split_pair(gc, gd, hi1[j], lo1[j]);
                    }
                    tmem_st8(a_hi + (c_base + c) / 2, hi0); tmem_st8(a_lo + (c_base + c) / 2, lo0);
                    tmem_st8(a_hi + (c_base + c) / 2 + 8, hi1); tmem_st8(a_lo + (c_base + c) / 2 + 8, lo1);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar0 + 8 * (BAR_A0 + s));
                // ---- backward stages B3, B2, B1: D -> mask -> A operand of the next stage
#pragma unroll 1
                for (int st = 0; st < 3; ++st) {
                    const int hw = st == 0 ? 48 : 64;                   // B3 feeds the 96-wide layer 2
                    const int cb = half * hw;
                    const uint64_t mm = st == 0 ? m2 : (st == 1 ? m1 : m0);
                    mbar_wait(bar0 + 8 * (BAR_ACC0 + s), ph_acc); ph_acc ^= 1;
                    tc_fence_after();
                    if (st == 0 && half == 1) {                         // columns 125..127 of g3 * W3: skip-path gradient wrt xyz
                        uint32_t t4[4];
                        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(t4[0]), "=r"(t4[1]), "=r"(t4[2]), "=r"(t4[3]) : "r"(acc + 124));
                        tmem_ld_wait();
                        gxs0 = __uint_as_float(t4[1]); gxs1 = __uint_as_float(t4[2]); gxs2 = __uint_as_float(t4[3]);
                    }
#pragma unroll 1
                    for (int c = 0; c < hw; c += 32) {
                        uint32_t v0[16], v1[16];
                        const bool two = c + 16 < hw;
                        tmem_ld16_nowait(acc + cb + c, v0);
                        if (two) tmem_ld16_nowait(acc + cb + c + 16, v1);
                        tmem_ld_wait();
                        convert_bwd16(v0, (uint32_t)(mm >> c) & 0xFFFFu, a_hi + (cb + c) / 2, a_lo + (cb + c) / 2);
                        if (two) convert_bwd16(v1, (uint32_t)(mm >> (c + 16)) & 0xFFFFu, a_hi + (cb + c) / 2 + 8, a_lo + (cb + c) / 2 + 8);
                    }
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar0 + 8 * (BAR_A0 + s));
                }
                // ---- B0 result: columns 29..31 = direct-path d r / d xyz
                mbar_wait(bar0 + 8 * (BAR_ACC0 + s), ph_acc); ph_acc ^= 1;
                tc_fence_after();
                if (half == 1) {
                    uint32_t t4[4];
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(t4[0]), "=r"(t4[1]), "=r"(t4[2]), "=r"(t4[3]) : "r"(acc + 28));
                    tmem_ld_wait();
                    gxs0 += __uint_as_float(t4[1]); gxs1 += __uint_as_float(t4[2]); gxs2 += __uint_as_float(t4[3]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + 8 * (BAR_E0 + s));        // this tile's accumulators are consumed
            // ---- residual / Jacobian / Huber / normal equations (half-1 warps own the rows), tracker.py:196-216
            if (half == 1) {
                float v[29];
#pragma unroll
                for (int j = 0; j < 29; ++j) v[j] = 0.f;
                if (valid) {
                    float w = 1.f;
                    if (a.huber_k > 0.f) { const float ar = fabsf(r); if (ar > a.huber_k) w = a.huber_k / ar; }
                    else if (a.huber_k < 0.f) { const float q = r / -a.huber_k, t1 = 1.f - q * q; w = fabsf(r) <= -a.huber_k ? t1 * t1 : 0.f; }      // Tukey (tracker.py:66-69)
                    v[27] = r * (r * w);
                    v[28] = 1.f;
                    if (a.want_grad) {
                        const int64_t i = tile * TILE + row;
                        const float* op = a.obs + (int64_t)a.obs_stride * i;
                        const float ox = __ldg(op), oy = __ldg(op + 1), oz = __ldg(op + 2);
                        const Pose& ps = fr.pose;
                        const float qx = fmaf(oz, ps.Rd[2], fmaf(oy, ps.Rd[1], ox * ps.Rd[0])) + ps.td[0];
                        const float qy = fmaf(oz, ps.Rd[5], fmaf(oy, ps.Rd[4], ox * ps.Rd[3])) + ps.td[1];
                        const float qz = fmaf(oz, ps.Rd[8], fmaf(oy, ps.Rd[7], ox * ps.Rd[6])) + ps.td[2];
                        const float gx = gxs0 / a.m.g.vs, gy = gxs1 / a.m.g.vs, gz = gxs2 / a.m.g.vs;
                        float J[6];
                        J[0] = gx * ps.Rl[0] + gy * ps.Rl[1] + gz * ps.Rl[2];
                        J[1] = gx * ps.Rl[3] + gy * ps.Rl[4] + gz * ps.Rl[5];
                        J[2] = gx * ps.Rl[6] + gy * ps.Rl[7] + gz * ps.Rl[8];
                        J[3] = qy * J[2] - qz * J[1];
                        J[4] = qz * J[0] - qx * J[2];
                        J[5] = qx * J[1] - qy * J[0];
                        int k = 0;
#pragma unroll
                        for (int p = 0; p < 6; ++p)
#pragma unroll
                            for (int q = p; q < 6; ++q) v[k++] = (w * J[p]) * J[q];
#pragma unroll
                        for (int p = 0; p < 6; ++p) v[21 + p] = J[p] * (r * w);
                    }
                }
#pragma unroll
                for (int j = 0; j < 29; ++j) {
                    const float t = warp_sum(v[j]);
                    if (lane == j) lane_sum += (double)t;
                }
            }
        }
        // no tile of this slot is left: its x tile is dead and becomes the CTA's reduction buffer [slot][quad][32] (fp64)
        if (half == 1) reinterpret_cast<double*>(smem + OFF_X + s * 2 * X_PLANE_B)[quad * 32 + lane] = lane_sum;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(512));
    // ---- deterministic reduction (no atomics on the sums): fixed-order sum of the CTA's 8 warps -> partials[cta][32];
    //      the last CTA to finish adds the CTAs' partials in a fixed order, scales by 1/M and expands the symmetric H
    __shared__ bool is_last;
    if (warp == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += reinterpret_cast<const double*>(smem + OFF_X + (w >> 2) * 2 * X_PLANE_B)[(w & 3) * 32 + lane];
        a.partials[(size_t)blockIdx.x * 32 + lane] = t;
        __threadfence();
        __syncwarp();
        if (lane == 0) is_last = atomicAdd(a.done_counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double* red = reinterpret_cast<double*>(smem + OFF_X);                 // [16 slices][32]
    if (threadIdx.x < 512) {
        double t = 0.0;
        for (unsigned c = threadIdx.x >> 5; c < gridDim.x; c += 16) t += __ldcg(a.partials + (size_t)c * 32 + lane);
        red[threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < 16; ++w) tot += red[w * 32 + threadIdx.x];
        const double M = __shfl_sync(0xffffffffu, tot, 28);
        const double scale = M > 0.0 ? 1.0 / M : 0.0;
        if (threadIdx.x < 21) {
            int p = 0, rem = threadIdx.x;
            while (rem >= 6 - p) { rem -= 6 - p; ++p; }
            const int q = p + rem;
            a.out[p * 6 + q] = tot * scale; a.out[q * 6 + p] = tot * scale;
        } else if (threadIdx.x < 27) a.out[36 + threadIdx.x - 21] = tot * scale;
        else if (threadIdx.x == 27) a.out[42] = tot * scale;
        else if (threadIdx.x == 28) a.out[43] = M;
        if (threadIdx.x == 0) *a.done_counter = 0u;       // leave the counter clean for the next call (zero-filled once by the caller)
    }
}

}  // namespace tc

}  // namespace dif
#include "icp_tc2.cuh"
namespace dif {

int launch_icp_tc(const void* decoder_prepared, const IcpTcArgs& a, cudaStream_t st) {
    const float* P = (const float*)decoder_prepared;
    const unsigned char* image = (const unsigned char*)decoder_prepared + (size_t)DecW::FP32_END * sizeof(float);
    const int64_t n_tiles = ((int64_t)a.n + tc::TILE - 1) / tc::TILE;
    const int grid = (int)(n_tiles < DIF_NUM_SMS ? (n_tiles > 0 ? n_tiles : 1) : DIF_NUM_SMS);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(tc::icp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::ICP_SMEM_B);
        cudaFuncSetAttribute(tc::icp_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::I2_SMEM_B);
        attr_set = true;
    }
    const char* v = getenv("DIF_ICP_V");                     // "1": the first pipeline (separate accumulator / operand regions), kept for A/B timing
    prof_begin(DIF_PROF_ICP, st);
    if (v && v[0] == '1') launch_pdl(tc::icp_tc_kernel, grid, tc::THREADS, tc::ICP_SMEM_B, st, image, P, a);
    else launch_pdl(tc::icp_tc2_kernel, grid, tc::I2_THREADS, tc::I2_SMEM_B, st, image, P, a);
    prof_end(DIF_PROF_ICP, st);
    DIF_COUNT_LAUNCH(1);
    return check_launch("icp_tc_kernel");
}

}  // namespace dif
