// Tensor-core point-to-implicit ICP linearisation, second pipeline.  Included by decode_tc.cu after icp_tc.cuh (same weight
// image, same argument block, same result contract as icp_tc_kernel).
//   replaces reference system/tracker.py:174-218 (compute_sdf_Hg) + system/map.py:559-579 (get_sdf) + autograd backward.
//
// What one launch is: a frame is ~34 k observations = ~270 tiles of 128 over 148 CTAs, i.e. ONE tile per slot and eight
// dependent GEMM stages per tile (F0..F3 forward, B3..B0 backward wrt xyz).  Nothing reaches a steady state, so the kernel
// is the sum of its latencies; this version removes them one by one (event trace of the first version: tools/icp_trace.py):
//
//   * in-place conversion + chunk hand-off (decode_tc2.cuh): every stage accumulates into the OTHER 128-column region of the
//     slot and its epilogue rewrites the accumulator in place as the next stage's A operand, 32 columns (= 2 K steps) at a
//     time; the next stage's MMAs are issued behind each pair of groups instead of behind the whole epilogue.
//         F0: x (smem) -> R0    F1: R0 -> R1    F2: R1 -> R0[0,96)   F3: [h2 | x parked in R0[96,128)] -> R1
//         B3: g3 (R1) * W3 -> R0 = [g2' (96) | d/d latent (29) | d/d xyz skip (3)]
//         B2: g2 (R0[0,96)) * W2 -> R1      B1: g1 (R1) * W1 -> R0      B0: g0 (R0) * W0[:, 16:32] -> R1[0,16) (cols 13..15 = d/d xyz)
//   * the weight image arrives in four groups with their own barriers (bias + W0, W1, W2, W3): F0 starts after 18 KB, the
//     other 180 KB stream in behind the gather and the first stages;
//   * the FIRST tile of a slot is gathered by the slot's 8 epilogue warps (idle until F0 completes): all 128 lookup chains
//     (point -> indexer -> obs count -> latent row) are in flight at once instead of four 32-row passes through one producer
//     warp; the producer warps start with the slot's second tile;
//   * the two heads are split over the two column halves (each thread: 64 columns, both dot products) and the partial sums
//     are exchanged through four dead TMEM columns of R0; g3 needs no second read of the accumulator (ReLU mask kept in
//     registers);
//   * the 29 per-row values are reduced with a transposing butterfly (31 shuffles per warp instead of 145).
//
// Barriers per slot: X / XF / E, ACCa (stages into R0) / ACCb (stages into R1), G0..G3 per 32-column epilogue group, consumed
// in a fixed order so every accumulation order - and therefore every output bit - is reproducible.
#pragma once

namespace dif {
namespace tc {

enum { I2_W = 0, I2_X = 4, I2_XF = 6, I2_E = 8, I2_ACCA = 10, I2_ACCB = 12, I2_G = 14, I2_NBAR = 22 };      // per-slot barriers: index + slot; G: 14 + 4*slot + g
constexpr uint32_t I2_OFF_BAR = OFF_X + 4 * X_PLANE_B;
constexpr uint32_t I2_OFF_TPTR = I2_OFF_BAR + 8 * I2_NBAR;
constexpr uint32_t I2_OFF_AUX = I2_OFF_TPTR + 16;             // per slot: validity byte of the tile's 128 samples (tiles gathered by a producer warp)
constexpr uint32_t I2_OFF_FRAME = I2_OFF_AUX + 2 * TILE;        // IcpFrame: pose + point count of the launch
constexpr uint32_t I2_SMEM_B = I2_OFF_FRAME + 144;
static_assert(sizeof(IcpFrame) <= 140 && I2_OFF_FRAME % 16 == 0, "frame block (+ the launch epoch at byte 140)");
static_assert(I2_SMEM_B + 64 <= 232448, "shared memory budget");

constexpr int I2_THREADS = 19 * 32;          // 16 epilogue warps, issuer, 2 producers: 104 registers per thread instead of 96 at 20 warps
#ifndef DIF_ICP2_EPI_SLEEP
#define DIF_ICP2_EPI_SLEEP 32          // ns between probes of an epilogue warp waiting for an accumulator (0: bare try_wait loop)
#endif
#ifndef DIF_ICP2_PREFETCH
#define DIF_ICP2_PREFETCH 1            // first DRAM hop of the first tile issued before the frame block is resolved
#endif
#ifndef DIF_ICP2_GRP_UNROLL
#define DIF_ICP2_GRP_UNROLL 1          // 1: two unrolled copies of the per-group conversion code
#endif
#if DIF_ICP2_GRP_UNROLL
#define I2_GRP_PRAGMA _Pragma("unroll")
#else
#define I2_GRP_PRAGMA _Pragma("unroll 1")
#endif
#ifndef DIF_ICP2_TAIL
#define DIF_ICP2_TAIL 2                // 0: fence.sc + atomic ticket, last CTA reduces; 1: acq_rel ticket; 2: epoch-tagged rows, polled by one reducer CTA
#endif
#ifndef DIF_ICP2_BWD_PASSES
#define DIF_ICP2_BWD_PASSES 3          // 2: backward A operands (gradients) rounded to ONE fp16 (to nearest), passes g*W_hi + g*W_lo
#endif

#ifdef DIF_TC_TRACE
#define I2_TRACE(ev) do { if (timing && lane == 0 && blockIdx.x == 0 && tn < 1024) \
    g_tc_timing[warp * 1024 + tn++] = ((unsigned long long)(ev) << 48) | ((unsigned long long)clock64() & 0xFFFFFFFFFFFFull); } while (0)
// per-CTA wall-clock stamps (every CTA's warp 0): [20*1024 + 300 + 8*cta + k]
#define I2_GT(k) do { if (timing && warp == 0 && lane == 0) { unsigned long long gt_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_)); \
    g_tc_timing[20 * 1024 + 300 + 8 * blockIdx.x + (k)] = gt_; } } while (0)
#else
#define I2_TRACE(ev) do { } while (0)
#define I2_GT(k) do { } while (0)
#endif

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory"); }

// one hand-off group of a stage: K steps KS0 (and KS1 if >= 0) x 3 passes (hi*hi, lo*hi, hi*lo); A chunk ks at a_base + 16 ks (hi) / + 8 (lo);
// STEP = weight-descriptor increment per K = 16 step (K-major forward slabs: N*32 >> 4; MN-major backward view: 256 >> 4)
template <uint32_t IDESC, uint32_t STEP, int KS0, int KS1, bool FIRST, bool A_LO = true>
__device__ __forceinline__ void issue_group_g(uint32_t acc, uint32_t a_base, uint64_t w_hi_d, uint64_t w_lo_d) {
    mma_ts(acc, a_base + 16 * KS0, w_hi_d + (uint64_t)(KS0 * STEP), IDESC, FIRST ? 0u : 1u);
    if constexpr (KS1 >= 0) mma_ts(acc, a_base + 16 * KS1, w_hi_d + (uint64_t)(KS1 * STEP), IDESC, 1u);
    if constexpr (A_LO) {
        mma_ts(acc, a_base + 16 * KS0 + 8, w_hi_d + (uint64_t)(KS0 * STEP), IDESC, 1u);
        if constexpr (KS1 >= 0) mma_ts(acc, a_base + 16 * KS1 + 8, w_hi_d + (uint64_t)(KS1 * STEP), IDESC, 1u);
    }
    mma_ts(acc, a_base + 16 * KS0, w_lo_d + (uint64_t)(KS0 * STEP), IDESC, 1u);
    if constexpr (KS1 >= 0) mma_ts(acc, a_base + 16 * KS1, w_lo_d + (uint64_t)(KS1 * STEP), IDESC, 1u);
}

// Stage 1..7 of one slot: groups (half 0, it 0) + (half 1, it 0) behind one wait, then (half 0, it 1) + (half 1, it 1) and the commit.
//   K-step sets per group:  A = 128 columns: {0,1} {4,5} | {2,3} {6,7};   A = 96 columns + 32 parked (F3): {0,1} {3,4} | {2,6} {5,7};
//                           A = 96 columns (B2): {0,1} {3,4} | {2} {5}
template <int STAGE>
__device__ __forceinline__ void icp2_issue_stage(uint32_t R0, uint32_t R1, uint64_t wh, uint64_t wl, uint32_t bg, uint32_t ph, uint32_t commit_bar) {
    constexpr bool BWD = STAGE >= 4;
    constexpr int N = STAGE == 2 ? 96 : (STAGE == 7 ? 16 : 128);
    constexpr uint32_t ID = BWD ? idesc_f16_bt(N) : idesc_f16(N);
    constexpr uint32_t ST = BWD ? (256u >> 4) : ((2u * N * 16u) >> 4);
    constexpr bool AL = !(BWD && DIF_ICP2_BWD_PASSES == 2);
    const uint32_t acc = (STAGE & 1) ? R1 : R0, ab = (STAGE & 1) ? R0 : R1;
    mbar_wait_spin(bg, ph);
    mbar_wait_spin(bg + 8, ph);
    tc_fence_after();
    if (elect_one()) {
        issue_group_g<ID, ST, 0, 1, true, AL>(acc, ab, wh, wl);
        if constexpr (STAGE == 3 || STAGE == 5) issue_group_g<ID, ST, 3, 4, false, AL>(acc, ab, wh, wl);
        else issue_group_g<ID, ST, 4, 5, false, AL>(acc, ab, wh, wl);
    }
    __syncwarp();
    mbar_wait_spin(bg + 16, ph);
    mbar_wait_spin(bg + 24, ph);
    tc_fence_after();
    if (elect_one()) {
        if constexpr (STAGE == 3) { issue_group_g<ID, ST, 2, 6, false, AL>(acc, ab, wh, wl); issue_group_g<ID, ST, 5, 7, false, AL>(acc, ab, wh, wl); }
        else if constexpr (STAGE == 5) { issue_group_g<ID, ST, 2, -1, false, AL>(acc, ab, wh, wl); issue_group_g<ID, ST, 5, -1, false, AL>(acc, ab, wh, wl); }
        else { issue_group_g<ID, ST, 2, 3, false, AL>(acc, ab, wh, wl); issue_group_g<ID, ST, 6, 7, false, AL>(acc, ab, wh, wl); }
        mma_commit(commit_bar);
    }
    __syncwarp();
}

// forward: 16 accumulator columns -> +bias, ReLU, hi/lo split -> the same 16 columns (one K = 16 A operand).  Returns the SIGN bits
// of the 16 pre-activations, column j at bit 15 - j (one funnel shift per column; relu'(h) = !sign, i.e. h = +0 counts as active -
// a measure-zero difference from "h > 0" that costs 1.5 instructions per column less than a compare + select + or)
__device__ __forceinline__ uint32_t cvt_fwd16(const uint32_t* v, const float* b, uint32_t col_addr) {
    uint32_t o[16], sg = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 bb = *reinterpret_cast<const float4*>(b + 4 * j);
        const float f0 = __uint_as_float(v[4 * j]) + bb.x, f1 = __uint_as_float(v[4 * j + 1]) + bb.y;
        const float f2 = __uint_as_float(v[4 * j + 2]) + bb.z, f3 = __uint_as_float(v[4 * j + 3]) + bb.w;
        sg = __funnelshift_l(__float_as_uint(f0), sg, 1);
        sg = __funnelshift_l(__float_as_uint(f1), sg, 1);
        sg = __funnelshift_l(__float_as_uint(f2), sg, 1);
        sg = __funnelshift_l(__float_as_uint(f3), sg, 1);
        relu_split_pair(f0, f1, o[2 * j], o[8 + 2 * j]);
        relu_split_pair(f2, f3, o[2 * j + 1], o[8 + 2 * j + 1]);
    }
    tmem_st16(col_addr, o);
    return sg & 0xFFFFu;
}
// backward: 16 gradient columns -> relu'(h) ? D : 0 -> hi/lo split in place (sg: sign bits as returned by cvt_fwd16)
__device__ __forceinline__ uint32_t pack_rn(float a, float b) { const __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<const uint32_t*>(&h); }
__device__ __forceinline__ void cvt_bwd16(const uint32_t* v, uint32_t sg, uint32_t col_addr) {
#if DIF_ICP2_BWD_PASSES == 2
    uint32_t o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
        o[j] = pack_rn((sg >> (15 - 2 * j)) & 1u ? 0.f : __uint_as_float(v[2 * j]), (sg >> (14 - 2 * j)) & 1u ? 0.f : __uint_as_float(v[2 * j + 1]));
    tmem_st8(col_addr, o);
#else
    uint32_t o[16];
#pragma unroll
    for (int j = 0; j < 8; ++j)
        split_pair((sg >> (15 - 2 * j)) & 1u ? 0.f : __uint_as_float(v[2 * j]), (sg >> (14 - 2 * j)) & 1u ? 0.f : __uint_as_float(v[2 * j + 1]), o[j], o[8 + j]);
    tmem_st16(col_addr, o);
#endif
}

// F3 result, 16 columns: h3 = relu(acc + b3) folded into both head dot products (weights read from the constant bank as vectors)
__device__ __forceinline__ void head_fma16(const uint32_t* v, const float* b3, const float* w4, const float* wu, float& ps, float& pu) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 bb = *reinterpret_cast<const float4*>(b3 + 4 * j);
        const float4 wa = *reinterpret_cast<const float4*>(w4 + 4 * j), wb = *reinterpret_cast<const float4*>(wu + 4 * j);
        const float r0 = fmaxf(__uint_as_float(v[4 * j]) + bb.x, 0.f), r1 = fmaxf(__uint_as_float(v[4 * j + 1]) + bb.y, 0.f);
        const float r2 = fmaxf(__uint_as_float(v[4 * j + 2]) + bb.z, 0.f), r3 = fmaxf(__uint_as_float(v[4 * j + 3]) + bb.w, 0.f);
        ps = fmaf(wa.x, r0, ps); pu = fmaf(wb.x, r0, pu);
        ps = fmaf(wa.y, r1, ps); pu = fmaf(wb.y, r1, pu);
        ps = fmaf(wa.z, r2, ps); pu = fmaf(wb.z, r2, pu);
        ps = fmaf(wa.w, r3, ps); pu = fmaf(wb.w, r3, pu);
    }
}
// g3' for 16 columns of the F3 accumulator, in place: pre-split constants (g3w: 8 hi pair words; + 64: lo; + 128: single-rounded)
// AND-ed with the pair's activity mask; PRMT with sign replication turns the sign bytes of two pre-activations into the mask
__device__ __forceinline__ uint32_t pair_neg_mask(float f0, float f1) {
    uint32_t m;
    asm("prmt.b32 %0, %1, %2, 0xFFBB;" : "=r"(m) : "r"(__float_as_uint(f0)), "r"(__float_as_uint(f1)));      // bytes 0,1 <- sign(f0), bytes 2,3 <- sign(f1)
    return m;
}
__device__ __forceinline__ void g3_signs16(const uint32_t* v, const float* b3, const uint32_t* g3w, uint32_t col_addr) {
#if DIF_ICP2_BWD_PASSES == 2
    uint32_t o[8];
#else
    uint32_t o[16];
#endif
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 bb = *reinterpret_cast<const float4*>(b3 + 4 * j);
        const uint32_t n0 = pair_neg_mask(__uint_as_float(v[4 * j]) + bb.x, __uint_as_float(v[4 * j + 1]) + bb.y);
        const uint32_t n1 = pair_neg_mask(__uint_as_float(v[4 * j + 2]) + bb.z, __uint_as_float(v[4 * j + 3]) + bb.w);
#if DIF_ICP2_BWD_PASSES == 2
        const uint2 h = *reinterpret_cast<const uint2*>(g3w + 128 + 2 * j);
        o[2 * j] = h.x & ~n0; o[2 * j + 1] = h.y & ~n1;
#else
        const uint2 h = *reinterpret_cast<const uint2*>(g3w + 2 * j), l = *reinterpret_cast<const uint2*>(g3w + 64 + 2 * j);
        o[2 * j] = h.x & ~n0; o[2 * j + 1] = h.y & ~n1;
        o[8 + 2 * j] = l.x & ~n0; o[8 + 2 * j + 1] = l.y & ~n1;
#endif
    }
#if DIF_ICP2_BWD_PASSES == 2
    tmem_st8(col_addr, o);
#else
    tmem_st16(col_addr, o);
#endif
}

// sum over the warp of v[j] for every j, delivered to lane j: transposing butterfly, 31 shuffles
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int j = 0; j < off; ++j) {
            const float send = up ? v[j] : v[j + off], keep = up ? v[j + off] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

// observation (ox, oy, oz) -> world point -> PLIVox lookup (map.py:565-575): latent row, validity, cell-relative coordinates
__device__ __forceinline__ void icp_lookup(const IcpTcArgs& a, const IcpFrame& fr, bool in_range, float ox, float oy, float oz,
                                           int64_t& lrow, bool& valid, float& rx, float& ry, float& rz) {
    valid = false; lrow = 0; rx = ry = rz = 0.f;
    if (in_range) {
        // cur = (last . delta) @ obs  (tracker.py:181, motion_util.py:322-327) -- same arithmetic as icp_linearize_kernel
        const float wx = fmaf(oz, fr.pose.Rc[2], fmaf(oy, fr.pose.Rc[1], ox * fr.pose.Rc[0])) + fr.pose.tc[0];
        const float wy = fmaf(oz, fr.pose.Rc[5], fmaf(oy, fr.pose.Rc[4], ox * fr.pose.Rc[3])) + fr.pose.tc[1];
        const float wz = fmaf(oz, fr.pose.Rc[8], fmaf(oy, fr.pose.Rc[7], ox * fr.pose.Rc[6])) + fr.pose.tc[2];
        const float3 p = normalize_point(a.m.g, wx, wy, wz);
        const int ix = (int)ceilf(p.x) - 1, iy = (int)ceilf(p.y) - 1, iz = (int)ceilf(p.z) - 1;
        if (p.x == p.x && p.y == p.y && p.z == p.z && in_grid(a.m.g, ix, iy, iz)) {
            const int64_t sl = a.m.indexer[lin_id(a.m.g, ix, iy, iz)];
            // the observation count and the latent row are both addressed by the slot alone: the caller's row load is issued
            // together with the count load (one dependent DRAM hop less than "count first, then the row")
            const int64_t slc = sl >= 0 ? sl : 0;
            const float oc = __ldg(a.m.obs + slc);
            lrow = a.m.row_of ? (int64_t)__ldg(a.m.row_of + slc) : slc;       // latent ROW (sharded map: -1 = not on this rank)
            valid = sl >= 0 && oc > a.m.ignore_th && lrow >= 0;
            if (lrow < 0) lrow = 0;
        }
        rx = __fsub_rn(__fsub_rn(p.x, (float)ix), 0.5f); ry = __fsub_rn(__fsub_rn(p.y, (float)iy), 0.5f); rz = __fsub_rn(__fsub_rn(p.z, (float)iz), 0.5f);
    }
}

__global__ void __launch_bounds__(I2_THREADS, 1) icp_tc2_kernel(const unsigned char* __restrict__ image, const float* __restrict__ P, IcpTcArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = sbase + I2_OFF_BAR;
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + I2_OFF_TPTR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    IcpFrame& fr = *reinterpret_cast<IcpFrame*>(smem + I2_OFF_FRAME);
#ifdef DIF_TC_TRACE
    const bool timing = g_tc_timing != nullptr;
    int tn = 0;
#endif
    I2_TRACE(200);
#ifdef DIF_TC_TRACE
    if (timing && threadIdx.x == 0) { unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); g_tc_timing[20 * 1024 + 2 * blockIdx.x] = gt; }
#endif

    if (threadIdx.x == 0) {
#pragma unroll
        for (int g = 0; g < 4; ++g) mbar_init(bar0 + 8 * (I2_W + g), 1);
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar0 + 8 * (I2_X + s), 1);          // x tile ready (first tile: the slot's epilogue warps; then its producer warp)
            mbar_init(bar0 + 8 * (I2_XF + s), 8);         // epilogue warps -> producer: x tile parked in TMEM, buffer free
            mbar_init(bar0 + 8 * (I2_E + s), 8);          // epilogue warps -> issuer: the tile's last accumulator has been read
            mbar_init(bar0 + 8 * (I2_ACCA + s), 1); mbar_init(bar0 + 8 * (I2_ACCB + s), 1);
#pragma unroll
            for (int g = 0; g < 4; ++g) mbar_init(bar0 + 8 * (I2_G + 4 * s + g), 4);
        }
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

    // The weight image is constant data: its load is issued BEFORE the programmatic-dependent-launch wait, in the order the
    // stages need it, one barrier per group.
    if (warp == MMA_WARP && lane == 0) {
        mbar_expect_tx(bar0 + 8 * (I2_W + 0), BIAS_B + 2 * W0_B);
        bulk_g2s(sbase + OFF_BIAS, image + OFF_BIAS, BIAS_B, bar0 + 8 * (I2_W + 0));
        bulk_g2s(sbase + OFF_W0, image + OFF_W0, W0_B, bar0 + 8 * (I2_W + 0));
        bulk_g2s(sbase + PLANE_B + OFF_W0, image + PLANE_B + OFF_W0, W0_B, bar0 + 8 * (I2_W + 0));
        mbar_expect_tx(bar0 + 8 * (I2_W + 1), 2 * W1_B);
        bulk_g2s(sbase + OFF_W1, image + OFF_W1, W1_B, bar0 + 8 * (I2_W + 1));
        bulk_g2s(sbase + PLANE_B + OFF_W1, image + PLANE_B + OFF_W1, W1_B, bar0 + 8 * (I2_W + 1));
        mbar_expect_tx(bar0 + 8 * (I2_W + 2), 2 * W2_B);
        bulk_g2s(sbase + OFF_W2, image + OFF_W2, W2_B, bar0 + 8 * (I2_W + 2));
        bulk_g2s(sbase + PLANE_B + OFF_W2, image + PLANE_B + OFF_W2, W2_B, bar0 + 8 * (I2_W + 2));
        mbar_expect_tx(bar0 + 8 * (I2_W + 3), 2 * W3_B);
        bulk_g2s(sbase + OFF_W3, image + OFF_W3, W3_B, bar0 + 8 * (I2_W + 3));
        bulk_g2s(sbase + PLANE_B + OFF_W3, image + PLANE_B + OFF_W3, W3_B, bar0 + 8 * (I2_W + 3));
    }
    pdl_wait(); pdl_launch_dependents();
    // first DRAM hop of the first tile's lookup chain, issued before the frame block is resolved (a.n bounds the buffer; the
    // frame's own count is applied afterwards): it overlaps the pose composition and the barrier below
    float ox0 = 0.f, oy0 = 0.f, oz0 = 0.f;
    if (DIF_ICP2_PREFETCH && warp < MMA_WARP) {
        const int64_t i0 = ((int64_t)blockIdx.x + (int64_t)gridDim.x * (warp >> 3)) * TILE + (warp & 3) * 32 + lane;
        if (i0 < (int64_t)a.n) {
            const float* op = a.obs + (int64_t)a.obs_stride * i0;
            ox0 = __ldg(op); oy0 = __ldg(op + 1); oz0 = __ldg(op + 2);
        }
    }
    if (threadIdx.x == 0) {
        icp_resolve_frame(a.frame, a.pose, a.n, fr);     // (a device-side frame block is only read after the dependency wait)
        const unsigned int e = *a.epoch + 1u;            // epoch of this launch (written back by the reducer; 0 = "never written")
        *reinterpret_cast<unsigned int*>(smem + I2_OFF_FRAME + 140) = e ? e : 1u;
    }
    __syncthreads();
    I2_TRACE(201);
    const int64_t n_tiles = ((int64_t)fr.n + TILE - 1) / TILE;
    const bool want_grad = a.want_grad != 0;

    if (warp == MMA_WARP) {
        // ===================================================== MMA issuer: one warp, the two slots strictly alternating stage by stage
        const uint64_t f0h = smem_desc(sbase + OFF_W0, 128 * 16, 128), f0l = smem_desc(sbase + PLANE_B + OFF_W0, 128 * 16, 128);
        const uint64_t f1h = smem_desc(sbase + OFF_W1, 128 * 16, 128), f1l = smem_desc(sbase + PLANE_B + OFF_W1, 128 * 16, 128);
        const uint64_t f2h = smem_desc(sbase + OFF_W2, 96 * 16, 128), f2l = smem_desc(sbase + PLANE_B + OFF_W2, 96 * 16, 128);
        const uint64_t f3h = smem_desc(sbase + OFF_W3, 128 * 16, 128), f3l = smem_desc(sbase + PLANE_B + OFF_W3, 128 * 16, 128);
        // backward view of the same slabs: MN-major, LBO = 128 B between 8-row groups, SBO = slab chunk stride (tools/tc_probe_mn.cu);
        // B0 only needs d/d x[16..31] (columns 29..31 = xyz): the view starts at K chunk 2 of W0 and N = 16
        const uint64_t b0h = smem_desc(sbase + OFF_W0 + 2 * 2048, 128, 2048), b0l = smem_desc(sbase + PLANE_B + OFF_W0 + 2 * 2048, 128, 2048);
        const uint64_t b1h = smem_desc(sbase + OFF_W1, 128, 2048), b1l = smem_desc(sbase + PLANE_B + OFF_W1, 128, 2048);
        const uint64_t b2h = smem_desc(sbase + OFF_W2, 128, 1536), b2l = smem_desc(sbase + PLANE_B + OFF_W2, 128, 1536);
        const uint64_t b3h = smem_desc(sbase + OFF_W3, 128, 2048), b3l = smem_desc(sbase + PLANE_B + OFF_W3, 128, 2048);
        const uint64_t xh0 = smem_desc(sbase + OFF_X, X_CHUNK_B, 128), xl0 = smem_desc(sbase + OFF_X + X_PLANE_B, X_CHUNK_B, 128);
        const uint64_t xh1 = smem_desc(sbase + OFF_X + 2 * X_PLANE_B, X_CHUNK_B, 128), xl1 = smem_desc(sbase + OFF_X + 3 * X_PLANE_B, X_CHUNK_B, 128);
        const uint32_t A0 = tmem, A1 = tmem + 128, B0 = tmem + 256, B1 = tmem + 384;              // slot A: R0 / R1, slot B: R0 / R1
        const uint32_t bgA = bar0 + 8 * I2_G, bgB = bar0 + 8 * (I2_G + 4);
        const uint32_t accaA = bar0 + 8 * I2_ACCA, accaB = accaA + 8, accbA = bar0 + 8 * I2_ACCB, accbB = accbA + 8;
        uint32_t ph_g = 0, ph_xA = 0, ph_xB = 0, ph_eA = 0, ph_eB = 0;
        bool l0A = false, l0B = false;                     // F0 of the current iteration already issued (early, behind the previous tile's last stage)
        auto issue_l0 = [&](const int s) {
            tc_fence_after();
            if (elect_one()) {
                if (s == 0) { issue_layer(idesc_f16(128), 128 * 16, 0, 2, A0, 0, 0, xh0, xl0, f0h, f0l); mma_commit(accaA); }
                else { issue_layer(idesc_f16(128), 128 * 16, 0, 2, B0, 0, 0, xh1, xl1, f0h, f0l); mma_commit(accaB); }
            }
            __syncwarp();
        };
        for (int64_t it = 0;; ++it) {
            const int64_t t0 = blockIdx.x + (int64_t)gridDim.x * (2 * it), t1 = t0 + gridDim.x;
            if (t0 >= n_tiles) break;
            const bool liveB = t1 < n_tiles;
            const bool nextA = t0 + 2 * (int64_t)gridDim.x < n_tiles, nextB = t1 + 2 * (int64_t)gridDim.x < n_tiles;
            // ---- F0: x tile (smem) -> R0.  Energy-only launches end a tile with the head exchange in R0: F0 waits for it.
            if (!l0A) {
                mbar_wait_spin(bar0 + 8 * I2_X, ph_xA); ph_xA ^= 1;
                if (it == 0) mbar_wait_spin(bar0 + 8 * (I2_W + 0), 0);
                else if (!want_grad) { mbar_wait_spin(bar0 + 8 * I2_E, ph_eA); ph_eA ^= 1; }
                I2_TRACE(1);
                issue_l0(0);
                I2_TRACE(2);
            }
            if (liveB && !l0B) {
                mbar_wait_spin(bar0 + 8 * (I2_X + 1), ph_xB); ph_xB ^= 1;
                if (it > 0 && !want_grad) { mbar_wait_spin(bar0 + 8 * (I2_E + 1), ph_eB); ph_eB ^= 1; }
                I2_TRACE(3);
                issue_l0(1);
                I2_TRACE(4);
            }
            l0A = l0B = false;
            // ---- F1 (R0 -> R1): the previous tile's last read of R1 must be over.  One copy of every stage's issue code serves
            //      both slots (run-time bases, compile-time K steps): the launch runs each instruction about once, so code
            //      size is instruction-fetch latency on the critical path.
            const int ns = liveB ? 2 : 1;
            if (it == 0) { mbar_wait_spin(bar0 + 8 * (I2_W + 1), 0); I2_TRACE(5); }
#pragma unroll 1
            for (int s = 0; s < ns; ++s) {
                if (it > 0 && want_grad) {
                    if (s == 0) { mbar_wait_spin(bar0 + 8 * I2_E, ph_eA); ph_eA ^= 1; } else { mbar_wait_spin(bar0 + 8 * (I2_E + 1), ph_eB); ph_eB ^= 1; }
                }
                icp2_issue_stage<1>(s ? B0 : A0, s ? B1 : A1, f1h, f1l, s ? bgB : bgA, ph_g, s ? accbB : accbA);
                I2_TRACE(10 + s);
            }
            if (it == 0) { mbar_wait_spin(bar0 + 8 * (I2_W + 2), 0); I2_TRACE(6); }
#pragma unroll 1
            for (int s = 0; s < ns; ++s) { icp2_issue_stage<2>(s ? B0 : A0, s ? B1 : A1, f2h, f2l, s ? bgB : bgA, ph_g ^ 1u, s ? accaB : accaA); I2_TRACE(12 + s); }
            if (it == 0) { mbar_wait_spin(bar0 + 8 * (I2_W + 3), 0); I2_TRACE(7); }
#pragma unroll 1
            for (int s = 0; s < ns; ++s) { icp2_issue_stage<3>(s ? B0 : A0, s ? B1 : A1, f3h, f3l, s ? bgB : bgA, ph_g, s ? accbB : accbA); I2_TRACE(14 + s); }
            if (want_grad) {
#pragma unroll 1
                for (int s = 0; s < ns; ++s) { icp2_issue_stage<4>(s ? B0 : A0, s ? B1 : A1, b3h, b3l, s ? bgB : bgA, ph_g ^ 1u, s ? accaB : accaA); I2_TRACE(16 + s); }
#pragma unroll 1
                for (int s = 0; s < ns; ++s) { icp2_issue_stage<5>(s ? B0 : A0, s ? B1 : A1, b2h, b2l, s ? bgB : bgA, ph_g, s ? accbB : accbA); I2_TRACE(18 + s); }
#pragma unroll 1
                for (int s = 0; s < ns; ++s) { icp2_issue_stage<6>(s ? B0 : A0, s ? B1 : A1, b1h, b1l, s ? bgB : bgA, ph_g ^ 1u, s ? accaB : accaA); I2_TRACE(20 + s); }
#pragma unroll 1
                for (int s = 0; s < ns; ++s) {
                    icp2_issue_stage<7>(s ? B0 : A0, s ? B1 : A1, b0h, b0l, s ? bgB : bgA, ph_g, s ? accbB : accbA);
                    I2_TRACE(22 + s);
                    // F0 of the slot's next tile right behind B0 when its x tile is ready (R0 is free once B0 is in the in-order pipe)
                    if (s == 0) { if (nextA && __all_sync(0xffffffffu, mbar_test(bar0 + 8 * I2_X, ph_xA))) { ph_xA ^= 1; issue_l0(0); l0A = true; } }
                    else if (nextB && __all_sync(0xffffffffu, mbar_test(bar0 + 8 * (I2_X + 1), ph_xB))) { ph_xB ^= 1; issue_l0(1); l0B = true; }
                }
            }
            ph_g ^= 1;
        }
    } else if (warp >= PRODUCER_WARP0) {
        // ===================================================== gather producer of slot s, from the slot's SECOND tile on (one observation per lane)
        const int s = warp - PRODUCER_WARP0;
        if (s < 2) {
            unsigned char* x_hi_p = smem + OFF_X + s * 2 * X_PLANE_B;
            unsigned char* aux = smem + I2_OFF_AUX + s * TILE;
            float xa[32], xb[32];
            bool va = false, vb = false;
            uint32_t ph_xf = 0;
            int64_t tile = blockIdx.x + (int64_t)gridDim.x * (2 + s);
            if (tile < n_tiles) icp_gather_row(a, fr, tile * TILE + lane, xa, va);
            while (tile < n_tiles) {
                mbar_wait(bar0 + 8 * (I2_XF + s), ph_xf); ph_xf ^= 1;        // the previous tile's inputs are parked in TMEM
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
                if (lane == 0) mbar_arrive(bar0 + 8 * (I2_X + s));
                tile = next;
            }
        }
    } else {
        // ===================================================== epilogue warps of slot s: lane quadrant x column half
        const int s = warp >> 3;
        const int quad = warp & 3, half = (warp >> 2) & 1;
        const int row = quad * 32 + lane;
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        const uint32_t R0 = tmem + s * 256 + lane_base, R1 = R0 + 128;
        const float* bias = reinterpret_cast<const float*>(smem + OFF_BIAS);
        const unsigned char* aux = smem + I2_OFF_AUX + s * TILE;
        const int head_slot = *reinterpret_cast<const int*>(image + IMAGE_B);
        const float* w4c = c_head_w[head_slot][0];
        const float* wuc = c_head_w[head_slot][1];
        const float b4 = __ldg(P + DecW::b4), bu = __ldg(P + DecW::bu);
        const uint32_t acc_a = bar0 + 8 * (I2_ACCA + s), acc_b = bar0 + 8 * (I2_ACCB + s);
        const uint32_t g_bar = bar0 + 8 * (I2_G + 4 * s + half);                     // + 16 bytes for iteration 1
        uint32_t ph_a = 0, ph_b = 0;
        double lane_sum = 0.0;                          // lane j of a half-1 warp: running sum of value j over this warp's tiles
        auto wait_acc = [&](bool b_side) {
#if DIF_ICP2_EPI_SLEEP == 0
            if (b_side) { mbar_wait_spin(acc_b, ph_b); ph_b ^= 1; } else { mbar_wait_spin(acc_a, ph_a); ph_a ^= 1; }
#else
            if (b_side) { mbar_wait_ns<DIF_ICP2_EPI_SLEEP>(acc_b, ph_b); ph_b ^= 1; } else { mbar_wait_ns<DIF_ICP2_EPI_SLEEP>(acc_a, ph_a); ph_a ^= 1; }
#endif
            tc_fence_after();
        };
        auto hand_off = [&](uint32_t bar) {
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar);
        };
        for (int64_t it = 0;; ++it) {
            const int64_t tile = blockIdx.x + (int64_t)gridDim.x * (2 * it + s);
            if (tile >= n_tiles) break;
            const int64_t i = tile * TILE + row;
            bool valid = false;
            float ox = ox0, oy = oy0, oz = oz0;
            if ((it > 0 || !DIF_ICP2_PREFETCH) && i < fr.n && (half == 1 || it == 0)) {
                const float* op = a.obs + (int64_t)a.obs_stride * i;
                ox = __ldg(op); oy = __ldg(op + 1); oz = __ldg(op + 2);
            }
            if (it == 0) {
                // ---- the slot's first tile: this thread looks up its own row and writes K chunks 2*half, 2*half+1 of the layer-0 A tile
                int64_t lrow; float rx, ry, rz;
                icp_lookup(a, fr, i < fr.n, ox, oy, oz, lrow, valid, rx, ry, rz);
                float x[16];
                const float* lp = a.m.latent + lrow * a.m.lat_stride + 16 * half;
                if (a.m.lat_stride == 32) {
                    const float4* q = reinterpret_cast<const float4*>(lp);
#pragma unroll
                    for (int j = 0; j < 4; ++j) { const float4 v = __ldg(q + j); x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w; }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) x[j] = (16 * half + j < DIF_L) ? __ldg(lp + j) : 0.f;
                }
                if (half == 1) { x[13] = rx; x[14] = ry; x[15] = rz; }
                unsigned char* x_hi_p = smem + OFF_X + s * 2 * X_PLANE_B;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint4 h, l;
                    split_pair(valid ? x[8 * c + 0] : 0.f, valid ? x[8 * c + 1] : 0.f, h.x, l.x); split_pair(valid ? x[8 * c + 2] : 0.f, valid ? x[8 * c + 3] : 0.f, h.y, l.y);
                    split_pair(valid ? x[8 * c + 4] : 0.f, valid ? x[8 * c + 5] : 0.f, h.z, l.z); split_pair(valid ? x[8 * c + 6] : 0.f, valid ? x[8 * c + 7] : 0.f, h.w, l.w);
                    *reinterpret_cast<uint4*>(x_hi_p + (2 * half + c) * X_CHUNK_B + row * 16) = h;
                    *reinterpret_cast<uint4*>(x_hi_p + X_PLANE_B + (2 * half + c) * X_CHUNK_B + row * 16) = l;
                }
                fence_async_smem();                          // generic-proxy stores -> visible to the tensor core (async proxy)
                named_bar_sync(9 + s, 256);
                if ((warp & 7) == 0 && lane == 0) mbar_arrive(bar0 + 8 * (I2_X + s));
                I2_TRACE(60); I2_GT(0);
                mbar_wait(bar0 + 8 * (I2_W + 0), 0);         // biases arrive with the first weight group
            }
            uint64_t m0 = 0, m1 = 0, m2 = 0;              // ReLU masks of this thread's columns of h0, h1, h2
            // ---- forward hidden layers 0..2: accumulator -> +bias, ReLU, split in place, 32 columns per hand-off group
#pragma unroll 1
            for (int layer = 0; layer < 3; ++layer) {
                const uint32_t R = layer == 1 ? R1 : R0;
                const int hw = layer == 2 ? 48 : 64;
                const int cb = half * hw;
                const float* b = bias + layer * 128 + cb;
                wait_acc(layer == 1);
                I2_TRACE(80 + layer); if (layer < 2) I2_GT(1 + layer);
                if (layer == 0 && it > 0) valid = aux[row] != 0;
                if (layer == 1) {
                    // F1 has completed, so R0 (its A operand) is dead until F2 writes columns 0..95: park this half's 16 inputs in
                    // R0[96 + 16*half ..) - the skip-connection K chunks 6 / 7 of F3 - and hand the x tile back to the producer
                    const unsigned char* xp = smem + OFF_X + s * 2 * X_PLANE_B + (2 * half) * X_CHUNK_B + row * 16;
                    const uint4 h0 = *reinterpret_cast<const uint4*>(xp), h1 = *reinterpret_cast<const uint4*>(xp + X_CHUNK_B);
                    const uint4 l0 = *reinterpret_cast<const uint4*>(xp + X_PLANE_B), l1 = *reinterpret_cast<const uint4*>(xp + X_PLANE_B + X_CHUNK_B);
                    const uint32_t xv[16] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w, l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
                    tmem_st16(R0 + 96 + 16 * half, xv);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar0 + 8 * (I2_XF + s));
                }
                // two hand-off groups of 32 columns (the second one of layer 2: 16 columns + the parked inputs, in place since E1);
                // ONE rolled copy of the code: the launch runs it about once per stage, and instruction fetch of cold code is
                // part of the critical path (trace: the first conversion of a launch takes 2x the later ones)
                uint64_t mm = 0;
                I2_GRP_PRAGMA
                for (int grp = 0; grp < 2; ++grp) {
                    const int c = cb + 32 * grp;
                    const bool two = !(layer == 2 && grp == 1);
                    uint32_t v0[16], v1[16];
                    tmem_ld16_nowait(R + c, v0);
                    if (two) tmem_ld16_nowait(R + c + 16, v1);
                    tmem_ld_wait();
                    uint32_t mk = cvt_fwd16(v0, b + 32 * grp, R + c);
                    if (two) mk |= cvt_fwd16(v1, b + 32 * grp + 16, R + c + 16) << 16;
                    mm |= (uint64_t)mk << (32 * grp);
                    hand_off(g_bar + 16 * grp);
                    I2_TRACE((grp ? 100 : 90) + layer);
                }
                if (layer == 0) m0 = mm; else if (layer == 1) m1 = mm; else m2 = mm;
            }
            // ---- F3 result.  The backward pass is LINEAR in the per-row seed d r / d pre_sdf = (1 - sdf^2) / std, so it runs with
            //      seed = 1: g3' = G3_SCALE * w4 * relu'(h3) (the split constants come from the constant bank, the ReLU mask from the
            //      sign bytes of h3 with one PRMT per pair) and the seed multiplies the three xyz gradients at the very end.  B3 therefore
            //      starts as soon as the SIGNS of h3 are known; the head dot products run behind each hand-off on the registers that
            //      are still live, and tanh / softplus / the divisions leave the chain altogether.  Each column half keeps its two
            //      partial sums; half 0 passes them to the row's owner (half 1) through two columns of B3's result that nobody
            //      needs (R0[96,98): d / d latent), written before its EB3 hand-off and read once B2 has completed - ordered by the
            //      G hand-off -> B2 -> ACCb chain, no extra synchronisation; the owner finishes the heads in the shadow of B1.
            wait_acc(true);
            I2_TRACE(83); I2_GT(3);
            float ps = 0.f, pu = 0.f;
            const int cb3 = half * 64;
            float gxs0 = 0.f, gxs1 = 0.f, gxs2 = 0.f;                  // d pre_sdf / d xyz * G3_SCALE (half-1 threads): skip path + direct path
            float r = 0.f, seed = 0.f;
            if (want_grad) {
                const uint32_t* g3c = c_head_g3[head_slot];
#pragma unroll 1
                for (int c = 0; c < 64; c += 32) {
                    uint32_t v0[16], v1[16];
                    tmem_ld16_nowait(R1 + cb3 + c, v0);
                    tmem_ld16_nowait(R1 + cb3 + c + 16, v1);
                    tmem_ld_wait();
                    const int k = cb3 + c;
                    g3_signs16(v0, bias + 352 + k, g3c + k / 2, R1 + k);
                    g3_signs16(v1, bias + 352 + k + 16, g3c + k / 2 + 8, R1 + k + 16);
                    hand_off(g_bar + (c ? 16 : 0));
                    head_fma16(v0, bias + 352 + k, w4c + k, wuc + k, ps, pu);
                    head_fma16(v1, bias + 352 + k + 16, w4c + k + 16, wuc + k + 16, ps, pu);
                }
                I2_TRACE(93);
                // ---- backward stages B3, B2, B1: D -> relu' mask -> A operand of the next stage, in place
#pragma unroll 1
                for (int st = 0; st < 3; ++st) {
                    const uint32_t R = st == 1 ? R1 : R0;
                    const int hw = st == 0 ? 48 : 64;                   // B3 feeds the 96-wide layer 2
                    const int cb = half * hw;
                    const uint64_t mm = st == 0 ? m2 : (st == 1 ? m1 : m0);
                    wait_acc(st == 1);
                    I2_TRACE(85 + st); if (st == 0) I2_GT(4);
                    if (st == 0 && half == 1) {                         // columns 125..127 of g3 * W3: skip-path gradient wrt xyz
                        uint32_t t4[4];
                        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(t4[0]), "=r"(t4[1]), "=r"(t4[2]), "=r"(t4[3]) : "r"(R0 + 124));
                        tmem_ld_wait();
                        gxs0 = __uint_as_float(t4[1]); gxs1 = __uint_as_float(t4[2]); gxs2 = __uint_as_float(t4[3]);
                    }
                    if (st == 0 && half == 0)                           // B3 has completed: its columns 96..123 (d / d latent) are not needed
                        asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" :: "r"(R0 + 96), "r"(__float_as_uint(ps)), "r"(__float_as_uint(pu)) : "memory");
                    if (st == 1 && half == 1) {                         // B2 has completed => half 0's EB3 hand-off (and the store before it) is visible
                        uint32_t q2[2];
                        asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(q2[0]), "=r"(q2[1]) : "r"(R0 + 96));
                        tmem_ld_wait();
                        ps = (__uint_as_float(q2[0]) + ps) + b4;
                        pu = (__uint_as_float(q2[1]) + pu) + bu;
                    }
                    I2_GRP_PRAGMA
                    for (int grp = 0; grp < 2; ++grp) {
                        const int c = cb + 32 * grp;
                        const bool two = !(st == 0 && grp == 1);
                        const uint32_t mk = (uint32_t)(mm >> (32 * grp));
                        uint32_t v0[16], v1[16];
                        tmem_ld16_nowait(R + c, v0);
                        if (two) tmem_ld16_nowait(R + c + 16, v1);
                        tmem_ld_wait();
                        cvt_bwd16(v0, mk & 0xFFFFu, R + c);
                        if (two) cvt_bwd16(v1, mk >> 16, R + c + 16);
                        hand_off(g_bar + 16 * grp);
                    }
                    I2_TRACE(95 + st);
                    if (st == 1 && half == 1) {                         // in the shadow of B1: residual, seed, Huber weight (tracker.py:186, :59-65)
                        const float sdf = tanhf(ps);
                        const float sd = 0.05f + 0.5f * softplus_ref(pu);
                        r = sdf / sd;
                        seed = ((1.f - sdf * sdf) / sd) * (1.f / G3_SCALE);       // d r / d pre_sdf (std detached), and the constant scale of g3'
                    }
                }
                // ---- B0 result: columns 13..15 of R1 = direct-path gradient
                wait_acc(true);
                I2_TRACE(88); I2_GT(5);
                if (half == 1) {
                    uint32_t t4[4];
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(t4[0]), "=r"(t4[1]), "=r"(t4[2]), "=r"(t4[3]) : "r"(R1 + 12));
                    tmem_ld_wait();
                    gxs0 = (gxs0 + __uint_as_float(t4[1])) * seed; gxs1 = (gxs1 + __uint_as_float(t4[2])) * seed; gxs2 = (gxs2 + __uint_as_float(t4[3])) * seed;   // d r / d xyz
                }
            } else {
                // energy only: both heads, partial sums exchanged through four columns of R0 (dead since F3 completed)
#pragma unroll 1
                for (int c = 0; c < 64; c += 32) {
                    uint32_t v0[16], v1[16];
                    tmem_ld16_nowait(R1 + cb3 + c, v0);
                    tmem_ld16_nowait(R1 + cb3 + c + 16, v1);
                    tmem_ld_wait();
                    const int k = cb3 + c;
                    head_fma16(v0, bias + 352 + k, w4c + k, wuc + k, ps, pu);
                    head_fma16(v1, bias + 352 + k + 16, w4c + k + 16, wuc + k + 16, ps, pu);
                }
                asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" :: "r"(R0 + 2 * half), "r"(__float_as_uint(ps)), "r"(__float_as_uint(pu)) : "memory");
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                named_bar_sync(1 + s * 4 + quad, 64);
                tc_fence_after();
                uint32_t q4[4];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(q4[0]), "=r"(q4[1]), "=r"(q4[2]), "=r"(q4[3]) : "r"(R0));
                tmem_ld_wait();
                ps = (__uint_as_float(q4[0]) + __uint_as_float(q4[2])) + b4;
                pu = (__uint_as_float(q4[1]) + __uint_as_float(q4[3])) + bu;
            }
            if (!want_grad && half == 1) r = tanhf(ps) / (0.05f + 0.5f * softplus_ref(pu));      // tracker.py:186
            I2_TRACE(84);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + 8 * (I2_E + s));        // this tile's accumulators are consumed
            // ---- residual / Jacobian / Huber / normal equations (half-1 warps own the rows), tracker.py:196-216
            if (half == 1) {
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.f;
                if (valid) {
                    float w = 1.f;
                    if (a.huber_k > 0.f) { const float ar = fabsf(r); if (ar > a.huber_k) w = a.huber_k / ar; }
                    else if (a.huber_k < 0.f) { const float q = r / -a.huber_k, t1 = 1.f - q * q; w = fabsf(r) <= -a.huber_k ? t1 * t1 : 0.f; }      // Tukey (tracker.py:66-69)
                    v[27] = r * (r * w);
                    v[28] = 1.f;
                    if (want_grad) {
                        const Pose& ps_ = fr.pose;
                        const float qx = fmaf(oz, ps_.Rd[2], fmaf(oy, ps_.Rd[1], ox * ps_.Rd[0])) + ps_.td[0];
                        const float qy = fmaf(oz, ps_.Rd[5], fmaf(oy, ps_.Rd[4], ox * ps_.Rd[3])) + ps_.td[1];
                        const float qz = fmaf(oz, ps_.Rd[8], fmaf(oy, ps_.Rd[7], ox * ps_.Rd[6])) + ps_.td[2];
                        const float gx = gxs0 / a.m.g.vs, gy = gxs1 / a.m.g.vs, gz = gxs2 / a.m.g.vs;      // (gxs: already d r / d xyz)
                        float J[6];
                        J[0] = gx * ps_.Rl[0] + gy * ps_.Rl[1] + gz * ps_.Rl[2];
                        J[1] = gx * ps_.Rl[3] + gy * ps_.Rl[4] + gz * ps_.Rl[5];
                        J[2] = gx * ps_.Rl[6] + gy * ps_.Rl[7] + gz * ps_.Rl[8];
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
                lane_sum += (double)warp_transpose_sum32(v, lane);
            }
            I2_TRACE(110); I2_GT(6);
        }
        // no tile of this slot is left: its x tile is dead and becomes the CTA's reduction buffer [slot][quad][32] (fp64)
        if (half == 1) reinterpret_cast<double*>(smem + OFF_X + s * 2 * X_PLANE_B)[quad * 32 + lane] = lane_sum;
    }
    tc_fence_before();
    __syncthreads();
    I2_TRACE(202);
    if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(512));
    // ---- deterministic grid reduction: fixed-order sum of the CTA's 8 warps -> one row of 32 fp64 partial sums per CTA; ONE
    //      reducer CTA adds the rows in a fixed order, scales by 1/M and expands the symmetric H.
    // The rows travel as self-validating words (the low-latency protocol of collective libraries): every 8-byte word carries
    // 32 bits of payload and the launch's 32-bit epoch, so a row needs no fence and no ticket - the reducer polls the table until
    // every word shows the current epoch.  The chain behind the CTA that finishes last is one store, one poll and the final sum
    // (the ticket version - release fence, atomic round trip, acquire, ten dependent loads - cost ~5 us of a 29 us launch).
    // The reducer is the LAST CTA of the grid (it has the fewest tiles); nobody waits for it and it waits for no CTA that is not
    // running or queued, so the wait cannot deadlock; it is bounded anyway (NaN results instead of a hung stream).
#if DIF_ICP2_TAIL == 2
    const unsigned int epoch = *reinterpret_cast<const unsigned int*>(smem + I2_OFF_FRAME + 140);
    const unsigned int reducer = gridDim.x - 1;
    double* red = reinterpret_cast<double*>(smem + OFF_X);                 // [19 slices][32]; overlays slot 0's [4][32] buffer once warp 0 has read it
    if (warp == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += reinterpret_cast<const double*>(smem + OFF_X + (w >> 2) * 2 * X_PLANE_B)[(w & 3) * 32 + lane];
        if (blockIdx.x != reducer) {
            const unsigned long long bits = (unsigned long long)__double_as_longlong(t), tag = (unsigned long long)epoch << 32;
            asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" :: "l"(a.ll + ((size_t)blockIdx.x * 32 + lane) * 2),
                         "l"(tag | (bits & 0xFFFFFFFFull)), "l"(tag | (bits >> 32)) : "memory");
        } else {
            reinterpret_cast<double*>(smem + I2_OFF_AUX)[lane] = t;      // (the validity bytes are dead: 256 B = 32 doubles)
        }
    }
    I2_TRACE(203);
#ifdef DIF_TC_TRACE
    if (timing && threadIdx.x == 0) { unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); g_tc_timing[20 * 1024 + 2 * blockIdx.x + 1] = gt; }
#endif
    if (blockIdx.x != reducer) return;
    {
        static_assert(DIF_NUM_SMS <= 19 * 8, "rows per thread");
        double t8[8];
        unsigned int pending = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { t8[j] = 0.0; if ((unsigned)(warp + 19 * j) < reducer) pending |= 1u << j; }
        for (int spin = 0; pending && spin < (1 << 20); ++spin) {
            unsigned long long w0[8], w1[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)                     // all pending rows in flight together: one L2 round trip per sweep
                if (pending >> j & 1u)
                    asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(w0[j]), "=l"(w1[j]) : "l"(a.ll + ((size_t)(warp + 19 * j) * 32 + lane) * 2) : "memory");
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if ((pending >> j & 1u) && (unsigned int)(w0[j] >> 32) == epoch && (unsigned int)(w1[j] >> 32) == epoch) {
                    t8[j] = __longlong_as_double((long long)((w0[j] & 0xFFFFFFFFull) | (w1[j] << 32)));
                    pending &= ~(1u << j);
                }
            }
        }
        double t = pending ? __longlong_as_double(0x7FF8000000000000ll) : 0.0;      // a row never arrived: NaN results, not a hang
#pragma unroll
        for (int j = 0; j < 8; ++j) t += t8[j];
        __syncthreads();                                                   // (the [slot][quad][32] buffers have been consumed by warp 0)
        red[threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        double tot = reinterpret_cast<const double*>(smem + I2_OFF_AUX)[threadIdx.x];      // the reducer's own row
#pragma unroll
        for (int w = 0; w < 19; ++w) tot += red[w * 32 + threadIdx.x];
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
        if (threadIdx.x == 0) *a.epoch = epoch;            // the next launch (stream order) uses the next epoch
    }
#else
    __shared__ bool is_last2;
    if (warp == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += reinterpret_cast<const double*>(smem + OFF_X + (w >> 2) * 2 * X_PLANE_B)[(w & 3) * 32 + lane];
#if DIF_ICP2_TAIL == 1
        __stcg(a.partials + (size_t)blockIdx.x * 32 + lane, t);
        __syncwarp();
        if (lane == 0) {
            unsigned int ticket;
            asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(ticket) : "l"(a.done_counter) : "memory");
            is_last2 = ticket == gridDim.x - 1;
        }
#else
        a.partials[(size_t)blockIdx.x * 32 + lane] = t;
        __threadfence();
        __syncwarp();
        if (lane == 0) is_last2 = atomicAdd(a.done_counter, 1u) == gridDim.x - 1;
#endif
    }
    __syncthreads();
    I2_TRACE(203);
#ifdef DIF_TC_TRACE
    if (timing && threadIdx.x == 0) { unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); g_tc_timing[20 * 1024 + 2 * blockIdx.x + 1] = gt; }
#endif
    if (!is_last2) return;
#if DIF_ICP2_TAIL != 1
    __threadfence();
#endif
    double* red = reinterpret_cast<double*>(smem + OFF_X);                 // [19 slices][32]
    {
        static_assert(DIF_NUM_SMS <= 19 * 8, "rows per thread");
        double t8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const unsigned c = warp + 19 * j;
            t8[j] = c < gridDim.x ? __ldcg(a.partials + (size_t)c * 32 + lane) : 0.0;
        }
        double t = 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j) t += t8[j];
        red[threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < 19; ++w) tot += red[w * 32 + threadIdx.x];
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
#endif
#ifdef DIF_TC_TRACE
    if (timing && threadIdx.x == 0) { unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); g_tc_timing[20 * 1024 + 2 * 148] = gt; g_tc_timing[20 * 1024 + 2 * 148 + 1] = blockIdx.x; }
#endif
}

}  // namespace tc
}  // namespace dif
