// Tensor-core (tcgen05 / TMEM) evaluation of the DI-Fusion point encoder with the 8-offset sample gather fused in front and
// the per-PLIVox scatter-accumulate fused behind it.
//   replaces reference network/di_encoder.py:26-30 (Conv1d(k=1)+BN+ReLU x3, Conv1d) evaluated on the materialised
//   (S, 6) sample tensor of system/map.py:419-446, and ext/indexing/indexing.cu:59-109 (groupby_sum).  SURVEY rows a-4..a-6.
//
// Persistent, one CTA per SM, 128-sample tiles (= TMEM lanes).  20 warps: 16 epilogue warps (TMEM lane quadrant x column
// quarter), one MMA issuer, two gather producers (64 samples of a tile each; the x tile is double buffered so they run one
// tile ahead).  Layers as tcgen05.mma M128 x N{32,64,256,32} x K16 with fp32 accumulation in TMEM and the same 3-pass fp16
// hi/lo split as the decoder (hi*hi + lo*hi + hi*lo); BN is folded into the weights at load.
//   L0  x[128x16 (6 real)]  (smem)  * W0^T -> acc[:, 0:32]    -> +b, ReLU, split -> h0 (A operand in TMEM)
//   L1  h0[128x32]          (TMEM)  * W1^T -> acc[:, 0:64]    -> h1
//   L2  h1[128x64]          (TMEM)  * W2^T -> acc[:, 0:256]   -> h2 (K = 256: the whole 256-column A region)
//   L3  h2[128x256]         (TMEM)  * W3^T -> acc[:, 0:32]    -> +b3 -> float4 atomicAdd into slot_sum[slot][0:32]  (or plain store)
// TMEM map (512 columns): accumulator [0, 256), A operand [256, 512).
#include <stdlib.h>

#include "tc_common.cuh"

namespace dif {
namespace enc {

using namespace tc;

// ---- weight image (bytes): hi plane, lo plane, biases -------------------------------------------------------------------
constexpr uint32_t EW0_B = 32 * 16 * 2, EW1_B = 64 * 32 * 2, EW2_B = 256 * 64 * 2, EW3_B = 32 * 256 * 2;
constexpr uint32_t EOFF_W0 = 0, EOFF_W1 = EOFF_W0 + EW0_B, EOFF_W2 = EOFF_W1 + EW1_B, EOFF_W3 = EOFF_W2 + EW2_B;
constexpr uint32_t EPLANE_B = EOFF_W3 + EW3_B;                    // 54272
constexpr uint32_t EOFF_BIAS = 2 * EPLANE_B;                      // b0[32] b1[64] b2[256] b3[32]
constexpr uint32_t EBIAS_B = 384 * 4;
constexpr uint32_t EIMAGE_B = EOFF_BIAS + EBIAS_B;                // 110080
constexpr uint32_t EX_CHUNK_B = TILE * 16 + 16;                   // one 8-wide K chunk of the x tile (+16 B bank skew)
constexpr uint32_t EX_PLANE_B = 2 * EX_CHUNK_B;                   // K = 16: chunk 0 = (rel xyz, normal, 0, 0), chunk 1 = zeros
constexpr uint32_t EOFF_X = EIMAGE_B;                             // [2 buffers][hi, lo]
constexpr uint32_t EOFF_SLOT = EOFF_X + 4 * EX_PLANE_B;           // [4 buffers][128] int32: target PLIVox slot of every sample
constexpr uint32_t EOFF_BAR = EOFF_SLOT + 4 * TILE * 4;             //   (4 deep: a tile's slots are read by its epilogue long after layer 0 freed the x buffer)
constexpr uint32_t ESMEM_B = EOFF_BAR + 96 + 16;
static_assert(EIMAGE_B % 16 == 0 && ESMEM_B <= 232448, "shared memory layout");

enum { EB_W = 0, EB_X0 = 1, EB_X1 = 2, EB_XF0 = 3, EB_XF1 = 4, EB_ACC = 5, EB_A = 6, EB_E = 7 };

struct EncodeArgs {
    // mode 0: samples come from the integrate sample list (point index, offset index, target slot); results are accumulated
    //         into slot_sum.  mode 1: explicit (n, 6) inputs, results stored to out (n, 29)  (dif_encode).
    int mode;
    Grid g; const float* p_hat; const float* normal; int normal_stride; const int32_t* s_pt; const int32_t* s_slot; const uint8_t* s_off;
    const int32_t* n_dev; float* slot_sum;
    const float* xyzn; int64_t n; float* out;
};

__device__ __forceinline__ void tmem_ld8_nowait(uint32_t addr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(addr));
}
__device__ __forceinline__ void tmem_st4(uint32_t addr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}

// +bias, ReLU, hi/lo split of 16 accumulator columns -> 8 + 8 packed A-operand columns
__device__ __forceinline__ void convert16(const uint32_t* v, const float* b, uint32_t a_hi, uint32_t a_lo) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float2 bb = *reinterpret_cast<const float2*>(b + 2 * j);
        split_pair(fmaxf(__uint_as_float(v[2 * j]) + bb.x, 0.f), fmaxf(__uint_as_float(v[2 * j + 1]) + bb.y, 0.f), hi[j], lo[j]);
    }
    tmem_st8(a_hi, hi); tmem_st8(a_lo, lo);
}

// TS stage: D (+)= A_tmem * B (three passes); w_step16 = descriptor increment per K=16 step
__device__ __forceinline__ void issue_ts(uint32_t idesc, uint32_t w_step16, int ksteps, uint32_t acc, uint32_t a_hi_t, uint32_t a_lo_t,
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

__global__ void __launch_bounds__(THREADS, 1) encode_tc_kernel(const unsigned char* __restrict__ image, EncodeArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = sbase + EOFF_BAR;
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + EOFF_BAR + 96);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_wait(); pdl_launch_dependents();                          // the sample count below is written by the previous kernel
    const int64_t n_total = a.mode == 0 ? (int64_t)*a.n_dev : a.n;
    const int64_t n_tiles = (n_total + TILE - 1) / TILE;
    if ((int64_t)blockIdx.x >= n_tiles) return;                   // nothing to do: skip the weight load altogether

    if (threadIdx.x == 0) {
        mbar_init(bar0 + 8 * EB_W, 1);
        mbar_init(bar0 + 8 * EB_X0, 2); mbar_init(bar0 + 8 * EB_X1, 2);        // two producer warps fill one x buffer
        mbar_init(bar0 + 8 * EB_XF0, 1); mbar_init(bar0 + 8 * EB_XF1, 1);      // tcgen05.commit after layer 0: buffer free
        mbar_init(bar0 + 8 * EB_ACC, 1);
        mbar_init(bar0 + 8 * EB_A, 16);
        mbar_init(bar0 + 8 * EB_E, 16);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // K chunk 1 of every x plane is identically zero (the encoder has 6 inputs, padded to K = 16): written once
    for (int i = threadIdx.x; i < 4 * (int)(EX_CHUNK_B / 16); i += THREADS)
        *reinterpret_cast<uint4*>(smem + EOFF_X + (i / (EX_CHUNK_B / 16)) * EX_PLANE_B + EX_CHUNK_B + (i % (EX_CHUNK_B / 16)) * 16) = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_ptr_s)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_ptr_s, 0);
    const uint32_t acc_c = tmem, a_c = tmem + 256;                // accumulator / A-operand column bases

    if (warp == MMA_WARP) {
        // ===================================================== weight load + MMA issuer (warp-uniform, instructions elected)
        if (lane == 0) {
            mbar_expect_tx(bar0 + 8 * EB_W, EIMAGE_B);
            constexpr uint32_t CH = 32768;
            for (uint32_t off = 0; off < EIMAGE_B; off += CH) bulk_g2s(sbase + off, image + off, (EIMAGE_B - off) < CH ? (EIMAGE_B - off) : CH, bar0 + 8 * EB_W);
        }
        __syncwarp();
        mbar_wait(bar0 + 8 * EB_W, 0);
        const uint64_t w0h = smem_desc(sbase + EOFF_W0, 32 * 16, 128), w0l = smem_desc(sbase + EPLANE_B + EOFF_W0, 32 * 16, 128);
        const uint64_t w1h = smem_desc(sbase + EOFF_W1, 64 * 16, 128), w1l = smem_desc(sbase + EPLANE_B + EOFF_W1, 64 * 16, 128);
        const uint64_t w2h = smem_desc(sbase + EOFF_W2, 256 * 16, 128), w2l = smem_desc(sbase + EPLANE_B + EOFF_W2, 256 * 16, 128);
        const uint64_t w3h = smem_desc(sbase + EOFF_W3, 32 * 16, 128), w3l = smem_desc(sbase + EPLANE_B + EOFF_W3, 32 * 16, 128);
        uint32_t ph_x = 0, ph_a = 0, ph_e = 0;
        int64_t it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int buf = (int)(it & 1);
            // ---- L0: A = x tile (smem), one K=16 step, N = 32
            mbar_wait_tight(bar0 + 8 * (EB_X0 + buf), (ph_x >> buf) & 1); ph_x ^= 1u << buf;
            if (it > 0) { mbar_wait_tight(bar0 + 8 * EB_E, ph_e); ph_e ^= 1; }       // previous tile's outputs were read
            tc_fence_after();
            if (elect_one()) {
                const uint64_t xh = smem_desc(sbase + EOFF_X + buf * 2 * EX_PLANE_B, EX_CHUNK_B, 128);
                const uint64_t xl = smem_desc(sbase + EOFF_X + buf * 2 * EX_PLANE_B + EX_PLANE_B, EX_CHUNK_B, 128);
                const uint32_t id = idesc_f16(32);
                mma_ss(acc_c, xh, w0h, id, 0); mma_ss(acc_c, xl, w0h, id, 1); mma_ss(acc_c, xh, w0l, id, 1);
                mma_commit(bar0 + 8 * EB_ACC);
                mma_commit(bar0 + 8 * (EB_XF0 + buf));
            }
            __syncwarp();
            // ---- L1 (K = 32, N = 64), L2 (K = 64, N = 256), L3 (K = 256, N = 32): A from TMEM
#pragma unroll 1
            for (int layer = 1; layer < 4; ++layer) {
                mbar_wait_tight(bar0 + 8 * EB_A, ph_a); ph_a ^= 1;
                tc_fence_after();
                if (elect_one()) {
                    if (layer == 1) issue_ts(idesc_f16(64), (2 * 64 * 16) >> 4, 2, acc_c, a_c, a_c + 16, w1h, w1l);
                    else if (layer == 2) issue_ts(idesc_f16(256), (2 * 256 * 16) >> 4, 4, acc_c, a_c + 32, a_c + 64, w2h, w2l);
                    else issue_ts(idesc_f16(32), (2 * 32 * 16) >> 4, 16, acc_c, a_c, a_c + 128, w3h, w3l);
                    mma_commit(bar0 + 8 * EB_ACC);
                }
                __syncwarp();
            }
        }
    } else if (warp >= PRODUCER_WARP0) {
        // ===================================================== gather producers: warp 17 rows 0..63, warp 18 rows 64..127
        const int pw = warp - PRODUCER_WARP0;
        if (pw < 2) {
            uint32_t ph_xf = 0;
            int64_t it = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const int buf = (int)(it & 1);
                if (it >= 2) { mbar_wait(bar0 + 8 * (EB_XF0 + buf), (ph_xf >> buf) & 1); ph_xf ^= 1u << buf; }   // layer 0 of tile it-2 has read this buffer
                unsigned char* x_hi_p = smem + EOFF_X + buf * 2 * EX_PLANE_B;
                int32_t* slot_p = reinterpret_cast<int32_t*>(smem + EOFF_SLOT) + (int)(it & 3) * TILE;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int row = pw * 64 + h * 32 + lane;
                    const int64_t si = tile * TILE + row;
                    float in[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    int slot = -1;
                    if (si < n_total) {
                        if (a.mode == 0) {
                            const int i = a.s_pt[si], k = a.s_off[si];
                            slot = a.s_slot[si];
                            const float px = a.p_hat[3 * i], py = a.p_hat[3 * i + 1], pz = a.p_hat[3 * i + 2];
                            const float ox = (k & 4) ? 0.5f : -0.5f, oy = (k & 2) ? 0.5f : -0.5f, oz = (k & 1) ? 0.5f : -0.5f;
                            const float cx = (float)clampi((int)ceilf(__fadd_rn(px, ox)) - 1, 0, a.g.nx - 1);
                            const float cy = (float)clampi((int)ceilf(__fadd_rn(py, oy)) - 1, 0, a.g.ny - 1);
                            const float cz = (float)clampi((int)ceilf(__fadd_rn(pz, oz)) - 1, 0, a.g.nz - 1);
                            // rel = p - cell - 0.5, two separately rounded subtractions as in map.py:425
                            in[0] = __fsub_rn(__fsub_rn(px, cx), 0.5f); in[1] = __fsub_rn(__fsub_rn(py, cy), 0.5f); in[2] = __fsub_rn(__fsub_rn(pz, cz), 0.5f);
                            const float* np_ = a.normal + (int64_t)a.normal_stride * i;
                            in[3] = np_[0]; in[4] = np_[1]; in[5] = np_[2];
                        } else {
                            slot = 0;
#pragma unroll
                            for (int j = 0; j < 6; ++j) in[j] = __ldg(a.xyzn + si * 6 + j);
                        }
                    }
                    uint4 hq, lq;
                    split_pair(in[0], in[1], hq.x, lq.x); split_pair(in[2], in[3], hq.y, lq.y);
                    split_pair(in[4], in[5], hq.z, lq.z); hq.w = 0u; lq.w = 0u;
                    *reinterpret_cast<uint4*>(x_hi_p + row * 16) = hq;
                    *reinterpret_cast<uint4*>(x_hi_p + EX_PLANE_B + row * 16) = lq;
                    slot_p[row] = slot;
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar0 + 8 * (EB_X0 + buf));
            }
        }
    } else {
        // ===================================================== epilogue warps: lanes 32*quad.., column quarter cq
        const int quad = warp & 3, cq = warp >> 2;
        const int row = quad * 32 + lane;
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        const uint32_t acc = acc_c + lane_base, ar = a_c + lane_base;
        const float* bias = reinterpret_cast<const float*>(smem + EOFF_BIAS);
        uint32_t ph_acc = 0;
        mbar_wait(bar0 + 8 * EB_W, 0);
        int64_t it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            // ---- E0: 32 columns, 8 per warp -> h0 (K = 32): hi @ A+0..15, lo @ A+16..31
            mbar_wait(bar0 + 8 * EB_ACC, ph_acc); ph_acc ^= 1;
            tc_fence_after();
            const int slot = reinterpret_cast<const int32_t*>(smem + EOFF_SLOT)[(int)(it & 3) * TILE + row];     // (written before the x tile was published)
            {
                uint32_t v[8], hi[4], lo[4];
                tmem_ld8_nowait(acc + 8 * cq, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    split_pair(fmaxf(__uint_as_float(v[2 * j]) + bias[8 * cq + 2 * j], 0.f), fmaxf(__uint_as_float(v[2 * j + 1]) + bias[8 * cq + 2 * j + 1], 0.f), hi[j], lo[j]);
                tmem_st4(ar + 4 * cq, hi); tmem_st4(ar + 16 + 4 * cq, lo);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + 8 * EB_A);
            // ---- E1: 64 columns, 16 per warp -> h1 (K = 64): hi @ A+32..63, lo @ A+64..95
            mbar_wait(bar0 + 8 * EB_ACC, ph_acc); ph_acc ^= 1;
            tc_fence_after();
            {
                uint32_t v[16];
                tmem_ld16_nowait(acc + 16 * cq, v);
                tmem_ld_wait();
                convert16(v, bias + 32 + 16 * cq, ar + 32 + 8 * cq, ar + 64 + 8 * cq);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + 8 * EB_A);
            // ---- E2: 256 columns, 64 per warp -> h2 (K = 256): hi @ A+0..127, lo @ A+128..255
            mbar_wait(bar0 + 8 * EB_ACC, ph_acc); ph_acc ^= 1;
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < 64; c += 32) {
                uint32_t v0[16], v1[16];
                const int col = 64 * cq + c;
                tmem_ld16_nowait(acc + col, v0);
                tmem_ld16_nowait(acc + col + 16, v1);
                tmem_ld_wait();
                convert16(v0, bias + 96 + col, ar + col / 2, ar + 128 + col / 2);
                convert16(v1, bias + 96 + col + 16, ar + col / 2 + 8, ar + 128 + col / 2 + 8);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + 8 * EB_A);
            // ---- E3: 32 columns (29 real), 8 per warp: + b3, then accumulate into the target PLIVox (or store)
            mbar_wait(bar0 + 8 * EB_ACC, ph_acc); ph_acc ^= 1;
            tc_fence_after();
            uint32_t v[8];
            tmem_ld8_nowait(acc + 8 * cq, v);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + 8 * EB_E);               // accumulator consumed: the next tile's layer 0 may start
            const int64_t si = tile * TILE + row;
            if (slot >= 0 && si < n_total) {
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = __uint_as_float(v[j]) + bias[352 + 8 * cq + j];      // (padding columns 29..31 are exactly 0)
                if (a.mode == 0) {
                    // two 16-byte vector reductions per thread into the 128-byte-strided sum row (4x fewer L2 atomic operations)
                    float4* dst = reinterpret_cast<float4*>(a.slot_sum + (int64_t)slot * DIF_SUM_STRIDE + 8 * cq);
                    atomicAdd(dst, make_float4(o[0], o[1], o[2], o[3]));
                    atomicAdd(dst + 1, make_float4(o[4], o[5], o[6], o[7]));
                } else {
                    float* dst = a.out + si * DIF_L;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (8 * cq + j < DIF_L) dst[8 * cq + j] = o[j];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(512));
}

// =====================================================================================================================
// Second pipeline (default): TWO tiles in flight, in-place activation conversion, group hand-off (see decode_tc2.cuh for the
// measurements behind the scheme).  The first kernel keeps ONE tile in flight because L2's 256-column accumulator plus its
// 256-column A operand fill the tensor memory; the tensor pipe then idles through every epilogue (ncu: 19.5 % active).
// Here L2 is computed in two N = 128 halves and L3 (K = 256) in the two matching K halves, so a tile needs 224 columns:
//   per slot (256 columns):  A1 = [0, 64)   h1, the A operand of L2 (acc of L1, converted in place)
//                            P  = [64, 192) acc of L0 (32 cols) / L2a / L2b, converted in place into the A operand of L1 / L3a / L3b
//                            Q  = [192, 224) acc of L3 (29 real outputs)
//   L0  x (smem, K = 16) -> P[0,32)        E0 in place (2 K chunks)          L1  P[0,32) -> A1 (N = 64)   E1 in place (4 K chunks)
//   L2a A1 -> P (N = 128, rows 0..127 of W2)   E2a in place, four 32-column groups      L3a P -> Q (K steps 0..7 of W3), issued per group pair
//   L2b A1 -> P (rows 128..255), issued right behind L3a (in-order pipe: P is read before it is overwritten)   E2b   L3b (K steps 8..15) accumulates into Q
//   E3  Q + b3 -> float4 atomics into the per-PLIVox sum rows (or plain stores)
// One issuer warp alternates strictly between the two slots stage by stage (anti-phase: one slot converts while the other's MMAs run);
// 8 epilogue warps per slot (lane quadrant x column half).  Producers and x / slot-id buffers are those of the first kernel
// (buffer = tile parity = slot).
constexpr uint32_t E2OFF_BAR = EOFF_SLOT + 4 * TILE * 4;
constexpr uint32_t E2SMEM_B = E2OFF_BAR + 272 + 16;
static_assert(E2SMEM_B <= 232448, "shared memory layout");
// barriers: 0 W; per slot s (index + s): X 1, XF 3, ACC0 5, ACC1 7, ACC2 9, ACC3 11, A0 13, A1 15, E 17; G 19 + 4 s + g  (.. 26)
enum { E2_W = 0, E2_X = 1, E2_XF = 3, E2_ACC0 = 5, E2_ACC1 = 7, E2_ACC2 = 9, E2_ACC3 = 11, E2_A0 = 13, E2_A1 = 15, E2_E = 17, E2_G = 19, E2_W2 = 27, E2_W3 = 28 };

__global__ void __launch_bounds__(THREADS, 1) encode_tc2_kernel(const unsigned char* __restrict__ image, EncodeArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = sbase + E2OFF_BAR;
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + E2OFF_BAR + 272);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_wait(); pdl_launch_dependents();                          // the sample count below is written by the previous kernel
    const int64_t n_total = a.mode == 0 ? (int64_t)*a.n_dev : a.n;
    const int64_t n_tiles = (n_total + TILE - 1) / TILE;
    if ((int64_t)blockIdx.x >= n_tiles) return;                   // nothing to do: skip the weight load altogether

    if (threadIdx.x == 0) {
        mbar_init(bar0 + 8 * E2_W, 1); mbar_init(bar0 + 8 * E2_W2, 1); mbar_init(bar0 + 8 * E2_W3, 1);
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar0 + 8 * (E2_X + s), 2);                  // two producer warps fill one x buffer
            mbar_init(bar0 + 8 * (E2_XF + s), 1);                 // tcgen05.commit after layer 0: buffer free
            mbar_init(bar0 + 8 * (E2_ACC0 + s), 1); mbar_init(bar0 + 8 * (E2_ACC1 + s), 1);
            mbar_init(bar0 + 8 * (E2_ACC2 + s), 1); mbar_init(bar0 + 8 * (E2_ACC3 + s), 1);
            mbar_init(bar0 + 8 * (E2_A0 + s), 8); mbar_init(bar0 + 8 * (E2_A1 + s), 8); mbar_init(bar0 + 8 * (E2_E + s), 8);
#pragma unroll
            for (int g = 0; g < 4; ++g) mbar_init(bar0 + 8 * (E2_G + 4 * s + g), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // K chunk 1 of every x plane is identically zero (the encoder has 6 inputs, padded to K = 16): written once
    for (int i = threadIdx.x; i < 4 * (int)(EX_CHUNK_B / 16); i += THREADS)
        *reinterpret_cast<uint4*>(smem + EOFF_X + (i / (EX_CHUNK_B / 16)) * EX_PLANE_B + EX_CHUNK_B + (i % (EX_CHUNK_B / 16)) * 16) = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_ptr_s)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_ptr_s, 0);

    if (warp == MMA_WARP) {
        // ===================================================== weight load + MMA issuer (warp-uniform, instructions elected)
        // three weight groups with their own barriers: biases + W0 + W1 (11.5 KB: L0 and L1 start on them), W2 (64 KB), W3 (32 KB)
        if (lane == 0) {
            mbar_expect_tx(bar0 + 8 * E2_W, EBIAS_B + 2 * (EW0_B + EW1_B));
            bulk_g2s(sbase + EOFF_BIAS, image + EOFF_BIAS, EBIAS_B, bar0 + 8 * E2_W);
            bulk_g2s(sbase + EOFF_W0, image + EOFF_W0, EW0_B + EW1_B, bar0 + 8 * E2_W);
            bulk_g2s(sbase + EPLANE_B + EOFF_W0, image + EPLANE_B + EOFF_W0, EW0_B + EW1_B, bar0 + 8 * E2_W);
            mbar_expect_tx(bar0 + 8 * E2_W2, 2 * EW2_B);
            bulk_g2s(sbase + EOFF_W2, image + EOFF_W2, EW2_B, bar0 + 8 * E2_W2);
            bulk_g2s(sbase + EPLANE_B + EOFF_W2, image + EPLANE_B + EOFF_W2, EW2_B, bar0 + 8 * E2_W2);
            mbar_expect_tx(bar0 + 8 * E2_W3, 2 * EW3_B);
            bulk_g2s(sbase + EOFF_W3, image + EOFF_W3, EW3_B, bar0 + 8 * E2_W3);
            bulk_g2s(sbase + EPLANE_B + EOFF_W3, image + EPLANE_B + EOFF_W3, EW3_B, bar0 + 8 * E2_W3);
        }
        __syncwarp();
        mbar_wait(bar0 + 8 * E2_W, 0);
        const uint64_t w0h = smem_desc(sbase + EOFF_W0, 32 * 16, 128), w0l = smem_desc(sbase + EPLANE_B + EOFF_W0, 32 * 16, 128);
        const uint64_t w1h = smem_desc(sbase + EOFF_W1, 64 * 16, 128), w1l = smem_desc(sbase + EPLANE_B + EOFF_W1, 64 * 16, 128);
        const uint64_t w2h = smem_desc(sbase + EOFF_W2, 256 * 16, 128), w2l = smem_desc(sbase + EPLANE_B + EOFF_W2, 256 * 16, 128);
        const uint64_t w3h = smem_desc(sbase + EOFF_W3, 32 * 16, 128), w3l = smem_desc(sbase + EPLANE_B + EOFF_W3, 32 * 16, 128);
        const uint64_t w2bh = desc_advance(w2h, 128 * 16), w2bl = desc_advance(w2l, 128 * 16);                 // rows 128..255 of W2
        const uint64_t w3bh = desc_advance(w3h, 8 * 2 * 32 * 16), w3bl = desc_advance(w3l, 8 * 2 * 32 * 16);   // K steps 8..15 of W3
        uint32_t ph_t = 0;                                  // parity of everything that completes once per tile (both slots in step)
        // L2 with N = 128 on a 256-row weight slab: the 8-wide k chunks are 256 * 16 bytes apart (LBO of the descriptor), the K = 16 step 2 * that
        constexpr uint32_t W2STEP = (2 * 256 * 16) >> 4;
        auto issue_l2 = [&](uint32_t acc, uint32_t a1, uint64_t wh, uint64_t wl) {        // 4 K steps x 3 passes, A operand = A1 chunks 0..3
            constexpr uint32_t id = idesc_f16(128);
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
                const uint32_t lo = pass == 1 ? 8u : 0u;
                const uint64_t w = pass == 2 ? wl : wh;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) mma_ts(acc, a1 + 16 * ks + lo, w + (uint64_t)(ks * W2STEP), id, (pass | ks) ? 1u : 0u);
            }
        };
        for (int64_t it = 0;; ++it) {
            const int64_t t0 = blockIdx.x + (int64_t)gridDim.x * (2 * it), t1 = t0 + gridDim.x;
            if (t0 >= n_tiles) break;
            const int nslots = t1 < n_tiles ? 2 : 1;
            // ---- L0: x tile (smem, one K = 16 step, N = 32) -> P[0, 32)
#pragma unroll 1
            for (int s = 0; s < nslots; ++s) {
                mbar_wait_spin(bar0 + 8 * (E2_X + s), ph_t);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t xh = smem_desc(sbase + EOFF_X + s * 2 * EX_PLANE_B, EX_CHUNK_B, 128);
                    const uint64_t xl = smem_desc(sbase + EOFF_X + s * 2 * EX_PLANE_B + EX_PLANE_B, EX_CHUNK_B, 128);
                    const uint32_t P = tmem + 256 * s + 64, id = idesc_f16(32);
                    mma_ss(P, xh, w0h, id, 0); mma_ss(P, xl, w0h, id, 1); mma_ss(P, xh, w0l, id, 1);
                    mma_commit(bar0 + 8 * (E2_ACC0 + s));
                    mma_commit(bar0 + 8 * (E2_XF + s));
                }
                __syncwarp();
            }
            // ---- L1: P[0, 32) (2 K chunks) -> A1 (N = 64)
#pragma unroll 1
            for (int s = 0; s < nslots; ++s) {
                mbar_wait_spin(bar0 + 8 * (E2_A0 + s), ph_t);
                tc_fence_after();
                if (elect_one()) { issue_group<64, 0, 1, true>(tmem + 256 * s, tmem + 256 * s + 64, w1h, w1l); mma_commit(bar0 + 8 * (E2_ACC1 + s)); }
                __syncwarp();
            }
            // ---- L2a: A1 (4 K chunks) -> P, rows 0..127 of W2
            if (it == 0) mbar_wait_spin(bar0 + 8 * E2_W2, 0);
#pragma unroll 1
            for (int s = 0; s < nslots; ++s) {
                mbar_wait_spin(bar0 + 8 * (E2_A1 + s), ph_t);
                tc_fence_after();
                if (elect_one()) { issue_l2(tmem + 256 * s + 64, tmem + 256 * s, w2h, w2l); mma_commit(bar0 + 8 * (E2_ACC2 + s)); }
                __syncwarp();
            }
            // ---- L3a: P (8 K chunks, two group pairs) -> Q, then L2b right behind it
            if (it == 0) mbar_wait_spin(bar0 + 8 * E2_W3, 0);
#pragma unroll 1
            for (int s = 0; s < nslots; ++s) {
                const uint32_t A1 = tmem + 256 * s, P = A1 + 64, Q = A1 + 192, bg = bar0 + 8 * (E2_G + 4 * s);
                mbar_wait_spin(bg, 0); mbar_wait_spin(bg + 8, 0);
                if (it > 0) mbar_wait_spin(bar0 + 8 * (E2_E + s), ph_t ^ 1u);        // the previous tile's outputs have been read from Q
                tc_fence_after();
                if (elect_one()) { issue_group<32, 0, 1, true>(Q, P, w3h, w3l); issue_group<32, 4, 5, false>(Q, P, w3h, w3l); }
                __syncwarp();
                mbar_wait_spin(bg + 16, 0); mbar_wait_spin(bg + 24, 0);
                tc_fence_after();
                if (elect_one()) {
                    issue_group<32, 2, 3, false>(Q, P, w3h, w3l); issue_group<32, 6, 7, false>(Q, P, w3h, w3l);
                    issue_l2(P, A1, w2bh, w2bl);                                     // L2b: rows 128..255 of W2 (P is read by the MMAs above first)
                    mma_commit(bar0 + 8 * (E2_ACC2 + s));
                }
                __syncwarp();
            }
            // ---- L3b: K steps 8..15 of W3 accumulate into Q
#pragma unroll 1
            for (int s = 0; s < nslots; ++s) {
                const uint32_t P = tmem + 256 * s + 64, Q = tmem + 256 * s + 192, bg = bar0 + 8 * (E2_G + 4 * s);
                mbar_wait_spin(bg, 1); mbar_wait_spin(bg + 8, 1);
                tc_fence_after();
                if (elect_one()) { issue_group<32, 0, 1, false>(Q, P, w3bh, w3bl); issue_group<32, 4, 5, false>(Q, P, w3bh, w3bl); }
                __syncwarp();
                mbar_wait_spin(bg + 16, 1); mbar_wait_spin(bg + 24, 1);
                tc_fence_after();
                if (elect_one()) {
                    issue_group<32, 2, 3, false>(Q, P, w3bh, w3bl); issue_group<32, 6, 7, false>(Q, P, w3bh, w3bl);
                    mma_commit(bar0 + 8 * (E2_ACC3 + s));
                }
                __syncwarp();
            }
            ph_t ^= 1;
        }
    } else if (warp >= PRODUCER_WARP0) {
        // ===================================================== gather producers: warp 17 rows 0..63, warp 18 rows 64..127 of every tile
        const int pw = warp - PRODUCER_WARP0;
        if (pw < 2) {
            uint32_t ph_xf = 0;
            int64_t it = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const int buf = (int)(it & 1);                     // == the slot the tile runs in
                if (it >= 2) { mbar_wait(bar0 + 8 * (E2_XF + buf), (ph_xf >> buf) & 1); ph_xf ^= 1u << buf; }   // layer 0 of tile it-2 has read this buffer
                unsigned char* x_hi_p = smem + EOFF_X + buf * 2 * EX_PLANE_B;
                int32_t* slot_p = reinterpret_cast<int32_t*>(smem + EOFF_SLOT) + (int)(it & 3) * TILE;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int row = pw * 64 + h * 32 + lane;
                    const int64_t si = tile * TILE + row;
                    float in[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    int slot = -1;
                    if (si < n_total) {
                        if (a.mode == 0) {
                            const int i = a.s_pt[si], k = a.s_off[si];
                            slot = a.s_slot[si];
                            const float px = a.p_hat[3 * i], py = a.p_hat[3 * i + 1], pz = a.p_hat[3 * i + 2];
                            const float ox = (k & 4) ? 0.5f : -0.5f, oy = (k & 2) ? 0.5f : -0.5f, oz = (k & 1) ? 0.5f : -0.5f;
                            const float cx = (float)clampi((int)ceilf(__fadd_rn(px, ox)) - 1, 0, a.g.nx - 1);
                            const float cy = (float)clampi((int)ceilf(__fadd_rn(py, oy)) - 1, 0, a.g.ny - 1);
                            const float cz = (float)clampi((int)ceilf(__fadd_rn(pz, oz)) - 1, 0, a.g.nz - 1);
                            // rel = p - cell - 0.5, two separately rounded subtractions as in map.py:425
                            in[0] = __fsub_rn(__fsub_rn(px, cx), 0.5f); in[1] = __fsub_rn(__fsub_rn(py, cy), 0.5f); in[2] = __fsub_rn(__fsub_rn(pz, cz), 0.5f);
                            const float* np_ = a.normal + (int64_t)a.normal_stride * i;
                            in[3] = np_[0]; in[4] = np_[1]; in[5] = np_[2];
                        } else {
                            slot = 0;
#pragma unroll
                            for (int j = 0; j < 6; ++j) in[j] = __ldg(a.xyzn + si * 6 + j);
                        }
                    }
                    uint4 hq, lq;
                    split_pair(in[0], in[1], hq.x, lq.x); split_pair(in[2], in[3], hq.y, lq.y);
                    split_pair(in[4], in[5], hq.z, lq.z); hq.w = 0u; lq.w = 0u;
                    *reinterpret_cast<uint4*>(x_hi_p + row * 16) = hq;
                    *reinterpret_cast<uint4*>(x_hi_p + EX_PLANE_B + row * 16) = lq;
                    slot_p[row] = slot;
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar0 + 8 * (E2_X + buf));
            }
        }
    } else {
        // ===================================================== epilogue warps of slot s: lane quadrant x column half
        const int s = warp >> 3, quad = warp & 3, half = (warp >> 2) & 1;
        const int row = quad * 32 + lane;
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        const uint32_t A1 = tmem + 256 * s + lane_base, P = A1 + 64, Q = A1 + 192;
        const float* bias = reinterpret_cast<const float*>(smem + EOFF_BIAS);          // b0 @0 (32), b1 @32 (64), b2 @96 (256), b3 @352 (32)
        const uint32_t g_bar = bar0 + 8 * (E2_G + 4 * s + half);
        uint32_t ph_t = 0;
        auto hand_off = [&](uint32_t bar) {
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar);
        };
        mbar_wait(bar0 + 8 * E2_W, 0);
        for (int64_t it = 0;; ++it) {
            const int64_t j = 2 * it + s;                          // the CTA's j-th tile
            const int64_t tile = blockIdx.x + (int64_t)gridDim.x * j;
            if (tile >= n_tiles) break;
            // ---- E0: P[0, 32): one K chunk per column half -> A operand of L1
            mbar_wait(bar0 + 8 * (E2_ACC0 + s), ph_t);
            tc_fence_after();
            const int slot = reinterpret_cast<const int32_t*>(smem + EOFF_SLOT)[(int)(j & 3) * TILE + row];     // (written before the x tile was published)
            {
                uint32_t v[16];
                tmem_ld16_nowait(P + 16 * half, v);
                tmem_ld_wait();
                convert_inplace16(v, bias + 16 * half, P + 16 * half);
            }
            hand_off(bar0 + 8 * (E2_A0 + s));
            // ---- E1: A1[0, 64): two K chunks per half -> A operand of L2
            mbar_wait(bar0 + 8 * (E2_ACC1 + s), ph_t);
            tc_fence_after();
            {
                uint32_t v0[16], v1[16];
                tmem_ld16_nowait(A1 + 32 * half, v0);
                tmem_ld16_nowait(A1 + 32 * half + 16, v1);
                tmem_ld_wait();
                convert_inplace16(v0, bias + 32 + 32 * half, A1 + 32 * half);
                convert_inplace16(v1, bias + 32 + 32 * half + 16, A1 + 32 * half + 16);
            }
            hand_off(bar0 + 8 * (E2_A1 + s));
            // ---- E2a / E2b: P[0, 128): four K chunks per half, handed off in two groups -> A operand of L3a / L3b
#pragma unroll 1
            for (int part = 0; part < 2; ++part) {
                mbar_wait(bar0 + 8 * (E2_ACC2 + s), (uint32_t)part);
                tc_fence_after();
                const float* b = bias + 96 + 128 * part + 64 * half;
#pragma unroll 1
                for (int i = 0; i < 2; ++i) {
                    uint32_t v0[16], v1[16];
                    const int c = 64 * half + 32 * i;
                    tmem_ld16_nowait(P + c, v0);
                    tmem_ld16_nowait(P + c + 16, v1);
                    tmem_ld_wait();
                    convert_inplace16(v0, b + 32 * i, P + c);
                    convert_inplace16(v1, b + 32 * i + 16, P + c + 16);
                    hand_off(g_bar + 16 * i);
                }
            }
            // ---- E3: Q[0, 32) (29 real outputs): + b3, accumulate into the target PLIVox (or store)
            mbar_wait(bar0 + 8 * (E2_ACC3 + s), ph_t);
            tc_fence_after();
            uint32_t v[16];
            tmem_ld16_nowait(Q + 16 * half, v);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + 8 * (E2_E + s));               // Q consumed: the next tile's L3a may start
            const int64_t si = tile * TILE + row;
            if (slot >= 0 && si < n_total) {
                float o[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) o[q] = __uint_as_float(v[q]) + bias[352 + 16 * half + q];      // (padding columns 29..31 are exactly 0)
                if (a.mode == 0) {
                    // 16-byte vector reductions into the 128-byte-strided sum row (4x fewer L2 atomic operations than scalar ones)
                    float4* dst = reinterpret_cast<float4*>(a.slot_sum + (int64_t)slot * DIF_SUM_STRIDE + 16 * half);
#pragma unroll
                    for (int q = 0; q < 4; ++q) atomicAdd(dst + q, make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]));
                } else {
                    float* dst = a.out + si * DIF_L;
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        if (16 * half + q < DIF_L) dst[16 * half + q] = o[q];
                }
            }
            ph_t ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(512));
}

// ---- weight image: fp16 hi/lo planes, no-swizzle K-major core-matrix slabs (element (n,k) at (k/8)*(N*16) + n*16 + (k%8)*2) ----
__global__ void prepare_encoder_tc_kernel(const float* __restrict__ P, unsigned char* __restrict__ image) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    // source: k-major fp32 copies of mlp_simt.cuh (EncW::W?t[k][n]); K of layer 0 padded 6 -> 16, N of layer 3 padded 29 -> 32
    const int srcs[4] = {EncW::W0t, EncW::W1t, EncW::W2t, EncW::W3t};
    const int Ns[4] = {32, 64, 256, 32}, Ks[4] = {16, 32, 64, 256}, Kreal[4] = {6, 32, 64, 256};
    const uint32_t offs[4] = {EOFF_W0, EOFF_W1, EOFF_W2, EOFF_W3};
    for (int l = 0; l < 4; ++l) {
        const int N = Ns[l], K = Ks[l];
        for (int i = tid; i < N * K; i += nth) {
            const int n = i / K, k = i % K;
            const float w = k < Kreal[l] ? P[srcs[l] + k * N + n] : 0.f;          // (W3t already has zero columns 29..31)
            const __half h = __float2half_rn(w);
            const __half lo = __float2half_rn(w - __half2float(h));
            const uint32_t o = offs[l] + (uint32_t)(k / 8) * (N * 16) + n * 16 + (k % 8) * 2;
            *reinterpret_cast<__half*>(image + o) = h;
            *reinterpret_cast<__half*>(image + EPLANE_B + o) = lo;
        }
    }
    float* b = reinterpret_cast<float*>(image + EOFF_BIAS);
    for (int i = tid; i < 256; i += nth) {
        b[96 + i] = P[EncW::b2 + i];
        if (i < 32) { b[i] = P[EncW::b0 + i]; b[352 + i] = i < DIF_L ? P[EncW::b3 + i] : 0.f; }
        if (i < 64) b[32 + i] = P[EncW::b1 + i];
    }
}

}  // namespace enc

size_t encoder_tc_image_bytes() { return enc::EIMAGE_B; }

int prepare_encoder_tc(const float* P, unsigned char* image, cudaStream_t st) {
    enc::prepare_encoder_tc_kernel<<<64, 256, 0, st>>>(P, image);
    DIF_COUNT_LAUNCH(1);
    return check_launch("prepare_encoder_tc_kernel");
}

static int launch(const unsigned char* image, const enc::EncodeArgs& a, int64_t max_samples, cudaStream_t st) {
    const int64_t max_tiles = (max_samples + tc::TILE - 1) / tc::TILE;
    const int grid = (int)(max_tiles < DIF_NUM_SMS ? (max_tiles > 0 ? max_tiles : 1) : DIF_NUM_SMS);
    static bool attr_set = false;        // once per process: keeps the launch path free of non-stream API calls (CUDA-graph capture)
    if (!attr_set) {
        cudaFuncSetAttribute(enc::encode_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)enc::ESMEM_B);
        cudaFuncSetAttribute(enc::encode_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)enc::E2SMEM_B);
        attr_set = true;
    }
    const char* v = getenv("DIF_ENCODE_V");                 // "1": the first pipeline (one tile in flight), kept for A/B timing
    prof_begin(DIF_PROF_ENCODE, st);
    if (v && v[0] == '1') launch_pdl(enc::encode_tc_kernel, grid, tc::THREADS, enc::ESMEM_B, st, image, a);
    else launch_pdl(enc::encode_tc2_kernel, grid, tc::THREADS, enc::E2SMEM_B, st, image, a);
    prof_end(DIF_PROF_ENCODE, st);
    DIF_COUNT_LAUNCH(1);
    return check_launch("encode_tc_kernel");
}

// integrate path: samples from the gather list (device-side count), accumulate into slot_sum
int launch_encode_accumulate_tc(const void* encoder_prepared, Grid g, const float* p_hat, const float* normal, int normal_stride, const int32_t* s_pt,
                                const int32_t* s_slot, const uint8_t* s_off, const int32_t* n_dev, int64_t max_samples, float* slot_sum,
                                cudaStream_t st) {
    const unsigned char* image = (const unsigned char*)encoder_prepared + (size_t)EncW::FP32_END * sizeof(float);
    enc::EncodeArgs a{0, g, p_hat, normal, normal_stride, s_pt, s_slot, s_off, n_dev, slot_sum, nullptr, 0, nullptr};
    return launch(image, a, max_samples, st);
}

// dif_encode path: explicit (n, 6) inputs -> (n, 29) outputs
int launch_encode_tc(const void* encoder_prepared, const float* xyzn, int64_t n, float* out, cudaStream_t st) {
    const unsigned char* image = (const unsigned char*)encoder_prepared + (size_t)EncW::FP32_END * sizeof(float);
    enc::EncodeArgs a{1, Grid{}, nullptr, nullptr, 3, nullptr, nullptr, nullptr, nullptr, nullptr, xyzn, n, out};
    return launch(image, a, n, st);
}

}  // namespace dif
