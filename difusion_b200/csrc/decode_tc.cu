// Tensor-core (tcgen05 / TMEM) forward evaluation of the DI-Fusion decoder MLP for sm_100a.
//   replaces reference network/di_decoder.py:55-86 evaluated through network/utility.py:61-126 (cuBLAS SGEMM chain)
//   SURVEY rows a-7, a-8 (forward), a-10;  precision scheme: SURVEY 7 "Hard parts" (3-pass fp16 split).
//
// One persistent CTA per SM.  A tile is 128 samples (= the 128 TMEM lanes, one sample per epilogue thread).  All four
// hidden layers run as tcgen05.mma M=128 x N=128(96) x K=16 instructions with fp32 accumulation in TMEM:
//
//   * weights:  fp16 hi + fp16 lo planes of every layer stay resident in shared memory for the life of the CTA
//               (192 KB, no-swizzle K-major core-matrix layout, loaded once with cp.async.bulk / UBLKCP);
//   * inputs :  the (latent, xyz) gather is fused into the A-operand stage: each epilogue thread gathers its sample's
//               29+3 inputs, splits them into fp16 hi/lo and writes the layer-0 A tile straight into shared memory;
//   * hidden :  activations never touch shared or global memory: TMEM accumulator -> registers (tcgen05.ld) ->
//               +bias, ReLU, hi/lo split -> packed fp16 A operand back into TMEM (tcgen05.st) -> next layer's MMA reads
//               A from TMEM;  the skip connection re-reads the layer-0 A tile from shared memory;
//   * 3 passes: D += A_hi*W_hi + A_lo*W_hi + A_hi*W_lo  (the dropped lo*lo term is ~2^-21 relative) keeps the result
//               within ~2e-6 of the fp32 reference, far inside the 1e-4 parity tolerance that single-pass fp16/bf16 fails;
//   * overlap : two tiles are in flight per CTA (TMEM columns 0-255 / 256-511); one elected thread issues all MMAs and
//               alternates between the two tiles layer by layer, so the tensor pipe runs tile B's layer while the
//               4 epilogue warps of tile A convert its accumulator (mbarrier hand-offs, tcgen05.commit).
//
// Warp roles: warps 0-3 = epilogue/gather warpgroup of slot 0, warps 4-7 = slot 1, warp 8 = TMEM allocator + MMA issuer.
#include <cuda_fp16.h>

#include "decode_args.cuh"
#include "mlp_simt.cuh"

namespace dif {

namespace tc {

constexpr int TILE = 128;
constexpr int THREADS = 9 * 32;

// ---- byte layout of the tensor-core section of the prepared decoder buffer == its image in shared memory ----------
constexpr uint32_t W0_B = 128 * 32 * 2, W1_B = 128 * 128 * 2, W2_B = 96 * 128 * 2, W3_B = 128 * 128 * 2;
constexpr uint32_t OFF_W0 = 0, OFF_W1 = OFF_W0 + W0_B, OFF_W2 = OFF_W1 + W1_B, OFF_W3 = OFF_W2 + W2_B;
constexpr uint32_t PLANE_B = OFF_W3 + W3_B;                 // 98304: one precision plane (hi or lo)
constexpr uint32_t OFF_BIAS = 2 * PLANE_B;                  // b0[128] b1[128] b2[96] b3[128] fp32
constexpr uint32_t BIAS_B = 480 * 4;
constexpr uint32_t IMAGE_B = OFF_BIAS + BIAS_B;             // 198528 bytes copied global -> shared per CTA
constexpr uint32_t OFF_X = IMAGE_B;                         // per slot: hi [4][128][8] halves, lo [4][128][8] halves
constexpr uint32_t X_PLANE_B = 4 * TILE * 16;               // 8192
constexpr uint32_t OFF_BAR = OFF_X + 4 * X_PLANE_B;         // 231296
constexpr uint32_t SMEM_B = OFF_BAR + 64 + 16;
static_assert(SMEM_B <= 232448, "shared memory budget");
static_assert(IMAGE_B % 16 == 0, "bulk copy granularity");

enum { BAR_W = 0, BAR_X0 = 1, BAR_X1 = 2, BAR_ACC0 = 3, BAR_ACC1 = 4, BAR_A0 = 5, BAR_A1 = 6 };

// ---- PTX wrappers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor, no swizzle, K-major: core matrix = 8 rows x 16 B contiguous;
// LBO = byte distance between the two 8-element K chunks of one MMA, SBO = byte distance between 8-row groups.
// (validated on hardware by tools/tc_probe.cu)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46);
}
// instruction descriptor, kind::f16: D=f32, A=B=f16, both K-major, M=128
__device__ __forceinline__ constexpr uint32_t idesc_f16(int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                 :: "r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                   "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                   "=r"(v[30]), "=r"(v[31]) : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t addr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
                    "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}

// fp32 -> (fp16 hi, fp16 lo) with hi = x truncated to 11 significant bits (exactly representable), lo = fp16(x - hi)
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    const float ah = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u), bh = __uint_as_float(__float_as_uint(b) & 0xFFFFE000u);
    const __half2 h = __floats2half2_rn(ah, bh), l = __floats2half2_rn(a - ah, b - bh);
    hi = *reinterpret_cast<const uint32_t*>(&h); lo = *reinterpret_cast<const uint32_t*>(&l);
}

// ---- MMA issue for one layer of one slot (single thread) -------------------------------------------------------------
// A planes (hi, lo) either in TMEM (column address) or in shared memory (x tile); three passes hi*hi, lo*hi, hi*lo.
template <int N, int KSTEPS_T, int KSTEPS_S>
__device__ __forceinline__ void issue_layer(uint32_t acc, uint32_t a_hi_t, uint32_t a_lo_t, uint32_t x_hi_s, uint32_t x_lo_s,
                                            uint32_t w_hi_s, uint32_t w_lo_s) {
    constexpr uint32_t idesc = idesc_f16(N);
    constexpr uint32_t WK = N * 16;                         // bytes per 8-wide K chunk of the weight slab
    constexpr uint32_t XK = TILE * 16;
    uint32_t accumulate = 0;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a_t = pass == 1 ? a_lo_t : a_hi_t, a_s = pass == 1 ? x_lo_s : x_hi_s, w = pass == 2 ? w_lo_s : w_hi_s;
#pragma unroll
        for (int ks = 0; ks < KSTEPS_T; ++ks) {
            mma_ts(acc, a_t + ks * 8, smem_desc(w + ks * 2 * WK, WK, 128), idesc, accumulate);
            accumulate = 1;
        }
#pragma unroll
        for (int ks = 0; ks < KSTEPS_S; ++ks) {
            mma_ss(acc, smem_desc(a_s + ks * 2 * XK, XK, 128), smem_desc(w + (KSTEPS_T + ks) * 2 * WK, WK, 128), idesc, accumulate);
            accumulate = 1;
        }
    }
}

__global__ void __launch_bounds__(THREADS, 1) decode_tc_kernel(const unsigned char* __restrict__ image, const float* __restrict__ P, DecodeArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = sbase + OFF_BAR;
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 64);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int64_t n_total = a.n_dev ? (int64_t)*a.n_dev : a.n;
    const int64_t n_tiles = (n_total + TILE - 1) / TILE;

    if (threadIdx.x == 0) {
        mbar_init(bar0 + 8 * BAR_W, 1);
        mbar_init(bar0 + 8 * BAR_X0, TILE); mbar_init(bar0 + 8 * BAR_X1, TILE);
        mbar_init(bar0 + 8 * BAR_ACC0, 1); mbar_init(bar0 + 8 * BAR_ACC1, 1);
        mbar_init(bar0 + 8 * BAR_A0, TILE); mbar_init(bar0 + 8 * BAR_A1, TILE);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_ptr_s)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr_s;

    if (warp == 8) {
        // ===================================================== weight load + MMA issuer (one elected lane)
        if (lane == 0) {
            mbar_expect_tx(bar0 + 8 * BAR_W, IMAGE_B);
            constexpr uint32_t CH = 32768;
            for (uint32_t off = 0; off < IMAGE_B; off += CH) bulk_g2s(sbase + off, image + off, (IMAGE_B - off) < CH ? (IMAGE_B - off) : CH, bar0 + 8 * BAR_W);
            mbar_wait(bar0 + 8 * BAR_W, 0);
            const uint32_t w_hi = sbase, w_lo = sbase + PLANE_B;
            uint32_t ph_x[2] = {0, 0}, ph_a[2] = {0, 0};
            for (int64_t it = 0;; ++it) {
                const int64_t t0 = blockIdx.x + (int64_t)gridDim.x * (2 * it), t1 = t0 + gridDim.x;
                const bool live[2] = {t0 < n_tiles, t1 < n_tiles};
                if (!live[0]) break;
#pragma unroll
                for (int layer = 0; layer < 4; ++layer) {
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        if (!live[s]) continue;
                        const uint32_t acc = tmem + s * 256, a_hi = acc + 128, a_lo = acc + 192;
                        const uint32_t x_hi = sbase + OFF_X + s * 2 * X_PLANE_B, x_lo = x_hi + X_PLANE_B;
                        if (layer == 0) { mbar_wait(bar0 + 8 * (BAR_X0 + s), ph_x[s]); ph_x[s] ^= 1; }
                        else { mbar_wait(bar0 + 8 * (BAR_A0 + s), ph_a[s]); ph_a[s] ^= 1; }
                        tc_fence_after();
                        if (layer == 0) issue_layer<128, 0, 2>(acc, 0, 0, x_hi, x_lo, w_hi + OFF_W0, w_lo + OFF_W0);
                        else if (layer == 1) issue_layer<128, 8, 0>(acc, a_hi, a_lo, 0, 0, w_hi + OFF_W1, w_lo + OFF_W1);
                        else if (layer == 2) issue_layer<96, 8, 0>(acc, a_hi, a_lo, 0, 0, w_hi + OFF_W2, w_lo + OFF_W2);
                        else issue_layer<128, 6, 2>(acc, a_hi, a_lo, x_hi, x_lo, w_hi + OFF_W3, w_lo + OFF_W3);
                        mma_commit(bar0 + 8 * (BAR_ACC0 + s));
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ===================================================== gather + epilogue warpgroup of slot s (one sample per thread)
        const int s = warp >> 2;
        const int row = threadIdx.x & 127;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t acc = tmem + s * 256 + lane_base, a_hi = acc + 128, a_lo = acc + 192;
        unsigned char* x_hi_p = smem + OFF_X + s * 2 * X_PLANE_B;
        unsigned char* x_lo_p = x_hi_p + X_PLANE_B;
        const float* bias = reinterpret_cast<const float*>(smem + OFF_BIAS);
        const int n3 = a.lat_n * a.lat_n * a.lat_n;
        uint32_t ph_acc = 0;
        bool weights_ready = false;
        for (int64_t it = 0;; ++it) {
            const int64_t tile = blockIdx.x + (int64_t)gridDim.x * (2 * it + s);
            if (tile >= n_tiles) break;
            const int64_t sidx = tile * TILE + row;
            int64_t src_row, out; int li;
            decode_sample_source(a, sidx, n_total, n3, src_row, out, li);
            {   // ---- gather the 32 inputs of this sample, split, and write the layer-0 A tile (k-chunk-major, conflict free)
                float x[32];
                if (src_row >= 0) {
                    const float* lp = a.latent + src_row * DIF_L;
#pragma unroll
                    for (int j = 0; j < DIF_L; ++j) x[j] = __ldg(lp + j);
                    if (a.mode == 0) {
                        x[29] = __ldg(a.xyz + sidx * 3); x[30] = __ldg(a.xyz + sidx * 3 + 1); x[31] = __ldg(a.xyz + sidx * 3 + 2);
                    } else {
                        const int nn = a.lat_n;
                        x[29] = lattice_coord(a, li / (nn * nn)); x[30] = lattice_coord(a, (li / nn) % nn); x[31] = lattice_coord(a, li % nn);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = 0.f;
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint4 h, l;
                    split_pair(x[8 * c + 0], x[8 * c + 1], h.x, l.x); split_pair(x[8 * c + 2], x[8 * c + 3], h.y, l.y);
                    split_pair(x[8 * c + 4], x[8 * c + 5], h.z, l.z); split_pair(x[8 * c + 6], x[8 * c + 7], h.w, l.w);
                    *reinterpret_cast<uint4*>(x_hi_p + c * (TILE * 16) + row * 16) = h;
                    *reinterpret_cast<uint4*>(x_lo_p + c * (TILE * 16) + row * 16) = l;
                }
                fence_async_smem();                         // generic-proxy stores -> visible to the tensor core (async proxy)
                mbar_arrive(bar0 + 8 * (BAR_X0 + s));
            }
            if (!weights_ready) { mbar_wait(bar0 + 8 * BAR_W, 0); weights_ready = true; }     // biases arrive with the weight image
            // ---- hidden layers 0..2: accumulator -> +bias, ReLU, split -> fp16 A operand in TMEM
#pragma unroll
            for (int layer = 0; layer < 3; ++layer) {
                const int ncol = layer == 2 ? 96 : 128;
                const float* b = bias + (layer == 0 ? 0 : (layer == 1 ? 128 : 256));
                mbar_wait(bar0 + 8 * (BAR_ACC0 + s), ph_acc); ph_acc ^= 1;
                tc_fence_after();
#pragma unroll
                for (int c0 = 0; c0 < 128; c0 += 32) {
                    if (c0 < ncol) {
                        uint32_t v[32], hi[16], lo[16];
                        tmem_ld32(acc + c0, v);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float2 bb = *reinterpret_cast<const float2*>(b + c0 + 2 * j);
                            const float f0 = fmaxf(__uint_as_float(v[2 * j]) + bb.x, 0.f), f1 = fmaxf(__uint_as_float(v[2 * j + 1]) + bb.y, 0.f);
                            split_pair(f0, f1, hi[j], lo[j]);
                        }
                        tmem_st16(a_hi + c0 / 2, hi);
                        tmem_st16(a_lo + c0 / 2, lo);
                    }
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                mbar_arrive(bar0 + 8 * (BAR_A0 + s));
            }
            // ---- layer 3 + the two heads on CUDA cores (std from the last layer's input, di_decoder.py:65-70)
            mbar_wait(bar0 + 8 * (BAR_ACC0 + s), ph_acc); ph_acc ^= 1;
            tc_fence_after();
            float p_sdf = 0.f, p_std = 0.f;
#pragma unroll
            for (int c0 = 0; c0 < 128; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(acc + c0, v);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 bb = *reinterpret_cast<const float4*>(bias + 352 + c0 + j);
                    const float4 w4 = __ldg(reinterpret_cast<const float4*>(P + DecW::w4 + c0 + j));
                    const float4 wu = __ldg(reinterpret_cast<const float4*>(P + DecW::wu + c0 + j));
                    const float h0 = fmaxf(__uint_as_float(v[j]) + bb.x, 0.f), h1 = fmaxf(__uint_as_float(v[j + 1]) + bb.y, 0.f);
                    const float h2 = fmaxf(__uint_as_float(v[j + 2]) + bb.z, 0.f), h3 = fmaxf(__uint_as_float(v[j + 3]) + bb.w, 0.f);
                    p_sdf = fmaf(w4.x, h0, p_sdf); p_sdf = fmaf(w4.y, h1, p_sdf); p_sdf = fmaf(w4.z, h2, p_sdf); p_sdf = fmaf(w4.w, h3, p_sdf);
                    p_std = fmaf(wu.x, h0, p_std); p_std = fmaf(wu.y, h1, p_std); p_std = fmaf(wu.z, h2, p_std); p_std = fmaf(wu.w, h3, p_std);
                }
            }
            tc_fence_before();                               // accumulator reads are complete before the next tile's MMA may overwrite it
            if (src_row >= 0) {
                a.sdf[out] = a.sdf_sign * tanhf(p_sdf + __ldg(P + DecW::b4));
                a.std[out] = 0.05f + 0.5f * softplus_ref(p_std + __ldg(P + DecW::bu));
            } else if (sidx < n_total && a.mode == 0 && !a.out_index) {
                a.sdf[out] = 0.f; a.std[out] = 0.f;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(512));
}

// ---- weight image: fp16 hi/lo planes in the no-swizzle K-major core-matrix layout, + biases -----------------------------
// element (n, k) of a layer with N rows lives at  (k/8)*(N*16) + n*16 + (k%8)*2  bytes inside its slab.
__global__ void prepare_tc_kernel(const float* __restrict__ P, unsigned char* __restrict__ image) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const int offW[4] = {DecW::W0, DecW::W1, DecW::W2, DecW::W3};
    const int Ns[4] = {128, 128, 96, 128}, Ks[4] = {32, 128, 128, 128};
    const uint32_t offs[4] = {OFF_W0, OFF_W1, OFF_W2, OFF_W3};
    for (int l = 0; l < 4; ++l) {
        const int N = Ns[l], K = Ks[l];
        for (int i = tid; i < N * K; i += nth) {
            const int n = i / K, k = i % K;
            const float w = P[offW[l] + i];
            const __half h = __float2half_rn(w);
            const __half lo = __float2half_rn(w - __half2float(h));
            const uint32_t o = offs[l] + (uint32_t)(k / 8) * (N * 16) + n * 16 + (k % 8) * 2;
            *reinterpret_cast<__half*>(image + o) = h;
            *reinterpret_cast<__half*>(image + PLANE_B + o) = lo;
        }
    }
    float* b = reinterpret_cast<float*>(image + OFF_BIAS);
    for (int i = tid; i < 128; i += nth) {
        b[i] = P[DecW::b0 + i]; b[128 + i] = P[DecW::b1 + i]; b[352 + i] = P[DecW::b3 + i];
        if (i < 96) b[256 + i] = P[DecW::b2 + i];
    }
}

}  // namespace tc

size_t decoder_tc_image_bytes() { return tc::IMAGE_B; }

int prepare_decoder_tc(const float* P, unsigned char* image, cudaStream_t st) {
    tc::prepare_tc_kernel<<<64, 256, 0, st>>>(P, image);
    DIF_COUNT_LAUNCH(1);
    return check_launch("prepare_tc_kernel");
}

int launch_decode_tc(const void* prepared, DecodeArgs a, int64_t n_max, cudaStream_t st) {
    if (n_max <= 0) return DIF_OK;
    const float* P = (const float*)prepared;
    const unsigned char* image = (const unsigned char*)prepared + (size_t)DecW::FP32_END * sizeof(float);
    const int64_t n_tiles = (n_max + tc::TILE - 1) / tc::TILE;
    const int grid = (int)(n_tiles < DIF_NUM_SMS ? n_tiles : DIF_NUM_SMS);
    cudaFuncSetAttribute(tc::decode_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_B);
    prof_begin(DIF_PROF_DECODE, st);
    tc::decode_tc_kernel<<<grid, tc::THREADS, tc::SMEM_B, st>>>(image, P, a);
    prof_end(DIF_PROF_DECODE, st);
    DIF_COUNT_LAUNCH(1);
    return check_launch("decode_tc_kernel");
}

}  // namespace dif
