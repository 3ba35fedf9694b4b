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
// Warp roles (20 warps): warps 0-7 = epilogue of slot 0, warps 8-15 = epilogue of slot 1, warp 16 = TMEM allocator + MMA
// issuer, warps 17/18 = gather producers of slot 0/1 (warp 19 idles).  Inside a slot, epilogue warp w owns TMEM lanes
// 32*(w&3).. (its 32 samples) and the column half (w>>2)&1 of every accumulator, so each SM sub-partition always has four
// epilogue warps to hide tcgen05.ld / conversion latency.  A producer warp gathers a whole 128-sample tile with
// coalesced loads (one sample's 29 latent floats + xyz per load instruction, lane = input column), three 16-row
// batches in flight, and runs one tile ahead of the tensor pipe.
#include <atomic>
#include <stdlib.h>

#include "tc_common.cuh"

namespace dif {

namespace tc {

// ---- gather producer helpers ------------------------------------------------------------------------------------------
// One producer warp per slot, one SAMPLE PER LANE: 32 independent loads per lane are in flight at once and the whole
// 32-row pass costs ~150 warp instructions (a warp-cooperative, coalesced variant needed ~70 instructions PER ROW of
// shuffles and address arithmetic and made the single producer warp the bottleneck of the kernel).
__device__ __forceinline__ void gather_load_row(const DecodeArgs& a, int64_t sidx, int64_t n_total, int n3, float inv_n, float (&x)[32], bool& valid) {
    int64_t row, out; int li;
    decode_sample_source(a, sidx, n_total, n3, row, out, li);
    valid = row >= 0;
    load_latent_row(a.latent, valid ? row : 0, a.lat_stride, x);
    if (a.mode == 0) {
        const float* xp = a.xyz + (valid ? sidx : 0) * 3;
        x[29] = __ldg(xp); x[30] = __ldg(xp + 1); x[31] = __ldg(xp + 2);
    } else {
        // lattice point -> (i, j, k) with exact float reciprocals (indices < 2^12), utility.py:143-147
        const int q1 = (int)(((float)li + 0.5f) * inv_n), q2 = (int)(((float)q1 + 0.5f) * inv_n);
        x[29] = lattice_coord(a, q2); x[30] = lattice_coord(a, q1 - q2 * a.lat_n); x[31] = lattice_coord(a, li - q1 * a.lat_n);
    }
}
// (hi, lo) fp16 split of one sample's 32 inputs -> its row of the layer-0 A tile (k-chunk-major core-matrix layout; consecutive
// lanes write consecutive 16-byte slots: conflict free)
__device__ __forceinline__ void gather_store_row(const float (&x)[32], bool valid, int row, unsigned char* x_hi_p) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint4 h, l;
        split_pair(valid ? x[8 * c + 0] : 0.f, valid ? x[8 * c + 1] : 0.f, h.x, l.x); split_pair(valid ? x[8 * c + 2] : 0.f, valid ? x[8 * c + 3] : 0.f, h.y, l.y);
        split_pair(valid ? x[8 * c + 4] : 0.f, valid ? x[8 * c + 5] : 0.f, h.z, l.z); split_pair(valid ? x[8 * c + 6] : 0.f, valid ? x[8 * c + 7] : 0.f, h.w, l.w);
        *reinterpret_cast<uint4*>(x_hi_p + c * X_CHUNK_B + row * 16) = h;
        *reinterpret_cast<uint4*>(x_hi_p + X_PLANE_B + c * X_CHUNK_B + row * 16) = l;
    }
}

// bias + ReLU + hi/lo split of NCOL (32 or 16) accumulator columns -> packed fp16 A-operand columns in TMEM
template <int NCOL>
__device__ __forceinline__ void convert_chunk(const uint32_t* v, const float* b, uint32_t a_hi, uint32_t a_lo) {
    uint32_t hi[NCOL / 2], lo[NCOL / 2];
#pragma unroll
    for (int j = 0; j < NCOL / 2; ++j) {
        const float2 bb = *reinterpret_cast<const float2*>(b + 2 * j);
        split_pair(fmaxf(__uint_as_float(v[2 * j]) + bb.x, 0.f), fmaxf(__uint_as_float(v[2 * j + 1]) + bb.y, 0.f), hi[j], lo[j]);
    }
    if constexpr (NCOL == 32) {
        tmem_st16(a_hi, *reinterpret_cast<const uint32_t(*)[16]>(hi)); tmem_st16(a_lo, *reinterpret_cast<const uint32_t(*)[16]>(lo));
    } else {
        tmem_st8(a_hi, hi); tmem_st8(a_lo, lo);
    }
}

__global__ void __launch_bounds__(THREADS, 1) decode_tc_kernel(const unsigned char* __restrict__ image, const float* __restrict__ P, DecodeArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = sbase + OFF_BAR;
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 96);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int64_t n_total = a.n_dev ? (int64_t)*a.n_dev : a.n;
    const int64_t n_tiles = (n_total + TILE - 1) / TILE;
    const bool timing = g_tc_timing != nullptr;
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tcur = timing ? clock64() : 0;

    if (threadIdx.x == 0) {
        mbar_init(bar0 + 8 * BAR_W, 1);
        mbar_init(bar0 + 8 * BAR_X0, 1); mbar_init(bar0 + 8 * BAR_X1, 1);          // producer warp -> MMA: layer-0 A tile ready
        mbar_init(bar0 + 8 * BAR_XF0, 8); mbar_init(bar0 + 8 * BAR_XF1, 8);        // epilogue warps -> producer: x tile free again
        mbar_init(bar0 + 8 * BAR_E0, 8); mbar_init(bar0 + 8 * BAR_E1, 8);          // epilogue warps -> MMA: last accumulator of the tile consumed
        mbar_init(bar0 + 8 * BAR_ACC0, 1); mbar_init(bar0 + 8 * BAR_ACC1, 1);
        mbar_init(bar0 + 8 * BAR_A0, 8); mbar_init(bar0 + 8 * BAR_A1, 8);          // one elected arrival per epilogue warp of the slot
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_ptr_s)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_ptr_s, 0);      // provably warp-uniform for the MMA issuer

    if (warp == MMA_WARP) {
        // ===================================================== weight load + MMA issuer
        // The whole warp runs this loop with warp-uniform values (descriptors live in uniform registers); only the
        // tcgen05 instructions themselves are executed by one elected lane.
        if (lane == 0) {
            mbar_expect_tx(bar0 + 8 * BAR_W, IMAGE_B);
            constexpr uint32_t CH = 32768;
            for (uint32_t off = 0; off < IMAGE_B; off += CH) bulk_g2s(sbase + off, image + off, (IMAGE_B - off) < CH ? (IMAGE_B - off) : CH, bar0 + 8 * BAR_W);
        }
        __syncwarp();
        mbar_wait(bar0 + 8 * BAR_W, 0);
        TC_ACC(0, tcur);                                   // [0] weight image load
        uint64_t wd_hi[4], wd_lo[4], xd_hi[2], xd_lo[2];
        {
            const uint32_t offs[4] = {OFF_W0, OFF_W1, OFF_W2, OFF_W3};
            const uint32_t wk[4] = {128 * 16, 128 * 16, 96 * 16, 128 * 16};
#pragma unroll
            for (int l = 0; l < 4; ++l) { wd_hi[l] = smem_desc(sbase + offs[l], wk[l], 128); wd_lo[l] = smem_desc(sbase + PLANE_B + offs[l], wk[l], 128); }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                xd_hi[q] = smem_desc(sbase + OFF_X + q * 2 * X_PLANE_B, X_CHUNK_B, 128);
                xd_lo[q] = smem_desc(sbase + OFF_X + q * 2 * X_PLANE_B + X_PLANE_B, X_CHUNK_B, 128);
            }
        }
        uint32_t ph_x = 0, ph_a = 0, ph_e = 0;             // bit s = phase parity of slot s
        for (int64_t it = 0;; ++it) {
            const int64_t t0 = blockIdx.x + (int64_t)gridDim.x * (2 * it), t1 = t0 + gridDim.x;
            const bool live[2] = {t0 < n_tiles, t1 < n_tiles};
            if (!live[0]) break;
#pragma unroll 1
            for (int layer = 0; layer < 4; ++layer) {
#pragma unroll 1
                for (int s = 0; s < 2; ++s) {
                    if (!live[s]) continue;
                    const uint32_t acc = tmem + s * 256, a_hi = acc + 128, a_lo = acc + 192;
                    if (layer == 0) {
                        mbar_wait_tight(bar0 + 8 * (BAR_X0 + s), (ph_x >> s) & 1); ph_x ^= 1u << s;
                        // the previous tile's head evaluation must have finished reading this slot's accumulator
                        if (it > 0) { mbar_wait_tight(bar0 + 8 * (BAR_E0 + s), (ph_e >> s) & 1); ph_e ^= 1u << s; }
                    }
                    else { mbar_wait_tight(bar0 + 8 * (BAR_A0 + s), (ph_a >> s) & 1); ph_a ^= 1u << s; }
                    TC_ACC(1, tcur);                       // [1] MMA warp waiting for operands
                    tc_fence_after();
                    if (elect_one()) {
                        if (layer == 0) issue_layer(idesc_f16(128), 128 * 16, 0, 2, acc, 0, 0, s ? xd_hi[1] : xd_hi[0], s ? xd_lo[1] : xd_lo[0], wd_hi[0], wd_lo[0]);
                        else if (layer == 1) issue_layer(idesc_f16(128), 128 * 16, 8, 0, acc, a_hi, a_lo, 0, 0, wd_hi[1], wd_lo[1]);
                        else if (layer == 2) issue_layer(idesc_f16(96), 96 * 16, 8, 0, acc, a_hi, a_lo, 0, 0, wd_hi[2], wd_lo[2]);
                        else issue_layer(idesc_f16(128), 128 * 16, 8, 0, acc, a_hi, a_lo, 0, 0, wd_hi[3], wd_lo[3]);     // [h2 | x] both in TMEM
                        mma_commit(bar0 + 8 * (BAR_ACC0 + s));
                    }
                    __syncwarp();
                    TC_ACC(2, tcur);                       // [2] MMA warp issuing
                }
            }
        }
    } else if (warp >= PRODUCER_WARP0) {
        // ===================================================== gather producer of slot s: global -> (hi, lo) fp16 -> layer-0 A tile in smem
        const int s = warp - PRODUCER_WARP0;
        if (s < 2) {
            unsigned char* x_hi_p = smem + OFF_X + s * 2 * X_PLANE_B;
            const int n3 = a.lat_n * a.lat_n * a.lat_n;
            const float inv_n = 1.0f / (float)(a.lat_n > 0 ? a.lat_n : 1);
            float xa[32], xb[32];                       // two 32-sample passes in flight (one sample per lane)
            bool va = false, vb = false;
            uint32_t ph_xf = 0;
            int64_t tile = blockIdx.x + (int64_t)gridDim.x * s;
            if (tile < n_tiles) gather_load_row(a, tile * TILE + lane, n_total, n3, inv_n, xa, va);
            for (int64_t it = 0; tile < n_tiles; ++it) {
                TC_ACC(0, tcur);                             // producer [0]: prefetch of the first pass
                if (it > 0) { mbar_wait(bar0 + 8 * (BAR_XF0 + s), ph_xf); ph_xf ^= 1; }      // previous tile's x is parked in TMEM (layer-2 epilogue)
                TC_ACC(1, tcur);                             // producer [1]: waiting for the x tile to be free
                gather_load_row(a, tile * TILE + 32 + lane, n_total, n3, inv_n, xb, vb);
                gather_store_row(xa, va, lane, x_hi_p);
                gather_load_row(a, tile * TILE + 64 + lane, n_total, n3, inv_n, xa, va);
                gather_store_row(xb, vb, 32 + lane, x_hi_p);
                gather_load_row(a, tile * TILE + 96 + lane, n_total, n3, inv_n, xb, vb);
                gather_store_row(xa, va, 64 + lane, x_hi_p);
                const int64_t next = tile + 2 * (int64_t)gridDim.x;
                if (next < n_tiles) gather_load_row(a, next * TILE + lane, n_total, n3, inv_n, xa, va);   // run ahead while the tensor pipe works
                gather_store_row(xb, vb, 96 + lane, x_hi_p);
                TC_ACC(3, tcur);                             // producer [3]: streaming the tile
                fence_async_smem();                          // generic-proxy stores -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(bar0 + 8 * (BAR_X0 + s));
                TC_ACC(4, tcur);                             // producer [4]: proxy fence + arrive
                tile = next;
            }
        }
    } else {
        // ===================================================== epilogue warps of slot s
        const int s = warp >> 3;
        const int quad = warp & 3, half = (warp >> 2) & 1;
        const int row = quad * 32 + lane;           // the sample this thread converts / outputs (TMEM lane)
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        const uint32_t acc = tmem + s * 256 + lane_base, a_hi = acc + 128, a_lo = acc + 192;
        const float* bias = reinterpret_cast<const float*>(smem + OFF_BIAS);
        const int n3 = a.lat_n * a.lat_n * a.lat_n;
        uint32_t ph_acc = 0;
        // the head this warp evaluates (half 0: sdf, half 1: std); weights come from the constant bank (c_head_w)
        const int head_slot = *reinterpret_cast<const int*>(image + IMAGE_B);         // which constant-memory slot holds this decoder's heads
        const float head_bias = __ldg(P + (half ? DecW::bu : DecW::b4));
        mbar_wait(bar0 + 8 * BAR_W, 0);               // biases arrive with the weight image
        for (int64_t it = 0;; ++it) {
            const int64_t tile = blockIdx.x + (int64_t)gridDim.x * (2 * it + s);
            if (tile >= n_tiles) break;
            TC_ACC(7, tcur);
            // ---- hidden layers 0..2: this warp converts its column half: accumulator -> +bias, ReLU, split -> fp16 A operand in TMEM
#pragma unroll 1
            for (int layer = 0; layer < 3; ++layer) {
                const int hw = layer == 2 ? 48 : 64;          // columns per half (layer 2 has 96 outputs)
                const int c_base = half * hw;
                const float* b = bias + layer * 128 + c_base;  // b0 @0, b1 @128, b2 @256
                mbar_wait(bar0 + 8 * (BAR_ACC0 + s), ph_acc); ph_acc ^= 1;
                TC_ACC(4, tcur);                               // [4] epilogue warps waiting for the accumulator
                tc_fence_after();
                if (layer == 2) {
                    // Layers 0-2 are done with the 64 A-operand columns.  Park this sample's 32 inputs (hi plane by half-0 warps, lo plane by
                    // half-1 warps) in the 16 A-operand columns that layer 2's 96 outputs leave free, so that the skip
                    // connection of layer 3 reads [h2 | x] entirely from TMEM and the producer may refill the x tile now
                    // (it then overlaps with layer 3 and the head evaluation instead of sitting on the critical path).
                    const unsigned char* xp = smem + OFF_X + s * 2 * X_PLANE_B + half * X_PLANE_B + row * 16;
                    uint32_t xv[16];
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const uint4 q = *reinterpret_cast<const uint4*>(xp + cc * X_CHUNK_B);
                        xv[4 * cc] = q.x; xv[4 * cc + 1] = q.y; xv[4 * cc + 2] = q.z; xv[4 * cc + 3] = q.w;
                    }
                    tmem_st16((half ? a_lo : a_hi) + 48, xv);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar0 + 8 * (BAR_XF0 + s));    // (our own generic-proxy reads of x are complete)
                }
#pragma unroll 1
                for (int c = 0; c < hw; c += 32) {             // 32 columns per iteration (the last one of layer 2 has 16)
                    uint32_t v0[16], v1[16];
                    const bool two = c + 16 < hw;
                    tmem_ld16_nowait(acc + c_base + c, v0);
                    if (two) tmem_ld16_nowait(acc + c_base + c + 16, v1);
                    tmem_ld_wait();
                    convert_chunk<16>(v0, b + c, a_hi + (c_base + c) / 2, a_lo + (c_base + c) / 2);
                    if (two) convert_chunk<16>(v1, b + c + 16, a_hi + (c_base + c) / 2 + 8, a_lo + (c_base + c) / 2 + 8);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar0 + 8 * (BAR_A0 + s));
                TC_ACC(5, tcur);                               // [5] hidden-layer conversion
            }
            // ---- layer 3 + one head per column-half warp on CUDA cores (half 0 -> sdf, half 1 -> std; di_decoder.py:65-70,84)
            mbar_wait(bar0 + 8 * (BAR_ACC0 + s), ph_acc); ph_acc ^= 1;
            TC_ACC(4, tcur);
            tc_fence_after();
            float p0 = 0.f, p1 = 0.f;
            const float* hw_c = c_head_w[head_slot][half];     // constant bank: the weight is an FFMA operand, no load instruction
#pragma unroll 1
            for (int c0 = 0; c0 < 128; c0 += 32) {
                uint32_t v0[16], v1[16];
                tmem_ld16_nowait(acc + c0, v0);
                tmem_ld16_nowait(acc + c0 + 16, v1);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 b0 = *reinterpret_cast<const float4*>(bias + 352 + c0 + j), b1 = *reinterpret_cast<const float4*>(bias + 352 + c0 + 16 + j);
                    p0 = fmaf(hw_c[c0 + j], fmaxf(__uint_as_float(v0[j]) + b0.x, 0.f), p0);
                    p0 = fmaf(hw_c[c0 + j + 1], fmaxf(__uint_as_float(v0[j + 1]) + b0.y, 0.f), p0);
                    p0 = fmaf(hw_c[c0 + j + 2], fmaxf(__uint_as_float(v0[j + 2]) + b0.z, 0.f), p0);
                    p0 = fmaf(hw_c[c0 + j + 3], fmaxf(__uint_as_float(v0[j + 3]) + b0.w, 0.f), p0);
                    p1 = fmaf(hw_c[c0 + 16 + j], fmaxf(__uint_as_float(v1[j]) + b1.x, 0.f), p1);
                    p1 = fmaf(hw_c[c0 + 16 + j + 1], fmaxf(__uint_as_float(v1[j + 1]) + b1.y, 0.f), p1);
                    p1 = fmaf(hw_c[c0 + 16 + j + 2], fmaxf(__uint_as_float(v1[j + 2]) + b1.z, 0.f), p1);
                    p1 = fmaf(hw_c[c0 + 16 + j + 3], fmaxf(__uint_as_float(v1[j + 3]) + b1.w, 0.f), p1);
                }
            }
            tc_fence_before();                               // accumulator reads complete: the next tile's layer 0 may overwrite it
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + 8 * (BAR_E0 + s));
            const int64_t sidx = tile * TILE + row;
            int64_t src_row, out; int li_unused;
            decode_sample_source(a, sidx, n_total, n3, src_row, out, li_unused);
            const float pre = p0 + p1 + head_bias;
            float* dst = half ? a.std : a.sdf;
            if (src_row >= 0) dst[out] = half ? 0.05f + 0.5f * softplus_ref(pre) : a.sdf_sign * tanhf(pre);
            else if (sidx < n_total && a.mode == 0 && !a.out_index) dst[out] = 0.f;
            TC_ACC(6, tcur);                                   // [6] last layer + heads + output
        }
    }
    if (timing && lane == 0) {
        for (int k = 0; k < 8; ++k) g_tc_timing[((size_t)blockIdx.x * 20 + warp) * 8 + k] = (unsigned long long)tacc[k];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(512));
}

// ---- weight image: fp16 hi/lo planes in the no-swizzle K-major core-matrix layout, + biases -----------------------------
// element (n, k) of a layer with N rows lives at  (k/8)*(N*16) + n*16 + (k%8)*2  bytes inside its slab.
__global__ void prepare_tc_kernel(const float* __restrict__ P, unsigned char* __restrict__ image, int head_slot) {
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
    if (tid == 0) *reinterpret_cast<int*>(image + IMAGE_B) = head_slot;
    if (tid < 64) {                                   // staging copy of c_head_g3 (copied into the constant bank by the host right after)
        uint32_t* g3 = reinterpret_cast<uint32_t*>(image + IMAGE_B + 16);
        const float wa = G3_SCALE * P[DecW::w4 + 2 * tid], wb = G3_SCALE * P[DecW::w4 + 2 * tid + 1];
        split_pair(wa, wb, g3[tid], g3[64 + tid]);
        const __half2 rn = __floats2half2_rn(wa, wb);
        g3[128 + tid] = *reinterpret_cast<const uint32_t*>(&rn);
    }
    float* b = reinterpret_cast<float*>(image + OFF_BIAS);
    for (int i = tid; i < 128; i += nth) {
        b[i] = P[DecW::b0 + i]; b[128 + i] = P[DecW::b1 + i]; b[352 + i] = P[DecW::b3 + i];
        if (i < 96) b[256 + i] = P[DecW::b2 + i];
    }
}

}  // namespace tc
}  // namespace dif

#include "decode_tc2.cuh"
#include "icp_args.cuh"
#include "icp_tc.cuh"

namespace dif {


size_t decoder_tc_image_bytes() { return tc::IMAGE_B + 16 + sizeof(tc::c_head_g3[0]); }

int set_tc_timing_buffer(unsigned long long* dev_buf) {
    return cudaMemcpyToSymbol(tc::g_tc_timing, &dev_buf, sizeof(dev_buf)) == cudaSuccess ? DIF_OK : DIF_E_LAUNCH;
}

int prepare_decoder_tc(const float* P, unsigned char* image, cudaStream_t st) {
    static std::atomic<int> next_slot{0};
    const int slot = next_slot.fetch_add(1) % tc::HEAD_SLOTS;
    // w4[128] and wu[128] are adjacent in the prepared fp32 section: one device-to-device copy into the constant bank
    if (cudaMemcpyToSymbolAsync(tc::c_head_w, P + DecW::w4, 2 * 128 * sizeof(float), (size_t)slot * 2 * 128 * sizeof(float),
                                cudaMemcpyDeviceToDevice, st) != cudaSuccess) return check_launch("cudaMemcpyToSymbolAsync(c_head_w)");
    tc::prepare_tc_kernel<<<64, 256, 0, st>>>(P, image, slot);
    DIF_COUNT_LAUNCH(1);
    if (cudaMemcpyToSymbolAsync(tc::c_head_g3, image + tc::IMAGE_B + 16, sizeof(tc::c_head_g3[0]), (size_t)slot * sizeof(tc::c_head_g3[0]),
                                cudaMemcpyDeviceToDevice, st) != cudaSuccess) return check_launch("cudaMemcpyToSymbolAsync(c_head_g3)");
    return check_launch("prepare_tc_kernel");
}

int launch_decode_tc(const void* prepared, DecodeArgs a, int64_t n_max, cudaStream_t st) {
    if (n_max <= 0) return DIF_OK;
    const float* P = (const float*)prepared;
    const unsigned char* image = (const unsigned char*)prepared + (size_t)DecW::FP32_END * sizeof(float);
    const int64_t n_tiles = (n_max + tc::TILE - 1) / tc::TILE;
    const int grid = (int)(n_tiles < DIF_NUM_SMS ? n_tiles : DIF_NUM_SMS);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(tc::decode_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_B);
        cudaFuncSetAttribute(tc::decode_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::D2_SMEM_B);
        attr_set = true;
    }
    const char* v = getenv("DIF_DECODE_V");                  // "1": the first pipeline (separate accumulator / operand regions), kept for A/B timing
    prof_begin(DIF_PROF_DECODE, st);
    if (v && v[0] == '1') tc::decode_tc_kernel<<<grid, tc::THREADS, tc::SMEM_B, st>>>(image, P, a);
    else tc::decode_tc2_kernel<<<grid, tc::THREADS, tc::D2_SMEM_B, st>>>(image, P, a);
    prof_end(DIF_PROF_DECODE, st);
    DIF_COUNT_LAUNCH(1);
    return check_launch("decode_tc_kernel");
}

}  // namespace dif
