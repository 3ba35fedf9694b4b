// Decoder forward on tcgen05, second pipeline: IN-PLACE activation conversion + chunk-granular hand-off to the next layer.
// Included by decode_tc.cu (same weight image, same gather producers, same heads as decode_tc_kernel).
//
// Why: with the accumulator and the A operand in separate TMEM regions, layer l+1 could not start before the whole epilogue
// of layer l had finished (measured on B200, tools/tc_timing.py: per tile 5.1 k cycles of MMA issue against 4.3 k cycles of
// the issuer WAITING for epilogues - the tensor pipe sat idle 46 % of the time although two tiles were in flight, because a
// layer's MMA -> epilogue -> MMA chain is ~3.6 k cycles of mostly fixed latencies against 1.3 k cycles of MMA work).
//
// How: each tile owns two 128-column TMEM regions R0 / R1 that swap roles every layer.  Layer l accumulates into one region;
// its epilogue converts that region IN PLACE, 16 fp32 columns -> the same 16 columns holding 8 packed fp16 "hi" + 8 packed
// fp16 "lo" columns = exactly the A operand of ONE K=16 MMA step of layer l+1.  Layer l+1 accumulates into the OTHER region
// (whose previous content - the A operand of layer l - is dead once layer l's MMAs have completed), so there is no
// write-after-read hazard and layer l+1's first K steps are issued as soon as the first 32-column group of the epilogue is
// in place: the tensor pipe works on layer l+1 while the rest of layer l's epilogue is still converting.
//
//   L0: x tile (smem) -> R0   E0: R0 in place      L1: R0 -> R1   E1: R1 in place
//   L2: R1 -> R0 (N = 96)     E2: R0[0,96) in place; the 32 inputs were parked in R0[96,128) at E1 (skip connection)
//   L3: R0 -> R1              E3: heads read R1.   Next tile: L0 -> R0 right behind L3 (in-order pipe), L1 waits for E3.
//
// Hand-off barriers per slot: ACCa (L0, L2) / ACCb (L1, L3) accumulator complete (tcgen05.commit); G0..G3 one per 32-column
// epilogue group (4 arrivals: the 4 lane quadrants), consumed in the fixed order (half 0, it 0), (half 1, it 0), (half 0, it 1),
// (half 1, it 1) so that the accumulation order - and therefore every output bit - is reproducible; X / XF / E as before.
// One issuer warp per slot (warps 16 and 19) walks its own tiles with blocking waits in a fixed order.
#pragma once

namespace dif {
namespace tc {

constexpr uint32_t D2_OFF_BAR = OFF_X + 4 * X_PLANE_B;
constexpr uint32_t D2_SMEM_B = D2_OFF_BAR + 176 + 16;           // 21 barriers (168 B) + TMEM base pointer
static_assert(D2_SMEM_B <= 232448, "shared memory budget");
enum { D2_W = 0, D2_X = 1, D2_XF = 3, D2_E = 5, D2_ACCA = 7, D2_ACCB = 9, D2_G = 11, D2_W1 = 19, D2_W2 = 20, D2_W3 = 21 };      // per-slot barriers: index + slot; G: 11 + 4*slot + g (.. 18); W, W1..W3: weight groups

#ifndef DIF_D2_ISSUERS
#define DIF_D2_ISSUERS 1               // 1: one issuer, slots strictly alternating (default); 2: one issuer warp per slot (A/B)
#endif
constexpr int MMA_WARP2 = 19;          // second issuer warp (DIF_D2_ISSUERS == 2)

// Development aid (tools/tc_trace.py, build with -DDIF_TC_TRACE): CTA 0 logs (event id, SM clock) per warp into the timing buffer.
#ifdef DIF_TC_TRACE
#define TC_TRACE(ev) do { if (timing && lane == 0 && blockIdx.x == 0 && tn < 1024) \
    g_tc_timing[warp * 1024 + tn++] = ((unsigned long long)(ev) << 48) | ((unsigned long long)clock64() & 0xFFFFFFFFFFFFull); } while (0)
#else
#define TC_TRACE(ev) do { } while (0)
#endif

// the four hand-off groups of hidden layer LAYER (1: R0 -> R1, 2: R1 -> R0 with N = 96, 3: [h2 | x] in R0 -> R1) for one slot
template <int LAYER, class Trace>
__device__ __forceinline__ void issue_hidden_layer(uint32_t R0, uint32_t R1, uint64_t wh, uint64_t wl, uint32_t bg, uint32_t ph_g,
                                                   uint32_t commit_bar, uint32_t e_bar, bool wait_e, uint32_t& ph_e, Trace&& trace) {
    constexpr int N = LAYER == 2 ? 96 : 128;
    const uint32_t acc = LAYER == 2 ? R0 : R1, a_base = LAYER == 2 ? R1 : R0;
    // group g = 2 i + h  (column half h, epilogue iteration i)  ->  K steps: layers 1, 2: {4h + 2i, +1}; layer 3: i = 0: {3h, 3h + 1}, i = 1: {3h + 2, 6 + h}
    // The two column halves finish an iteration at about the same time, so their groups are awaited and issued together (12 MMAs
    // behind one wait / fence / elect sequence: that sequence costs ~250 cycles, too much to pay per 384 cycles of tensor work).
    mbar_wait_spin(bg, ph_g);
    mbar_wait_spin(bg + 8, ph_g);
    if (LAYER == 1 && wait_e) { mbar_wait_spin(e_bar, ph_e); ph_e ^= 1; }             // heads of the previous tile have read R1
    trace(16 + 4 * LAYER);
    tc_fence_after();
    if (elect_one()) {
        issue_group<N, 0, 1, true>(acc, a_base, wh, wl);
        if (LAYER < 3) issue_group<N, 4, 5, false>(acc, a_base, wh, wl); else issue_group<N, 3, 4, false>(acc, a_base, wh, wl);
    }
    __syncwarp();
    trace(49 + 4 * LAYER);
    mbar_wait_spin(bg + 16, ph_g);
    mbar_wait_spin(bg + 24, ph_g);
    trace(18 + 4 * LAYER);
    tc_fence_after();
    if (elect_one()) {
        if (LAYER < 3) { issue_group<N, 2, 3, false>(acc, a_base, wh, wl); issue_group<N, 6, 7, false>(acc, a_base, wh, wl); }
        else { issue_group<N, 2, 6, false>(acc, a_base, wh, wl); issue_group<N, 5, 7, false>(acc, a_base, wh, wl); }
        mma_commit(commit_bar);
    }
    __syncwarp();
    trace(51 + 4 * LAYER);
}

__global__ void __launch_bounds__(THREADS, 1) decode_tc2_kernel(const unsigned char* __restrict__ image, const float* __restrict__ P, DecodeArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = sbase + D2_OFF_BAR;
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + D2_OFF_BAR + 176);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int64_t n_total = a.n_dev ? (int64_t)*a.n_dev : a.n;
    const int64_t n_tiles = (n_total + TILE - 1) / TILE;
    const bool timing = g_tc_timing != nullptr;
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tcur = timing ? clock64() : 0;
    int tn = 0; (void)tn;

    if (threadIdx.x == 0) {
        mbar_init(bar0 + 8 * D2_W, 1); mbar_init(bar0 + 8 * D2_W1, 1); mbar_init(bar0 + 8 * D2_W2, 1); mbar_init(bar0 + 8 * D2_W3, 1);
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar0 + 8 * (D2_X + s), 1);          // producer warp -> issuer: layer-0 A tile ready
            mbar_init(bar0 + 8 * (D2_XF + s), 8);         // epilogue warps -> producer: x tile parked in TMEM, buffer free
            mbar_init(bar0 + 8 * (D2_E + s), 8);          // epilogue warps -> issuer: heads have read R1
            mbar_init(bar0 + 8 * (D2_ACCA + s), 1); mbar_init(bar0 + 8 * (D2_ACCB + s), 1);
#pragma unroll
            for (int g = 0; g < 4; ++g) mbar_init(bar0 + 8 * (D2_G + 4 * s + g), 4);
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

#if DIF_D2_ISSUERS == 1
    if (warp == MMA_WARP) {
        // ===================================================== MMA issuer: ONE warp, the two slots STRICTLY ALTERNATING layer by layer
        // (L0 A, L0 B, L1 A, L1 B, ...), every hand-off group awaited with a blocking wait in a fixed order.  With one issuer per
        // slot the two tiles ran in lockstep (a shared in-order pipe synchronises its clients: whoever is behind gets the pipe
        // alone and catches up), so both slots converted at the same time and both wanted the tensor pipe at the same time -
        // measured ~5 k cycles per layer pair for 3.1 k cycles of MMA work.  Alternation puts them in anti-phase: while slot A's
        // layer l+1 is issued group by group behind its epilogue, slot B's epilogue warps have the ALUs to themselves, and vice versa.
        // The weight image arrives in four groups with their own barriers (biases + W0, W1, W2, W3: as icp_tc2_kernel): L0 of the first
        // tile starts after 18 KB of the 198 KB; the rest streams in behind the first layers (a small batch is ONE round of tiles, so
        // the prologue is a large part of the launch).
        if (lane == 0) {
            mbar_expect_tx(bar0 + 8 * D2_W, BIAS_B + 2 * W0_B);
            bulk_g2s(sbase + OFF_BIAS, image + OFF_BIAS, BIAS_B, bar0 + 8 * D2_W);
            bulk_g2s(sbase + OFF_W0, image + OFF_W0, W0_B, bar0 + 8 * D2_W);
            bulk_g2s(sbase + PLANE_B + OFF_W0, image + PLANE_B + OFF_W0, W0_B, bar0 + 8 * D2_W);
            mbar_expect_tx(bar0 + 8 * D2_W1, 2 * W1_B);
            bulk_g2s(sbase + OFF_W1, image + OFF_W1, W1_B, bar0 + 8 * D2_W1);
            bulk_g2s(sbase + PLANE_B + OFF_W1, image + PLANE_B + OFF_W1, W1_B, bar0 + 8 * D2_W1);
            mbar_expect_tx(bar0 + 8 * D2_W2, 2 * W2_B);
            bulk_g2s(sbase + OFF_W2, image + OFF_W2, W2_B, bar0 + 8 * D2_W2);
            bulk_g2s(sbase + PLANE_B + OFF_W2, image + PLANE_B + OFF_W2, W2_B, bar0 + 8 * D2_W2);
            mbar_expect_tx(bar0 + 8 * D2_W3, 2 * W3_B);
            bulk_g2s(sbase + OFF_W3, image + OFF_W3, W3_B, bar0 + 8 * D2_W3);
            bulk_g2s(sbase + PLANE_B + OFF_W3, image + PLANE_B + OFF_W3, W3_B, bar0 + 8 * D2_W3);
        }
        __syncwarp();
        mbar_wait(bar0 + 8 * D2_W, 0);
        TC_ACC(0, tcur);                                   // [0] first weight group
        const uint64_t w0h = smem_desc(sbase + OFF_W0, 128 * 16, 128), w0l = smem_desc(sbase + PLANE_B + OFF_W0, 128 * 16, 128);
        const uint64_t w1h = smem_desc(sbase + OFF_W1, 128 * 16, 128), w1l = smem_desc(sbase + PLANE_B + OFF_W1, 128 * 16, 128);
        const uint64_t w2h = smem_desc(sbase + OFF_W2, 96 * 16, 128), w2l = smem_desc(sbase + PLANE_B + OFF_W2, 96 * 16, 128);
        const uint64_t w3h = smem_desc(sbase + OFF_W3, 128 * 16, 128), w3l = smem_desc(sbase + PLANE_B + OFF_W3, 128 * 16, 128);
        const uint64_t xh0 = smem_desc(sbase + OFF_X, X_CHUNK_B, 128), xl0 = smem_desc(sbase + OFF_X + X_PLANE_B, X_CHUNK_B, 128);
        const uint64_t xh1 = smem_desc(sbase + OFF_X + 2 * X_PLANE_B, X_CHUNK_B, 128), xl1 = smem_desc(sbase + OFF_X + 3 * X_PLANE_B, X_CHUNK_B, 128);
        const uint32_t A0 = tmem, A1 = tmem + 128, B0 = tmem + 256, B1 = tmem + 384;              // slot A: R0 / R1, slot B: R0 / R1
        const uint32_t bgA = bar0 + 8 * D2_G, bgB = bar0 + 8 * (D2_G + 4);
        uint32_t ph_eA = 0, ph_eB = 0, ph_g = 0;           // ph_g: all group barriers flip once per hidden layer, both slots in step
        auto trace = [&](int ev) { TC_TRACE(ev); TC_ACC((ev < 48 ? 1 : 2), tcur); };
        // L0 of a slot's NEXT tile is issued right behind its L3 when the producer has the x tile ready by then (one probe, no
        // wait): R0 is free as soon as L3 is in the in-order pipe, and the tile boundary (heads + refill) stops being a bubble.
        bool l0A = false, l0B = false;                     // L0 of the current iteration already issued (early, in the previous one)
        auto issue_l0 = [&](const int s) {
            tc_fence_after();
            if (elect_one()) {
                if (s == 0) { issue_layer(idesc_f16(128), 128 * 16, 0, 2, A0, 0, 0, xh0, xl0, w0h, w0l); mma_commit(bar0 + 8 * D2_ACCA); }
                else { issue_layer(idesc_f16(128), 128 * 16, 0, 2, B0, 0, 0, xh1, xl1, w0h, w0l); mma_commit(bar0 + 8 * (D2_ACCA + 1)); }
            }
            __syncwarp();
        };
        uint32_t ph_xA = 0, ph_xB = 0;
        for (int64_t it = 0;; ++it) {
            const int64_t t0 = blockIdx.x + (int64_t)gridDim.x * (2 * it), t1 = t0 + gridDim.x;
            if (t0 >= n_tiles) break;
            const bool liveB = t1 < n_tiles;
            const bool nextA = t0 + 2 * (int64_t)gridDim.x < n_tiles, nextB = t1 + 2 * (int64_t)gridDim.x < n_tiles;
            // ---- L0: x tile (smem) -> R0
            if (!l0A) {
                mbar_wait_spin(bar0 + 8 * D2_X, ph_xA); ph_xA ^= 1;
                TC_TRACE(1); TC_ACC(1, tcur);
                issue_l0(0);
                TC_TRACE(2); TC_ACC(2, tcur);
            }
            if (liveB && !l0B) {
                mbar_wait_spin(bar0 + 8 * (D2_X + 1), ph_xB); ph_xB ^= 1;
                TC_ACC(1, tcur);
                issue_l0(1);
                TC_ACC(2, tcur);
            }
            l0A = l0B = false;
            // ---- L1 (R0 -> R1), L2 (R1 -> R0, N = 96), L3 (R0 -> R1): four hand-off groups each, constant operands
            if (it == 0) mbar_wait_spin(bar0 + 8 * D2_W1, 0);
            issue_hidden_layer<1>(A0, A1, w1h, w1l, bgA, ph_g, bar0 + 8 * D2_ACCB, bar0 + 8 * D2_E, it > 0, ph_eA, trace);
            if (liveB) issue_hidden_layer<1>(B0, B1, w1h, w1l, bgB, ph_g, bar0 + 8 * (D2_ACCB + 1), bar0 + 8 * (D2_E + 1), it > 0, ph_eB, trace);
            if (it == 0) mbar_wait_spin(bar0 + 8 * D2_W2, 0);
            issue_hidden_layer<2>(A0, A1, w2h, w2l, bgA, ph_g ^ 1u, bar0 + 8 * D2_ACCA, 0u, false, ph_eA, trace);
            if (liveB) issue_hidden_layer<2>(B0, B1, w2h, w2l, bgB, ph_g ^ 1u, bar0 + 8 * (D2_ACCA + 1), 0u, false, ph_eB, trace);
            if (it == 0) mbar_wait_spin(bar0 + 8 * D2_W3, 0);
            issue_hidden_layer<3>(A0, A1, w3h, w3l, bgA, ph_g, bar0 + 8 * D2_ACCB, 0u, false, ph_eA, trace);
            if (nextA && __all_sync(0xffffffffu, mbar_test(bar0 + 8 * D2_X, ph_xA))) { ph_xA ^= 1; TC_TRACE(1); issue_l0(0); TC_TRACE(2); l0A = true; }
            if (liveB) {
                issue_hidden_layer<3>(B0, B1, w3h, w3l, bgB, ph_g, bar0 + 8 * (D2_ACCB + 1), 0u, false, ph_eB, trace);
                if (nextB && __all_sync(0xffffffffu, mbar_test(bar0 + 8 * (D2_X + 1), ph_xB))) { ph_xB ^= 1; issue_l0(1); l0B = true; }
            }
            ph_g ^= 1;
        }
#else
    if (warp == MMA_WARP || warp == MMA_WARP2) {
        // ===================================================== MMA issuer of slot s (warp 16: slot 0 + weight load, warp 19: slot 1)
        // One issuer warp PER SLOT, each walking its own tiles with blocking waits in a fixed order: a single issuer polling both
        // slots' hand-off barriers (mbarrier.test_wait ~150 cycles a probe) spent ~740 cycles per 6-MMA group, twice the 384
        // cycles the tensor pipe needs for them, and became the bottleneck.  The whole warp runs the loop with warp-uniform
        // values; only the tcgen05 instructions are executed by one elected lane.
        const int s = warp == MMA_WARP ? 0 : 1;
        if (warp == MMA_WARP && lane == 0) {
            mbar_expect_tx(bar0 + 8 * D2_W, IMAGE_B);
            constexpr uint32_t CH = 32768;
            for (uint32_t off = 0; off < IMAGE_B; off += CH) bulk_g2s(sbase + off, image + off, (IMAGE_B - off) < CH ? (IMAGE_B - off) : CH, bar0 + 8 * D2_W);
        }
        __syncwarp();
        mbar_wait(bar0 + 8 * D2_W, 0);
        TC_ACC(0, tcur);                                   // [0] weight image load
        const uint64_t w0h = smem_desc(sbase + OFF_W0, 128 * 16, 128), w0l = smem_desc(sbase + PLANE_B + OFF_W0, 128 * 16, 128);
        const uint64_t w1h = smem_desc(sbase + OFF_W1, 128 * 16, 128), w1l = smem_desc(sbase + PLANE_B + OFF_W1, 128 * 16, 128);
        const uint64_t w2h = smem_desc(sbase + OFF_W2, 96 * 16, 128), w2l = smem_desc(sbase + PLANE_B + OFF_W2, 96 * 16, 128);
        const uint64_t w3h = smem_desc(sbase + OFF_W3, 128 * 16, 128), w3l = smem_desc(sbase + PLANE_B + OFF_W3, 128 * 16, 128);
        const uint64_t xh = smem_desc(sbase + OFF_X + s * 2 * X_PLANE_B, X_CHUNK_B, 128);
        const uint64_t xl = smem_desc(sbase + OFF_X + s * 2 * X_PLANE_B + X_PLANE_B, X_CHUNK_B, 128);
        const uint32_t R0 = tmem + s * 256, R1 = R0 + 128;
        const uint32_t bx = bar0 + 8 * (D2_X + s), be = bar0 + 8 * (D2_E + s), bg = bar0 + 8 * (D2_G + 4 * s);
        const uint32_t ba = bar0 + 8 * (D2_ACCA + s), bb = bar0 + 8 * (D2_ACCB + s);
        uint32_t ph_x = 0, ph_e = 0, ph_g = 0;             // ph_g: the 4 group barriers flip together (once per hidden layer)
        int64_t it = 0;
        for (int64_t tile = blockIdx.x + (int64_t)gridDim.x * s; tile < n_tiles; tile += 2 * (int64_t)gridDim.x, ++it) {
            // ---- L0: x tile (smem) -> R0
            mbar_wait_spin(bx, ph_x); ph_x ^= 1;
            TC_TRACE(1);
            TC_ACC(1, tcur);                               // [1] waiting for operands
            tc_fence_after();
            if (elect_one()) { issue_layer(idesc_f16(128), 128 * 16, 0, 2, R0, 0, 0, xh, xl, w0h, w0l); mma_commit(ba); }
            __syncwarp();
            TC_TRACE(2);
            TC_ACC(2, tcur);                               // [2] issuing
            // ---- L1 (R0 -> R1), L2 (R1 -> R0, N = 96), L3 (R0 -> R1): four hand-off groups each, fully unrolled with constant operands
            auto trace = [&](int ev) { TC_TRACE(ev); TC_ACC((ev < 48 ? 1 : 2), tcur); };
            issue_hidden_layer<1>(R0, R1, w1h, w1l, bg, ph_g, bb, be, it > 0, ph_e, trace);
            issue_hidden_layer<2>(R0, R1, w2h, w2l, bg, ph_g ^ 1u, ba, be, false, ph_e, trace);
            issue_hidden_layer<3>(R0, R1, w3h, w3l, bg, ph_g, bb, be, false, ph_e, trace);
            ph_g ^= 1;
        }
#endif
    } else if (warp >= PRODUCER_WARP0) {
        // ===================================================== gather producer of slot s: one sample per lane, two stages
        // Stage A resolves WHERE a tile's 128 samples come from (latent row index + xyz: 4 samples x 4 registers per lane) one
        // tile ahead; stage B fetches the latent rows, two 32-sample passes in flight.  The first stage reads streamed arrays
        // (DRAM latency), the second the L2-resident table; chained per pass they put ~1.5 k cycles x 4 passes between "x tile
        // free" and "x tile ready" (measured: the issuer waited ~2.2 k cycles per tile for the producer).  Resolved one tile
        // ahead, only two L2 round trips remain behind the hand-off.
        const int s = warp - PRODUCER_WARP0;
        if (s < 2) {
            unsigned char* x_hi_p = smem + OFF_X + s * 2 * X_PLANE_B;
            const int n3 = a.lat_n * a.lat_n * a.lat_n;
            const float inv_n = 1.0f / (float)(a.lat_n > 0 ? a.lat_n : 1);
            int r0, r1, r2, r3;                              // latent row of the lane's sample in pass 0..3 (-1: padding)
            float3 p0, p1, p2, p3;                           // its xyz
            auto resolve = [&](int64_t sidx, int& r, float3& p) {
                int64_t row, out; int li;
                decode_sample_source(a, sidx, n_total, n3, row, out, li);
                r = (int)row;
                if (a.mode == 0) {
                    const float* xp = a.xyz + (sidx < n_total ? sidx : 0) * 3;      // (address independent of the loaded row: no load chain)
                    p = make_float3(__ldg(xp), __ldg(xp + 1), __ldg(xp + 2));
                } else {
                    const int q1 = (int)(((float)li + 0.5f) * inv_n), q2 = (int)(((float)q1 + 0.5f) * inv_n);
                    p = make_float3(lattice_coord(a, q2), lattice_coord(a, q1 - q2 * a.lat_n), lattice_coord(a, li - q1 * a.lat_n));
                }
            };
            auto fetch = [&](int r, const float3& p, float (&x)[32]) {
                load_latent_row(a.latent, r >= 0 ? r : 0, a.lat_stride, x);
                x[29] = p.x; x[30] = p.y; x[31] = p.z;
            };
            float xa[32], xb[32];
            uint32_t ph_xf = 0;
            int64_t tile = blockIdx.x + (int64_t)gridDim.x * s;
            if (tile < n_tiles) {
                resolve(tile * TILE + lane, r0, p0); resolve(tile * TILE + 32 + lane, r1, p1);
                resolve(tile * TILE + 64 + lane, r2, p2); resolve(tile * TILE + 96 + lane, r3, p3);
            }
            for (int64_t it = 0; tile < n_tiles; ++it) {
                fetch(r0, p0, xa); fetch(r1, p1, xb);
                if (it > 0) { mbar_wait(bar0 + 8 * (D2_XF + s), ph_xf); ph_xf ^= 1; }
                TC_TRACE(120);
                gather_store_row(xa, r0 >= 0, lane, x_hi_p);
                TC_TRACE(122);
                fetch(r2, p2, xa);
                gather_store_row(xb, r1 >= 0, 32 + lane, x_hi_p);
                TC_TRACE(123);
                fetch(r3, p3, xb);
                const bool v2 = r2 >= 0, v3 = r3 >= 0;
                const int64_t next = tile + 2 * (int64_t)gridDim.x;
                if (next < n_tiles) {                        // stage A of the next tile: in flight while this tile is stored and computed
                    resolve(next * TILE + lane, r0, p0); resolve(next * TILE + 32 + lane, r1, p1);
                    resolve(next * TILE + 64 + lane, r2, p2); resolve(next * TILE + 96 + lane, r3, p3);
                }
                TC_TRACE(124);
                gather_store_row(xa, v2, 64 + lane, x_hi_p);
                TC_TRACE(125);
                gather_store_row(xb, v3, 96 + lane, x_hi_p);
                TC_TRACE(126);
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar0 + 8 * (D2_X + s));
                TC_TRACE(121);
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
        const int n3 = a.lat_n * a.lat_n * a.lat_n;
        uint32_t ph_a = 0, ph_b = 0;
        const uint32_t acc_a = bar0 + 8 * (D2_ACCA + s), acc_b = bar0 + 8 * (D2_ACCB + s);
        const uint32_t g_bar = bar0 + 8 * (D2_G + 4 * s + half);                     // + 16 bytes for iteration 1
        const int head_slot = *reinterpret_cast<const int*>(image + IMAGE_B);
        const float head_bias = __ldg(P + (half ? DecW::bu : DecW::b4));
        mbar_wait(bar0 + 8 * D2_W, 0);                 // biases arrive with the weight image
        float pend_pre = 0.f; int64_t pend_tile = -1;  // outputs of the previous tile, not yet written
        auto finish_output = [&](int64_t t, float pre) {
            const int64_t sidx = t * TILE + row;
            int64_t src_row, out; int li_unused;
            decode_sample_source(a, sidx, n_total, n3, src_row, out, li_unused);
            float* dst = half ? a.std : a.sdf;
            if (src_row >= 0) dst[out] = half ? 0.05f + 0.5f * softplus_ref(pre) : a.sdf_sign * tanhf(pre);
            else if (sidx < n_total && a.mode == 0 && !a.out_index) dst[out] = 0.f;
        };
        for (int64_t it = 0;; ++it) {
            const int64_t tile = blockIdx.x + (int64_t)gridDim.x * (2 * it + s);
            if (tile >= n_tiles) break;
            TC_ACC(7, tcur);
#pragma unroll 1
            for (int layer = 0; layer < 3; ++layer) {
                const uint32_t R = layer == 1 ? R1 : R0;
                const int hw = layer == 2 ? 48 : 64;
                const int c_base = half * hw;
                const float* b = bias + layer * 128 + c_base;
                if (layer == 1) { mbar_wait(acc_b, ph_b); ph_b ^= 1; } else { mbar_wait(acc_a, ph_a); ph_a ^= 1; }
                TC_TRACE(80 + layer);
                TC_ACC(4, tcur);                               // [4] waiting for the accumulator
                tc_fence_after();
                if (layer == 1) {
                    // Layer 1 has completed, so R0 (its A operand) is dead until layer 2 writes columns 0..95: park this half's 16 inputs
                    // (k = 16*half .. +15: hi plane k-chunks 2*half, 2*half+1, then the lo plane's) in R0[96 + 16*half ..) NOW - they are the
                    // skip-connection K chunks 6 / 7 of layer 3 - and hand the x tile back to the producer three layers before it is
                    // needed again (parked at E2 the refill came ~3 k cycles too late for the next tile's layer 0).
                    const unsigned char* xp = smem + OFF_X + s * 2 * X_PLANE_B + (2 * half) * X_CHUNK_B + row * 16;
                    const uint4 h0 = *reinterpret_cast<const uint4*>(xp), h1 = *reinterpret_cast<const uint4*>(xp + X_CHUNK_B);
                    const uint4 l0 = *reinterpret_cast<const uint4*>(xp + X_PLANE_B), l1 = *reinterpret_cast<const uint4*>(xp + X_PLANE_B + X_CHUNK_B);
                    const uint32_t xv[16] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w, l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
                    tmem_st16(R0 + 96 + 16 * half, xv);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar0 + 8 * (D2_XF + s));       // our generic-proxy reads of the x tile are complete
                }
                // ---- group 0: 32 columns = 2 K chunks
                {
                    uint32_t v0[16], v1[16];
                    tmem_ld16_nowait(R + c_base, v0);
                    tmem_ld16_nowait(R + c_base + 16, v1);
                    tmem_ld_wait();
                    convert_inplace16(v0, b, R + c_base);
                    convert_inplace16(v1, b + 16, R + c_base + 16);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(g_bar);
                TC_TRACE(90 + layer);
                // ---- group 1: layers 0/1: 32 more columns; layer 2: 16 columns + this half's 16 parked inputs (skip connection)
                if (layer < 2) {
                    uint32_t v0[16], v1[16];
                    tmem_ld16_nowait(R + c_base + 32, v0);
                    tmem_ld16_nowait(R + c_base + 48, v1);
                    tmem_ld_wait();
                    convert_inplace16(v0, b + 32, R + c_base + 32);
                    convert_inplace16(v1, b + 48, R + c_base + 48);
                } else {
                    uint32_t v0[16];
                    tmem_ld16_nowait(R + c_base + 32, v0);
                    tmem_ld_wait();
                    convert_inplace16(v0, b + 32, R + c_base + 32);           // (this group's second K chunk, the parked inputs, is in place since E1)
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(g_bar + 16);
                TC_TRACE(100 + layer);
                TC_ACC(5, tcur);                               // [5] hidden-layer conversion
                if (layer == 0 && pend_tile >= 0) { finish_output(pend_tile, pend_pre); pend_tile = -1; }
            }
            // ---- layer 3 + one head per column-half warp on CUDA cores (half 0 -> sdf, half 1 -> std; di_decoder.py:65-70,84)
            mbar_wait(acc_b, ph_b); ph_b ^= 1;
            TC_TRACE(83);
            TC_ACC(4, tcur);
            tc_fence_after();
            float p0 = 0.f, p1 = 0.f;
            const float* hw_c = c_head_w[head_slot][half];
#pragma unroll 1
            for (int c0 = 0; c0 < 128; c0 += 32) {
                uint32_t v0[16], v1[16];
                tmem_ld16_nowait(R1 + c0, v0);
                tmem_ld16_nowait(R1 + c0 + 16, v1);
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
            tc_fence_before();                               // R1 reads complete: the next tile's layer 1 may overwrite it
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + 8 * (D2_E + s));
            TC_TRACE(110);
            // the activation + store of this tile's outputs is deferred until the NEXT tile's first conversion is handed off: the
            // chain heads(t) -> E0(t+1) -> L1(t+1) is what the tensor pipe waits for at a tile boundary
            pend_pre = p0 + p1 + head_bias; pend_tile = tile;
            TC_ACC(6, tcur);                                   // [6] last layer + heads
        }
        if (pend_tile >= 0) finish_output(pend_tile, pend_pre);
    }
#ifndef DIF_TC_TRACE
    if (timing && lane == 0) {
        for (int k = 0; k < 8; ++k) g_tc_timing[((size_t)blockIdx.x * 20 + warp) * 8 + k] = (unsigned long long)tacc[k];
    }
#endif
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(512));
}

}  // namespace tc
}  // namespace dif
