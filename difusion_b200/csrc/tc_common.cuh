// Shared building blocks of the tcgen05 / TMEM MLP kernels (decode_tc.cu, icp_tc.cu): shared-memory image layout of the
// decoder weights, PTX wrappers (mbarrier, cp.async.bulk, tcgen05.mma/ld/st/commit), UMMA descriptors, fp16 hi/lo splitting.
#pragma once
#include <cuda_fp16.h>

#include "decode_args.cuh"
#include "mlp_simt.cuh"

namespace dif {
namespace tc {


constexpr int TILE = 128;
constexpr int THREADS = 20 * 32;
constexpr int MMA_WARP = 16;
constexpr int PRODUCER_WARP0 = 17;

// ---- byte layout of the tensor-core section of the prepared decoder buffer == its image in shared memory ----------
constexpr uint32_t W0_B = 128 * 32 * 2, W1_B = 128 * 128 * 2, W2_B = 96 * 128 * 2, W3_B = 128 * 128 * 2;
constexpr uint32_t OFF_W0 = 0, OFF_W1 = OFF_W0 + W0_B, OFF_W2 = OFF_W1 + W1_B, OFF_W3 = OFF_W2 + W2_B;
constexpr uint32_t PLANE_B = OFF_W3 + W3_B;                 // 98304: one precision plane (hi or lo)
constexpr uint32_t OFF_BIAS = 2 * PLANE_B;                  // b0[128] b1[128] b2[96] b3[128] fp32
constexpr uint32_t BIAS_B = 480 * 4;
constexpr uint32_t IMAGE_B = OFF_BIAS + BIAS_B;             // 198528 bytes copied global -> shared per CTA
constexpr uint32_t OFF_X = IMAGE_B;                         // per slot: hi [4 k-chunks][128 rows][8 halves], then lo
constexpr uint32_t X_CHUNK_B = TILE * 16 + 16;              // 2064: +16 B skews the banks of the 4 k-chunks (conflict-free gather stores)
constexpr uint32_t X_PLANE_B = 4 * X_CHUNK_B;               // 8256 (hi -> lo plane lands 16 banks away)
constexpr uint32_t OFF_BAR = OFF_X + 4 * X_PLANE_B;
constexpr uint32_t SMEM_B = OFF_BAR + 96 + 16;            // 11 barriers (88 B) + TMEM base pointer
static_assert(SMEM_B <= 232448, "shared memory budget");
static_assert(IMAGE_B % 16 == 0, "bulk copy granularity");

enum { BAR_W = 0, BAR_X0 = 1, BAR_X1 = 2, BAR_ACC0 = 3, BAR_ACC1 = 4, BAR_A0 = 5, BAR_A1 = 6, BAR_XF0 = 7, BAR_XF1 = 8, BAR_E0 = 9, BAR_E1 = 10 };

// Head weights (sdf head w4, std head wu) of up to 8 prepared decoders live in the constant bank, so the last layer's dot
// products use them as FFMA operands.  A prepared decoder records its slot in the word that follows its weight image.
constexpr int HEAD_SLOTS = 8;
__constant__ float c_head_w[HEAD_SLOTS][2][128];
// Backward seed of the ICP kernel: G3_SCALE * w4 as fp16 pair words (pair p = columns 2p, 2p+1): [0,64) hi (truncated to 11
// bits), [64,128) lo, [128,192) rounded to nearest (single-pass variant).  The power-of-two scale keeps the small gradients of
// the deeper backward stages inside fp16's normal range; it is divided out exactly with the per-row seed.
constexpr float G3_SCALE = 16.f;
__constant__ uint32_t c_head_g3[HEAD_SLOTS][192];

// ---- PTX wrappers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
// Waiting warps must not steal issue / MIO slots from the working warps of their sub-partition: the hardware suspend of
// try_wait is only ~40 cycles, so a bare retry loop makes 16+ idle warps hammer the barrier (ncu: 62 M try_wait executions
// per launch, producer warp at 0.1 IPC).  The loop therefore backs off with nanosleep between probes.
template <int SLEEP_NS>
__device__ __forceinline__ void mbar_wait_ns(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE;\n\t"
                 "RETRY:\n\t"
                 "nanosleep.u32 %2;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@!p bra RETRY;\n\t"
                 "DONE:\n\t}\n" :: "r"(bar), "r"(parity), "n"(SLEEP_NS) : "memory");
}
// Alternative: let the hardware suspend the thread on the barrier (try_wait with a suspend-time hint): no polling at all and
// the wake-up comes with the phase flip.
__device__ __forceinline__ void mbar_wait_suspend(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "WAIT:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
                 "@!p bra WAIT;\n\t}\n" :: "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}
#ifndef DIF_WAIT_MODE
#define DIF_WAIT_MODE 0
#endif
#if DIF_WAIT_MODE == 0
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) { mbar_wait_ns<96>(bar, parity); }
__device__ __forceinline__ void mbar_wait_tight(uint32_t bar, uint32_t parity) { mbar_wait_ns<20>(bar, parity); }   // MMA issuer
#elif DIF_WAIT_MODE == 1
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) { mbar_wait_ns<32>(bar, parity); }
__device__ __forceinline__ void mbar_wait_tight(uint32_t bar, uint32_t parity) { mbar_wait_ns<20>(bar, parity); }
#elif DIF_WAIT_MODE == 2
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) { mbar_wait_suspend(bar, parity); }
__device__ __forceinline__ void mbar_wait_tight(uint32_t bar, uint32_t parity) { mbar_wait_suspend(bar, parity); }
#else
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) { mbar_wait_suspend(bar, parity); }
__device__ __forceinline__ void mbar_wait_tight(uint32_t bar, uint32_t parity) { mbar_wait_ns<20>(bar, parity); }
#endif
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
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

__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                 :: "r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
// descriptor + byte offset (the start-address field holds addr >> 4 in the low 14 bits; offsets never carry out of it)
__device__ __forceinline__ uint64_t desc_advance(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }
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
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t addr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                   "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                   "=r"(v[30]), "=r"(v[31]) : "r"(addr));
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t addr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(addr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
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

// ---- MMA issue for one layer of one slot (one elected lane) ------------------------------------------------------------
// A planes (hi, lo) either in TMEM (column address) or in shared memory (x tile); three passes hi*hi, lo*hi, hi*lo.
// Deliberately ROLLED loops: the whole kernel has to stay inside the instruction cache (an earlier fully unrolled version
// was 350 KB of SASS and ran every warp at ~0.1 IPC); per MMA the loop costs a handful of uniform-datapath instructions,
// far below the 48-64 cycles the tensor pipe needs per instruction.
__device__ __forceinline__ void issue_layer(uint32_t idesc, uint32_t wk_bytes, int ksteps_t, int ksteps_s, uint32_t acc,
                                            uint32_t a_hi_t, uint32_t a_lo_t, uint64_t x_hi_d, uint64_t x_lo_d, uint64_t w_hi_d, uint64_t w_lo_d) {
    const uint32_t w_step = (2 * wk_bytes) >> 4, x_step = (2 * X_CHUNK_B) >> 4;      // descriptor increments per K=16 step
    uint32_t accumulate = 0;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a_t = pass == 1 ? a_lo_t : a_hi_t;
        const uint64_t a_s = pass == 1 ? x_lo_d : x_hi_d;
        uint64_t w = pass == 2 ? w_lo_d : w_hi_d;
#pragma unroll 4
        for (int ks = 0; ks < ksteps_t; ++ks) { mma_ts(acc, a_t + ks * 8, w, idesc, accumulate); accumulate = 1; w += w_step; }
#pragma unroll 2
        for (int ks = 0; ks < ksteps_s; ++ks) { mma_ss(acc, a_s + (uint64_t)(ks * x_step), w, idesc, accumulate); accumulate = 1; w += w_step; }
    }
}

// ---- second-generation pipeline helpers (decode_tc2.cuh, encode_tc.cu): in-place conversion, group hand-off ----------------------
// issuer-side wait: no back-off - try_wait suspends the thread in hardware and wakes it ~60 cycles after the phase flips
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "WAITSPIN:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@!p bra WAITSPIN;\n\t}\n" :: "r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

// ReLU + fp16 hi/lo split of a pair with the ReLU folded into the two conversions (cvt.*.relu): for x >= 0 this is split_pair
// (hi = x truncated to 11 significant bits, lo = fp16(x - hi)); for x < 0 the truncated hi is <= 0 in magnitude order and
// x - hi is <= 0, so both conversions clamp to +0 - the same result as splitting max(x, 0), with two FMNMX fewer per pair.
__device__ __forceinline__ void relu_split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    const float ah = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u), bh = __uint_as_float(__float_as_uint(b) & 0xFFFFE000u);
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(bh), "f"(ah));          // (first source -> upper half)
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - bh), "f"(a - ah));
}

// 16 accumulator columns -> +bias, ReLU, hi/lo split -> the same 16 columns: [8 packed hi | 8 packed lo] = one K=16 A operand
__device__ __forceinline__ void convert_inplace16(const uint32_t* v, const float* b, uint32_t col_addr) {
    uint32_t o[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 bb = *reinterpret_cast<const float4*>(b + 4 * j);
        relu_split_pair(__uint_as_float(v[4 * j]) + bb.x, __uint_as_float(v[4 * j + 1]) + bb.y, o[2 * j], o[8 + 2 * j]);
        relu_split_pair(__uint_as_float(v[4 * j + 2]) + bb.z, __uint_as_float(v[4 * j + 3]) + bb.w, o[2 * j + 1], o[8 + 2 * j + 1]);
    }
    tmem_st16(col_addr, o);
}

// one hand-off group: 2 K steps x 3 passes (hi*hi, lo*hi, hi*lo) of one layer; A chunk ks lives at a_base + 16 ks (hi) / + 8 (lo).
// Everything but the bases is a compile-time constant: with run-time K-step indices every operand had to be moved into a uniform
// register right before its UTCHMMA and the issuer needed ~80-100 cycles per MMA (measured: 400-600 cycles per 6-MMA group against
// the 384 cycles the tensor pipe needs for them).
template <int N, int KS0, int KS1, bool FIRST>
__device__ __forceinline__ void issue_group(uint32_t acc, uint32_t a_base, uint64_t w_hi_d, uint64_t w_lo_d) {
    constexpr uint32_t idesc = idesc_f16(N);
    constexpr uint32_t step = (2 * N * 16) >> 4;                 // descriptor increment per K = 16 step
    mma_ts(acc, a_base + 16 * KS0, w_hi_d + (uint64_t)(KS0 * step), idesc, FIRST ? 0u : 1u);
    mma_ts(acc, a_base + 16 * KS1, w_hi_d + (uint64_t)(KS1 * step), idesc, 1u);
    mma_ts(acc, a_base + 16 * KS0 + 8, w_hi_d + (uint64_t)(KS0 * step), idesc, 1u);
    mma_ts(acc, a_base + 16 * KS1 + 8, w_hi_d + (uint64_t)(KS1 * step), idesc, 1u);
    mma_ts(acc, a_base + 16 * KS0, w_lo_d + (uint64_t)(KS0 * step), idesc, 1u);
    mma_ts(acc, a_base + 16 * KS1, w_lo_d + (uint64_t)(KS1 * step), idesc, 1u);
}

// Optional phase timing (tools/tc_timing.py): cycles spent per role and phase, summed per warp into a global buffer.
__device__ unsigned long long* g_tc_timing = nullptr;       // [gridDim][20 warps][8 counters]
#define TC_T0() const long long _t0 = timing ? clock64() : 0
#define TC_ACC(slot_idx, t_from) do { if (timing) { const long long _n = clock64(); tacc[slot_idx] += _n - (t_from); (t_from) = _n; } } while (0)


}  // namespace tc
}  // namespace dif
