// fp32 SIMT evaluation of the DI-Fusion encoder / decoder MLPs on tiles of T samples held in shared memory.
// This is the exact-fp32 path: it is the numerical yardstick for the tensor-core (tcgen05) kernels and the fallback
// for sample counts too small to fill a 128-row MMA tile.
//
//   decoder: reference network/di_decoder.py:55-86 (eval, weight-norm folded) - SURVEY A.7
//   encoder: reference network/di_encoder.py:26-30 ('cnp', BN folded)          - SURVEY A.6
//
// Activations are k-major in shared memory: act[k][TP] with TP = T + 4 floats (16-byte aligned rows; the +4 makes
// the float4 column stores of consecutive neurons hit distinct banks).  One thread owns one output neuron for TS
// consecutive samples, so the weight is read once (coalesced over neurons, L1/L2 resident) and the activations are
// warp-broadcast float4 loads.
#pragma once
#include "common.cuh"

namespace dif {

constexpr int MLP_T = 32;                 // samples per tile
constexpr int MLP_TP = MLP_T + 4;         // padded row
constexpr int MLP_THREADS = 256;

// ------------------------------------------------------------------------------------------------------------
// prepared-weight layouts (float offsets)
struct DecW {       // decoder, see include/difusion_b200.h for the blob layout
    static constexpr int W0t = 0;                        // [32][128]   k-major (forward)
    static constexpr int W1t = W0t + 32 * 128;           // [128][128]
    static constexpr int W2t = W1t + 128 * 128;          // [128][96]
    static constexpr int W3t = W2t + 128 * 96;           // [128][128]
    static constexpr int W0 = W3t + 128 * 128;           // [128][32]   row-major (backward)
    static constexpr int W1 = W0 + 128 * 32;             // [128][128]
    static constexpr int W2 = W1 + 128 * 128;            // [96][128]
    static constexpr int W3 = W2 + 96 * 128;             // [128][128]
    static constexpr int b0 = W3 + 128 * 128;
    static constexpr int b1 = b0 + 128;
    static constexpr int b2 = b1 + 128;
    static constexpr int b3 = b2 + 96;
    static constexpr int w4 = b3 + 128;                  // [128] sdf head
    static constexpr int wu = w4 + 128;                  // [128] std head
    static constexpr int b4 = wu + 128;                  // [1]
    static constexpr int bu = b4 + 1;                    // [1]
    static constexpr int FP32_END = ((bu + 1 + 63) / 64) * 64;
};

struct EncW {
    static constexpr int W0t = 0;                        // [6][32]
    static constexpr int W1t = W0t + 6 * 32;             // [32][64]
    static constexpr int W2t = W1t + 32 * 64;            // [64][256]
    static constexpr int W3t = W2t + 64 * 256;           // [256][32]  (29 real columns, 3 zero)
    static constexpr int b0 = W3t + 256 * 32;
    static constexpr int b1 = b0 + 32;
    static constexpr int b2 = b1 + 64;
    static constexpr int b3 = b2 + 256;                  // [32] (29 real)
    static constexpr int FP32_END = ((b3 + 32 + 63) / 64) * 64;
};

constexpr int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// out[n][t] = epi(n, t, sum_k Wt[k*LDW + n] * in[k][t])   for n < N, t < T.  Ends WITHOUT a barrier.
template <int K, int N, int LDW, class Epi>
__device__ __forceinline__ void dense_tile(const float* __restrict__ Wt, const float* in, Epi epi) {
    constexpr int NP = next_pow2(N) < 32 ? 32 : next_pow2(N);
    constexpr int G = MLP_THREADS / NP;
    constexpr int TS = MLP_T / G;
    static_assert(TS >= 4 && TS % 4 == 0, "tile shape");
    const int n = threadIdx.x % NP, sg = threadIdx.x / NP;
    if (n >= N) return;
    float acc[TS];
#pragma unroll
    for (int t = 0; t < TS; ++t) acc[t] = 0.f;
    const float* col = in + sg * TS;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float w = __ldg(Wt + k * LDW + n);
        const float4* a4 = reinterpret_cast<const float4*>(col + k * MLP_TP);
#pragma unroll
        for (int q = 0; q < TS / 4; ++q) {
            const float4 a = a4[q];
            acc[4 * q + 0] = fmaf(w, a.x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(w, a.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(w, a.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(w, a.w, acc[4 * q + 3]);
        }
    }
#pragma unroll
    for (int q = 0; q < TS / 4; ++q) epi(n, sg * TS + 4 * q, make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]));
}

__device__ __forceinline__ float4 relu4(float4 v, float b) {
    return make_float4(fmaxf(v.x + b, 0.f), fmaxf(v.y + b, 0.f), fmaxf(v.z + b, 0.f), fmaxf(v.w + b, 0.f));
}
__device__ __forceinline__ void st4(float* base, int n, int t, float4 v) { *reinterpret_cast<float4*>(base + n * MLP_TP + t) = v; }
__device__ __forceinline__ float4 ld4(const float* base, int n, int t) { return *reinterpret_cast<const float4*>(base + n * MLP_TP + t); }

// torch.nn.functional.softplus (beta=1, threshold=20), di_decoder.py:68
__device__ __forceinline__ float softplus_ref(float x) { return x > 20.f ? x : log1pf(expf(x)); }

// ------------------------------------------------------------------------------------------------------------
// Decoder tile.  Shared memory: cat[128][TP] (rows 0..95 = h2, rows 96..127 = the 32 inputs: latent 0..28, x, y, z),
// h0, h1, h3 [128][TP] each, small[] for head partials.
struct DecoderSmem {
    float cat[128 * MLP_TP];
    float h0[128 * MLP_TP];
    float h1[128 * MLP_TP];
    float h3[128 * MLP_TP];
    float part[4 * 2 * MLP_T];      // head partial sums
    float pre[2 * MLP_T];           // pre-activations of the two heads
    float gx[3 * MLP_T];            // d F / d xyz
    float seed[MLP_T];              // dF/d(head pre-activation) per sample
};

// Forward on the tile whose inputs are already in s.cat rows 96..127.  On return s.pre holds (pre_sdf, pre_std) per
// sample and h0,h1,cat[0:96],h3 hold the post-ReLU activations.  Ends with a barrier.
__device__ __forceinline__ void decoder_forward_tile(const float* __restrict__ P, DecoderSmem& s) {
    const float* x = s.cat + 96 * MLP_TP;
    dense_tile<32, 128, 128>(P + DecW::W0t, x, [&](int n, int t, float4 v) { st4(s.h0, n, t, relu4(v, __ldg(P + DecW::b0 + n))); });
    __syncthreads();
    dense_tile<128, 128, 128>(P + DecW::W1t, s.h0, [&](int n, int t, float4 v) { st4(s.h1, n, t, relu4(v, __ldg(P + DecW::b1 + n))); });
    __syncthreads();
    dense_tile<128, 96, 96>(P + DecW::W2t, s.h1, [&](int n, int t, float4 v) { st4(s.cat, n, t, relu4(v, __ldg(P + DecW::b2 + n))); });
    __syncthreads();
    dense_tile<128, 128, 128>(P + DecW::W3t, s.cat, [&](int n, int t, float4 v) { st4(s.h3, n, t, relu4(v, __ldg(P + DecW::b3 + n))); });
    __syncthreads();
    {   // two heads, 4 partial chains of 32 each: thread = (part, head, t)
        const int t = threadIdx.x % MLP_T, head = (threadIdx.x / MLP_T) % 2, part = threadIdx.x / (2 * MLP_T);
        const float* w = P + (head ? DecW::wu : DecW::w4) + part * 32;
        const float* h = s.h3 + part * 32 * MLP_TP + t;
        float a = 0.f;
#pragma unroll 8
        for (int k = 0; k < 32; ++k) a = fmaf(__ldg(w + k), h[k * MLP_TP], a);
        s.part[(part * 2 + head) * MLP_T + t] = a;
    }
    __syncthreads();
    if (threadIdx.x < 2 * MLP_T) {
        const int t = threadIdx.x % MLP_T, head = threadIdx.x / MLP_T;
        float a = __ldg(P + (head ? DecW::bu : DecW::b4));
#pragma unroll
        for (int p = 0; p < 4; ++p) a += s.part[(p * 2 + head) * MLP_T + t];
        s.pre[head * MLP_T + t] = a;
    }
    __syncthreads();
}

// Backward of one head wrt the xyz inputs.  s.seed[t] = dF/d(pre_head); destroys h3,cat[0:96],h1,h0 (they become the
// masked gradients).  Result in s.gx[c][t] (c = x,y,z).  Ends with a barrier.
__device__ __forceinline__ void decoder_backward_tile(const float* __restrict__ P, DecoderSmem& s, int head) {
    {   // g3 = seed * w_head * relu'(h3)
        const float* w = P + (head ? DecW::wu : DecW::w4);
        for (int i = threadIdx.x; i < 128 * MLP_T; i += MLP_THREADS) {
            const int n = i / MLP_T, t = i % MLP_T;
            float* p = s.h3 + n * MLP_TP + t;
            *p = (*p > 0.f) ? s.seed[t] * __ldg(w + n) : 0.f;
        }
    }
    __syncthreads();
    // [g2 ; gx(skip)] = W3^T g3, masked by relu'(h2) on the first 96 rows; rows 125..127 are the xyz columns of the skip
    dense_tile<128, 128, 128>(P + DecW::W3, s.h3, [&](int k, int t, float4 v) {
        if (k < 96) {
            const float4 h = ld4(s.cat, k, t);
            st4(s.cat, k, t, make_float4(h.x > 0.f ? v.x : 0.f, h.y > 0.f ? v.y : 0.f, h.z > 0.f ? v.z : 0.f, h.w > 0.f ? v.w : 0.f));
        } else if (k >= 96 + 29) {
            *reinterpret_cast<float4*>(s.gx + (k - 125) * MLP_T + t) = v;
        }
    });
    __syncthreads();
    dense_tile<96, 128, 128>(P + DecW::W2, s.cat, [&](int k, int t, float4 v) {
        const float4 h = ld4(s.h1, k, t);
        st4(s.h1, k, t, make_float4(h.x > 0.f ? v.x : 0.f, h.y > 0.f ? v.y : 0.f, h.z > 0.f ? v.z : 0.f, h.w > 0.f ? v.w : 0.f));
    });
    __syncthreads();
    dense_tile<128, 128, 128>(P + DecW::W1, s.h1, [&](int k, int t, float4 v) {
        const float4 h = ld4(s.h0, k, t);
        st4(s.h0, k, t, make_float4(h.x > 0.f ? v.x : 0.f, h.y > 0.f ? v.y : 0.f, h.z > 0.f ? v.z : 0.f, h.w > 0.f ? v.w : 0.f));
    });
    __syncthreads();
    if (threadIdx.x < 3 * MLP_T) {   // gx[c] += sum_n W0[n][29+c] g0[n]
        const int t = threadIdx.x % MLP_T, c = threadIdx.x / MLP_T;
        const float* w = P + DecW::W0 + 29 + c;
        float a0 = 0.f, a1 = 0.f;
#pragma unroll 4
        for (int n = 0; n < 128; n += 2) {
            a0 = fmaf(__ldg(w + n * 32), s.h0[n * MLP_TP + t], a0);
            a1 = fmaf(__ldg(w + (n + 1) * 32), s.h0[(n + 1) * MLP_TP + t], a1);
        }
        s.gx[c * MLP_T + t] += a0 + a1;
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------------------
// Encoder tile: in[6][TP] -> out[32][TP] (29 real rows).
struct EncoderSmem {
    float in[8 * MLP_TP];
    float a0[32 * MLP_TP];
    float a1[64 * MLP_TP];
    float a2[256 * MLP_TP];
    float out[32 * MLP_TP];
};

__device__ __forceinline__ void encoder_forward_tile(const float* __restrict__ P, EncoderSmem& s) {
    dense_tile<6, 32, 32>(P + EncW::W0t, s.in, [&](int n, int t, float4 v) { st4(s.a0, n, t, relu4(v, __ldg(P + EncW::b0 + n))); });
    __syncthreads();
    dense_tile<32, 64, 64>(P + EncW::W1t, s.a0, [&](int n, int t, float4 v) { st4(s.a1, n, t, relu4(v, __ldg(P + EncW::b1 + n))); });
    __syncthreads();
    dense_tile<64, 256, 256>(P + EncW::W2t, s.a1, [&](int n, int t, float4 v) { st4(s.a2, n, t, relu4(v, __ldg(P + EncW::b2 + n))); });
    __syncthreads();
    dense_tile<256, 32, 32>(P + EncW::W3t, s.a2, [&](int n, int t, float4 v) {
        const float b = __ldg(P + EncW::b3 + n);
        st4(s.out, n, t, make_float4(v.x + b, v.y + b, v.z + b, v.w + b));
    });
    __syncthreads();
}

}  // namespace dif
