// Latent optimisation: gradient of the reference's latent-refinement loss wrt the latent rows.
//   replaces the autograd pass of reference system/map.py:80-117 (OptimizeProcess.do_optimize: decoder forward through
//   network/utility.py:61-126 forward_model with loss_func, Normal log-likelihood, backward into latent_vecs_unique).
// The branch is disabled in the shipped loop (main.py:85-86), so this is the exact-fp32 SIMT tile (mlp_simt.cuh), not a tensor-core
// pipeline: forward -> per-sample loss seeds for BOTH heads -> backward wrt all 32 inputs -> atomic accumulation into the unique
// latent rows.  The Adam step on the (U, 29) rows stays in torch (plumbing on a few hundred kilobytes).
#include "mlp_simt.cuh"

namespace dif {

struct LatentOptSmem {
    DecoderSmem d;
    float gin[32 * MLP_TP];         // d loss / d input (latent 0..28, xyz 29..31)
    float seed_u[MLP_T];            // d loss / d pre_std (d.seed holds d loss / d pre_sdf)
};

// loss = sum_i -log N(clamp(gt_i); clamp(sdf_i), std_i) / n_div   (map.py:88-97; torch.distributions.Normal.log_prob)
__global__ void __launch_bounds__(MLP_THREADS) latent_grad_kernel(const float* __restrict__ P, const float* __restrict__ latent_u,
                                                                  const int64_t* __restrict__ inv, const float* __restrict__ rel_xyz,
                                                                  const float* __restrict__ gt_sdf, int n, float inv_div,
                                                                  float* __restrict__ grad_u, double* __restrict__ loss_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LatentOptSmem& s = *reinterpret_cast<LatentOptSmem*>(smem_raw);
    __shared__ int64_t row_s[MLP_T];
    double loss_acc = 0.0;                                   // thread t < T only
    const int n_tiles = (n + MLP_T - 1) / MLP_T;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int base = tile * MLP_T;
        if (threadIdx.x < MLP_T) { const int i = base + threadIdx.x; row_s[threadIdx.x] = i < n ? inv[i] : -1; }
        __syncthreads();
        for (int idx = threadIdx.x; idx < MLP_T * 32; idx += MLP_THREADS) {
            const int t = idx / 32, j = idx % 32;
            const int64_t r = row_s[t];
            float v = 0.f;
            if (r >= 0) v = j < DIF_L ? __ldg(latent_u + r * DIF_L + j) : __ldg(rel_xyz + (int64_t)(base + t) * 3 + (j - DIF_L));
            s.d.cat[(96 + j) * MLP_TP + t] = v;
        }
        __syncthreads();
        decoder_forward_tile(P, s.d);
        if (threadIdx.x < MLP_T) {
            const int t = threadIdx.x;
            float ss = 0.f, su = 0.f;
            if (row_s[t] >= 0) {
                const float pre_u = s.d.pre[MLP_T + t];
                const float sdf = tanhf(s.d.pre[t]);
                const float sd = 0.05f + 0.5f * softplus_ref(pre_u);
                const float gt = fminf(fmaxf(gt_sdf[base + t], -0.2f), 0.2f);
                const float sc = fminf(fmaxf(sdf, -0.2f), 0.2f);
                const float d = sc - gt;
                loss_acc += (double)(logf(sd) + 0.91893853320467274f + d * d / (2.f * sd * sd));
                const float dl_dsdf = (sdf >= -0.2f && sdf <= 0.2f) ? d / (sd * sd) : 0.f;        // clamp passes the gradient inside [min, max]
                const float dl_dstd = 1.f / sd - d * d / (sd * sd * sd);
                const float sig = pre_u > 20.f ? 1.f : 1.f / (1.f + expf(-pre_u));                 // softplus'(x), threshold 20
                ss = dl_dsdf * (1.f - sdf * sdf) * inv_div;
                su = dl_dstd * 0.5f * sig * inv_div;
            }
            s.d.seed[t] = ss; s.seed_u[t] = su;
        }
        __syncthreads();
        // ---- backward wrt all 32 inputs, both heads at once
        for (int i = threadIdx.x; i < 128 * MLP_T; i += MLP_THREADS) {
            const int nn = i / MLP_T, t = i % MLP_T;
            float* p = s.d.h3 + nn * MLP_TP + t;
            *p = (*p > 0.f) ? s.d.seed[t] * __ldg(P + DecW::w4 + nn) + s.seed_u[t] * __ldg(P + DecW::wu + nn) : 0.f;
        }
        __syncthreads();
        dense_tile<128, 128, 128>(P + DecW::W3, s.d.h3, [&](int k, int t, float4 v) {
            if (k < 96) {
                const float4 h = ld4(s.d.cat, k, t);
                st4(s.d.cat, k, t, make_float4(h.x > 0.f ? v.x : 0.f, h.y > 0.f ? v.y : 0.f, h.z > 0.f ? v.z : 0.f, h.w > 0.f ? v.w : 0.f));
            } else st4(s.gin, k - 96, t, v);                           // skip connection: the 32 inputs feed layer 3 directly
        });
        __syncthreads();
        dense_tile<96, 128, 128>(P + DecW::W2, s.d.cat, [&](int k, int t, float4 v) {
            const float4 h = ld4(s.d.h1, k, t);
            st4(s.d.h1, k, t, make_float4(h.x > 0.f ? v.x : 0.f, h.y > 0.f ? v.y : 0.f, h.z > 0.f ? v.z : 0.f, h.w > 0.f ? v.w : 0.f));
        });
        __syncthreads();
        dense_tile<128, 128, 128>(P + DecW::W1, s.d.h1, [&](int k, int t, float4 v) {
            const float4 h = ld4(s.d.h0, k, t);
            st4(s.d.h0, k, t, make_float4(h.x > 0.f ? v.x : 0.f, h.y > 0.f ? v.y : 0.f, h.z > 0.f ? v.z : 0.f, h.w > 0.f ? v.w : 0.f));
        });
        __syncthreads();
        dense_tile<128, 32, 32>(P + DecW::W0, s.d.h0, [&](int k, int t, float4 v) {
            const float4 g = ld4(s.gin, k, t);
            st4(s.gin, k, t, make_float4(g.x + v.x, g.y + v.y, g.z + v.z, g.w + v.w));
        });
        __syncthreads();
        for (int idx = threadIdx.x; idx < MLP_T * 32; idx += MLP_THREADS) {
            const int t = idx / 32, j = idx % 32;
            const int64_t r = row_s[t];
            if (r >= 0 && j < DIF_L) atomicAdd(grad_u + r * DIF_L + j, s.gin[j * MLP_TP + t]);
        }
        __syncthreads();
    }
    if (loss_out && threadIdx.x < MLP_T) {
        for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
        if (threadIdx.x == 0) atomicAdd(loss_out, loss_acc * (double)inv_div);
    }
}

}  // namespace dif

using namespace dif;

extern "C" int dif_latent_grad(const void* decoder_prepared, const float* latent_u, const int64_t* inv, const float* rel_xyz, const float* gt_sdf,
                               int64_t n, int64_t n_div, float* grad_u, double* loss_out, void* stream) {
    if (!decoder_prepared || n < 0 || n >= (int64_t(1) << 31) || n_div <= 0) return DIF_E_INVALID;
    if (n == 0) return DIF_OK;
    if (!latent_u || !inv || !rel_xyz || !gt_sdf || !grad_u) return DIF_E_INVALID;
    const size_t smem = sizeof(LatentOptSmem);
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(latent_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_set = true; }
    const int64_t n_tiles = (n + MLP_T - 1) / MLP_T;
    const int grid = (int)(n_tiles < DIF_NUM_SMS * 2 ? n_tiles : DIF_NUM_SMS * 2);
    latent_grad_kernel<<<grid, MLP_THREADS, smem, (cudaStream_t)stream>>>((const float*)decoder_prepared, latent_u, inv, rel_xyz, gt_sdf, (int)n,
                                                                          1.0f / (float)n_div, grad_u, loss_out);
    DIF_COUNT_LAUNCH(1);
    return check_launch("latent_grad_kernel");
}
