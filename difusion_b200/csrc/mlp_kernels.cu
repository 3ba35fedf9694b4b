// Network entry points of the C ABI: weight preparation, dif_decode, dif_encode (fp32 SIMT tiles).
//   replaces reference network/utility.py:61-126 (forward_model) + di_decoder.py:55-86 / di_encoder.py:26-30.
#include "mlp_simt.cuh"
#include "decode_args.cuh"
#include <stdlib.h>

namespace dif {

thread_local char g_last_error[256] = "";
bool pdl_enabled() {
    const char* e = getenv("DIF_PDL");                                  // read per call so one process can time both settings
    return !(e && e[0] == '0');
}
thread_local ProfHook g_prof[DIF_PROF_COUNT] = {};
thread_local uint64_t g_launches = 0;

// ---------------------------------------------------------------------------------------------- prepare
__global__ void prepare_decoder_kernel(const float* __restrict__ blob, float* __restrict__ P) {
    // blob: W0[128][32] b0 W1[128][128] b1 W2[96][128] b2 W3[128][128] b3 w4[128] b4 wu[128] bu
    const int oW0 = 0, ob0 = oW0 + 128 * 32, oW1 = ob0 + 128, ob1 = oW1 + 128 * 128, oW2 = ob1 + 128, ob2 = oW2 + 96 * 128,
              oW3 = ob2 + 96, ob3 = oW3 + 128 * 128, ow4 = ob3 + 128, ob4 = ow4 + 128, owu = ob4 + 1, obu = owu + 128;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int i = tid; i < 128 * 32; i += nth) { int n = i / 32, k = i % 32; float w = blob[oW0 + i]; P[DecW::W0 + i] = w; P[DecW::W0t + k * 128 + n] = w; }
    for (int i = tid; i < 128 * 128; i += nth) { int n = i / 128, k = i % 128; float w = blob[oW1 + i]; P[DecW::W1 + i] = w; P[DecW::W1t + k * 128 + n] = w; }
    for (int i = tid; i < 96 * 128; i += nth) { int n = i / 128, k = i % 128; float w = blob[oW2 + i]; P[DecW::W2 + i] = w; P[DecW::W2t + k * 96 + n] = w; }
    for (int i = tid; i < 128 * 128; i += nth) { int n = i / 128, k = i % 128; float w = blob[oW3 + i]; P[DecW::W3 + i] = w; P[DecW::W3t + k * 128 + n] = w; }
    for (int i = tid; i < 128; i += nth) {
        P[DecW::b0 + i] = blob[ob0 + i]; P[DecW::b1 + i] = blob[ob1 + i]; P[DecW::b3 + i] = blob[ob3 + i];
        P[DecW::w4 + i] = blob[ow4 + i]; P[DecW::wu + i] = blob[owu + i];
        if (i < 96) P[DecW::b2 + i] = blob[ob2 + i];
    }
    if (tid == 0) { P[DecW::b4] = blob[ob4]; P[DecW::bu] = blob[obu]; }
}

__global__ void prepare_encoder_kernel(const float* __restrict__ blob, float* __restrict__ P) {
    const int oW0 = 0, ob0 = oW0 + 32 * 6, oW1 = ob0 + 32, ob1 = oW1 + 64 * 32, oW2 = ob1 + 64, ob2 = oW2 + 256 * 64,
              oW3 = ob2 + 256, ob3 = oW3 + 29 * 256;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int i = tid; i < 32 * 6; i += nth) { int n = i / 6, k = i % 6; P[EncW::W0t + k * 32 + n] = blob[oW0 + i]; }
    for (int i = tid; i < 64 * 32; i += nth) { int n = i / 32, k = i % 32; P[EncW::W1t + k * 64 + n] = blob[oW1 + i]; }
    for (int i = tid; i < 256 * 64; i += nth) { int n = i / 64, k = i % 64; P[EncW::W2t + k * 256 + n] = blob[oW2 + i]; }
    for (int i = tid; i < 256 * 32; i += nth) { int k = i / 32, n = i % 32; P[EncW::W3t + i] = n < 29 ? blob[oW3 + n * 256 + k] : 0.f; }
    for (int i = tid; i < 256; i += nth) {
        P[EncW::b2 + i] = blob[ob2 + i];
        if (i < 32) { P[EncW::b0 + i] = blob[ob0 + i]; P[EncW::b3 + i] = i < 29 ? blob[ob3 + i] : 0.f; }
        if (i < 64) P[EncW::b1 + i] = blob[ob1 + i];
    }
}

// ---------------------------------------------------------------------------------------------- decode
template <bool GRAD>
__global__ void __launch_bounds__(MLP_THREADS) decode_simt_kernel(DecodeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DecoderSmem& s = *reinterpret_cast<DecoderSmem*>(smem_raw);
    __shared__ int64_t row_s[MLP_T], out_s[MLP_T];
    __shared__ int lat_i[MLP_T];
    const int64_t n_total = a.n_dev ? (int64_t)*a.n_dev : a.n;
    const int64_t n_tiles = (n_total + MLP_T - 1) / MLP_T;
    const int n3 = a.lat_n * a.lat_n * a.lat_n;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * MLP_T;
        if (threadIdx.x < MLP_T) {
            const int64_t sidx = base + threadIdx.x;
            int64_t row = -1, o = sidx; int li = 0;
            if (sidx < n_total) {
                if (a.mode == 0) {
                    row = a.rows ? (int64_t)a.rows[sidx] : sidx;
                    o = a.out_index ? (int64_t)a.out_index[sidx] : sidx;
                } else {
                    const int64_t g = a.mode == 1 ? sidx : (int64_t)a.list[sidx];
                    row = a.rows[g / n3]; li = (int)(g % n3); o = g;
                    if (a.row_map && row >= 0) row = a.row_map[row];
                }
            }
            row_s[threadIdx.x] = row; out_s[threadIdx.x] = o; lat_i[threadIdx.x] = li;
        }
        __syncthreads();
        // stage inputs: thread (t, j) loads input column j of sample t  (a latent row is 116 contiguous bytes)
        for (int i = threadIdx.x; i < MLP_T * 32; i += MLP_THREADS) {
            const int t = i / 32, j = i % 32;
            const int64_t row = row_s[t];
            float v = 0.f;
            if (row >= 0) {
                if (j < DIF_L) v = __ldg(a.latent + row * a.lat_stride + j);
                else if (a.mode == 0) v = __ldg(a.xyz + (base + t) * 3 + (j - DIF_L));
                else {
                    const int li = lat_i[t], c = j - DIF_L, nn = a.lat_n;
                    v = lattice_coord(a, c == 0 ? li / (nn * nn) : (c == 1 ? (li / nn) % nn : li % nn));
                }
            }
            s.cat[(96 + j) * MLP_TP + t] = v;
        }
        __syncthreads();
        decoder_forward_tile(a.P, s);
        if (threadIdx.x < MLP_T) {
            const int t = threadIdx.x;
            const int64_t sidx = base + t;
            const float sdf = tanhf(s.pre[t]);
            const float u = s.pre[MLP_T + t];
            const float sd = 0.05f + 0.5f * softplus_ref(u);
            const bool live = row_s[t] >= 0;
            // padding samples (row < 0) write zeros in place, but never scatter
            if (sidx < n_total && (live || (a.mode == 0 && !a.out_index))) {
                const int64_t o = out_s[t];
                a.sdf[o] = live ? a.sdf_sign * sdf : 0.f; a.std[o] = live ? sd : 0.f;
            }
            if (GRAD) s.seed[t] = a.grad_head == 0 ? a.sdf_sign * (1.f - sdf * sdf) : 0.5f / (1.f + expf(-u));
        }
        if (GRAD) {
            __syncthreads();
            decoder_backward_tile(a.P, s, a.grad_head);
            if (threadIdx.x < 3 * MLP_T) {
                const int t = threadIdx.x / 3, c = threadIdx.x % 3;
                const int64_t sidx = base + t;
                const bool live = row_s[t] >= 0;
                if (sidx < n_total && (live || (a.mode == 0 && !a.out_index))) a.grad[out_s[t] * 3 + c] = live ? s.gx[c * MLP_T + t] : 0.f;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------- encode
__global__ void __launch_bounds__(MLP_THREADS) encode_simt_kernel(const float* __restrict__ P, const float* __restrict__ xyzn,
                                                                  int64_t n, float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EncoderSmem& s = *reinterpret_cast<EncoderSmem*>(smem_raw);
    const int64_t n_tiles = (n + MLP_T - 1) / MLP_T;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * MLP_T;
        if (threadIdx.x < MLP_T * 6) {
            const int t = threadIdx.x / 6, j = threadIdx.x % 6;
            s.in[j * MLP_TP + t] = (base + t < n) ? __ldg(xyzn + (base + t) * 6 + j) : 0.f;
        }
        __syncthreads();
        encoder_forward_tile(P, s);
        for (int i = threadIdx.x; i < MLP_T * DIF_L; i += MLP_THREADS) {
            const int t = i / DIF_L, j = i % DIF_L;
            if (base + t < n) out[(base + t) * DIF_L + j] = s.out[j * MLP_TP + t];
        }
        __syncthreads();
    }
}

static int grid_for_tiles(int64_t n_tiles, int ctas_per_sm) {
    int64_t cap = (int64_t)DIF_NUM_SMS * ctas_per_sm;
    return (int)(n_tiles < cap ? (n_tiles > 0 ? n_tiles : 1) : cap);
}

size_t decoder_tc_image_bytes();
int prepare_decoder_tc(const float* P, unsigned char* image, cudaStream_t st);
size_t encoder_tc_image_bytes();
int prepare_encoder_tc(const float* P, unsigned char* image, cudaStream_t st);
int launch_encode_tc(const void* encoder_prepared, const float* xyzn, int64_t n, float* out, cudaStream_t st);

// 0 = auto (tensor cores for forward-only launches of >= 1024 samples), 1 = force fp32 SIMT, 2 = force tensor cores
static int decode_path_override() {      // read per call so tests can compare both paths
    const char* e = getenv("DIF_DECODE_PATH");
    return !e ? 0 : (e[0] == 's' ? 1 : (e[0] == 't' ? 2 : 0));
}

int launch_decode(DecodeArgs a, int64_t n_max, cudaStream_t st) {
    if (n_max <= 0) return DIF_OK;
    const int ov = decode_path_override();
    if (!a.grad && ov != 1 && (ov == 2 || n_max >= 1024)) return launch_decode_tc(a.P, a, n_max, st);
    const int64_t n_tiles = (n_max + MLP_T - 1) / MLP_T;
    const size_t smem = sizeof(DecoderSmem);
    prof_begin(DIF_PROF_DECODE, st);
    if (a.grad) {
        cudaFuncSetAttribute(decode_simt_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        decode_simt_kernel<true><<<grid_for_tiles(n_tiles, 3), MLP_THREADS, smem, st>>>(a);
    } else {
        cudaFuncSetAttribute(decode_simt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        decode_simt_kernel<false><<<grid_for_tiles(n_tiles, 3), MLP_THREADS, smem, st>>>(a);
    }
    prof_end(DIF_PROF_DECODE, st);
    DIF_COUNT_LAUNCH(1);
    return check_launch("decode_simt_kernel");
}

int launch_decode_explicit(const float* P, const float* latent, int lat_stride, const int32_t* rows, const float* xyz, int64_t n,
                           const int32_t* out_index, float sdf_sign, float* sdf, float* std, float* grad, int grad_head, cudaStream_t st) {
    DecodeArgs a{P, latent, lat_stride, rows, xyz, n, out_index, sdf_sign, sdf, std, grad, grad_head, 0, 1, 0.f, 0.f, nullptr, nullptr, nullptr};
    return launch_decode(a, n, st);
}

// lattice decode for mesh extraction; list == nullptr: all n_blocks * lat_n^3 points, else the first *n_dev entries of list
int launch_decode_lattice(const float* P, const float* latent, int lat_stride, const int32_t* row_map, const int32_t* block_slots, int64_t n_blocks,
                          int lat_n, float lat_step, float lat_a, const uint32_t* list, const int32_t* n_dev, int64_t n_max, float sdf_sign,
                          float* sdf, float* std, cudaStream_t st) {
    DecodeArgs a{P, latent, lat_stride, block_slots, nullptr, n_blocks * lat_n * lat_n * lat_n, nullptr, sdf_sign, sdf, std, nullptr, 0,
                 list ? 2 : 1, lat_n, lat_step, lat_a, list, n_dev, row_map};
    return launch_decode(a, n_max, st);
}

}  // namespace dif

using namespace dif;

extern "C" {

int dif_debug_tc_timing(void* dev_buf) { return dif::set_tc_timing_buffer((unsigned long long*)dev_buf); }
int dif_abi_version(void) { return DIF_ABI_VERSION; }
int dif_profile_hook(int which, void* start_event, void* stop_event) {
    if (which < 0 || which >= DIF_PROF_COUNT) return DIF_E_INVALID;
    dif::g_prof[which].start = (cudaEvent_t)start_event; dif::g_prof[which].stop = (cudaEvent_t)stop_event;
    return DIF_OK;
}
uint64_t dif_launch_count(int reset) { const uint64_t v = dif::g_launches; if (reset) dif::g_launches = 0; return v; }
const char* dif_last_error(void) { return dif::g_last_error; }

size_t dif_decoder_prepared_bytes(void) { return (size_t)DecW::FP32_END * sizeof(float) + decoder_tc_image_bytes(); }
size_t dif_encoder_prepared_bytes(void) { return (size_t)EncW::FP32_END * sizeof(float) + encoder_tc_image_bytes(); }

int dif_prepare_decoder(const float* blob_dev, void* prepared_dev, void* stream) {
    if (!blob_dev || !prepared_dev) return DIF_E_INVALID;
    prepare_decoder_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(blob_dev, (float*)prepared_dev);
    DIF_COUNT_LAUNCH(1);
    int rc = check_launch("prepare_decoder_kernel");
    if (rc) return rc;
    return prepare_decoder_tc((const float*)prepared_dev, (unsigned char*)prepared_dev + (size_t)DecW::FP32_END * sizeof(float), (cudaStream_t)stream);
}

int dif_prepare_encoder(const float* blob_dev, void* prepared_dev, void* stream) {
    if (!blob_dev || !prepared_dev) return DIF_E_INVALID;
    prepare_encoder_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(blob_dev, (float*)prepared_dev);
    DIF_COUNT_LAUNCH(1);
    int rc = check_launch("prepare_encoder_kernel");
    if (rc) return rc;
    return prepare_encoder_tc((const float*)prepared_dev, (unsigned char*)prepared_dev + (size_t)EncW::FP32_END * sizeof(float), (cudaStream_t)stream);
}

int dif_decode(const void* decoder_prepared, const float* latent, int latent_stride, const int32_t* rows, const float* xyz, int64_t n,
               const int32_t* out_index, float sdf_sign, float* sdf, float* std, float* dsdf_dxyz, float* dstd_dxyz, void* stream) {
    if (n < 0 || !decoder_prepared || latent_stride < DIF_L || (n > 0 && (!latent || !xyz || !sdf || !std))) return DIF_E_INVALID;
    if (latent_stride == 32 && ((uintptr_t)latent & 15)) return DIF_E_INVALID;          // 32-float rows are read with 16-byte loads
    const int lat_stride = latent_stride;
    const float* P = (const float*)decoder_prepared;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = launch_decode_explicit(P, latent, lat_stride, rows, xyz, n, out_index, sdf_sign, sdf, std, dsdf_dxyz, 0, st);
    if (rc == DIF_OK && dstd_dxyz) rc = launch_decode_explicit(P, latent, lat_stride, rows, xyz, n, out_index, sdf_sign, sdf, std, dstd_dxyz, 1, st);
    return rc;
}

int dif_encode(const void* encoder_prepared, const float* xyzn, int64_t n, float* latent_out, void* stream) {
    if (n < 0 || !encoder_prepared || (n > 0 && (!xyzn || !latent_out))) return DIF_E_INVALID;
    if (n == 0) return DIF_OK;
    {   // tensor cores for anything beyond a few tiles; DIF_ENCODE_PATH=simt forces the exact-fp32 kernel
        const char* e = getenv("DIF_ENCODE_PATH");
        if (!(e && e[0] == 's') && n >= 1024) return launch_encode_tc(encoder_prepared, xyzn, n, latent_out, (cudaStream_t)stream);
    }
    const size_t smem = sizeof(EncoderSmem);
    cudaFuncSetAttribute(encode_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t n_tiles = (n + MLP_T - 1) / MLP_T;
    encode_simt_kernel<<<grid_for_tiles(n_tiles, 4), MLP_THREADS, smem, (cudaStream_t)stream>>>((const float*)encoder_prepared, xyzn, n, latent_out);
    DIF_COUNT_LAUNCH(1);
    return check_launch("encode_simt_kernel");
}

}  // extern "C"
