// Hash-sharded map: pack / unpack of the per-frame latent exchange, so that the exchange is ONE fixed-size NCCL
// all-gather per frame with no host synchronisation (sizes live in a header row, not on the host).
//   The reference is single-GPU; there is nothing to cite.  Protocol: each rank sends [1 + cap_rows][32] floats:
//     row 0      header: word 0 = number of rows this rank wants to publish (may exceed cap_rows -> overflow, see unpack)
//     row 1 + i  word 0 = PLIVox slot (int32 bits), words 1..29 = latent row, words 30..31 = 0        (128-byte rows)
//   Slots are global because the integer map state is replicated (dif_map_view.shard_*).  HBM-bound copies of 128-byte rows.
#include "common.cuh"

namespace dif {

constexpr int XROW = 32;

__global__ void shard_pack_kernel(const float* __restrict__ latent, int lat_stride, const int32_t* __restrict__ xchg_slots, const int32_t* __restrict__ n_xchg,
                                  int64_t cap_rows, float* __restrict__ send) {
    const int count = *n_xchg;
    const int64_t rows = count < cap_rows ? count : cap_rows;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;          // one thread per word
    if (e < XROW) send[e] = e == 0 ? __int_as_float(count) : 0.f;
    const int64_t r = e / XROW; const int w = (int)(e % XROW);
    if (r >= rows) return;
    const int slot = xchg_slots[r];
    float v = 0.f;
    if (w == 0) v = __int_as_float(slot);
    else if (w <= DIF_L) v = latent[(int64_t)slot * lat_stride + (w - 1)];
    send[(r + 1) * XROW + w] = v;
}

__global__ void shard_unpack_kernel(float* __restrict__ latent, int lat_stride, int64_t capacity, const float* __restrict__ gathered, int world, int my_rank,
                                    int64_t cap_rows, int32_t* __restrict__ overflow) {
    const int src = blockIdx.y;
    const float* buf = gathered + (int64_t)src * (cap_rows + 1) * XROW;
    const int count = __float_as_int(buf[0]);
    if (count > cap_rows && blockIdx.x == 0 && threadIdx.x == 0) atomicMax(overflow, count); // every rank sees every header: same value everywhere
    if (src == my_rank) return;                                                 // own rows are already in place
    const int64_t rows = count < cap_rows ? count : cap_rows;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = e / XROW; const int w = (int)(e % XROW);
    if (r >= rows || w == 0 || w > DIF_L) return;
    const int slot = __float_as_int(buf[(r + 1) * XROW]);
    if (slot >= 0 && slot < capacity) latent[(int64_t)slot * lat_stride + (w - 1)] = buf[(r + 1) * XROW + w];
}

}  // namespace dif

using namespace dif;

extern "C" {

size_t dif_shard_xchg_bytes(int64_t cap_rows) { return (size_t)(cap_rows + 1) * XROW * sizeof(float); }

int dif_shard_pack(const dif_map_view* map, const int32_t* n_xchg_dev, int64_t cap_rows, float* send_buf, void* stream) {
    if (!map || !map->xchg_slots || !n_xchg_dev || cap_rows <= 0 || !send_buf) return DIF_E_INVALID;
    const int64_t words = (cap_rows + 1) * XROW;
    shard_pack_kernel<<<(unsigned)((words + 255) / 256), 256, 0, (cudaStream_t)stream>>>(map->latent_vecs, map->latent_stride > 0 ? map->latent_stride : DIF_L, map->xchg_slots, n_xchg_dev, cap_rows, send_buf);
    DIF_COUNT_LAUNCH(1);
    return check_launch("shard_pack_kernel");
}

int dif_shard_unpack(const dif_map_view* map, const float* gathered, int world, int64_t cap_rows, int32_t* overflow_dev, void* stream) {
    if (!map || !gathered || world < 1 || cap_rows <= 0 || !overflow_dev) return DIF_E_INVALID;
    const int64_t words = cap_rows * XROW;
    const dim3 grid((unsigned)((words + 255) / 256), (unsigned)world);
    shard_unpack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(map->latent_vecs, map->latent_stride > 0 ? map->latent_stride : DIF_L, map->capacity, gathered, world, map->shard_rank, cap_rows, overflow_dev);
    DIF_COUNT_LAUNCH(1);
    return check_launch("shard_unpack_kernel");
}

}  // extern "C"
