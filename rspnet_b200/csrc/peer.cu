// Shuffle-BN over NVLink peer memory (reference: _batch_shuffle_ddp, moco/builder_diffspeed_diffloss.py:361-387).
//
// The reference all_gathers every rank's key clips on every rank (W x the batch) and keeps B rows.  Here every rank
// keeps its key clips in a buffer that its peers have mapped (CUDA IPC over NVLink / NVSwitch), and ONE kernel per rank
// pulls exactly the B rows the permutation assigns to it straight out of the owners' memory: no pack pass, no
// all_to_all with host-side split sizes, no unpack pass, and the permutation never has to be known on the host.
#include <cstring>

#include "common.cuh"
#include "rspnet_b200.h"

namespace rsp {

// dst[i] = peer[index[i] / rows_per_peer] row (index[i] % rows_per_peer); rows are `vec_per_row` 16-byte vectors.
// Four independent 16-byte loads per thread keep enough requests in flight to cover the NVLink round trip; peer data
// is read with L1 no-allocate (it is written by another device between launches).
__global__ void __launch_bounds__(256) gather_rows_peer_kernel(const uint4* const* __restrict__ peers,
                                                               const int64_t* __restrict__ index,
                                                               uint4* __restrict__ dst, int rows_per_peer,
                                                               size_t vec_per_row) {
  const size_t row = blockIdx.y;
  const int64_t g = index[row];
  const int owner = static_cast<int>(g / rows_per_peer);
  const uint4* s = peers[owner] + static_cast<size_t>(g - static_cast<int64_t>(owner) * rows_per_peer) * vec_per_row;
  uint4* o = dst + row * vec_per_row;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < vec_per_row; i += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4* a = s + i + j * stride;
      asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(v[j].x), "=r"(v[j].y), "=r"(v[j].z), "=r"(v[j].w)
                   : "l"(a));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i + j * stride] = v[j];
  }
  for (; i < vec_per_row; i += stride) {
    uint4 v;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(s + i));
    o[i] = v;
  }
}

// inv[perm[i]] = i  (idx_unshuffle = argsort(idx_shuffle) for a permutation, builder:381)
__global__ void invert_permutation_kernel(const int64_t* __restrict__ perm, int64_t* __restrict__ inv, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) inv[perm[i]] = i;
}

}  // namespace rsp

using namespace rsp;

extern "C" {

int rsp_peer_alloc(int64_t bytes, void** dev_ptr) {
  RSP_REQUIRE(bytes > 0 && dev_ptr, "peer_alloc: bad arguments");
  // a dedicated cudaMalloc allocation (not the caller's caching allocator) so that the IPC handle covers exactly it
  cudaError_t e = cudaMalloc(dev_ptr, static_cast<size_t>(bytes));
  if (e != cudaSuccess) {
    set_error("peer_alloc: cudaMalloc(%lld) failed: %s", static_cast<long long>(bytes), cudaGetErrorString(e));
    return RSP_ERR_CUDA;
  }
  return RSP_OK;
}

int rsp_peer_free(void* dev_ptr) {
  cudaError_t e = cudaFree(dev_ptr);
  if (e != cudaSuccess) {
    set_error("peer_free: %s", cudaGetErrorString(e));
    return RSP_ERR_CUDA;
  }
  return RSP_OK;
}

int rsp_peer_export(void* dev_ptr, uint8_t* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, dev_ptr);
  if (e != cudaSuccess) {
    set_error("peer_export: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return RSP_ERR_CUDA;
  }
  memcpy(handle64, &h, 64);
  return RSP_OK;
}

int rsp_peer_open(const uint8_t* handle64, void** mapped) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(mapped, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    set_error("peer_open: cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
    return RSP_ERR_CUDA;
  }
  return RSP_OK;
}

int rsp_peer_close(void* mapped) {
  cudaError_t e = cudaIpcCloseMemHandle(mapped);
  if (e != cudaSuccess) {
    set_error("peer_close: %s", cudaGetErrorString(e));
    return RSP_ERR_CUDA;
  }
  return RSP_OK;
}

int rsp_gather_rows_peer(const void* const* peer_bases, const int64_t* index, void* dst, int64_t n_rows,
                         int32_t rows_per_peer, int64_t row_bytes, void* stream) {
  if (n_rows == 0) return RSP_OK;
  RSP_REQUIRE(row_bytes % 16 == 0 && n_rows <= 65535 && rows_per_peer > 0,
              "gather_rows_peer: row_bytes %% 16 != 0, too many rows or rows_per_peer <= 0");
  const size_t vec = static_cast<size_t>(row_bytes) / 16;
  // ~4 vectors per thread; at most 64 CTAs per row so that B rows give a few thousand CTAs in flight without
  // monopolising the SMs the concurrent encoder passes run on
  unsigned gx = static_cast<unsigned>((vec + 1023) / 1024);
  if (gx > 64) gx = 64;
  if (gx == 0) gx = 1;
  dim3 grid(gx, static_cast<unsigned>(n_rows));
  gather_rows_peer_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4* const*>(peer_bases), index, static_cast<uint4*>(dst), rows_per_peer, vec);
  return check_launch("gather_rows_peer");
}

int rsp_invert_permutation(const int64_t* perm, int64_t* inv, int32_t n, void* stream) {
  RSP_REQUIRE(n > 0, "invert_permutation: empty permutation");
  invert_permutation_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(perm, inv, n);
  return check_launch("invert_permutation");
}

}  // extern "C"
