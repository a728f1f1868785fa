// Probe: MN-major SWIZZLE_128B UMMA A operand whose two 64-wide M panels OVERLAP (LBO = a few 128-byte pixel rows).
// Needed by a direct (im2col-free) wgrad: rows of the smem buffer are pixels (K), the 64 channels of a pixel are one 128 B
// swizzled row (M); the second M panel is the same pixel run shifted by `lbo_rows` pixels (the next filter tap).
//   D[m][n] = sum_k A[k][m] * B[k][n],  B[k][n] = (n == k)  ->  D[m][n] = A[n][m]   (n < 16)
//   expected: m < 64 : buf[shift + n][m];  m >= 64 : buf[shift + lbo_rows + n][m - 64]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../rspnet_b200/csrc/common.cuh"
using namespace rsp;

__global__ void probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int shift, int lbo_rows) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;               // 256 pixel rows * 128 B = 32 KB
  uint8_t* sb = smem + 32768;       // 64 pixel rows * 128 B = 8 KB (only 16 used)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768 + 8192);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  int t = threadIdx.x;
  for (int i = t; i < 256 * 8; i += blockDim.x) {
    int row = i >> 3, ch = i & 7;
    *reinterpret_cast<uint4*>(sa + row * 128 + ((ch ^ (row & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + row * 64 + ch * 8);
  }
  for (int i = t; i < 64 * 8; i += blockDim.x) {
    int row = i >> 3, ch = i & 7;
    *reinterpret_cast<uint4*>(sb + row * 128 + ((ch ^ (row & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + row * 64 + ch * 8);
  }
  if (t == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (t < 32) tmem_alloc(slot, 64);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tm = *slot;
  if (t == 0) {
    uint32_t astart = smem_u32(sa) + shift * 128;
    uint64_t adesc = make_smem_desc_sw128(astart, lbo_rows * 128, 1024);   // LBO: next 64-wide M panel; SBO: next 8 pixels
    uint64_t bdesc = make_smem_desc_sw128(smem_u32(sb), 8192, 1024);
    uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
    umma_bf16(tm, adesc, bdesc, idesc, 0);    // K = 16 pixels
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after_sync();
  if (t < 128) {
    int warp = t >> 5;
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tm + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) D[t * 64 + c0 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (t < 32) tmem_dealloc(tm, 64);
}
namespace rsp { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; }
int make_tmap_bf16(CUtensorMap*, const void*, int, const unsigned long long*, const unsigned long long*, const unsigned*) { return 0; } }

int main() {
  std::vector<__nv_bfloat16> hA(256 * 64), hB(64 * 64);
  for (int r = 0; r < 256; ++r) for (int c = 0; c < 64; ++c) hA[r * 64 + c] = __float2bfloat16(float((r * 7 + c * 3) % 97) - 48.f);
  for (int k = 0; k < 64; ++k) for (int n = 0; n < 64; ++n) hB[k * 64 + n] = __float2bfloat16(n == k ? 1.f : 0.f);
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  std::vector<float> hD(128 * 64);
  const int lbos[] = {64, 1, 2, 3, 8, 28, 56};
  for (int lbo : lbos)
    for (int shift = 0; shift < 10; ++shift) {
      probe<<<1, 128, 44 * 1024>>>(dA, dB, dD, shift, lbo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("lbo %d shift %d: CUDA error %s\n", lbo, shift, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) {
        int row = shift + n + (m >= 64 ? lbo : 0);
        if (hD[m * 64 + n] != __bfloat162float(hA[row * 64 + (m & 63)])) ++bad;
      }
      printf("MN-major A, LBO = %d pixel rows, start row %d: %s (%d mismatches)\n", lbo, shift, bad ? "WRONG" : "ok", bad);
    }
  return 0;
}
