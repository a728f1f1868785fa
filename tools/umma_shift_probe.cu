// Probe: can a K-major SWIZZLE_128B UMMA operand start at an arbitrary 128-byte row of a larger swizzled buffer?
// A buffer: 256 rows x 64 bf16 (128 B per row), written with the address-based 128B swizzle (chunk ^= (row & 7)).
// For shift s in 0..9 the MMA reads rows [s, s+128) as the A tile; B = [64 x 64] chosen so that D[m][n] = A[m][n].
// Variants: base_offset field = 0, or = (start_addr >> 7) & 7.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../rspnet_b200/csrc/common.cuh"
using namespace rsp;

__global__ void probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int shift, int variant) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;               // 256 rows * 128 B = 32 KB
  uint8_t* sb = smem + 32768;       // 64 rows * 128 B = 8 KB
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
    uint64_t adesc = make_smem_desc_sw128(astart, 16, 1024);
    if (variant == 1) adesc |= static_cast<uint64_t>((astart >> 7) & 7) << 49;
    uint64_t bdesc = make_smem_desc_sw128(smem_u32(sb), 16, 1024);
    uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    for (int k = 0; k < 4; ++k) umma_bf16(tm, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
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
namespace rsp { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }

int main() {
  std::vector<__nv_bfloat16> hA(256 * 64), hB(64 * 64);
  for (int r = 0; r < 256; ++r) for (int c = 0; c < 64; ++c) hA[r * 64 + c] = __float2bfloat16(float((r * 7 + c * 3) % 97) - 48.f);
  for (int n = 0; n < 64; ++n) for (int k = 0; k < 64; ++k) hB[n * 64 + k] = __float2bfloat16(n == k ? 1.f : 0.f);  // D = A
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  std::vector<float> hD(128 * 64);
  for (int variant = 0; variant < 2; ++variant)
    for (int shift = 0; shift < 10; ++shift) {
      probe<<<1, 128, 44 * 1024>>>(dA, dB, dD, shift, variant);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("variant %d shift %d: CUDA error %s\n", variant, shift, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n)
        if (hD[m * 64 + n] != __bfloat162float(hA[(m + shift) * 64 + n])) ++bad;
      printf("variant %d (base_offset=%s) shift %d: %s (%d mismatches)\n", variant, variant ? "(addr>>7)&7" : "0", shift, bad ? "WRONG" : "ok", bad);
    }
  return 0;
}
