// tcgen05.mma (SS form, M=128 K=16) issue rate by operand major-ness, with the 128-byte-swizzled layouts the conv
// kernels use: K-major (fprop / dgrad: rows = M or N, 128 B = 64 K elements) against MN-major (wgrad: rows = K pixels,
// 128 B = 64 M or N elements, 64-wide panels LBO apart).  One CTA per SM, 2048 back-to-back MMAs on zeroed operands.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_mn_rate_probe tools/umma_mn_rate_probe.cu rspnet_b200/csrc/common.cu -lcudart
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../rspnet_b200/csrc/common.cuh"
using namespace rsp;

template <int N, int AMN, int BMN>
__global__ void rate_probe(long long* out, int iters) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  constexpr int PANEL = 8192;                    // 64 rows x 128 B
  constexpr int A_BYTES = 2 * PANEL;             // K-major: 128 rows x 64 K;  MN-major: 2 panels of 64 pixels x 64 M
  constexpr int B_BYTES = (N / 64) * PANEL;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + A_BYTES + B_BYTES);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int t = threadIdx.x;
  for (int i = t; i < (A_BYTES + B_BYTES) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (t == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (t < 32) tmem_alloc(slot, N < 32 ? 32 : N);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = *slot;
  if (t < 32) {
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      const uint64_t adesc = AMN ? make_smem_desc_sw128(smem_u32(smem), PANEL, 1024) : make_smem_desc_sw128(smem_u32(smem), 16, 1024);
      const uint64_t bdesc = BMN ? make_smem_desc_sw128(smem_u32(smem) + A_BYTES, PANEL, 1024)
                                 : make_smem_desc_sw128(smem_u32(smem) + A_BYTES, 16, 1024);
      constexpr uint32_t idesc = make_idesc_bf16(128, N, AMN, BMN);
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k)   // K-major: +32 B inside the swizzle row; MN-major: +2 groups of 8 pixels (2048 B)
          umma_bf16(tm, adesc + (AMN ? 128 * k : 2 * k), bdesc + (BMN ? 128 * k : 2 * k), idesc, 1);
      }
      umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    t1 = clock64();
    if (t == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (t < 32) tmem_dealloc(tm, N < 32 ? 32 : N);
}

template <int N, int AMN, int BMN>
void run(long long* d) {
  const int iters = 512, grid = 148;
  const int smem = 16384 + (N / 64) * 8192 + 1024 + 64;
  cudaFuncSetAttribute(rate_probe<N, AMN, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  rate_probe<N, AMN, BMN><<<grid, 128, smem>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d A %s B %s: CUDA error %s\n", N, AMN ? "MN" : "K", BMN ? "MN" : "K", cudaGetErrorString(e)); exit(1); }
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), d, grid * 8, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (auto v : h) mx = v > mx ? v : mx;
  const double per = double(mx) / (iters * 4);
  printf("SS M=128 N=%3d K=16  A %2s-major  B %2s-major: %6.1f clk per MMA (floor %3d) -> %3.0f%% of the tensor pipe\n", N,
         AMN ? "MN" : "K", BMN ? "MN" : "K", per, N / 2, 100.0 * (N / 2) / per);
}

int main() {
  long long* d;
  cudaMalloc(&d, 1024 * 8);
  run<64, 0, 0>(d);  run<64, 1, 0>(d);  run<64, 0, 1>(d);  run<64, 1, 1>(d);
  run<128, 0, 0>(d); run<128, 1, 0>(d); run<128, 0, 1>(d); run<128, 1, 1>(d);
  run<256, 0, 0>(d); run<256, 1, 0>(d); run<256, 0, 1>(d); run<256, 1, 1>(d);
  return 0;
}
