// Probe for tcgen05.mma with the A operand in tensor memory (TS form):
//  1. layout: A written with tcgen05.st.32x32b (lane = row, column j = bf16 pair (2j, 2j+1)); B = identity in smem;
//     D must reproduce A.
//  2. throughput: clocks per MMA for the SS form (A and B from shared memory) and the TS form, N = 64 / 128 / 256,
//     2048 back-to-back MMAs per CTA, one CTA per SM.  Tells how much of the tensor pipe the shared-memory operand fetch
//     can feed (SS) and what the registers -> TMEM route for A would buy.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../rspnet_b200/csrc/common.cuh"
using namespace rsp;

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void layout_probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sb = smem;  // 64 rows * 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8192);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  int t = threadIdx.x;
  for (int i = t; i < 64 * 8; i += blockDim.x) {
    int row = i >> 3, ch = i & 7;
    *reinterpret_cast<uint4*>(sb + row * 128 + ((ch ^ (row & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + row * 64 + ch * 8);
  }
  if (t == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (t < 32) tmem_alloc(slot, 128);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tm = *slot;
  {  // thread t = row t: 16 bf16 = 8 packed registers -> TMEM columns 64..71 of lane t
    uint32_t r[8];
    const uint32_t* src = reinterpret_cast<const uint32_t*>(A + t * 16);
    for (int j = 0; j < 8; ++j) r[j] = src[j];
    tmem_st_32x32b_x8(tm + (static_cast<uint32_t>((t >> 5) * 32) << 16) + 64, r);
    tmem_st_wait();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (t == 0) {
    uint64_t bdesc = make_smem_desc_sw128(smem_u32(sb), 16, 1024);
    uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    umma_bf16_ts(tm, tm + 64, bdesc, idesc, 0);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after_sync();
  {
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
  if (t < 32) tmem_dealloc(tm, 128);
}

template <int N, int TS>
__global__ void rate_probe(long long* out, int iters, int shift_rows) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;            // (128 + 16) rows * 128 B
  uint8_t* sb = smem + 20480;    // N rows * 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 20480 + 256 * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  int t = threadIdx.x;
  for (int i = t; i < (20480 + 256 * 128) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (t == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (t < 32) tmem_alloc(slot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tm = *slot;
  long long t0 = 0, t1 = 0;
  if (t < 32) {
    if (elect_one()) {
      const uint64_t adesc = make_smem_desc_sw128(smem_u32(sa) + shift_rows * 128, 16, 1024);
      const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sb), 16, 1024);
      constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (TS) umma_bf16_ts(tm, tm + 256 + 8 * k, bdesc + 2 * k, idesc, 1);
          else umma_bf16(tm, adesc + 2 * k, bdesc + 2 * k, idesc, 1);
        }
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
  if (t < 32) tmem_dealloc(tm, 512);
}
namespace rsp { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; }
int make_tmap_bf16(CUtensorMap*, const void*, int, const unsigned long long*, const unsigned long long*, const unsigned*) { return 0; } }

template <int N, int TS>
void run_rate(long long* d, int grid, int shift_rows = 0) {
  const int iters = 512;
  cudaFuncSetAttribute(rate_probe<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  rate_probe<N, TS><<<grid, 128, 56 * 1024>>>(d, iters, shift_rows);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("rate N=%d TS=%d: CUDA error %s\n", N, TS, cudaGetErrorString(e)); exit(1); }
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), d, grid * 8, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (auto v : h) mx = v > mx ? v : mx;
  double per = double(mx) / (iters * 4);
  printf("%s  M=128 N=%3d K=16 grid %3d A start row %d: %.1f clk per MMA (floor %d) -> %.0f%% of the tensor pipe\n",
         TS ? "TS (A in TMEM)" : "SS (A in smem)", N, grid, shift_rows, per, N / 2, 100.0 * (N / 2) / per);
}

int main() {
  std::vector<__nv_bfloat16> hA(128 * 16), hB(64 * 64);
  for (int r = 0; r < 128; ++r) for (int c = 0; c < 16; ++c) hA[r * 16 + c] = __float2bfloat16(float((r * 7 + c * 3) % 97) - 48.f);
  for (int n = 0; n < 64; ++n) for (int k = 0; k < 64; ++k) hB[n * 64 + k] = __float2bfloat16(n == k ? 1.f : 0.f);
  __nv_bfloat16 *dA, *dB; float* dD; long long* dT;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * 64 * 4); cudaMalloc(&dT, 1024 * 8);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  layout_probe<<<1, 128, 16 * 1024>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("layout probe: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> hD(128 * 64);
  cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) if (hD[m * 64 + n] != __bfloat162float(hA[m * 16 + n])) ++bad;
  printf("TS layout (lane = row, column j = elements 2j, 2j+1): %s (%d mismatches)\n", bad ? "WRONG" : "ok", bad);
  if (bad) for (int n = 0; n < 16; ++n) printf("  D[1][%d] = %g, A[1][%d] = %g\n", n, hD[64 + n], n, __bfloat162float(hA[16 + n]));
  for (int grid : {1, 148}) {
    run_rate<64, 0>(dT, grid); run_rate<64, 1>(dT, grid);
    run_rate<128, 0>(dT, grid); run_rate<128, 1>(dT, grid);
    run_rate<256, 0>(dT, grid); run_rate<256, 1>(dT, grid);
  }
  // does an A tile that starts in the middle of a 1024-byte swizzle atom (direct conv: arbitrary pixel-row offsets) cost more?
  for (int shift : {1, 3, 4, 8, 9}) {
    run_rate<64, 0>(dT, 148, shift);
    run_rate<128, 0>(dT, 148, shift);
  }
  return 0;
}
