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

// TILES distinct A tiles (128 rows x 128 B: four K=16 steps each) and B tiles (N rows x 128 B) are cycled so that no operand
// is fetched twice in a row; TILES = 1 repeats the same operands (what an operand cache would hide).
template <int N, int TS, int TILES>
__global__ void rate_probe(long long* out, int iters, int shift_rows) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  constexpr int A_TILE = 16384, B_TILE = N * 128;
  uint8_t* sa = smem;                              // TILES A tiles (+ 2 KB slack for the row shift)
  uint8_t* sb = smem + TILES * A_TILE + 2048;      // TILES B tiles
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + TILES * B_TILE);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  int t = threadIdx.x;
  for (int i = t; i < (TILES * (A_TILE + B_TILE) + 2048) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
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
      const uint64_t adesc0 = make_smem_desc_sw128(smem_u32(sa) + shift_rows * 128, 16, 1024);
      const uint64_t bdesc0 = make_smem_desc_sw128(smem_u32(sb), 16, 1024);
      constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const int tile = i % TILES;
        const uint64_t adesc = adesc0 + static_cast<uint64_t>(tile * (A_TILE >> 4));
        const uint64_t bdesc = bdesc0 + static_cast<uint64_t>(tile * (B_TILE >> 4));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (TS) umma_bf16_ts(tm, tm + 256 + 32 * (tile & 3) + 8 * k, bdesc + 2 * k, idesc, 1);
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
// Interference probe: SS-form MMAs (N = 64, 4 distinct tiles) with parts of a real mainloop switched on one at a time:
//  flags & 1 : tcgen05.commit to an mbarrier after every 16 MMAs            flags & 2 : rotate over 4 accumulators
//  flags & 4 : warps 1-3 keep reading TMEM (tcgen05.ld, what an epilogue does)
//  flags & 8 : warps 1-3 keep writing shared memory (st.shared.v4 into a scratch ring, what the producers' traffic does)
//  flags & 16: mbarrier try_wait + tcgen05.fence::after_thread_sync before every 16 MMAs
//  flags & 32: warps 1-3 keep issuing warp shuffles (epilogue statistics)
template <int N>
__global__ void interference_probe(long long* out, int iters, int flags, float* sink) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  constexpr int TILES = 4, A_TILE = 16384, B_TILE = N * 128;
  uint8_t* sa = smem;
  uint8_t* sb = smem + TILES * A_TILE;
  uint8_t* scratch = sb + TILES * B_TILE;            // 32 KB written by the interfering warps
  uint64_t* bar = reinterpret_cast<uint64_t*>(scratch + 32768);
  uint64_t* dummy = bar + 1;
  uint64_t* ready = bar + 2;
  volatile int* stop = reinterpret_cast<volatile int*>(bar + 3);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
  int t = threadIdx.x;
  for (int i = t; i < (TILES * (A_TILE + B_TILE) + 32768) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (t == 0) { mbar_init(bar, 1); mbar_init(dummy, 1 << 20); mbar_init(ready, 1); mbar_arrive(ready); *stop = 0; fence_mbar_init(); }
  if (t < 32) tmem_alloc(slot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tm = *slot;
  if (t < 32) {
    long long t0 = 0, t1 = 0;
    const uint64_t adesc0 = make_smem_desc_sw128(smem_u32(sa), 16, 1024);
    const uint64_t bdesc0 = make_smem_desc_sw128(smem_u32(sb), 16, 1024);
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if ((flags & 16) && (i & 3) == 0) {
        mbar_wait(ready, 0);
        tc_fence_after_sync();
      }
      if (elect_one()) {
        const int tile = i % TILES;
        const uint64_t adesc = adesc0 + static_cast<uint64_t>(tile * (A_TILE >> 4));
        const uint64_t bdesc = bdesc0 + static_cast<uint64_t>(tile * (B_TILE >> 4));
        const uint32_t d = tm + ((flags & 2) ? (i & 3) * N : 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(d, adesc + 2 * k, bdesc + 2 * k, idesc, 1);
        if ((flags & 1) && (i & 3) == 3) umma_commit(dummy);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(bar);
    __syncwarp();
    mbar_wait(bar, 0);
    t1 = clock64();
    *stop = 1;
    if (t == 0) out[blockIdx.x] = t1 - t0;
  } else {
    float acc = 0.f;
    const int w = t >> 5;
    uint32_t it = 0;
    while (!*stop) {
      if (flags & 4) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tm + (static_cast<uint32_t>(w * 32) << 16) + 256 + (it & 3) * 32, v);
        tmem_ld_wait();
        acc += __uint_as_float(v[it & 31]);
      }
      if (flags & 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(scratch + ((it * 8 + j) & 63) * 512 + (t & 31) * 16) = make_uint4(it, j, t, 0);
      }
      if (flags & 32) {
#pragma unroll
        for (int j = 0; j < 16; ++j) acc += __shfl_xor_sync(0xffffffffu, acc + j, 1 + (j & 15));
      }
      ++it;
    }
    if (acc == 123.456f) sink[t] = acc;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (t < 32) tmem_dealloc(tm, 512);
}

template <int N>
void run_interference(long long* d, float* sink, int flags, const char* what) {
  const int iters = 512, grid = 148;
  const int smem = 4 * (16384 + N * 128) + 32768 + 1024 + 128;
  cudaFuncSetAttribute(interference_probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  interference_probe<N><<<grid, 128, smem>>>(d, iters, flags, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("interference flags %d: CUDA error %s\n", flags, cudaGetErrorString(e)); exit(1); }
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), d, grid * 8, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (auto v : h) mx = v > mx ? v : mx;
  printf("SS N=%3d + %-58s: %.1f clk per MMA\n", N, what, double(mx) / (iters * 4));
}

// The RGB-stem operand layout: A = raw input rows, K-major WITHOUT swizzle, rows 16 B apart (LBO 16, SBO 128: overlapping
// windows); B = filter slab, K-major without swizzle (LBO 1024, SBO 128).  Same MMA shape (128 x 64 x 16).
__global__ void stem_layout_rate_probe(long long* out, int iters, int swizzled_b) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;               // 14 rows * 1024 B
  uint8_t* sb = smem + 14336;       // 7 * 4096 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 14336 + 28672);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  int t = threadIdx.x;
  for (int i = t; i < (14336 + 28672) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (t == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (t < 32) tmem_alloc(slot, 256);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tm = *slot;
  if (t < 32) {
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      auto nosw = [](uint32_t saddr, uint32_t lbo, uint32_t sbo) {
        uint64_t d = 0;
        d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
        d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
        d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
        d |= 1ull << 46;
        return d;
      };
      const uint64_t abase = nosw(smem_u32(sa), 16, 128);
      const uint64_t bbase = swizzled_b ? make_smem_desc_sw128(smem_u32(sb), 16, 1024) : nosw(smem_u32(sb), 1024, 128);
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int b = 0; b < 7; ++b)
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            const int j = 4 * m + b;
            const int arow = ((j & 1) * 7 + (j >> 1)) * 1024;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              umma_bf16(tm + m * 64, abase + static_cast<uint64_t>((arow + ks * 32) >> 4),
                        bbase + static_cast<uint64_t>(swizzled_b ? (b * 512 + ks * 2) : ((b * 4096 + ks * 2048) >> 4)), idesc, 1);
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
  if (t < 32) tmem_dealloc(tm, 256);
}

namespace rsp { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; }
int make_tmap_bf16(CUtensorMap*, const void*, int, const unsigned long long*, const unsigned long long*, const unsigned*) { return 0; } }

template <int N, int TS, int TILES>
void run_rate(long long* d, int grid, int shift_rows = 0) {
  const int iters = 512;
  const int smem = TILES * (16384 + N * 128) + 2048 + 1024 + 64;
  cudaFuncSetAttribute(rate_probe<N, TS, TILES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  rate_probe<N, TS, TILES><<<grid, 128, smem>>>(d, iters, shift_rows);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("rate N=%d TS=%d: CUDA error %s\n", N, TS, cudaGetErrorString(e)); exit(1); }
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), d, grid * 8, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (auto v : h) mx = v > mx ? v : mx;
  double per = double(mx) / (iters * 4);
  printf("%s  M=128 N=%3d K=16 grid %3d, %d distinct operand tiles, A start row %d: %.1f clk per MMA (floor %d) -> %.0f%% of the tensor pipe\n",
         TS ? "TS (A in TMEM)" : "SS (A in smem)", N, grid, TILES, shift_rows, per, N / 2, 100.0 * (N / 2) / per);
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
    run_rate<64, 0, 1>(dT, grid); run_rate<64, 1, 1>(dT, grid);
    run_rate<128, 0, 1>(dT, grid); run_rate<128, 1, 1>(dT, grid);
    run_rate<256, 0, 1>(dT, grid); run_rate<256, 1, 1>(dT, grid);
  }
  run_rate<64, 0, 4>(dT, 148); run_rate<64, 1, 4>(dT, 148);
  run_rate<128, 0, 4>(dT, 148); run_rate<128, 1, 4>(dT, 148);
  run_rate<256, 0, 4>(dT, 148); run_rate<256, 1, 4>(dT, 148);
  // does an A tile that starts in the middle of a 1024-byte swizzle atom (direct conv: arbitrary pixel-row offsets) cost more?
  for (int shift : {1, 9}) {
    run_rate<64, 0, 4>(dT, 148, shift);
    run_rate<128, 0, 4>(dT, 148, shift);
  }
  for (int sw = 0; sw < 2; ++sw) {
    cudaFuncSetAttribute(stem_layout_rate_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    stem_layout_rate_probe<<<148, 128, 45 * 1024>>>(dT, 64, sw);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e2 != cudaSuccess) { printf("stem layout probe: CUDA error %s\n", cudaGetErrorString(e2)); return 1; }
    std::vector<long long> h(148);
    cudaMemcpy(h.data(), dT, 148 * 8, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (auto v : h) mx = v > mx ? v : mx;
    printf("stem operand layout (A: no swizzle, 16-byte row pitch; B: %s): %.1f clk per MMA (M=128 N=64 K=16)\n",
           sw ? "128B swizzle" : "no swizzle, LBO 1024", double(mx) / (64 * 28));
  }
  float* sink; cudaMalloc(&sink, 4096);
  run_interference<64>(dT, sink, 0, "nothing");
  run_interference<64>(dT, sink, 1, "commit every 16 MMAs");
  run_interference<64>(dT, sink, 2, "4 accumulators in rotation");
  run_interference<64>(dT, sink, 4, "3 warps reading TMEM");
  run_interference<64>(dT, sink, 8, "3 warps writing shared memory");
  run_interference<64>(dT, sink, 16, "mbarrier wait + tcgen05.fence every 16 MMAs");
  run_interference<64>(dT, sink, 32, "3 warps shuffling");
  run_interference<64>(dT, sink, 1 | 2 | 4 | 16, "commit + accumulators + TMEM reads + waits");
  run_interference<128>(dT, sink, 0, "nothing");
  run_interference<128>(dT, sink, 8, "3 warps writing shared memory");
  run_interference<128>(dT, sink, 1 | 2 | 4 | 16, "commit + accumulators + TMEM reads + waits");
  return 0;
}
