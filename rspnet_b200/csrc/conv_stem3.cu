// Direct tcgen05 convolution for the 3x3x3, unit-stride RGB stem: 4 stored input channels (RGB + pad) -> 64 channels
// (reference: models/c3d.py:21-25 conv1, 3 -> 64, kernel 3, padding 1, on 16 x 112 x 112 clips).
//
// The layer is HBM-bound (2 x 64 bytes written per 8 bytes read); the generic gather kernel spends ~40 instructions per
// 8-byte pixel and runs 10x slower than the output write.  Here the raw-row trick of conv_stem.cu is applied to a
// w-stride of 1: in the K-major no-swizzle UMMA layout consecutive A rows are 16 bytes = TWO pixels apart, so A row j of
// a raw input row in shared memory is the 4-pixel window (2j .. 2j+3), K = 16 = 4 pixel slots x 4 channels.  That window
// serves TWO output pixels with two differently packed copies of the filter:
//   odd  pixel 2j+1: taps (w-1, w, w+1) = window slots 0, 1, 2   (filter set O, slot 3 zero)
//   even pixel 2j  : row j-1, i.e. the window (2j-2 .. 2j+1): slots 1, 2, 3 (filter set E, slot 0 zero)
// Rows sit 1 KB apart with the pixels at byte 16 (the 16 bytes in front and everything behind the row stay zero: the
// left / right padding), so M = 128 = 64 windows of output row h and 64 of row h+1 (56 of 64 used).  Each input row is one
// bulk copy (cp.async.bulk); rows outside the clip are copied from a zero row stored behind the packed filter.
// Both filter sets (2 x 9 x 2 KB) stay resident in shared memory.
//
// CTA (persistent, 1/SM): warps 0-3 epilogue, warp 4 producer, warp 5 MMA lane.  One iteration = 4 output rows of one
// (n, to) = 4 accumulators (row pair x parity) of 128 x 64, double buffered in TMEM (512 columns); stage = one frame tap.
#include "common.cuh"
#include "rspnet_b200.h"

namespace rsp {

int device_sm_count();

struct Stem3Params {
  const __nv_bfloat16* x;    // [N][Ti][Hi][Wi][4]
  const __nv_bfloat16* wst;  // [2 sets][kt][kh][2 kchunk][8 co-group][8 co][8 k], then 512 zeros
  __nv_bfloat16* y;          // [N][To][Ho][Wo][64]
  const float* bias;
  float* stats;
  int N, Ti, Hi, Wi, To, Ho, Wo;
  int hq;        // ceil(Ho / 4)
  int numIters;  // N * To * hq
};

constexpr int kS3Threads = 192;
constexpr int kS3Stages = 4;
constexpr int kS3OutRows = 4;
constexpr int kS3Rows = kS3OutRows + 2;             // input rows per stage
constexpr int kS3StageBytes = kS3Rows * 1024;
constexpr int kS3SetBytes = 9 * 2048;               // (frame tap, filter row) x [64 co x 16 k] bf16
constexpr int kS3FilterBytes = 2 * kS3SetBytes;     // set O, set E

__device__ __forceinline__ uint64_t s3_desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  return d;
}

__device__ __forceinline__ void s3_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(kS3Threads, 1) conv_stem3_kernel(const Stem3Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* filt = smem + kS3Stages * kS3StageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(filt + kS3FilterBytes);
  uint64_t* empty_bar = full_bar + kS3Stages;
  uint64_t* acc_full = empty_bar + kS3Stages;   // [2]
  uint64_t* acc_empty = acc_full + 2;           // [2]
  uint64_t* filt_bar = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(filt_bar + 1);

  const int t = threadIdx.x;
  const int warp = t >> 5;
  // zero the row slots once: the 16 bytes in front of a row and everything behind it are never written afterwards
  for (int i = t; i < kS3Stages * kS3StageBytes / 16; i += kS3Threads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (t == 0) {
    for (int s = 0; s < kS3Stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 128);
    }
    mbar_init(filt_bar, 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ------------------------------------------------------------------ producer: lane r copies input row r of the stage
    const int lane = t & 31;
    const uint32_t rowBytes = static_cast<uint32_t>(p.Wi) * 8u;
    const __nv_bfloat16* zero_row = p.wst + kS3FilterBytes / 2;
    if (lane == 0) {
      mbar_arrive_expect_tx(filt_bar, kS3FilterBytes);
      s3_bulk_g2s(smem_u32(filt), p.wst, kS3FilterBytes, filt_bar);
    }
    int s = 0;
    uint32_t ph = 1;
    for (int it = blockIdx.x; it < p.numIters; it += gridDim.x) {
      const int hq = it % p.hq;
      const int q = it / p.hq;
      const int to = q % p.To, n = q / p.To;
      const int hi = hq * kS3OutRows - 1 + lane;
      for (int a = 0; a < 3; ++a) {
        mbar_wait(&empty_bar[s], ph);
        const int ti = to - 1 + a;
        if (lane == 0) mbar_arrive_expect_tx(&full_bar[s], kS3Rows * rowBytes);
        __syncwarp();
        if (lane < kS3Rows) {
          const bool ok = ti >= 0 && ti < p.Ti && hi >= 0 && hi < p.Hi;
          const __nv_bfloat16* src = ok ? p.x + ((static_cast<size_t>(n) * p.Ti + ti) * p.Hi + hi) * p.Wi * 4 : zero_row;
          s3_bulk_g2s(smem_u32(smem + s * kS3StageBytes) + lane * 1024 + 16, src, rowBytes, &full_bar[s]);
        }
        if (++s == kS3Stages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp < 4) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp, lane = t & 31;
    float ssum[2] = {0.f, 0.f}, ssq[2] = {0.f, 0.f};  // running channel sums: lane l <-> channels l and 32 + l
    uint32_t iter_ctr = 0;
    for (int it = blockIdx.x; it < p.numIters; it += gridDim.x, ++iter_ctr) {
      const int buf = iter_ctr & 1;
      const uint32_t ph = (iter_ctr >> 1) & 1;
      const int hq = it % p.hq;
      const int q = it / p.hq;
      const int to = q % p.To, n = q / p.To;
      mbar_wait(&acc_full[buf], ph);
      tc_fence_after_sync();
      // columns outer, tiles inner: BN statistics are reduced across lanes once per 32 columns and iteration
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        float ra[32], qa[32];
#pragma unroll
        for (int jx = 0; jx < 32; ++jx) ra[jx] = qa[jx] = 0.f;
#pragma unroll 1
        for (int tile = 0; tile < 4; ++tile) {   // tile = row pair * 2 + parity
          const int rp = tile >> 1, par = tile & 1;
          const int ho = hq * kS3OutRows + 2 * rp + (ew >> 1);
          const int ow = 2 * ((ew & 1) * 32 + lane) + par;
          const bool ok = ho < p.Ho && ow < p.Wo;
          __nv_bfloat16* orow =
              p.y + ((((static_cast<size_t>(n) * p.To + to) * p.Ho + (ok ? ho : 0)) * p.Wo) + (ok ? ow : 0)) * 64;
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + buf * 256 + tile * 64 + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int jx = 0; jx < 32; jx += 8) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              f[e] = __uint_as_float(v[jx + e]);
              if (p.bias) f[e] += __ldg(p.bias + c0 + jx + e);
            }
            uint4 o;
            o.x = pack_bf16x2(f[0], f[1]);
            o.y = pack_bf16x2(f[2], f[3]);
            o.z = pack_bf16x2(f[4], f[5]);
            o.w = pack_bf16x2(f[6], f[7]);
            if (ok) *reinterpret_cast<uint4*>(orow + c0 + jx) = o;
            if (p.stats && ok) {
              const uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float lo = __uint_as_float(w[e] << 16), hi = __uint_as_float(w[e] & 0xffff0000u);
                ra[jx + 2 * e] += lo;
                ra[jx + 2 * e + 1] += hi;
                qa[jx + 2 * e] = fmaf(lo, lo, qa[jx + 2 * e]);
                qa[jx + 2 * e + 1] = fmaf(hi, hi, qa[jx + 2 * e + 1]);
              }
            }
          }
        }
        if (p.stats) {
          warp_column_sums(ra, lane);
          warp_column_sums(qa, lane);
          ssum[c0 >> 5] += ra[0];
          ssq[c0 >> 5] += qa[0];
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&acc_empty[buf]);
    }
    if (p.stats) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        atomicAdd(p.stats + h * 32 + lane, ssum[h]);
        atomicAdd(p.stats + 64 + h * 32 + lane, ssq[h]);
      }
    }
  } else {
    // ------------------------------------------------------------------ MMA lane
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      mbar_wait(filt_bar, 0);
      const uint64_t bbase = s3_desc_nosw(smem_u32(filt), 1024, 128);
      int s = 0;
      uint32_t ph = 0, iter_ctr = 0;
      for (int it = blockIdx.x; it < p.numIters; it += gridDim.x, ++iter_ctr) {
        const int buf = iter_ctr & 1;
        mbar_wait(&acc_empty[buf], ((iter_ctr >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t dcol = tmem_base + buf * 256;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after_sync();
          const uint64_t abase = s3_desc_nosw(smem_u32(smem + s * kS3StageBytes), 16, 128);
#pragma unroll
          for (int b = 0; b < 3; ++b) {
#pragma unroll
            for (int tile = 0; tile < 4; ++tile) {
              const int rp = tile >> 1, par = tile & 1;
              // A: stage rows (2*rp + b) and (2*rp + b + 1) = output rows 2*rp, 2*rp + 1 for filter row b; odd pixels start at
              // the row's pixel 0 (byte 16) with filter set O, even pixels one window earlier (byte 0) with filter set E
              umma_bf16(dcol + tile * 64, abase + static_cast<uint64_t>(((2 * rp + b) * 1024 + par * 16) >> 4),
                        bbase + static_cast<uint64_t>(((1 - par) * kS3SetBytes + (a * 3 + b) * 2048) >> 4), idesc,
                        (a | b) != 0);
            }
          }
          umma_commit(&empty_bar[s]);
          if (a == 2) umma_commit(&acc_full[buf]);
          if (++s == kS3Stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
    __syncwarp();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, 512);
}

// w fp32 [64][Ci<=4][3][3][3] -> wst bf16 [set][a][b][2][8][8][8] + 512 zeros; K slot q = kchunk*8 + e: window pixel q/4,
// channel q%4; set 0 (odd pixels): kw tap = slot (slot 3 zero); set 1 (even pixels): kw tap = slot - 1 (slot 0 zero)
__global__ void pack_weight_stem3_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wst, int Co, int Ci) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * 9 * 1024 + 512) return;
  float v = 0.f;
  if (idx < 2 * 9 * 1024) {
    const int set = idx / (9 * 1024), rem = idx - set * 9 * 1024;
    const int e = rem & 7, r = (rem >> 3) & 7, g = (rem >> 6) & 7, j = (rem >> 9) & 1;
    const int ab = rem >> 10;
    const int a = ab / 3, b = ab - a * 3;
    const int co = g * 8 + r;
    const int qk = j * 8 + e;
    const int c = (qk >> 2) - set, ch = qk & 3;
    if (co < Co && ch < Ci && c >= 0 && c < 3) v = w[(((static_cast<size_t>(co) * Ci + ch) * 3 + a) * 3 + b) * 3 + c];
  }
  wst[idx] = __float2bfloat16(v);
}

// =====================================================================================================================
// wgrad of the same layer:  dW[co][ch][a][b][c] = sum_pixels dY[pixel][co] * X[pixel + (a-1, b-1, c-1)][ch]
//   A = dY of one output row, pixels of ONE parity (TMA box with a traversal stride of 2 along w: 64 pixel rows x 128 B,
//       128B-swizzled, pixels >= Wo zero)                                  [K = 64 pixels] x [M = 64 co], MN-major
//   B = the 9 raw input rows (3 frame taps x 3 filter rows, 1 KB apart, pixels at byte 16): K = pixel pairs 16 B apart,
//       N chunk = one 2-pixel window (8 elements) of every row (SBO = 1 KB) -> N = 72; chunk j = window slots 2j, 2j+1
//   odd pixels w = 2k+1: window = pixels 2k .. 2k+3 (slot s <-> tap c = s); even pixels w = 2k: window 2k-2 .. 2k+1
//   (slot s <-> tap c = s-1), i.e. the same rows read 16 bytes earlier.
//   D[parity][j] = [64 co] x [72]: 4 accumulators of 72 TMEM columns kept for the whole kernel, added into dW with atomics.
// grid: one persistent CTA per SM over the output rows (n, t, h).
// =====================================================================================================================
struct Stem3WgradParams {
  CUtensorMap tmapDy;       // dy as {64, Wo, N*To*Ho}, box {64, 128 (stride 2 -> 64 pixels), 1}
  const __nv_bfloat16* x;   // [N][Ti][Hi][Wi][4]
  const __nv_bfloat16* zero_row;
  float* dw;                // [Co][Ci][3][3][3] fp32, accumulated atomically
  int N, Ti, Hi, Wi;
  int Co, Ci;               // logical
  int numRows;              // N*Ti*Hi
};

constexpr int kS3WStages = 6;
constexpr int kS3WDyBytes = 2 * 8192;                       // even pixels, odd pixels
constexpr int kS3WStageBytes = kS3WDyBytes + 9 * 1024;

__global__ void __launch_bounds__(192, 1) conv_stem3_wgrad_kernel(const __grid_constant__ Stem3WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // 1 KB of zeros follows the stages: windows of pixel pairs >= Wi/2 of the last row are read (against zero dY rows)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kS3WStages * kS3WStageBytes + 1024);
  uint64_t* empty_bar = full_bar + kS3WStages;
  uint64_t* accum_bar = empty_bar + kS3WStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
  const int t = threadIdx.x;
  const int warp = t >> 5;
  int iters = 0;
  for (int r = blockIdx.x; r < p.numRows; r += gridDim.x) ++iters;
  for (int i = t; i < (kS3WStages * kS3WStageBytes + 1024) / 16; i += 192) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (t == 0) {
    for (int s = 0; s < kS3WStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (iters > 0) {
    if (warp == 4) {
      // ---------------- producer: two TMA boxes (even / odd dY pixels) + one bulk copy per raw input row (lane = row)
      const int lane = t & 31;
      const uint32_t rowBytes = static_cast<uint32_t>(p.Wi) * 8u;
      const int al = lane / 3, b = lane - al * 3;
      if (lane == 0) tma_prefetch_desc(&p.tmapDy);
      int s = 0;
      uint32_t ph = 1;
      int h = blockIdx.x % p.Hi, tt = (blockIdx.x / p.Hi) % p.Ti, n = blockIdx.x / (p.Hi * p.Ti);
      const int dh = gridDim.x % p.Hi, dt = (gridDim.x / p.Hi) % p.Ti, dn = gridDim.x / (p.Hi * p.Ti);
      for (int r = blockIdx.x; r < p.numRows; r += gridDim.x) {
        mbar_wait(&empty_bar[s], ph);
        const uint32_t stage = smem_u32(smem + s * kS3WStageBytes);
        if (lane == 0) {
          mbar_arrive_expect_tx(&full_bar[s], kS3WDyBytes + 9 * rowBytes);
#pragma unroll
          for (int par = 0; par < 2; ++par)
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                ::"r"(stage + par * 8192), "l"(&p.tmapDy), "r"(smem_u32(&full_bar[s])), "r"(0), "r"(par), "r"(r)
                : "memory");
        }
        __syncwarp();
        if (lane < 9) {
          const int ti = tt - 1 + al, hi = h - 1 + b;
          const bool ok = ti >= 0 && ti < p.Ti && hi >= 0 && hi < p.Hi;
          const __nv_bfloat16* src = ok ? p.x + ((static_cast<size_t>(n) * p.Ti + ti) * p.Hi + hi) * p.Wi * 4 : p.zero_row;
          s3_bulk_g2s(stage + kS3WDyBytes + lane * 1024 + 16, src, rowBytes, &full_bar[s]);
        }
        if (++s == kS3WStages) {
          s = 0;
          ph ^= 1;
        }
        h += dh;
        tt += dt;
        n += dn;
        if (h >= p.Hi) {
          h -= p.Hi;
          ++tt;
        }
        if (tt >= p.Ti) {
          tt -= p.Ti;
          ++n;
        }
      }
    } else if (warp < 4) {
      // ---------------- epilogue: TMEM lanes of an M=64 accumulator: co = 16*warp + lane, lanes 16..31 unused
      mbar_wait(accum_bar, 0);
      tc_fence_after_sync();
      const int lane = t & 31;
      const int co = warp * 16 + lane;
      for (int acc = 0; acc < 4; ++acc) {          // acc = parity * 2 + j
        const int par = acc >> 1, j = acc & 1;
        for (int c0 = 0; c0 < 72; c0 += 8) {
          uint32_t v[8];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                       : "r"(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + acc * 72 + c0)
                       : "memory");
          tmem_ld_wait();
          if (lane < 16 && co < p.Co) {
            const int fr = c0 >> 3;
            const int a = fr / 3, b = fr - a * 3;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int slot = 2 * j + (e >> 2), ch = e & 3, c = slot - (1 - par);
              if (ch < p.Ci && c >= 0 && c < 3)
                atomicAdd(p.dw + (((static_cast<size_t>(co) * p.Ci + ch) * 3 + a) * 3 + b) * 3 + c, __uint_as_float(v[e]));
            }
          }
        }
      }
    } else {
      // ---------------- MMA lane
      if (elect_one()) {
        constexpr uint32_t idesc = make_idesc_bf16(64, 72, 1, 1);
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after_sync();
          const uint32_t stage = smem_u32(smem + s * kS3WStageBytes);
#pragma unroll
          for (int par = 0; par < 2; ++par) {
            // A: 16 pixels = two 8-row groups of the swizzled dY panel.  B: rows 1 KB apart (SBO), pixel pairs 16 B apart
            // with 8-pair groups 128 B apart (LBO); odd pixels start at the row's pixel 0 (byte 16), even pixels 16 B earlier
            const uint64_t abase = make_smem_desc_sw128(stage + par * 8192, 8192, 1024);
            const uint64_t bbase = s3_desc_nosw(stage + kS3WDyBytes + par * 16, 128, 1024);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                umma_bf16(tmem_base + (par * 2 + j) * 72, abase + static_cast<uint64_t>((ks * 2048) >> 4),
                          bbase + static_cast<uint64_t>((j * 16 + ks * 256) >> 4), idesc, (it | ks) != 0);
              }
            }
          }
          umma_commit(&empty_bar[s]);
          if (it == iters - 1) umma_commit(accum_bar);
          if (++s == kS3WStages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, 512);
}

int launch_stem3_wgrad(const rsp_conv3d_desc* d, int Ci_logical, int Co_logical, const void* x, const void* dy,
                       void* zero_row_1k, float* dw, int accumulate, int sm_count, cudaStream_t stream) {
  Stem3WgradParams p{};
  p.x = static_cast<const __nv_bfloat16*>(x);
  // rows outside the clip are copied from 1 KB of zeros at the head of the caller's workspace
  if (rsp::zero_async(zero_row_1k, 1024, stream) != cudaSuccess) {
    set_error("stem3 wgrad: workspace memset failed");
    return RSP_ERR_CUDA;
  }
  p.zero_row = static_cast<const __nv_bfloat16*>(zero_row_1k);
  p.dw = dw;
  p.N = d->N; p.Ti = d->Ti; p.Hi = d->Hi; p.Wi = d->Wi;
  p.Co = Co_logical; p.Ci = Ci_logical;
  p.numRows = p.N * p.Ti * p.Hi;
  const size_t nw = static_cast<size_t>(Co_logical) * Ci_logical * 27;
  if (!accumulate) {
    cudaError_t e = rsp::zero_async(dw, nw * sizeof(float), stream);
    if (e != cudaSuccess) {
      set_error("stem3 wgrad memset: %s", cudaGetErrorString(e));
      return RSP_ERR_CUDA;
    }
  }
  {
    const unsigned long long dims[3] = {64, static_cast<unsigned long long>(p.Wi), static_cast<unsigned long long>(p.numRows)};
    const unsigned long long strides[2] = {128, static_cast<unsigned long long>(p.Wi) * 128};
    const unsigned box[3] = {64, 128, 1};
    const unsigned estr[3] = {1, 2, 1};
    int rc = make_tmap_bf16_strided(&p.tmapDy, dy, 3, dims, strides, box, estr);
    if (rc != RSP_OK) return rc;
  }
  constexpr int smem = kS3WStages * kS3WStageBytes + 1024 + 1024 + 256;
  cudaError_t e = cudaFuncSetAttribute(conv_stem3_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(conv_stem3_wgrad): %s", cudaGetErrorString(e));
    return RSP_ERR_CUDA;
  }
  const int grid = p.numRows < sm_count ? p.numRows : sm_count;
  conv_stem3_wgrad_kernel<<<grid, 192, smem, stream>>>(p);
  return check_launch("conv_stem3_wgrad");
}

bool stem3_supported(const rsp_conv3d_desc* d) {
  return d->Ci == 4 && d->Co == 64 && d->kt == 3 && d->kh == 3 && d->kw == 3 && d->st == 1 && d->sh == 1 && d->sw == 1 &&
         d->pt == 1 && d->ph == 1 && d->pw == 1 && (d->Wi % 2) == 0 && d->Wi <= 124;
}

int pack_stem3(int Ci_logical, int Co_logical, const float* w, void* wst, cudaStream_t stream) {
  pack_weight_stem3_kernel<<<(2 * 9 * 1024 + 512 + 255) / 256, 256, 0, stream>>>(w, static_cast<__nv_bfloat16*>(wst), Co_logical,
                                                                       Ci_logical);
  return check_launch("pack_weight_stem3");
}

int launch_stem3(const rsp_conv3d_desc* d, const void* x, const void* wst, const float* bias, void* y, float* stats,
                 int sm_count, cudaStream_t stream) {
  Stem3Params p{};
  p.x = static_cast<const __nv_bfloat16*>(x);
  p.wst = static_cast<const __nv_bfloat16*>(wst);
  p.y = static_cast<__nv_bfloat16*>(y);
  p.bias = bias;
  p.stats = stats;
  p.N = d->N; p.Ti = d->Ti; p.Hi = d->Hi; p.Wi = d->Wi;
  p.To = d->Ti; p.Ho = d->Hi; p.Wo = d->Wi;
  p.hq = (p.Ho + kS3OutRows - 1) / kS3OutRows;
  p.numIters = p.N * p.To * p.hq;
  constexpr int smem = kS3Stages * kS3StageBytes + kS3FilterBytes + 1024 + 256;
  cudaError_t e = cudaFuncSetAttribute(conv_stem3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(conv_stem3): %s", cudaGetErrorString(e));
    return RSP_ERR_CUDA;
  }
  const int grid = p.numIters < sm_count ? p.numIters : sm_count;
  conv_stem3_kernel<<<grid, kS3Threads, smem, stream>>>(p);
  return check_launch("conv_stem3");
}

}  // namespace rsp
