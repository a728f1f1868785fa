// Direct tcgen05 convolution for the 3x3x3, unit-stride RGB stem: 4 stored input channels (RGB + pad) -> 64 channels
// (reference: models/c3d.py:21-25 conv1, 3 -> 64, kernel 3, padding 1, on 16 x 112 x 112 clips).
//
// The layer is HBM-bound (2 x 64 bytes written per 8 bytes read); the generic gather kernel spends ~40 instructions per
// 8-byte pixel and runs 10x slower than the output write.  Here the raw-row trick of conv_stem.cu is applied to a
// w-stride of 1: in the K-major no-swizzle UMMA layout consecutive A rows are 16 bytes = TWO pixels apart, so one raw input
// row in shared memory is the A operand of every second output pixel.  Each input row is therefore staged twice:
//   copy E (pixel x at byte 8 + 8x, i.e. the row starts at pixel -1): A row j = pixels 2j-1 .. 2j+2 -> output pixel 2j
//   copy O (pixel x at byte 8x):                                      A row j = pixels 2j .. 2j+3   -> output pixel 2j+1
// Both copies come from the same 4-D tensor map over {W, H, T, N} with 8-byte elements (one pixel = one element): the box
// starts at x = -1 or x = 0, is 128 pixels wide (the 1 KB smem row pitch) and R+2 rows high; everything outside the image
// (left/right halo, rows above/below, frames before/after the clip) is zero-filled by the TMA unit.
//   K = 16 = 4 pixel slots x 4 channels: slots 0..2 are the kw taps, slot 3 multiplies a zero filter entry.
//   M = 128 = 64 pixel pairs of output row h and 64 of row h+1 (input rows are 1 KB apart, SBO 128 B, 56 of 64 used).
// The whole filter (27 taps x 64 x 4, packed 18 KB) stays resident in shared memory.
//
// CTA (persistent, 1/SM): warps 0-3 epilogue, warp 4 producer lane, warp 5 MMA lane.  One iteration = 4 output rows of one
// (n, to) = 4 accumulators (row pair x parity) of 128 x 64, double buffered in TMEM (512 columns); stage = one frame tap.
#include "common.cuh"
#include "rspnet_b200.h"

namespace rsp {

int device_sm_count();

struct Stem3Params {
  CUtensorMap tmapX;         // {Wi, Hi, Ti, N} of 8-byte pixels, box {128, kS3Rows, 1, 1}, no swizzle
  const __nv_bfloat16* wst;  // [kt][kh][2 kchunk][8 co-group][8 co][8 k]
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
constexpr int kS3CopyBytes = kS3Rows * 1024;        // one copy (E or O) of the rows
constexpr int kS3StageBytes = 2 * kS3CopyBytes;
constexpr int kS3FilterBytes = 9 * 2048;            // (frame tap, filter row) x [64 co x 16 k] bf16

__device__ __forceinline__ uint64_t s3_desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  return d;
}

__device__ __forceinline__ void s3_tma_load_4d(uint32_t dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void s3_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(kS3Threads, 1) conv_stem3_kernel(const __grid_constant__ Stem3Params p) {
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
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ------------------------------------------------------------------ producer lane
    if (elect_one()) {
      tma_prefetch_desc(&p.tmapX);
      mbar_arrive_expect_tx(filt_bar, kS3FilterBytes);
      s3_bulk_g2s(smem_u32(filt), p.wst, kS3FilterBytes, filt_bar);
      int s = 0;
      uint32_t ph = 1;
      for (int it = blockIdx.x; it < p.numIters; it += gridDim.x) {
        const int hq = it % p.hq;
        const int q = it / p.hq;
        const int to = q % p.To, n = q / p.To;
        const int hi0 = hq * kS3OutRows - 1;
        for (int a = 0; a < 3; ++a) {
          mbar_wait(&empty_bar[s], ph);
          const uint32_t dst = smem_u32(smem + s * kS3StageBytes);
          mbar_arrive_expect_tx(&full_bar[s], kS3StageBytes);
          s3_tma_load_4d(dst, &p.tmapX, &full_bar[s], -1, hi0, to - 1 + a, n);                 // copy E
          s3_tma_load_4d(dst + kS3CopyBytes, &p.tmapX, &full_bar[s], 0, hi0, to - 1 + a, n);   // copy O
          if (++s == kS3Stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
    __syncwarp();
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
#pragma unroll 1
      for (int tile = 0; tile < 4; ++tile) {   // tile = row pair * 2 + parity
        const int rp = tile >> 1, par = tile & 1;
        const int ho = hq * kS3OutRows + 2 * rp + (ew >> 1);
        const int ow = 2 * ((ew & 1) * 32 + lane) + par;
        const bool ok = ho < p.Ho && ow < p.Wo;
        __nv_bfloat16* orow =
            p.y + ((((static_cast<size_t>(n) * p.To + to) * p.Ho + (ok ? ho : 0)) * p.Wo) + (ok ? ow : 0)) * 64;
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + buf * 256 + tile * 64 + c0, v);
          tmem_ld_wait();
          float r[32];
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
            if (p.stats) {
              const uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                r[jx + 2 * e] = ok ? __uint_as_float(w[e] << 16) : 0.f;
                r[jx + 2 * e + 1] = ok ? __uint_as_float(w[e] & 0xffff0000u) : 0.f;
              }
            }
          }
          if (p.stats) {
            float qq[32];
#pragma unroll
            for (int jx = 0; jx < 32; ++jx) qq[jx] = r[jx] * r[jx];
            warp_column_sums(r, lane);
            warp_column_sums(qq, lane);
            ssum[c0 >> 5] += r[0];
            ssq[c0 >> 5] += qq[0];
          }
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
              // A: copy `par`, input slab rows (2*rp + b) and (2*rp + b + 1) = output rows 2*rp, 2*rp + 1 for filter row b
              umma_bf16(dcol + tile * 64, abase + static_cast<uint64_t>((par * kS3CopyBytes + (2 * rp + b) * 1024) >> 4),
                        bbase + static_cast<uint64_t>(((a * 3 + b) * 2048) >> 4), idesc, (a | b) != 0);
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

// w fp32 [64][Ci<=4][3][3][3] -> wst bf16 [a][b][2][8][8][8]; K slot q = kchunk*8 + e: pixel slot q/4 (= kw tap, slot 3 is
// always zero), channel q%4
__global__ void pack_weight_stem3_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wst, int Co, int Ci) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 9 * 1024) return;
  const int e = idx & 7, r = (idx >> 3) & 7, g = (idx >> 6) & 7, j = (idx >> 9) & 1;
  const int ab = idx >> 10;
  const int a = ab / 3, b = ab - a * 3;
  const int co = g * 8 + r;
  const int qk = j * 8 + e;
  const int c = qk >> 2, ch = qk & 3;
  float v = 0.f;
  if (co < Co && ch < Ci && c < 3) v = w[(((static_cast<size_t>(co) * Ci + ch) * 3 + a) * 3 + b) * 3 + c];
  wst[idx] = __float2bfloat16(v);
}

bool stem3_supported(const rsp_conv3d_desc* d) {
  return d->Ci == 4 && d->Co == 64 && d->kt == 3 && d->kh == 3 && d->kw == 3 && d->st == 1 && d->sh == 1 && d->sw == 1 &&
         d->pt == 1 && d->ph == 1 && d->pw == 1 && (d->Wi % 2) == 0 && d->Wi <= 126;
}

int pack_stem3(int Ci_logical, int Co_logical, const float* w, void* wst, cudaStream_t stream) {
  pack_weight_stem3_kernel<<<(9 * 1024 + 255) / 256, 256, 0, stream>>>(w, static_cast<__nv_bfloat16*>(wst), Co_logical,
                                                                       Ci_logical);
  return check_launch("pack_weight_stem3");
}

int make_tmap_u64_rows(CUtensorMap* out, const void* base, int rank, const unsigned long long* dims,
                       const unsigned long long* strides_bytes, const unsigned* box);

int launch_stem3(const rsp_conv3d_desc* d, const void* x, const void* wst, const float* bias, void* y, float* stats,
                 int sm_count, cudaStream_t stream) {
  Stem3Params p{};
  p.wst = static_cast<const __nv_bfloat16*>(wst);
  p.y = static_cast<__nv_bfloat16*>(y);
  p.bias = bias;
  p.stats = stats;
  p.N = d->N; p.Ti = d->Ti; p.Hi = d->Hi; p.Wi = d->Wi;
  p.To = d->Ti; p.Ho = d->Hi; p.Wo = d->Wi;
  p.hq = (p.Ho + kS3OutRows - 1) / kS3OutRows;
  p.numIters = p.N * p.To * p.hq;
  {
    const unsigned long long W = p.Wi, H = p.Hi, T = p.Ti, N = p.N;
    const unsigned long long dims[4] = {W, H, T, N};
    const unsigned long long strides[3] = {W * 8, H * W * 8, T * H * W * 8};
    const unsigned box[4] = {128, kS3Rows, 1, 1};
    int rc = make_tmap_u64_rows(&p.tmapX, x, 4, dims, strides, box);
    if (rc != RSP_OK) return rc;
  }
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
