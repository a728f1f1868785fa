// Direct (im2col-free) tcgen05 convolution for the RGB stem: kt x kh x 7 filters, stride (st, sh, 2), 4 stored input
// channels (RGB + pad), 64 output channels  (reference: models/resnet.py:130-136 conv1 7x7x7 s(1,2,2);
// models/r2plus1d_vcop.py conv1.spatial_conv 1x7x7 s(1,2,2)).
//
// Observation: with 8 B per input pixel and a w-stride of 2, the patches of two neighbouring output pixels start
// exactly 16 B apart inside a raw input row.  That is the row pitch of the *non-swizzled K-major* UMMA operand layout
// (core matrix = 8 rows x 16 B), so a raw input row sitting in shared memory already IS the A operand of
//   D[ow, co] += sum_{q<8 px, c<4} X[row, 2*ow - 4 + q, c] * W[co, c, a, b, q-1]
// (rows overlap, which a descriptor does not mind): descriptor start = row address, LBO = 16 B (next 2 pixels),
// SBO = 128 B (next 8 output pixels).  Rows are stored 1024 B apart per h-stride phase so that output row ho+1
// (input row + sh) is the second half of an M=128 tile.  No gather, no index math, no im2col traffic: each input
// row is read once per (tile, frame tap) with ONE bulk copy (cp.async.bulk, 8*Wi bytes) issued by a lane of the producer
// warp, the filter slab of one frame tap (kh*4 KB) with another; rows outside the image are zeroed in place.
//
// CTA (persistent, 1/SM): 4 epilogue warps, 1 producer warp, 1 MMA warp.  One iteration = 8 output rows of one
// (n, to) = two 128x128 accumulators in TMEM, double buffered (512 columns) so the epilogue of iteration i overlaps the
// MMAs of i+1.  Pipeline stage = one frame tap a: {21 input rows, kh*4 KB filter slab}, 4 stages.  (All 148 SMs stream
// the same 196 KB filter from L2 over and over: eight rows per slab fetch instead of four halves that traffic.)
//
// Row pairing (N = 128).  An SS-form M128 N64 K16 MMA holds the pipe for 48 clocks (operand fetch: 6 KB at 128 B/clk) for
// 32 clocks of math; N = 128 costs 64 for 64.  Input row r feeds output row h with filter row b = r - 2h AND output row
// h + 1 with filter row b - 2, so ONE MMA with B = [W[b]; W[b-2]] (128 filter columns) serves both from a single A fetch.
// A tile covers 4 output rows: M halves = input rows (r, r + 4) (rows are stored in four h-phases, 1 KB apart within a
// phase), N halves = output rows (+0, +1):  D[i][k] = output row 4m + 2i + k.  Per tile and frame tap the input rows
// rho = r - 2*h0 = 0..8 are visited: rho 2..6 as N = 128, rho 0,1 (only W[rho] -> k = 0) and rho 7,8 (only W[rho-2] ->
// k = 1) as N = 64 MMAs on one half of the accumulator: (5*64 + 4*48) * 2 clocks per 4 rows instead of 7*2*2*48 (-24 %).
// The filter slab keeps W[b-2] exactly 1 KB (= 8 co-groups * SBO) behind W[b]: [kchunk 4][b: 6,4,2,0,5,3,1][8][8][8].
// Frame taps that fall outside the clip (pt = 3: 12 of 112 (to, a) pairs) are skipped instead of multiplied by zeros.
#include "common.cuh"
#include "rspnet_b200.h"

namespace rsp {

int make_tmap_u64_rows(CUtensorMap* out, const void* base, int rank, const unsigned long long* dims,
                       const unsigned long long* strides_bytes, const unsigned* box);
int make_tmap_u64_rows_strided(CUtensorMap* out, const void* base, int rank, const unsigned long long* dims,
                               const unsigned long long* strides_bytes, const unsigned* box, const unsigned* estr);

struct StemParams {
  CUtensorMap tmapX;         // x as 8-byte pixels {Wi, Hi, Ti, N}, box {128, 24 at row stride 4 = 6 rows, 1, 1}, no swizzle
  const __nv_bfloat16* x;    // [N][Ti][Hi][Wi][4]
  const __nv_bfloat16* wst;  // [kt][4 kchunk][7 slots: b = 6,4,2,0,5,3,1][8 co-group][8 co][8 k] bf16
  __nv_bfloat16* y;          // [N][To][Ho][Wo][ldc], already offset to this launch's 64-channel group
  const float* bias;         // offset likewise
  float* stats;              // optional [2][ldc] (offset likewise): += per-channel sum / sum of squares of the stored output
  int ldc;                   // stored output channels (a multiple of 64; one launch per 64-channel group)
  int x0, ow0, wlim;         // column tile of this launch: input pixel held by slot 0 (2*ow0 - 4), first output column,
                             // number of output columns (<= 60); one launch per tile, a single tile up to Wi = 120
  int N, Ti, Hi, Wi, To, Ho, Wo;
  int kt, kh, st, sh, pt, ph;
  int hq;        // ceil(Ho / kStemOutRows)
  int numIters;  // N * To * hq
};

constexpr int kStemThreads = 192;
constexpr int kStemStages = 4;
constexpr int kStemRowBytes = 1024;   // 128 pixel slots of 8 B; slot s holds input pixel s - 4
constexpr int kStemOutRows = 8;       // output rows per iteration: two 128x128 tiles (four rows each) share one filter slab
constexpr int kStemPerPhase = 6;      // rows of one h-phase (input row mod 4), stored 1024 B apart
constexpr int kStemMaxRows = 4 * kStemPerPhase;  // row slots per A slab (21 used)
constexpr int kStemKChunk = 7 * 1024; // bytes of one 8-element K chunk of a filter slab (7 filter rows x 64 co x 16 B)
constexpr int kStemASlab = kStemMaxRows * kStemRowBytes;
constexpr int kStemBSlabMax = 7 * 4096;
constexpr int kStemStageBytes = kStemASlab + kStemBSlabMax;

__device__ __forceinline__ uint64_t make_smem_desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  return d;  // layout type 0: no swizzle
}

__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// frame taps a in [a_lo, a_hi) read a frame inside the clip; the others contribute nothing and are skipped by the
// producer and the MMA lane alike (an empty range cannot happen for pt < kt, but is mapped to "all taps, zero rows")
__device__ __forceinline__ void stem_tap_range(const StemParams& p, int to, int& a_lo, int& a_hi) {
  const int t0 = to * p.st - p.pt;
  a_lo = t0 < 0 ? -t0 : 0;
  a_hi = p.Ti - t0 < p.kt ? p.Ti - t0 : p.kt;
  if (a_lo >= a_hi) {
    a_lo = 0;
    a_hi = p.kt;
  }
}

__global__ void __launch_bounds__(kStemThreads, 1) conv_stem_kernel(const __grid_constant__ StemParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStemStages * kStemStageBytes);
  uint64_t* empty_bar = full_bar + kStemStages;
  uint64_t* acc_full = empty_bar + kStemStages;   // [2]
  uint64_t* acc_empty = acc_full + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int t = threadIdx.x;
  const int warp = t >> 5;
  const uint32_t bslab_bytes = static_cast<uint32_t>(p.kh) * 4096u;

  // zero the A slabs once: halo pixel slots are never written afterwards
  for (int i = t; i < kStemStages * kStemStageBytes / 16; i += kStemThreads)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (t == 0) {
    for (int s = 0; s < kStemStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 128);
    }
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ------------------------------------------------------------------ producer (one lane)
    // per stage: one bulk copy of the filter slab and one TMA box per h-phase: {128 pixel slots from pixel -4, every 4th
    // input row, 6 rows} lands as the 6 rows of that phase 1 KB apart; halo pixels and rows outside the image are zero-filled
    // by the TMA unit (the 21 per-row cp.async.bulk requests this replaces cost ~100 clocks each in the copy unit)
    if (elect_one()) {
      tma_prefetch_desc(&p.tmapX);
      int s = 0;
      uint32_t ph = 0;
      const uint32_t tx = bslab_bytes + 4u * kStemPerPhase * kStemRowBytes;
      for (int it = blockIdx.x; it < p.numIters; it += gridDim.x) {
        const int hq = it % p.hq;
        const int q = it / p.hq;
        const int to = q % p.To, n = q / p.To;
        const int hi0 = hq * kStemOutRows * p.sh - p.ph;
        int a_lo, a_hi;
        stem_tap_range(p, to, a_lo, a_hi);
        for (int a = a_lo; a < a_hi; ++a) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          const uint32_t aslab = smem_u32(smem + s * kStemStageBytes);
          const int ti = to * p.st - p.pt + a;
          mbar_arrive_expect_tx(&full_bar[s], tx);
          bulk_copy_g2s(aslab + kStemASlab, p.wst + static_cast<size_t>(a) * p.kh * 2048, bslab_bytes, &full_bar[s]);
#pragma unroll
          for (int phase = 0; phase < 4; ++phase) {
            asm volatile(
                "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                ::"r"(aslab + phase * kStemPerPhase * kStemRowBytes), "l"(&p.tmapX), "r"(smem_u32(&full_bar[s])),
                  "r"(p.x0), "r"(hi0 + phase), "r"(ti), "r"(n)
                : "memory");
          }
          if (++s == kStemStages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp < 4) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp;              // TMEM lane quarter
    const int lane = t & 31;
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
      // columns outer, tiles inner: the per-lane partial sums of the BN statistics run over all tiles of the iteration and
      // are reduced across lanes once per 32 columns (62 shuffles) instead of once per tile
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        float ra[32], qa[32];
#pragma unroll
        for (int jx = 0; jx < 32; ++jx) ra[jx] = qa[jx] = 0.f;
#pragma unroll 1
        for (int m = 0; m < kStemOutRows / 2; ++m) {   // 64-column block m: tile m / 2, N half m % 2
          const int ho = hq * kStemOutRows + 4 * (m >> 1) + 2 * (ew >> 1) + (m & 1);
          const int ow = (ew & 1) * 32 + lane;
          const bool ok = ho < p.Ho && ow < p.wlim;
          __nv_bfloat16* orow =
              p.y + ((((static_cast<size_t>(n) * p.To + to) * p.Ho + (ok ? ho : 0)) * p.Wo) + (ok ? p.ow0 + ow : 0)) * p.ldc;
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + buf * 256 + m * 64 + c0, v);
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
        atomicAdd(p.stats + p.ldc + h * 32 + lane, ssq[h]);
      }
    }
  } else {
    // ------------------------------------------------------------------ MMA issuer
    // one elected lane waits, issues and commits (no warp-level re-convergence between stages)
    constexpr uint32_t idesc128 = make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc64 = make_idesc_bf16(128, 64, 0, 0);
    uint32_t iter_ctr = 0;
    int s = 0;
    uint32_t ph = 0;
    const bool leader = elect_one();
    for (int it = blockIdx.x; leader && it < p.numIters; it += gridDim.x, ++iter_ctr) {
      const int buf = iter_ctr & 1;
      const uint32_t aph = (iter_ctr >> 1) & 1;
      const int to = (it / p.hq) % p.To;
      int a_lo, a_hi;
      stem_tap_range(p, to, a_lo, a_hi);
      mbar_wait(&acc_empty[buf], aph ^ 1);
      tc_fence_after_sync();
      for (int a = a_lo; a < a_hi; ++a) {
        mbar_wait(&full_bar[s], ph);
        fence_proxy_async_smem();
        tc_fence_after_sync();
        // kh == 7 and sh == 2 (stem_filter_ok): every descriptor is a compile-time offset from two bases, so the
        // issuing thread spends ~2 instructions per MMA instead of rebuilding 64-bit descriptors
        const uint32_t aslab = smem_u32(smem + s * kStemStageBytes);
        const uint64_t abase = make_smem_desc_nosw(aslab, 16, 128);
        const uint64_t bbase = make_smem_desc_nosw(aslab + kStemASlab, kStemKChunk, 128);
        const uint32_t dcol = tmem_base + buf * 256;
        const uint32_t fresh = (a == a_lo) ? 0u : 1u;
        // input rows rho = 2..6 first: the N = 128 MMA of rho = 2 is the one that may overwrite both halves of a tile
#pragma unroll
        for (int o = 0; o < 9; ++o) {
          constexpr int kOrder[9] = {2, 3, 4, 5, 6, 0, 1, 7, 8};
          const int rho = kOrder[o];
          // filter row whose 1 KB block the descriptor starts at, and the accumulator half an N = 64 MMA writes
          const int b0 = rho <= 6 ? rho : rho - 2;
          const int bslot = (b0 & 1) ? 4 + (5 - b0) / 2 : (6 - b0) / 2;
          const bool wide = rho >= 2 && rho <= 6;
          const int khalf = rho >= 7 ? 64 : 0;
#pragma unroll
          for (int m = 0; m < kStemOutRows / 4; ++m) {
            const int j = 8 * m + rho;                         // input row of the tile's first M half
            const int arow = ((j & 3) * kStemPerPhase + (j >> 2)) * kStemRowBytes;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              umma_bf16(dcol + m * 128 + khalf, abase + static_cast<uint64_t>((arow + ks * 32) >> 4),
                        bbase + static_cast<uint64_t>((ks * 2 * kStemKChunk + bslot * 1024) >> 4),
                        wide ? idesc128 : idesc64, (o | ks) != 0 ? 1u : fresh);
            }
          }
        }
        umma_commit(&empty_bar[s]);
        if (a == a_hi - 1) umma_commit(&acc_full[buf]);
        if (++s == kStemStages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, 512);
}

// w fp32 [64][Ci<=4][kt][7][7] -> wst bf16 [kt][4 kchunk][7 slots][8 co-group][8 co][8 k]; slot order b = 6,4,2,0,5,3,1 so
// that W[b-2] lies 1 KB behind W[b]; K slot q = kchunk*8 + e: pixel slot q/4 (kw = slot-1), ch q%4
__global__ void pack_weight_stem_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wst, int Co, int Ci,
                                        int kt, int kh, int kw) {
  size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  size_t total = static_cast<size_t>(kt) * kh * 2048;
  if (idx >= total) return;
  int e = idx & 7, r = (idx >> 3) & 7, g = (idx >> 6) & 7;
  int rest = static_cast<int>(idx >> 9);        // (a * 4 + kchunk) * 7 + slot
  int slot = rest % 7;
  int j = (rest / 7) & 3;
  int a = rest / 28;
  int b = slot < 4 ? 6 - 2 * slot : 5 - 2 * (slot - 4);
  int co = g * 8 + r;
  int qk = j * 8 + e;
  int pslot = qk >> 2, ch = qk & 3;
  int c = pslot - 1;
  float v = 0.f;
  if (co < Co && ch < Ci && c >= 0 && c < kw)
    v = w[(((static_cast<size_t>(co) * Ci + ch) * kt + a) * kh + b) * kw + c];
  wst[idx] = __float2bfloat16(v);
}

static bool stem_filter_ok(const rsp_conv3d_desc* d) {
  return d->Ci == 4 && d->kw == 7 && d->sw == 2 && d->pw == 3 && d->kh == 7 && d->sh == 2 && (d->Wi % 2) == 0 &&
         ((kStemOutRows - 1) * d->sh + d->kh) <= kStemMaxRows;
}

// fprop: any number of 64-channel output groups (R(2+1)D's 1x7x7 stem has 83 -> 128 stored channels) and any number of
// column tiles of up to 60 output pixels (S3D-G at 224 x 224: two tiles of 56) — one launch per (group, tile)
bool stem_fprop_supported(const rsp_conv3d_desc* d) {
  const int Wo = (d->Wi + 2 * d->pw - d->kw) / d->sw + 1;
  return stem_filter_ok(d) && d->Co % 64 == 0 && d->Co >= 64 && d->Co <= 256 && Wo >= 1 && Wo <= 240;
}

// wgrad (below): one launch per (64-channel group of dY, column tile of up to 60 output pixels)
bool stem_wgrad_supported(const rsp_conv3d_desc* d) {
  return stem_fprop_supported(d) && d->kh * 2 * 32 <= 448;
}

int launch_stem(const rsp_conv3d_desc* d, const void* x, const void* wst, const float* bias, void* y, float* stats,
                int sm_count, cudaStream_t stream) {
  StemParams p{};
  p.x = static_cast<const __nv_bfloat16*>(x);
  p.wst = static_cast<const __nv_bfloat16*>(wst);
  p.y = static_cast<__nv_bfloat16*>(y);
  p.bias = bias;
  p.stats = stats;
  p.N = d->N; p.Ti = d->Ti; p.Hi = d->Hi; p.Wi = d->Wi;
  p.To = (d->Ti + 2 * d->pt - d->kt) / d->st + 1;
  p.Ho = (d->Hi + 2 * d->ph - d->kh) / d->sh + 1;
  p.Wo = (d->Wi + 2 * d->pw - d->kw) / d->sw + 1;
  p.kt = d->kt; p.kh = d->kh; p.st = d->st; p.sh = d->sh; p.pt = d->pt; p.ph = d->ph;
  p.hq = (p.Ho + kStemOutRows - 1) / kStemOutRows;
  p.numIters = p.N * p.To * p.hq;
  {
    const unsigned long long W = p.Wi, H = p.Hi, T = p.Ti;
    const unsigned long long xdims[4] = {W, H, T, static_cast<unsigned long long>(p.N)};
    const unsigned long long xstrides[3] = {W * 8, H * W * 8, T * H * W * 8};
    const unsigned xbox[4] = {128, 4 * kStemPerPhase, 1, 1};
    const unsigned xestr[4] = {1, 4, 1, 1};
    int rc = make_tmap_u64_rows_strided(&p.tmapX, x, 4, xdims, xstrides, xbox, xestr);
    if (rc != RSP_OK) return rc;
  }
  constexpr int smem = kStemStages * kStemStageBytes + 1024 + 256;
  cudaError_t e = cudaFuncSetAttribute(conv_stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(conv_stem): %s", cudaGetErrorString(e));
    return RSP_ERR_CUDA;
  }
  int grid = p.numIters < sm_count ? p.numIters : sm_count;
  p.ldc = d->Co;
  const int wtiles = (p.Wo + 59) / 60, wt = (p.Wo + wtiles - 1) / wtiles;
  for (int g = 0; g < d->Co / 64; ++g) {   // one launch per 64-channel group (filter slabs are packed group by group)
    p.wst = static_cast<const __nv_bfloat16*>(wst) + static_cast<size_t>(g) * d->kt * d->kh * 2048;
    p.y = static_cast<__nv_bfloat16*>(y) + g * 64;
    p.bias = bias ? bias + g * 64 : nullptr;
    p.stats = stats ? stats + g * 64 : nullptr;
    for (int w = 0; w < wtiles; ++w) {     // and per column tile
      p.ow0 = w * wt;
      p.x0 = 2 * p.ow0 - 4;
      p.wlim = p.Wo - p.ow0 < wt ? p.Wo - p.ow0 : wt;
      conv_stem_kernel<<<grid, kStemThreads, smem, stream>>>(p);
      int rc = check_launch("conv_stem");
      if (rc != RSP_OK) return rc;
    }
  }
  return RSP_OK;
}

int pack_stem(const rsp_conv3d_desc* d, int Ci_logical, int Co_logical, const float* w, void* wst,
              cudaStream_t stream) {
  size_t total = static_cast<size_t>(d->kt) * d->kh * 2048;
  for (int g = 0; g < d->Co / 64; ++g) {   // one slab set per 64-channel output group
    const int co_left = Co_logical - g * 64;
    pack_weight_stem_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
        w + static_cast<size_t>(g) * 64 * Ci_logical * d->kt * d->kh * d->kw,
        static_cast<__nv_bfloat16*>(wst) + g * total, co_left < 0 ? 0 : (co_left > 64 ? 64 : co_left), Ci_logical, d->kt,
        d->kh, d->kw);
    int rc = check_launch("pack_weight_stem");
    if (rc != RSP_OK) return rc;
  }
  return RSP_OK;
}


// =====================================================================================================================
// Stem wgrad with the same raw-row trick:  dW[co][c][a][b][kw] = sum_pixels dY[pixel][co] * X[patch(pixel)][kw, c].
//   A = raw input rows of up to 32 filter rows (a, b)   [64 px (K)] x [M = 16 rows x 8 k-slots], MN-major, no swizzle:
//       pixel pitch 16 B (K; overlapping windows), 8-pixel K groups 128 B apart (LBO), the same 2-pixel window of the
//       next filter row 1 KB further (SBO) -> one instruction covers 16 filter rows (M = 128)
//   B = dY row panel   [64 px (K)] x [64 co (N)]   MN-major, 128 B swizzle (one pixel = one 128 B row; TMA box, pixels
//       >= Wo zero-filled)
//   D[set][j] = [16 rows x 8 slots] x [64 co]: window chunk j (slots 2j, 2j+1) of filter-row set `set`; a CTA keeps
//       2 sets x 4 chunks = 512 TMEM columns while it streams output rows, then adds them into dW with fp32 atomics.
//   N = 128 by shifting dY: sum_k X[window(k) chunk j] dY[k-1] = sum_k X[window(k+1) chunk j] dY[k] and window(k+1) chunk j
//       is window(k) chunk j+1 (one output pixel = two input pixels = one chunk).  So the B operand {dY[k-1] | dY[k]} —
//       the same panel, second half one 128-byte row further — makes ONE M128 x N128 MMA (64 clocks, full rate) produce
//       chunks j+1 and j where two N = 64 MMAs took 2 x 48.  The dY box starts at pixel -1 (zero-filled) for that.
// The 49 filter rows are split over two groups of CTAs (25 + 24 rows), so dY is read twice (an M = 64 formulation with
// the roles swapped needs four groups and runs the tensor core at half rate).
// grid: x = workers over output rows within a group, y = group.
// =====================================================================================================================
struct StemWgradParams {
  CUtensorMap tmapDy;       // dy as {Co stored, Wo, N*To*Ho}, box {64, 72, 1} from (co0, pixel -1): one output row per copy, zero-filled outside
  CUtensorMap tmapX;        // x as 8-byte pixels {Wi, Hi, Ti, N}, box {128, kh, 1, 1}, no swizzle
  const __nv_bfloat16* x;   // [N][Ti][Hi][Wi][4]
  float* dw;                // [Co][Ci][kt][kh][kw] fp32, accumulated atomically
  int N, Ti, Hi, Wi, To, Ho, Wo;
  int kt, kh, kw, st, sh, pt, ph;
  int Co, Ci;               // logical (Co: channels of this launch's 64-channel group, dw points at its first filter)
  int co0;                  // first stored dY channel of the group
  int x0;                   // input pixel held by slot 0 of a raw row: 2 * (first output column of the tile) - 4
  int numRows;              // N*To*Ho
  int rowsPerGroup;         // filter rows (a, b) per CTA group (<= 32)
};

constexpr int kSWStages = 5;
constexpr int kSWDyRows = 72;              // pixel -1 .. 70 of one output row (64 K rows + the shifted panel's extra row)
constexpr int kSWDyBytes = kSWDyRows * 128;
constexpr int kSWRowSlots = 32;
constexpr int kSWStageBytes = kSWDyBytes + kSWRowSlots * kStemRowBytes;

__global__ void __launch_bounds__(192, 1) conv_stem_wgrad_kernel(const __grid_constant__ StemWgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // 1 KB of zeros follows the stages: pixels ow >= Wo of the last raw row are read (and multiplied by zero dY rows),
  // so whatever lies behind it must be finite bf16 data, never barrier words
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kSWStages * kSWStageBytes + 1024);
  uint64_t* empty_bar = full_bar + kSWStages;
  uint64_t* accum_bar = empty_bar + kSWStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int t = threadIdx.x;
  const int warp = t >> 5;
  const int frTotal = p.kt * p.kh;
  const int f0 = blockIdx.y * p.rowsPerGroup;                               // first filter row of this CTA
  const int nr = (frTotal - f0) < p.rowsPerGroup ? (frTotal - f0) : p.rowsPerGroup;
  const int nsets = (nr + 15) >> 4;                                          // 1 or 2 sets of 16 filter rows
  int iters = 0;
  for (int r = blockIdx.x; r < p.numRows; r += gridDim.x) ++iters;

  // zero everything once: halo pixel slots and unused row slots are never written afterwards
  for (int i = t; i < (kSWStages * kSWStageBytes + 1024) / 16; i += 192)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (t == 0) {
    for (int s = 0; s < kSWStages; ++s) {
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

  if (iters > 0 && nr > 0) {
    if (warp == 4) {
      // ---------------- producer (one lane): one TMA box for the dY row, one per frame tap for its kh raw input rows
      // ({128 pixel slots from pixel -4, kh rows}: rows land 1 KB apart in filter-row order, everything outside the clip —
      // halo pixels, rows above / below the image, frames before / after the clip — is zero-filled by the TMA unit).
      // (25 per-row cp.async.bulk requests per stage bounded this kernel at ~100 clocks per request.)
      if (elect_one()) {
        tma_prefetch_desc(&p.tmapDy);
        tma_prefetch_desc(&p.tmapX);
        const int a0 = f0 / p.kh, na = nr / p.kh;
        const uint32_t tx = kSWDyBytes + static_cast<uint32_t>(na * p.kh) * kStemRowBytes;
        int s = 0;
        uint32_t ph = 0;
        int ho = blockIdx.x % p.Ho, to = (blockIdx.x / p.Ho) % p.To, n = blockIdx.x / (p.Ho * p.To);
        const int dho = gridDim.x % p.Ho, dto = (gridDim.x / p.Ho) % p.To, dn = gridDim.x / (p.Ho * p.To);
        for (int r = blockIdx.x; r < p.numRows; r += gridDim.x) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          const uint32_t stage = smem_u32(smem + s * kSWStageBytes);
          mbar_arrive_expect_tx(&full_bar[s], tx);
          asm volatile(
              "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
              ::"r"(stage), "l"(&p.tmapDy), "r"(smem_u32(&full_bar[s])), "r"(p.co0), "r"(-1), "r"(r)
              : "memory");
          for (int al = 0; al < na; ++al) {
            asm volatile(
                "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                ::"r"(stage + kSWDyBytes + al * p.kh * kStemRowBytes), "l"(&p.tmapX), "r"(smem_u32(&full_bar[s])),
                  "r"(p.x0), "r"(ho * p.sh - p.ph), "r"(to * p.st - p.pt + a0 + al), "r"(n)
                : "memory");
          }
          if (++s == kSWStages) {
            s = 0;
            ph ^= 1;
          }
          ho += dho;
          to += dto;
          n += dn;
          if (ho >= p.Ho) {
            ho -= p.Ho;
            ++to;
          }
          if (to >= p.To) {
            to -= p.To;
            ++n;
          }
        }
      }
    } else if (warp < 4) {
      // ---------------- epilogue: D lane = (filter row within the set) * 8 + k-slot element, columns = co; the 64-column
      // block `blk` of a set holds window chunk blk ^ 1 (the shifted dY panel comes first)
      mbar_wait(accum_bar, 0);
      tc_fence_after_sync();
      const int lane = t & 31;
      const int l = warp * 32 + lane;
      const int frl = l >> 3, e = l & 7;
      const int ch = e & 3;
      for (int set = 0; set < nsets; ++set) {
        const int fr = f0 + set * 16 + frl;
        const bool row_ok = (set * 16 + frl) < nr;
        const int a = fr / p.kh, b = fr - a * p.kh;
        for (int blk = 0; blk < 4; ++blk) {
          const int j = blk ^ 1;
          const int c = 2 * j + (e >> 2) - 1;           // window slot 2j + e/4 holds tap c = slot - 1
          const bool ok = row_ok && ch < p.Ci && c >= 0 && c < p.kw;
          float* dst = p.dw + ((static_cast<size_t>(ok ? ch : 0) * p.kt + (ok ? a : 0)) * p.kh + (ok ? b : 0)) * p.kw + (ok ? c : 0);
          const size_t co_stride = static_cast<size_t>(p.Ci) * p.kt * p.kh * p.kw;
#pragma unroll 1
          for (int c0 = 0; c0 < 64; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + (set * 4 + blk) * 64 + c0, v);
            tmem_ld_wait();
            if (ok) {
#pragma unroll
              for (int q = 0; q < 32; ++q)
                if (c0 + q < p.Co) atomicAdd(dst + static_cast<size_t>(c0 + q) * co_stride, __uint_as_float(v[q]));
            }
          }
        }
      }
    } else {
      // ---------------- MMA lane
      constexpr uint32_t idesc = make_idesc_bf16(128, 128, 1, 1);
      const bool leader = elect_one();
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; leader && it < iters; ++it) {
        mbar_wait(&full_bar[s], ph);
        fence_proxy_async_smem();
        tc_fence_after_sync();
        const uint32_t stage = smem_u32(smem + s * kSWStageBytes);
        // A: raw rows; M chunk = the same 2-pixel window (16 B) of 16 consecutive filter rows (rows are 1 KB apart -> SBO),
        //    K: pixels 16 B apart, 8-pixel groups 128 B apart (LBO).  B: 16 pixels = two 8-row groups of the dY panel;
        //    N panel 0 = rows k (pixel k - 1), N panel 1 = rows k + 1 (pixel k): LBO = one 128-byte row.
        const uint64_t abase = make_smem_desc_nosw(stage + kSWDyBytes, 128, kStemRowBytes);
        const uint64_t bbase = make_smem_desc_sw128(stage, 128, 1024);
        for (int set = 0; set < nsets; ++set) {
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {   // chunks 2jj + 1 (shifted panel) and 2jj
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              umma_bf16(tmem_base + (set * 2 + jj) * 128,
                        abase + static_cast<uint64_t>((set * 16 * kStemRowBytes + jj * 32 + ks * 256) >> 4),
                        bbase + static_cast<uint64_t>((ks * 2048) >> 4), idesc, (it | ks) != 0);
            }
          }
        }
        umma_commit(&empty_bar[s]);
        if (it == iters - 1) umma_commit(accum_bar);
        if (++s == kSWStages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, 512);
}

int launch_stem_wgrad(const rsp_conv3d_desc* d, int Ci_logical, int Co_logical, const void* x, const void* dy,
                      float* dw, int accumulate, int sm_count, cudaStream_t stream) {
  StemWgradParams p{};
  p.x = static_cast<const __nv_bfloat16*>(x);
  p.dw = dw;
  p.N = d->N; p.Ti = d->Ti; p.Hi = d->Hi; p.Wi = d->Wi;
  p.To = (d->Ti + 2 * d->pt - d->kt) / d->st + 1;
  p.Ho = (d->Hi + 2 * d->ph - d->kh) / d->sh + 1;
  p.Wo = (d->Wi + 2 * d->pw - d->kw) / d->sw + 1;
  p.kt = d->kt; p.kh = d->kh; p.kw = d->kw; p.st = d->st; p.sh = d->sh; p.pt = d->pt; p.ph = d->ph;
  p.Co = Co_logical; p.Ci = Ci_logical;
  p.numRows = p.N * p.To * p.Ho;
  const size_t nw = static_cast<size_t>(Co_logical) * Ci_logical * d->kt * d->kh * d->kw;
  if (!accumulate) {
    cudaError_t e = rsp::zero_async(dw, nw * sizeof(float), stream);
    if (e != cudaSuccess) {
      set_error("stem wgrad memset: %s", cudaGetErrorString(e));
      return RSP_ERR_CUDA;
    }
  }
  constexpr int smem = kSWStages * kSWStageBytes + 1024 + 1024 + 256;
  cudaError_t e = cudaFuncSetAttribute(conv_stem_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(conv_stem_wgrad): %s", cudaGetErrorString(e));
    return RSP_ERR_CUDA;
  }
  // whole frame taps per CTA group (a raw-row box holds the kh rows of one frame tap): 4 + 3 taps for kt = 7
  const int tapsPerGroup = kSWRowSlots / d->kh < d->kt ? kSWRowSlots / d->kh : d->kt;
  const int groups = (d->kt + tapsPerGroup - 1) / tapsPerGroup;
  p.rowsPerGroup = tapsPerGroup * d->kh;
  int workers = sm_count / groups;
  if (workers < 1) workers = 1;
  if (workers > p.numRows) workers = p.numRows;
  {
    const unsigned long long W = p.Wi, H = p.Hi, T = p.Ti;
    const unsigned long long xdims[4] = {W, H, T, static_cast<unsigned long long>(p.N)};
    const unsigned long long xstrides[3] = {W * 8, H * W * 8, T * H * W * 8};
    const unsigned xbox[4] = {128, static_cast<unsigned>(d->kh), 1, 1};
    int rc = make_tmap_u64_rows(&p.tmapX, x, 4, xdims, xstrides, xbox);
    if (rc != RSP_OK) return rc;
  }
  dim3 grid(workers, groups);
  const size_t perCo = static_cast<size_t>(Ci_logical) * d->kt * d->kh * d->kw;
  const unsigned long long Cst = static_cast<unsigned long long>(d->Co);
  const int wtiles = (p.Wo + 59) / 60, wt = (p.Wo + wtiles - 1) / wtiles;
  for (int w = 0; w < wtiles; ++w) {
    // column tile: the dY map is restricted to the tile's pixels (base shifted, W = tile width), so that the box from
    // pixel -1 zero-fills its first row and everything right of the tile — every output pixel is summed exactly once
    const int ow0 = w * wt, wlim = p.Wo - ow0 < wt ? p.Wo - ow0 : wt;
    const unsigned long long dims[3] = {Cst, static_cast<unsigned long long>(wlim), static_cast<unsigned long long>(p.numRows)};
    const unsigned long long strides[2] = {Cst * 2, static_cast<unsigned long long>(p.Wo) * Cst * 2};
    const unsigned box[3] = {64, kSWDyRows, 1};
    int rc = make_tmap_bf16(&p.tmapDy, static_cast<const __nv_bfloat16*>(dy) + static_cast<size_t>(ow0) * Cst, 3, dims,
                            strides, box);
    if (rc != RSP_OK) return rc;
    p.x0 = 2 * ow0 - 4;
    for (int g = 0; g * 64 < Co_logical; ++g) {   // one launch per 64-channel group of dY (x is re-read per group)
      p.co0 = g * 64;
      p.Co = Co_logical - g * 64 < 64 ? Co_logical - g * 64 : 64;
      p.dw = dw + static_cast<size_t>(g) * 64 * perCo;
      conv_stem_wgrad_kernel<<<grid, 192, smem, stream>>>(p);
      rc = check_launch("conv_stem_wgrad");
      if (rc != RSP_OK) return rc;
    }
  }
  return RSP_OK;
}

}  // namespace rsp
