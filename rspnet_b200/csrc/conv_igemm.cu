// Implicit-GEMM 3-D convolution on Blackwell tensor cores (tcgen05.mma, accumulators in TMEM).
//
// Replaces the cuDNN calls behind nn.Conv3d on the RSPNet pretraining path
// (reference call sites: models/resnet.py:21-27,130-136,170-175; models/c3d.py:21-50) and their
// autograd (convolution_backward).  Activations are bf16 NDHWC, accumulation is fp32.
//
// One smem image serves all three GEMMs: a "panel" is R pixel rows x 128 B (64 bf16 channels of one
// filter tap), 128-byte swizzled.
//   fprop : D[pixel, cout] = sum_k  A[pixel, k] * Wp[cout, k]     A,B K-major    (k = tap*Cin + ci)
//   dgrad : same kernel with the transposed gather (src = (dst + pad - tap)/stride) and Wd[ci, tap*Co+co]
//   wgrad : D[k, cout]     = sum_pixel A[pixel, k] * dY[pixel, cout]   A,B MN-major, split over pixels,
//           fp32 red.global.add into dWt[k, cout]
// Producer warps gather rows with zero-filling cp.async (LDGSTS) that complete on an mbarrier; one elected
// thread issues tcgen05.mma; the producer warps turn into the TMEM->register->global epilogue at the end.
//
// Two gather modes:
//   GENERIC: Cs % 64 == 0, one K-block = one tap x 64 channels (one 128 B row segment)
//   SMALLC : Cs == 4 (RGB padded), one K-block = S (kt,kh) rows x PXS w-pixels x 4 channels
#include "common.cuh"
#include "rspnet_b200.h"

namespace rsp {

int device_sm_count();

constexpr int kProducerThreads = 128;
constexpr int kThreads = 160;  // 4 producer/epilogue warps + 1 MMA warp

struct GatherGeom {
  const __nv_bfloat16* src;  // [N][Ts][Hs][Ws][Cs]
  int N, Ts, Hs, Ws, Cs;
  int Td, Hd, Wd;            // pixel grid that indexes GEMM rows
  int kt, kh, kw;
  int st, sh, sw;            // forward strides (for transposed: must be powers of two)
  int pt, ph, pw;
  int transposed;            // 0: src = d*s - p + k ; 1: src = (d + p - k)/s when divisible
  int pxs;                   // SMALLC: w-pixels per (kt,kh) segment (4 or 8)
  int numKb;                 // number of 64-wide K blocks that are walked
  int kbSkip;                // K blocks in front of them that are skipped: frame taps that read padding for EVERY pixel
                             // (a 3x3x3 filter on a one-frame tensor — R3D-18 layer4 — only ever sees its middle frame tap)
  long long M;               // N*Td*Hd*Wd
  int fast;                  // 1: source offset is linear in the tap (fprop, or dgrad with unit strides) and k <= 8
  unsigned long long mulW, mulH, mulT;  // ceil(2^sh / d) for row decoding without integer division
  int shW, shH, shT;
  // strided dgrad runs once per parity class of the destination grid (unit-stride problem in class space):
  int cls;                   // 1: the fields below are active
  int wa0, was, wb0, wbs, wc0, wcs, wkh, wkw;   // class tap (a,b,c) -> real filter tap (wa0+was*a, ...), real kh / kw
  int oT, oH, oW;            // full destination grid
  int ost, osh, osw, oot, ooh, oow;             // class pixel (t,h,w) -> real pixel (ost*t + oot, ...)
};

__device__ __forceinline__ uint32_t fdiv(uint32_t n, unsigned long long mul, int sh) {
  return static_cast<uint32_t>((static_cast<unsigned long long>(n) * mul) >> sh);
}

// (n, td, hd, wd) of GEMM row m (< 2^31) using the precomputed reciprocals
__device__ __forceinline__ void decode_fast(const GatherGeom& g, uint32_t m, int& n, int& td, int& hd, int& wd) {
  uint32_t q = fdiv(m, g.mulW, g.shW);
  wd = static_cast<int>(m - q * g.Wd);
  uint32_t q2 = fdiv(q, g.mulH, g.shH);
  hd = static_cast<int>(q - q2 * g.Hd);
  uint32_t q3 = fdiv(q2, g.mulT, g.shT);
  td = static_cast<int>(q2 - q3 * g.Td);
  n = static_cast<int>(q3);
}

// Per-row gather state of the fast path: element offset of the row's (n, t0, h0, w0, chunk) and per-tap validity bits
// (bit a: t tap a in range; bit 8+b: h tap b; bit 16+c: w tap c).  Source of tap (a,b,c) = base + dir*((a*Hs+b)*Ws+c)*Cs.
__device__ __forceinline__ void row_state(const GatherGeom& g, long long m, int chunk, int& base, uint32_t& mask) {
  if (m >= g.M) {
    base = 0;
    mask = 0;
    return;
  }
  int n, td, hd, wd;
  decode_fast(g, static_cast<uint32_t>(m), n, td, hd, wd);
  const int dir = g.transposed ? -1 : 1;
  const int t0 = g.transposed ? td + g.pt : td * g.st - g.pt;
  const int h0 = g.transposed ? hd + g.ph : hd * g.sh - g.ph;
  const int w0 = g.transposed ? wd + g.pw : wd * g.sw - g.pw;
  uint32_t mk = 0;
  for (int a = 0; a < g.kt; ++a) mk |= (static_cast<unsigned>(t0 + dir * a) < static_cast<unsigned>(g.Ts)) << a;
  for (int b = 0; b < g.kh; ++b) mk |= (static_cast<unsigned>(h0 + dir * b) < static_cast<unsigned>(g.Hs)) << (8 + b);
  for (int c = 0; c < g.kw; ++c) mk |= (static_cast<unsigned>(w0 + dir * c) < static_cast<unsigned>(g.Ws)) << (16 + c);
  mask = mk;
  base = (((n * g.Ts + t0) * g.Hs + h0) * g.Ws + w0) * g.Cs + chunk * 8;
}

struct ConvParams {
  CUtensorMap tmapW;         // wgt as a 2-D tensor {K, Nout}, box {64, NT}, 128 B swizzle: one TMA per K block
  GatherGeom g;
  const __nv_bfloat16* wgt;  // [Nout][numKb*64]
  __nv_bfloat16* out;        // [M][Nout]
  const float* bias;         // [Nout] or null
  int Nout;
  int wgtKb;                 // K blocks per filter row of `wgt` (== g.numKb unless a parity class uses a tap subset)
  float* acc;                // split-K: fp32 [M][Nout] accumulation buffer (zeroed by the host), else null
  int kbPerSplit;            // K blocks handled by one z-slice
  float* stats;              // optional [2][Nout]: += per-channel sum / sum of squares of the stored (bf16) output
};

// All parity classes of a strided dgrad in ONE launch: blockIdx.x walks the M tiles of the classes back to back
// (tileEnd = running tile count), each class with its own geometry; the filter operand (and its tensor map) is shared.
constexpr int kMaxClasses = 8;
struct ConvParamsMulti {
  ConvParams base;
  int ncls;
  int tileEnd[kMaxClasses];
  GatherGeom gx[kMaxClasses];
};
__device__ __forceinline__ const ConvParams& base_of(const ConvParams& p) { return p; }
__device__ __forceinline__ const ConvParams& base_of(const ConvParamsMulti& p) { return p.base; }
template <typename P> struct is_multi { static constexpr bool value = false; };
template <> struct is_multi<ConvParamsMulti> { static constexpr bool value = true; };

struct WgradParams {
  CUtensorMap tmapDy;        // dy as a 2-D tensor {Nout, M}, box {64, 64}, 128 B swizzle: one TMA per 64-channel panel
  GatherGeom g;
  const __nv_bfloat16* dy;   // [M][Nout]
  float* dwt;                // [numKb*64][Nout], accumulated with atomics
  int Nout;
  int pbPerSplit;            // 64-pixel blocks per z-slice
};

enum { MODE_GENERIC = 0, MODE_SMALLC = 1 };

// ---------------------------------------------------------------------------------------------
// row gather: where does (row pixel, K-block kb, 16B/8B chunk) come from?
// ---------------------------------------------------------------------------------------------
struct RowCoord {
  int base;  // pixel index of (n,0,0,0) in src
  int td, hd, wd;
  int ok;    // row < M
};

__device__ __forceinline__ RowCoord decode_row(const GatherGeom& g, long long m) {
  RowCoord r;
  r.ok = m < g.M;
  uint32_t mm = r.ok ? static_cast<uint32_t>(m) : 0u;  // host guarantees M < 2^31
  uint32_t q = mm / static_cast<uint32_t>(g.Wd);
  int wd = static_cast<int>(mm - q * g.Wd);
  uint32_t q2 = q / static_cast<uint32_t>(g.Hd);
  int hd = static_cast<int>(q - q2 * g.Hd);
  uint32_t q3 = q2 / static_cast<uint32_t>(g.Td);
  int td = static_cast<int>(q2 - q3 * g.Td);
  int n = static_cast<int>(q3);
  r.base = n * g.Ts * g.Hs * g.Ws;
  r.td = td; r.hd = hd; r.wd = wd;
  return r;
}

__device__ __forceinline__ bool map_coord(int d, int k, int s, int p, int transposed, int limit, int& out) {
  if (!transposed) {
    out = d * s - p + k;
    return out >= 0 && out < limit;
  }
  int num = d + p - k;
  if (num < 0 || (num & (s - 1))) return false;
  out = num >> (31 - __clz(s));
  return out < limit;
}

// GENERIC: one 16-byte chunk (8 channels) of a row for K-block kb.
__device__ __forceinline__ const __nv_bfloat16* generic_src(const GatherGeom& g, const RowCoord& r, int a, int b,
                                                           int c, int cc, int chunk, bool& valid) {
  int ts, hs, ws;
  bool v = r.ok;
  v = map_coord(r.td, a, g.st, g.pt, g.transposed, g.Ts, ts) && v;
  v = map_coord(r.hd, b, g.sh, g.ph, g.transposed, g.Hs, hs) && v;
  v = map_coord(r.wd, c, g.sw, g.pw, g.transposed, g.Ws, ws) && v;
  valid = v;
  if (!v) return g.src;
  size_t pix = static_cast<size_t>(r.base) + (static_cast<size_t>(ts) * g.Hs + hs) * g.Ws + ws;
  return g.src + pix * g.Cs + cc * 64 + chunk * 8;
}

// SMALLC: one 8-byte chunk (one RGBx pixel) of a row for K-block kb. sub in [0,16).
__device__ __forceinline__ const __nv_bfloat16* smallc_src(const GatherGeom& g, const RowCoord& r, int kb, int sub,
                                                          bool& valid) {
  int seg = sub / g.pxs, px = sub - seg * g.pxs;
  int segs = 16 / g.pxs;
  int rr = kb * segs + seg;
  bool v = r.ok && rr < g.kt * g.kh && px < g.kw;
  int a = rr / g.kh, b = rr - a * g.kh;
  int ts = r.td * g.st - g.pt + a, hs = r.hd * g.sh - g.ph + b, ws = r.wd * g.sw - g.pw + px;
  v = v && ts >= 0 && ts < g.Ts && hs >= 0 && hs < g.Hs && ws >= 0 && ws < g.Ws;
  valid = v;
  if (!v) return g.src;
  size_t pix = static_cast<size_t>(r.base) + (static_cast<size_t>(ts) * g.Hs + hs) * g.Ws + ws;
  return g.src + pix * 4;
}

// byte offset of logical (row, byte b) inside a 128B-swizzled panel whose base is 1024 B aligned
__device__ __forceinline__ uint32_t swz(int row, int b) {
  return static_cast<uint32_t>(row * 128 + ((((b >> 4) ^ (row & 7)) << 4) | (b & 15)));
}

// Gather ROWS rows of K-block kb into a panel. rows[] were decoded by the caller with the SAME thread mapping:
//   GENERIC: chunk = t & 7,  row_i = (t >> 3) + 16*i, i < ROWS/16
//   SMALLC : sub   = t & 15, row_i = (t >> 4) + 8*i,  i < ROWS/8
template <int MODE, int ROWS>
__device__ __forceinline__ void gather_panel(const GatherGeom& g, uint32_t panel, int kb, int t,
                                             const RowCoord* rows) {
  if (MODE == MODE_GENERIC) {
    const int cchunks = g.Cs >> 6;
    int tap = kb / cchunks, cc = kb - tap * cchunks;
    bool kbok = kb < g.numKb;
    int a = tap / (g.kh * g.kw);
    int rem = tap - a * g.kh * g.kw;
    int b = rem / g.kw, c = rem - b * g.kw;
    const int chunk = t & 7;
#pragma unroll
    for (int i = 0; i < ROWS / 16; ++i) {
      int row = (t >> 3) + 16 * i;
      bool valid;
      const __nv_bfloat16* s = generic_src(g, rows[i], a, b, c, cc, chunk, valid);
      valid = valid && kbok;
      cp_async16(panel + swz(row, chunk * 16), valid ? s : g.src, valid ? 16u : 0u);
    }
  } else {
    const int sub = t & 15;
    bool kbok = kb < g.numKb;
#pragma unroll
    for (int i = 0; i < ROWS / 8; ++i) {
      int row = (t >> 4) + 8 * i;
      bool valid;
      const __nv_bfloat16* s = smallc_src(g, rows[i], kb, sub, valid);
      valid = valid && kbok;
      cp_async8(panel + swz(row, sub * 8), valid ? s : g.src, valid ? 8u : 0u);
    }
  }
}

// Plain rows: ROWS rows of 128 B taken from a row-major matrix (weights or dY).
template <int ROWS>
__device__ __forceinline__ void load_rows(uint32_t panel, const __nv_bfloat16* base, long long row0, long long nrows,
                                          size_t ld, int t) {
  const int chunk = t & 7;
#pragma unroll
  for (int i = 0; i < ROWS / 16; ++i) {
    int row = (t >> 3) + 16 * i;
    bool valid = (row0 + row) < nrows;
    const __nv_bfloat16* s = base + (valid ? static_cast<size_t>(row0 + row) * ld + chunk * 8 : 0);
    cp_async16(panel + swz(row, chunk * 16), s, valid ? 16u : 0u);
  }
}

// ---------------------------------------------------------------------------------------------
// fprop / dgrad kernel:  out[M][Nout] = gather(src)[M][K] * wgt[Nout][K]^T (+ bias)
// ---------------------------------------------------------------------------------------------
template <int NT, int STAGES, int MODE, typename P>
__global__ void __launch_bounds__(kThreads) conv_igemm_kernel(const __grid_constant__ P pp) {
  const ConvParams& p = base_of(pp);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int A_BYTES = 128 * 128;
  constexpr int B_BYTES = NT * 128;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int ROWS_PER_THREAD = (MODE == MODE_GENERIC) ? 8 : 16;

  // dynamic smem is only guaranteed 16 B aligned: round up to the 1024 B the 128B swizzle needs
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
  float* red = reinterpret_cast<float*>(tmem_slot + 2);  // [2][NT] per-CTA channel sums (BN statistics)

  const int t = threadIdx.x;
  const int warp = t >> 5;
  unsigned tile = blockIdx.x;
  const GatherGeom* gp = &p.g;
  if constexpr (is_multi<P>::value) {
    int cls = 0;
    while (cls + 1 < pp.ncls && tile >= static_cast<unsigned>(pp.tileEnd[cls])) ++cls;
    if (cls) tile -= pp.tileEnd[cls - 1];
    gp = &pp.gx[cls];
  }
  const GatherGeom& g = *gp;
  const long long m0 = static_cast<long long>(tile) * 128;
  const int n0 = blockIdx.y * NT;
  // split-K: z-slice handles K blocks [kbBegin, kbBegin + numKb)
  const int kbBegin = blockIdx.z * p.kbPerSplit;
  const int numKb = (g.numKb - kbBegin) < p.kbPerSplit ? (g.numKb - kbBegin) : p.kbPerSplit;

  if (t == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], kProducerThreads);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (p.stats) {
    for (int i = t; i < 2 * NT; i += kThreads) red[i] = 0.f;
  }
  if (warp == 4) tmem_alloc(tmem_slot, NT);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ---------------- producers ----------------
    if (MODE == MODE_GENERIC && g.fast) {
      // fast path: ~10 instructions per 16 B copy (row state precomputed, taps advanced incrementally)
      const int chunk = t & 7;
      int rbase[8];
      uint32_t rmask[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) row_state(g, m0 + (t >> 3) + 16 * i, chunk, rbase[i], rmask[i]);
      const uint32_t dst0 = swz(t >> 3, chunk * 16);  // rows (t>>3)+16i share the swizzle phase: + i*2048
      const int dirCs = g.transposed ? -g.Cs : g.Cs;
      const int cchunks = g.Cs >> 6;
      int cc = (kbBegin + g.kbSkip) % cchunks;
      int tap = (kbBegin + g.kbSkip) / cchunks;
      int c = tap % g.kw;
      tap /= g.kw;
      int b = tap % g.kh;
      int a = tap / g.kh;
      for (int kb = 0; kb < numKb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        const int koff = ((a * g.Hs + b) * g.Ws + c) * dirCs + cc * 64;
        const int sb = 8 + b, sc = 16 + c;
        mbar_wait(&empty_bar[s], ph ^ 1);
        const uint32_t a_panel = smem_u32(smem + s * STAGE_BYTES) + dst0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t v = (rmask[i] >> a) & (rmask[i] >> sb) & (rmask[i] >> sc) & 1u;
          const __nv_bfloat16* src = g.src + (v ? static_cast<ptrdiff_t>(rbase[i] + koff) : 0);
          cp_async16(a_panel + i * 2048, src, v << 4);
        }
        const int kbw = g.cls ? (((g.wa0 + g.was * a) * g.wkh + (g.wb0 + g.wbs * b)) * g.wkw + (g.wc0 + g.wcs * c)) *
                                        cchunks + cc
                              : kbBegin + kb + g.kbSkip;
        cp_async_mbar_arrive(&full_bar[s]);
        if (t == 0) {   // the filter tile of this K block: one bulk tensor copy, counted in bytes on the same barrier
          mbar_arrive_expect_tx(&full_bar[s], B_BYTES);
          tma_load_2d(smem_u32(smem + s * STAGE_BYTES) + A_BYTES, &p.tmapW, &full_bar[s], kbw * 64, n0);
        } else {
          mbar_arrive(&full_bar[s]);
        }
        if (++cc == cchunks) {
          cc = 0;
          if (++c == g.kw) {
            c = 0;
            if (++b == g.kh) {
              b = 0;
              ++a;
            }
          }
        }
      }
    } else {
      RowCoord rows[ROWS_PER_THREAD];
#pragma unroll
      for (int i = 0; i < ROWS_PER_THREAD; ++i) {
        int row = (MODE == MODE_GENERIC) ? (t >> 3) + 16 * i : (t >> 4) + 8 * i;
        rows[i] = decode_row(g, m0 + row);
      }
      for (int kb = 0; kb < numKb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        const uint32_t a_panel = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t b_panel = a_panel + A_BYTES;
        gather_panel<MODE, 128>(g, a_panel, kbBegin + kb + g.kbSkip, t, rows);
        cp_async_mbar_arrive(&full_bar[s]);
        if (t == 0) {
          mbar_arrive_expect_tx(&full_bar[s], B_BYTES);
          tma_load_2d(b_panel, &p.tmapW, &full_bar[s], (kbBegin + kb + g.kbSkip) * 64, n0);
        } else {
          mbar_arrive(&full_bar[s]);
        }
      }
    }
    // ---------------- epilogue ----------------
    mbar_wait(accum_bar, 0);
    tc_fence_after_sync();
    const long long row = m0 + warp * 32 + (t & 31);
    const bool row_ok = row < g.M;
    size_t opix = static_cast<size_t>(row_ok ? row : 0);
    if (g.cls && row_ok) {
      int n, td, hd, wd;
      decode_fast(g, static_cast<uint32_t>(row), n, td, hd, wd);
      opix = ((static_cast<size_t>(n) * g.oT + td * g.ost + g.oot) * g.oH + hd * g.osh + g.ooh) * g.oW + wd * g.osw +
             g.oow;
    }
    __nv_bfloat16* orow = p.out + opix * p.Nout + n0;
#pragma unroll 1
    for (int c0 = 0; c0 < NT; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
      if (row_ok && p.acc) {
        float* arow = p.acc + opix * p.Nout + n0 + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(arow + j), "f"(__uint_as_float(v[j])),
                       "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                       : "memory");
        }
      } else {
        float r[32];  // the values as stored (bias added, rounded to bf16); zero for rows beyond M
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            f[e] = __uint_as_float(v[j + e]);
            if (p.bias) f[e] += __ldg(p.bias + n0 + c0 + j + e);
          }
          uint4 o;
          o.x = pack_bf16x2(f[0], f[1]);
          o.y = pack_bf16x2(f[2], f[3]);
          o.z = pack_bf16x2(f[4], f[5]);
          o.w = pack_bf16x2(f[6], f[7]);
          if (row_ok) *reinterpret_cast<uint4*>(orow + c0 + j) = o;
          if (p.stats) {
            const uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              r[j + 2 * e] = row_ok ? __uint_as_float(w[e] << 16) : 0.f;
              r[j + 2 * e + 1] = row_ok ? __uint_as_float(w[e] & 0xffff0000u) : 0.f;
            }
          }
        }
        if (p.stats) {
          float q[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) q[j] = r[j] * r[j];
          warp_column_sums(r, t & 31);
          warp_column_sums(q, t & 31);
          atomicAdd(&red[c0 + (t & 31)], r[0]);
          atomicAdd(&red[NT + c0 + (t & 31)], q[0]);
        }
      }
    }
    if (p.stats) {
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps only
      for (int i = t; i < NT; i += 128) {
        atomicAdd(p.stats + n0 + i, red[i]);
        atomicAdd(p.stats + p.Nout + n0 + i, red[NT + i]);
      }
    }
  } else {
    // ---------------- MMA issuer ----------------
    // one elected lane waits, issues and commits (no warp-level re-convergence between K blocks)
    constexpr uint32_t idesc = make_idesc_bf16(128, NT, 0, 0);
    if (elect_one()) {
      const uint64_t adesc0 = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
      const uint64_t bdesc0 = make_smem_desc_sw128(smem_u32(smem) + A_BYTES, 16, 1024);
      int s = 0;
      uint32_t ph = 0, acc = 0;
      uint64_t adesc = adesc0, bdesc = bdesc0;
      for (int kb = 0; kb < numKb; ++kb) {
        mbar_wait(&full_bar[s], ph);
        fence_proxy_async_smem();
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // +32 B per K=16 step inside the 128 B swizzle row (descriptor address is in 16 B units)
          umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, k == 0 ? acc : 1u);
        }
        umma_commit(&empty_bar[s]);
        acc = 1;
        adesc += STAGE_BYTES >> 4;
        bdesc += STAGE_BYTES >> 4;
        if (++s == STAGES) {
          s = 0;
          ph ^= 1;
          adesc = adesc0;
          bdesc = bdesc0;
        }
      }
      umma_commit(accum_bar);
    }
    __syncwarp();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, NT);
}

// ---------------------------------------------------------------------------------------------
// wgrad kernel: dwt[kb*64 + kk][cout] += sum_pixel gather(src)[pixel][kb*64+kk] * dy[pixel][cout]
// grid: x = pairs of K blocks (M tile = 128 rows of the K axis), y = cout tile, z = pixel split
// ---------------------------------------------------------------------------------------------
// threads: warps 0-3 gather + epilogue, warp 4 issues the MMAs, warps 5-8 gather only (fast path): one producer warp per
// scheduler retires ~0.5 instructions per clock, which made the gather the bound; dy arrives by TMA.
constexpr int kWgradThreads = 288;
template <int NT, int STAGES, int MODE>
__global__ void __launch_bounds__(kWgradThreads) conv_wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int PANEL = 64 * 128;  // 64 pixel rows x 128 B
  constexpr int NB = NT / 64;
  constexpr int STAGE_BYTES = (2 + NB) * PANEL;
  constexpr int ROWS_PER_THREAD = (MODE == MODE_GENERIC) ? 4 : 8;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int t = threadIdx.x;
  const int warp = t >> 5;
  const GatherGeom& g = p.g;
  const int kb0 = blockIdx.x * 2;
  const int n0 = blockIdx.y * NT;
  const long long totalPb = (g.M + 63) / 64;
  const long long pbBegin = static_cast<long long>(blockIdx.z) * p.pbPerSplit;
  long long pbEnd = pbBegin + p.pbPerSplit;
  if (pbEnd > totalPb) pbEnd = totalPb;
  const int iters = pbEnd > pbBegin ? static_cast<int>(pbEnd - pbBegin) : 0;

  const bool fastpath = MODE == MODE_GENERIC && !g.transposed && g.fast;
  if (t == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], fastpath ? 2 * kProducerThreads : kProducerThreads);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, NT);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (iters > 0) {
    if (warp != 4) {
      if (fastpath) {
        const int pt = warp < 4 ? t : t - 32;          // producer index 0..255: pixel rows (pt>>3) + 32 i, 16 B chunk pt&7
        const int chunk = pt & 7;
        const uint32_t dst0 = swz(pt >> 3, chunk * 16);
        const int cchunks = g.Cs >> 6;
        int ta[2], tb[2], tc[2], koff[2];
        bool kbok[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int kb = kb0 + j + g.kbSkip;
          kbok[j] = kb0 + j < g.numKb;
          const int tap = kb / cchunks, cc = kb - tap * cchunks;
          ta[j] = tap / (g.kh * g.kw);
          const int rem = tap - ta[j] * g.kh * g.kw;
          tb[j] = rem / g.kw;
          tc[j] = rem - tb[j] * g.kw;
          koff[j] = ((ta[j] * g.Hs + tb[j]) * g.Ws + tc[j]) * g.Cs + cc * 64;
        }
        for (int it = 0; it < iters; ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          const uint32_t prow0 = static_cast<uint32_t>((pbBegin + it) * 64) + (pt >> 3);
          int base[2], t0[2], h0[2], w0[2];
          bool ok[2];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const uint32_t m = prow0 + 32 * i;
            ok[i] = m < static_cast<uint32_t>(g.M);
            int n, td, hd, wd;
            decode_fast(g, ok[i] ? m : 0u, n, td, hd, wd);
            t0[i] = td * g.st - g.pt;
            h0[i] = hd * g.sh - g.ph;
            w0[i] = wd * g.sw - g.pw;
            base[i] = (((n * g.Ts + t0[i]) * g.Hs + h0[i]) * g.Ws + w0[i]) * g.Cs + chunk * 8;
          }
          mbar_wait(&empty_bar[s], ph ^ 1);
          const uint32_t stage0 = smem_u32(smem + s * STAGE_BYTES);
          if (pt == 0) {   // the dy panels of this pixel block: rows beyond M are zero-filled by the TMA unit
            mbar_arrive_expect_tx(&full_bar[s], NB * PANEL);
#pragma unroll
            for (int j = 0; j < NB; ++j)
              tma_load_2d(stage0 + (2 + j) * PANEL, &p.tmapDy, &full_bar[s], n0 + j * 64,
                          static_cast<int>((pbBegin + it) * 64));
          }
          const uint32_t stage = stage0 + dst0;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const bool v = ok[i] && kbok[j] && static_cast<unsigned>(t0[i] + ta[j]) < static_cast<unsigned>(g.Ts) &&
                             static_cast<unsigned>(h0[i] + tb[j]) < static_cast<unsigned>(g.Hs) &&
                             static_cast<unsigned>(w0[i] + tc[j]) < static_cast<unsigned>(g.Ws);
              cp_async16(stage + j * PANEL + i * 4096, g.src + (v ? static_cast<ptrdiff_t>(base[i] + koff[j]) : 0),
                         v ? 16u : 0u);
            }
          }
          cp_async_mbar_arrive(&full_bar[s]);
          if (pt != 0) mbar_arrive(&full_bar[s]);
        }
      } else if (warp < 4)
      for (int it = 0; it < iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        const long long prow0 = (pbBegin + it) * 64;
        RowCoord rows[ROWS_PER_THREAD];
#pragma unroll
        for (int i = 0; i < ROWS_PER_THREAD; ++i) {
          int row = (MODE == MODE_GENERIC) ? (t >> 3) + 16 * i : (t >> 4) + 8 * i;
          rows[i] = decode_row(g, prow0 + row);
        }
        mbar_wait(&empty_bar[s], ph ^ 1);
        const uint32_t stage = smem_u32(smem + s * STAGE_BYTES);
        if (t == 0) {
          mbar_arrive_expect_tx(&full_bar[s], NB * PANEL);
#pragma unroll
          for (int j = 0; j < NB; ++j)
            tma_load_2d(stage + (2 + j) * PANEL, &p.tmapDy, &full_bar[s], n0 + j * 64, static_cast<int>(prow0));
        }
        gather_panel<MODE, 64>(g, stage, kb0 + g.kbSkip, t, rows);
        gather_panel<MODE, 64>(g, stage + PANEL, kb0 + 1 + g.kbSkip, t, rows);
        cp_async_mbar_arrive(&full_bar[s]);
        if (t != 0) mbar_arrive(&full_bar[s]);
      }
      if (warp < 4) {
      // epilogue: D row = K index inside the pair of K blocks, columns = cout
      mbar_wait(accum_bar, 0);
      tc_fence_after_sync();
      const int krow = kb0 * 64 + warp * 32 + (t & 31);
      const bool row_ok = krow < g.numKb * 64;
      float* drow = p.dwt + static_cast<size_t>(row_ok ? krow + g.kbSkip * 64 : 0) * p.Nout + n0;
#pragma unroll 1
      for (int c0 = 0; c0 < NT; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + c0 + j),
                         "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])),
                         "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                         : "memory");
          }
        }
      }
      }
    } else {
      constexpr uint32_t idesc = make_idesc_bf16(128, NT, 1, 1);
      if (elect_one()) {
        // MN-major: 64-wide panels PANEL bytes apart (LBO), 8-pixel groups 1024 B apart (SBO)
        const uint64_t adesc0 = make_smem_desc_sw128(smem_u32(smem), PANEL, 1024);
        const uint64_t bdesc0 = make_smem_desc_sw128(smem_u32(smem) + 2 * PANEL, PANEL, 1024);
        int s = 0;
        uint32_t ph = 0, acc = 0;
        uint64_t adesc = adesc0, bdesc = bdesc0;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(&full_bar[s], ph);
          fence_proxy_async_smem();
          tc_fence_after_sync();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // K=16 pixels = two 8-pixel groups = 2048 B
            umma_bf16(tmem_base, adesc + 128 * k, bdesc + 128 * k, idesc, k == 0 ? acc : 1u);
          }
          umma_commit(&empty_bar[s]);
          acc = 1;
          adesc += STAGE_BYTES >> 4;
          bdesc += STAGE_BYTES >> 4;
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
            adesc = adesc0;
            bdesc = bdesc0;
          }
        }
        umma_commit(accum_bar);
      }
      __syncwarp();
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, NT);
}

// ---------------------------------------------------------------------------------------------
// weight packing / gradient unpacking (tiny, elementwise)
// ---------------------------------------------------------------------------------------------
// K index -> (ci, kt, kh, kw) of the PyTorch filter [Co][Ci][kt][kh][kw]; returns false for padding slots.
__device__ __forceinline__ bool k_to_filter(int mode, int k, int Cs, int Ci, int kt, int kh, int kw, int pxs,
                                            int& ci, int& a, int& b, int& c) {
  if (mode == MODE_GENERIC) {
    int tap = k / Cs;
    ci = k - tap * Cs;
    a = tap / (kh * kw);
    int rem = tap - a * kh * kw;
    b = rem / kw;
    c = rem - b * kw;
    return ci < Ci && a < kt;
  }
  int kb = k >> 6, e = k & 63;
  int sub = e >> 2;
  ci = e & 3;
  int seg = sub / pxs, px = sub - seg * pxs;
  int rr = kb * (16 / pxs) + seg;
  a = rr / kh;
  b = rr - a * kh;
  c = px;
  return ci < Ci && rr < kt * kh && px < kw;
}

// w fp32 [Co][Ci][kt][kh][kw] -> wp bf16 [CoPad][Kpad]  (fprop B operand).  One block per output channel: the
// filter of that channel (Ci*taps contiguous floats) is staged in smem with coalesced reads, then written in K order.
__global__ void __launch_bounds__(256) pack_weight_fprop_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp,
                                                                int mode, int Co, int Ci, int Cs, int kt, int kh, int kw,
                                                                int pxs, int Kpad) {
  extern __shared__ float sw[];
  const int co = blockIdx.x;
  const int per = Ci * kt * kh * kw;
  if (co < Co)
    for (int i = threadIdx.x; i < per; i += 256) sw[i] = w[static_cast<size_t>(co) * per + i];
  __syncthreads();
  const int taps = kt * kh * kw;
  for (int k = threadIdx.x; k < Kpad; k += 256) {
    int ci, a, b, c;
    float v = 0.f;
    if (co < Co && k_to_filter(mode, k, Cs, Ci, kt, kh, kw, pxs, ci, a, b, c))
      v = sw[ci * taps + (a * kh + b) * kw + c];
    wp[static_cast<size_t>(co) * Kpad + k] = __float2bfloat16(v);
  }
}

// The same for up to kPackBatch filters in one launch (a whole encoder is re-packed after every optimizer / EMA step):
// blockIdx.x walks the output channels of all jobs back to back.
constexpr int kPackBatch = 24;
struct PackJob {
  const float* w;
  __nv_bfloat16* wp;
  int mode, Co, CoPad, Ci, Cs, kt, kh, kw, pxs, Kpad, blk0;
};
struct PackBatch {
  PackJob j[kPackBatch];
  int n;
};

__global__ void __launch_bounds__(256) pack_weight_fprop_batch_kernel(const __grid_constant__ PackBatch b) {
  extern __shared__ float sw[];
  int ji = 0;
  while (ji + 1 < b.n && static_cast<int>(blockIdx.x) >= b.j[ji + 1].blk0) ++ji;
  const PackJob& J = b.j[ji];
  const int co = blockIdx.x - J.blk0;
  const int taps = J.kt * J.kh * J.kw;
  const int per = J.Ci * taps;
  if (co < J.Co) {
    const float* src = J.w + static_cast<size_t>(co) * per;
    if ((per & 3) == 0) {
      for (int i = threadIdx.x; i < (per >> 2); i += 256)
        reinterpret_cast<float4*>(sw)[i] = reinterpret_cast<const float4*>(src)[i];
    } else {
      for (int i = threadIdx.x; i < per; i += 256) sw[i] = src[i];
    }
  }
  __syncthreads();
  __nv_bfloat16* dst = J.wp + static_cast<size_t>(co) * J.Kpad;
  if (J.mode == MODE_GENERIC && (J.Cs & 1) == 0) {
    // k = tap * Cs + ci, walked without divisions; one thread packs two adjacent channels (one 32-bit store)
    int tap = 0, ci = 2 * threadIdx.x;
    while (ci >= J.Cs) { ci -= J.Cs; ++tap; }
    for (int k = 2 * threadIdx.x; k < J.Kpad; k += 512) {
      float v0 = 0.f, v1 = 0.f;
      if (co < J.Co && tap < taps) {
        if (ci < J.Ci) v0 = sw[ci * taps + tap];
        if (ci + 1 < J.Ci) v1 = sw[(ci + 1) * taps + tap];
      }
      *reinterpret_cast<uint32_t*>(dst + k) = pack_bf16x2(v0, v1);
      ci += 512;
      while (ci >= J.Cs) { ci -= J.Cs; ++tap; }
    }
    return;
  }
  for (int k = threadIdx.x; k < J.Kpad; k += 256) {
    int ci, a, bb, c;
    float v = 0.f;
    if (co < J.Co && k_to_filter(J.mode, k, J.Cs, J.Ci, J.kt, J.kh, J.kw, J.pxs, ci, a, bb, c))
      v = sw[ci * taps + (a * J.kh + bb) * J.kw + c];
    dst[k] = __float2bfloat16(v);
  }
}

// w fp32 [Co][Ci][kt][kh][kw] -> wd bf16 [CiPad][taps*CoPad]  (dgrad B operand: K index = tap*CoPad + co).
// Seen as matrices this is a plain transpose: w is [Co] x [R = Ci*taps] (r = ci*taps + tap) and wd is
// [CiPad*taps] x [CoPad] with the same r on its rows, so a 64 x 64 smem tile gives coalesced reads (along r) and
// coalesced bf16 writes (along co); rows / columns beyond the logical extents are zero.
__global__ void __launch_bounds__(256) pack_weight_dgrad_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wd,
                                                                int Co, int CoPad, int R, int Rpad) {
  __shared__ float tile[64][65];
  const int r0 = blockIdx.x * 64, co0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int co = co0 + ty + 4 * i, r = r0 + tx;
    tile[ty + 4 * i][tx] = (co < Co && r < R) ? w[static_cast<size_t>(co) * R + r] : 0.f;
  }
  __syncthreads();
  // each thread writes two adjacent output channels (one 32-bit store); a warp covers one 128 B row segment
  const int cx = (threadIdx.x & 31) * 2, ry = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + ry + 8 * i;
    if (r < Rpad && co0 + cx < CoPad)
      *reinterpret_cast<uint32_t*>(wd + static_cast<size_t>(r) * CoPad + co0 + cx) =
          pack_bf16x2(tile[cx][ry + 8 * i], tile[cx + 1][ry + 8 * i]);
  }
}

// dwt fp32 [Kpad][CoPad] -> dw fp32 [Co][Ci][kt][kh][kw]   (accumulate != 0: +=)
__global__ void unpack_wgrad_kernel(const float* __restrict__ dwt, float* __restrict__ dw, int mode, int Co, int CoPad,
                                    int Ci, int Cs, int kt, int kh, int kw, int pxs, int Kpad, int accumulate) {
  size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  size_t total = static_cast<size_t>(CoPad) * Kpad;
  if (idx >= total) return;
  // co fastest so that reads of dwt are coalesced
  int k = static_cast<int>(idx / CoPad), co = static_cast<int>(idx % CoPad);
  int ci, a, b, c;
  if (co >= Co || !k_to_filter(mode, k, Cs, Ci, kt, kh, kw, pxs, ci, a, b, c)) return;
  size_t o = (((static_cast<size_t>(co) * Ci + ci) * kt + a) * kh + b) * kw + c;
  float v = dwt[idx];
  dw[o] = accumulate ? dw[o] + v : v;
}

// The generic-mode unpack as a tiled transpose: dwt rows are k = tap * Cs + ci, dw rows are co with r = ci * taps + tap
// contiguous.  A CTA moves 32 output channels x ciTile input channels x all taps through shared memory: 128-byte
// coalesced reads along co, contiguous ciTile * taps float runs on the way out.
__global__ void __launch_bounds__(256) unpack_wgrad_tiled_kernel(const float* __restrict__ dwt, float* __restrict__ dw,
                                                                 int Co, int CoPad, int Ci, int Cs, int taps, int ciTile,
                                                                 int accumulate) {
  extern __shared__ float tile[];
  const int R = ciTile * taps, pitch = R | 1;
  const int co0 = blockIdx.x * 32, ci0 = blockIdx.y * ciTile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j = warp; j < R; j += 8) {
    const int tap = j / ciTile, cl = j - tap * ciTile;
    const int ci = ci0 + cl;
    float v = 0.f;
    if (ci < Ci && co0 + lane < CoPad) v = dwt[(static_cast<size_t>(tap) * Cs + ci) * CoPad + co0 + lane];
    tile[lane * pitch + cl * taps + tap] = v;
  }
  __syncthreads();
  const int valid = (Ci - ci0 < ciTile ? Ci - ci0 : ciTile) * taps;
  for (int row = warp; row < 32; row += 8) {
    const int co = co0 + row;
    if (co >= Co) break;
    float* dst = dw + (static_cast<size_t>(co) * Ci + ci0) * taps;
    const float* srow = tile + row * pitch;
    if (accumulate)
      for (int r = lane; r < valid; r += 32) dst[r] += srow[r];
    else
      for (int r = lane; r < valid; r += 32) dst[r] = srow[r];
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

static int fill_geom(GatherGeom& g, const rsp_conv3d_desc* d, int mode, int transposed) {
  // forward geometry
  const int To = (d->Ti + 2 * d->pt - d->kt) / d->st + 1;
  const int Ho = (d->Hi + 2 * d->ph - d->kh) / d->sh + 1;
  const int Wo = (d->Wi + 2 * d->pw - d->kw) / d->sw + 1;
  RSP_REQUIRE(To > 0 && Ho > 0 && Wo > 0, "conv3d: empty output");
  g.N = d->N;
  g.kt = d->kt; g.kh = d->kh; g.kw = d->kw;
  g.st = d->st; g.sh = d->sh; g.sw = d->sw;
  g.pt = d->pt; g.ph = d->ph; g.pw = d->pw;
  g.transposed = transposed;
  g.pxs = 0;
  g.kbSkip = 0;
  if (!transposed) {
    g.Ts = d->Ti; g.Hs = d->Hi; g.Ws = d->Wi; g.Cs = d->Ci;
    g.Td = To; g.Hd = Ho; g.Wd = Wo;
  } else {
    RSP_REQUIRE(is_pow2(d->st) && is_pow2(d->sh) && is_pow2(d->sw), "conv3d dgrad: strides must be powers of two");
    g.Ts = To; g.Hs = Ho; g.Ws = Wo; g.Cs = d->Co;
    g.Td = d->Ti; g.Hd = d->Hi; g.Wd = d->Wi;
  }
  g.M = static_cast<long long>(g.N) * g.Td * g.Hd * g.Wd;
  if (mode == MODE_GENERIC) {
    RSP_REQUIRE(g.Cs % 64 == 0, "conv3d: gathered channel count %d must be a multiple of 64", g.Cs);
    g.numKb = d->kt * d->kh * d->kw * (g.Cs / 64);
  } else {
    RSP_REQUIRE(!transposed, "conv3d: small-channel mode has no dgrad");
    RSP_REQUIRE(g.Cs == 4, "conv3d: small-channel mode needs Ci == 4 (got %d)", g.Cs);
    RSP_REQUIRE(d->kw <= 8, "conv3d: small-channel mode needs kw <= 8");
    g.pxs = d->kw <= 4 ? 4 : 8;
    const int segs = 16 / g.pxs;
    g.numKb = (d->kt * d->kh + segs - 1) / segs;
  }
  RSP_REQUIRE(static_cast<long long>(g.N) * g.Ts * g.Hs * g.Ws * g.Cs < (1ll << 31) && g.M < (1ll << 31),
              "conv3d: tensor too large");
  auto recip = [](int d, unsigned long long& mul, int& sh) {
    int l = 0;
    while ((1ll << l) < d) ++l;
    sh = 32 + l;  // exact for dividends < 2^31 (Granlund-Montgomery: shift >= 31 + ceil(log2 d))
    mul = ((1ull << sh) + d - 1) / d;
  };
  recip(g.Wd, g.mulW, g.shW);
  recip(g.Hd, g.mulH, g.shH);
  recip(g.Td, g.mulT, g.shT);
  g.fast = (mode == MODE_GENERIC) && d->kt <= 8 && d->kh <= 8 && d->kw <= 8 &&
           (!transposed || (d->st == 1 && d->sh == 1 && d->sw == 1));
  return RSP_OK;
}

static int conv_mode(const rsp_conv3d_desc* d) { return d->Ci == 4 ? MODE_SMALLC : MODE_GENERIC; }

// Frame taps that read temporal padding for every pixel of the problem contribute exact zeros: drop their K blocks
// (they form a prefix and a suffix of the tap order).  Returns the K blocks of the full filter row.
static int skip_dead_frame_taps(GatherGeom& g) {
  const int full = g.numKb;
  if (g.cls || (g.transposed && (g.st != 1 || g.sh != 1 || g.sw != 1))) return full;
  int aLo = g.kt, aHi = 0;
  for (int a = 0; a < g.kt; ++a) {
    bool live = false;
    for (int td = 0; td < g.Td && !live; ++td) {
      const int ts = g.transposed ? td + g.pt - a : td * g.st - g.pt + a;
      live = ts >= 0 && ts < g.Ts;
    }
    if (live) {
      if (a < aLo) aLo = a;
      aHi = a + 1;
    }
  }
  if (aLo >= aHi || (aLo == 0 && aHi == g.kt)) return full;
  const int perTap = g.kh * g.kw * (g.Cs / 64);
  g.kbSkip = aLo * perTap;
  g.numKb = (aHi - aLo) * perTap;
  return full;
}

// split-K finalize: out = bf16(acc + bias)
__global__ void __launch_bounds__(256) splitk_finalize_kernel(const float4* __restrict__ acc,
                                                              const float* __restrict__ bias,
                                                              uint2* __restrict__ out, size_t n4, int Nout) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * 256) {
    float4 v = acc[i];
    if (bias) {
      int c = static_cast<int>((i * 4) % Nout);
      v.x += bias[c]; v.y += bias[c + 1]; v.z += bias[c + 2]; v.w += bias[c + 3];
    }
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    out[i] = o;
  }
}

static int choose_splits(long long M, int Nout, int NT, int numKb, int sm_count) {
  const long long ctas = ((M + 127) / 128) * (Nout / NT);
  if (ctas * 2 > sm_count || numKb < 16) return 1;     // at least half a wave already, or K too short to split
  long long s = (2ll * sm_count + ctas - 1) / ctas;    // aim at ~2 CTAs per SM
  if (s > numKb / 8) s = numKb / 8;
  if (s > 16) s = 16;
  return s < 1 ? 1 : static_cast<int>(s);
}

template <int NT, int STAGES, int MODE>
static int launch_igemm(ConvParams& p, cudaStream_t stream) {
  if (p.g.M == 0) return RSP_OK;
  int splits = 1;
  if (p.acc && !p.g.cls) splits = choose_splits(p.g.M, p.Nout, NT, p.g.numKb, device_sm_count());
  float* acc = splits > 1 ? p.acc : nullptr;
  p.acc = acc;
  if (acc) p.stats = nullptr;  // split-K partial tiles cannot produce statistics of the final values
  p.kbPerSplit = (p.g.numKb + splits - 1) / splits;
  splits = (p.g.numKb + p.kbPerSplit - 1) / p.kbPerSplit;
  if (acc) {
    cudaError_t e = rsp::zero_async(acc, static_cast<size_t>(p.g.M) * p.Nout * sizeof(float), stream);
    if (e != cudaSuccess) {
      set_error("split-K memset: %s", cudaGetErrorString(e));
      return RSP_ERR_CUDA;
    }
  }
  {
    const unsigned long long dims[2] = {static_cast<unsigned long long>(p.wgtKb) * 64, static_cast<unsigned long long>(p.Nout)};
    const unsigned long long strides[1] = {static_cast<unsigned long long>(p.wgtKb) * 64 * 2};
    const unsigned box[2] = {64, NT};
    int rc = make_tmap_bf16(&p.tmapW, p.wgt, 2, dims, strides, box);
    if (rc != RSP_OK) return rc;
  }
  constexpr int smem = STAGES * (128 * 128 + NT * 128) + 1024 + 256 + 2 * NT * 4;
  auto kern = conv_igemm_kernel<NT, STAGES, MODE, ConvParams>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(conv_igemm): %s", cudaGetErrorString(e));
    return RSP_ERR_CUDA;
  }
  dim3 grid(static_cast<unsigned>((p.g.M + 127) / 128), static_cast<unsigned>(p.Nout / NT),
            static_cast<unsigned>(splits));
  kern<<<grid, kThreads, smem, stream>>>(p);
  int rc = check_launch("conv_igemm");
  if (rc != RSP_OK || !acc) return rc;
  const size_t n4 = static_cast<size_t>(p.g.M) * p.Nout / 4;
  unsigned fg = static_cast<unsigned>((n4 + 255) / 256);
  if (fg > 148u * 8) fg = 148u * 8;
  splitk_finalize_kernel<<<fg, 256, 0, stream>>>(reinterpret_cast<const float4*>(acc), p.bias,
                                                 reinterpret_cast<uint2*>(p.out), n4, p.Nout);
  return check_launch("splitk_finalize");
}

template <int MODE>
static int dispatch_igemm(ConvParams& p, cudaStream_t stream, int wgtKb = 0) {
  p.wgtKb = wgtKb > 0 ? wgtKb : p.g.numKb;
  if (p.Nout % 128 == 0) return launch_igemm<128, 3, MODE>(p, stream);
  return launch_igemm<64, 4, MODE>(p, stream);
}

template <int NT, int STAGES>
static int launch_igemm_multi(ConvParamsMulti& pm, cudaStream_t stream) {
  ConvParams& p = pm.base;
  p.acc = nullptr;
  p.stats = nullptr;
  p.kbPerSplit = 0;
  for (int i = 0; i < pm.ncls; ++i)
    if (pm.gx[i].numKb > p.kbPerSplit) p.kbPerSplit = pm.gx[i].numKb;
  const unsigned long long dims[2] = {static_cast<unsigned long long>(p.wgtKb) * 64, static_cast<unsigned long long>(p.Nout)};
  const unsigned long long strides[1] = {static_cast<unsigned long long>(p.wgtKb) * 64 * 2};
  const unsigned box[2] = {64, NT};
  int rc = make_tmap_bf16(&p.tmapW, p.wgt, 2, dims, strides, box);
  if (rc != RSP_OK) return rc;
  constexpr int smem = STAGES * (128 * 128 + NT * 128) + 1024 + 256 + 2 * NT * 4;
  auto kern = conv_igemm_kernel<NT, STAGES, MODE_GENERIC, ConvParamsMulti>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(conv_igemm multi): %s", cudaGetErrorString(e));
    return RSP_ERR_CUDA;
  }
  dim3 grid(static_cast<unsigned>(pm.tileEnd[pm.ncls - 1]), static_cast<unsigned>(p.Nout / NT), 1);
  kern<<<grid, kThreads, smem, stream>>>(pm);
  return check_launch("conv_igemm (dgrad classes)");
}

template <int NT, int STAGES, int MODE>
static int launch_wgrad(WgradParams& p, int sm_count, cudaStream_t stream) {
  constexpr int smem = STAGES * (2 + NT / 64) * 64 * 128 + 1024 + 256;
  auto kern = conv_wgrad_kernel<NT, STAGES, MODE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(conv_wgrad): %s", cudaGetErrorString(e));
    return RSP_ERR_CUDA;
  }
  const int mt = (p.g.numKb + 1) / 2, nt = p.Nout / NT;
  const long long totalPb = (p.g.M + 63) / 64;
  // pixel splits: the CTAs run in waves of 2 per SM, so pick the split count that minimises
  // (number of waves) x (pixel blocks per CTA + a fixed per-CTA cost for the prologue and the atomic epilogue);
  // at least 8 pixel blocks per split.  (602 CTAs on 592 slots is three waves, 588 is two.)
  const long long slots = 2ll * sm_count, tiles = static_cast<long long>(mt) * nt;
  long long maxSplits = (totalPb + 7) / 8;
  if (maxSplits < 1) maxSplits = 1;
  if (maxSplits > 4096) maxSplits = 4096;
  long long splits = 1, best = -1;
  for (long long sp = 1; sp <= maxSplits; ++sp) {
    const long long waves = (tiles * sp + slots - 1) / slots;
    const long long cost = waves * ((totalPb + sp - 1) / sp + 24);
    if (best < 0 || cost < best) {
      best = cost;
      splits = sp;
    }
    if (tiles * sp > 8 * slots) break;
  }
  p.pbPerSplit = static_cast<int>((totalPb + splits - 1) / splits);
  splits = (totalPb + p.pbPerSplit - 1) / p.pbPerSplit;
  {
    const unsigned long long dims[2] = {static_cast<unsigned long long>(p.Nout), static_cast<unsigned long long>(p.g.M)};
    const unsigned long long strides[1] = {static_cast<unsigned long long>(p.Nout) * 2};
    const unsigned box[2] = {64, 64};
    int rc = make_tmap_bf16(&p.tmapDy, p.dy, 2, dims, strides, box);
    if (rc != RSP_OK) return rc;
  }
  dim3 grid(mt, nt, static_cast<unsigned>(splits));
  kern<<<grid, kWgradThreads, smem, stream>>>(p);
  return check_launch("conv_wgrad");
}

template <int MODE>
static int dispatch_wgrad(WgradParams& p, int sm_count, cudaStream_t stream) {
  if (p.Nout % 128 == 0) return launch_wgrad<128, 3, MODE>(p, sm_count, stream);
  return launch_wgrad<64, 4, MODE>(p, sm_count, stream);
}

int device_sm_count();
bool stem_fprop_supported(const rsp_conv3d_desc* d);
bool stem_wgrad_supported(const rsp_conv3d_desc* d);
int launch_stem(const rsp_conv3d_desc* d, const void* x, const void* wst, const float* bias, void* y, float* stats,
                int sm_count, cudaStream_t stream);
int pack_stem(const rsp_conv3d_desc* d, int Ci_logical, int Co_logical, const float* w, void* wst,
              cudaStream_t stream);
int launch_stem_wgrad(const rsp_conv3d_desc* d, int Ci_logical, int Co_logical, const void* x, const void* dy,
                      float* dw, int accumulate, int sm_count, cudaStream_t stream);
bool stem3_supported(const rsp_conv3d_desc* d);
int pack_stem3(int Ci_logical, int Co_logical, const float* w, void* wst, cudaStream_t stream);
int launch_stem3(const rsp_conv3d_desc* d, const void* x, const void* wst, const float* bias, void* y, float* stats,
                 int sm_count, cudaStream_t stream);
int launch_stem3_wgrad(const rsp_conv3d_desc* d, int Ci_logical, int Co_logical, const void* x, const void* dy,
                       void* zero_row_1k, float* dw, int accumulate, int sm_count, cudaStream_t stream);
bool wgrad_direct_supported(const rsp_conv3d_desc* d);
int launch_wgrad_direct(const rsp_conv3d_desc* d, const void* x, const void* dy, float* dwt, cudaStream_t stream);
bool direct_supported(const rsp_conv3d_desc* d, int transposed);
int launch_direct(const rsp_conv3d_desc* d, int transposed, const void* x, const void* wgt, const float* bias, void* y,
                  float* stats, cudaStream_t stream);

}  // namespace rsp

using namespace rsp;

extern "C" {

int rsp_conv3d_kpad(const rsp_conv3d_desc* d, int which) {
  GatherGeom g{};
  const int mode = conv_mode(d);
  if (which == 1) {  // dgrad weights: K = taps * Co
    return d->kt * d->kh * d->kw * d->Co;
  }
  if (fill_geom(g, d, mode, 0) != RSP_OK) return RSP_ERR_INVALID;
  return g.numKb * 64;
}

int64_t rsp_conv3d_packed_elems(const rsp_conv3d_desc* d, int which) {
  const int kpad = rsp_conv3d_kpad(d, which);
  if (kpad < 0) return kpad;
  if (which == 1) return static_cast<int64_t>(d->Ci) * kpad;
  int64_t n = static_cast<int64_t>(d->Co) * kpad;
  if (stem_fprop_supported(d)) n += static_cast<int64_t>(d->Co / 64) * d->kt * d->kh * 2048;  // direct-conv filter slabs appended
  if (stem3_supported(d)) n += 2 * 9 * 1024 + 512;   // two filter sets + a zero row
  return n;
}

int rsp_conv3d_pack_weight(const rsp_conv3d_desc* d, int Ci_logical, int Co_logical, const float* w, void* wp,
                           int which, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int mode = conv_mode(d);
  if (which == 0) {
    GatherGeom g{};
    int rc = fill_geom(g, d, mode, 0);
    if (rc != RSP_OK) return rc;
    const int Kpad = g.numKb * 64;
    size_t total = static_cast<size_t>(d->Co) * Kpad;
    const size_t per = static_cast<size_t>(Ci_logical) * d->kt * d->kh * d->kw * sizeof(float);
    RSP_REQUIRE(per <= 200 * 1024, "pack_weight: one filter (%zu bytes) does not fit in shared memory", per);
    static bool attr_done = false;
    if (!attr_done) {
      cudaFuncSetAttribute(pack_weight_fprop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      attr_done = true;
    }
    pack_weight_fprop_kernel<<<d->Co, 256, per, stream>>>(w, static_cast<__nv_bfloat16*>(wp), mode, Co_logical,
                                                          Ci_logical, d->Ci, d->kt, d->kh, d->kw, g.pxs, Kpad);
    rc = check_launch("pack_weight_fprop");
    if (rc != RSP_OK) return rc;
    if (stem_fprop_supported(d))
      return pack_stem(d, Ci_logical, Co_logical, w, static_cast<__nv_bfloat16*>(wp) + total, stream);
    if (stem3_supported(d)) return pack_stem3(Ci_logical, Co_logical, w, static_cast<__nv_bfloat16*>(wp) + total, stream);
    return RSP_OK;
  }
  RSP_REQUIRE(mode == MODE_GENERIC, "pack_weight(dgrad): small-channel convs have no dgrad");
  const int taps = d->kt * d->kh * d->kw;
  RSP_REQUIRE(d->Co % 64 == 0 && d->Ci % 64 == 0, "pack_weight(dgrad): stored channel counts must be multiples of 64");
  const int R = Ci_logical * taps, Rpad = d->Ci * taps;
  dim3 grid((Rpad + 63) / 64, d->Co / 64);
  pack_weight_dgrad_kernel<<<grid, 256, 0, stream>>>(w, static_cast<__nv_bfloat16*>(wp), Co_logical, d->Co, R, Rpad);
  return check_launch("pack_weight_dgrad");
}

int rsp_conv3d_pack_weights(int32_t n, const rsp_conv3d_desc* descs, const int32_t* ci_logical,
                            const int32_t* co_logical, const float* const* w, void* const* wp, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(pack_weight_fprop_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_done = true;
  }
  for (int base = 0; base < n; base += kPackBatch) {
    PackBatch b{};
    b.n = (n - base) < kPackBatch ? (n - base) : kPackBatch;
    int blocks = 0;
    size_t smem = 0;
    for (int i = 0; i < b.n; ++i) {
      const rsp_conv3d_desc* d = descs + base + i;
      const int mode = conv_mode(d);
      GatherGeom g{};
      int rc = fill_geom(g, d, mode, 0);
      if (rc != RSP_OK) return rc;
      PackJob& J = b.j[i];
      J.w = w[base + i];
      J.wp = static_cast<__nv_bfloat16*>(wp[base + i]);
      J.mode = mode;
      J.Co = co_logical[base + i];
      J.CoPad = d->Co;
      J.Ci = ci_logical[base + i];
      J.Cs = d->Ci;
      J.kt = d->kt; J.kh = d->kh; J.kw = d->kw;
      J.pxs = g.pxs;
      J.Kpad = g.numKb * 64;
      J.blk0 = blocks;
      blocks += d->Co;
      const size_t per = static_cast<size_t>(J.Ci) * d->kt * d->kh * d->kw * sizeof(float);
      RSP_REQUIRE(per <= 200 * 1024, "pack_weights: one filter (%zu bytes) does not fit in shared memory", per);
      if (per > smem) smem = per;
    }
    pack_weight_fprop_batch_kernel<<<blocks, 256, smem, stream>>>(b);
    int rc = check_launch("pack_weight_fprop_batch");
    if (rc != RSP_OK) return rc;
    for (int i = 0; i < b.n; ++i) {
      const rsp_conv3d_desc* d = descs + base + i;
      if (stem_fprop_supported(d)) {
        rc = pack_stem(d, b.j[i].Ci, b.j[i].Co, b.j[i].w, b.j[i].wp + static_cast<size_t>(d->Co) * b.j[i].Kpad, stream);
        if (rc != RSP_OK) return rc;
      }
      if (stem3_supported(d)) {
        rc = pack_stem3(b.j[i].Ci, b.j[i].Co, b.j[i].w, b.j[i].wp + static_cast<size_t>(d->Co) * b.j[i].Kpad, stream);
        if (rc != RSP_OK) return rc;
      }
    }
  }
  return RSP_OK;
}

int64_t rsp_conv3d_workspace_bytes(const rsp_conv3d_desc* d, int which) {
  // which = 0: fprop (M = output pixels, N = Co); which = 1: dgrad (M = input pixels, N = Ci). 0 when split-K is unused.
  const long long To = (d->Ti + 2 * d->pt - d->kt) / d->st + 1, Ho = (d->Hi + 2 * d->ph - d->kh) / d->sh + 1,
                  Wo = (d->Wi + 2 * d->pw - d->kw) / d->sw + 1;
  const long long M = which == 0 ? d->N * To * Ho * Wo : static_cast<long long>(d->N) * d->Ti * d->Hi * d->Wi;
  const int Nout = which == 0 ? d->Co : d->Ci;
  const int Cs = which == 0 ? d->Ci : d->Co;
  if (Cs % 64 != 0 || Nout % 64 != 0) return 0;
  if (which == 1 && (d->st > 1 || d->sh > 1 || d->sw > 1)) return 0;
  if (direct_supported(d, which)) return 0;
  const int NT = Nout % 128 == 0 ? 128 : 64;
  const int numKb = d->kt * d->kh * d->kw * (Cs / 64);
  if (choose_splits(M, Nout, NT, numKb, device_sm_count()) <= 1) return 0;
  return M * Nout * static_cast<long long>(sizeof(float));
}

int rsp_conv3d_fprop(const rsp_conv3d_desc* d, const void* x, const void* wp, const float* bias, void* y,
                     void* workspace, float* stats, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int mode = conv_mode(d);
  ConvParams p{};
  int rc = fill_geom(p.g, d, mode, 0);
  if (rc != RSP_OK) return rc;
  RSP_REQUIRE(d->Co % 64 == 0, "conv3d fprop: Co=%d must be a multiple of 64", d->Co);
  if (stem_fprop_supported(d)) {
    const __nv_bfloat16* wst = static_cast<const __nv_bfloat16*>(wp) + static_cast<size_t>(d->Co) * p.g.numKb * 64;
    return launch_stem(d, x, wst, bias, y, stats, device_sm_count(), stream);
  }
  if (stem3_supported(d)) {
    const __nv_bfloat16* wst = static_cast<const __nv_bfloat16*>(wp) + static_cast<size_t>(d->Co) * p.g.numKb * 64;
    return launch_stem3(d, x, wst, bias, y, stats, device_sm_count(), stream);
  }
  if (mode == MODE_GENERIC && direct_supported(d, 0)) return launch_direct(d, 0, x, wp, bias, y, stats, stream);
  p.g.src = static_cast<const __nv_bfloat16*>(x);
  p.wgt = static_cast<const __nv_bfloat16*>(wp);
  p.out = static_cast<__nv_bfloat16*>(y);
  p.bias = bias;
  p.Nout = d->Co;
  p.acc = mode == MODE_GENERIC ? static_cast<float*>(workspace) : nullptr;
  p.stats = stats;
  if (mode != MODE_GENERIC) return dispatch_igemm<MODE_SMALLC>(p, stream);
  const int fullKb = skip_dead_frame_taps(p.g);
  return dispatch_igemm<MODE_GENERIC>(p, stream, fullKb);
}

// One parity class of a strided dgrad: destination pixels (st*t'+par_t, ...) only see the taps a = a0 + st*i with
// a0 = (par_t + pt) % st, and source (dY) index t' + (par_t + pt - a0)/st - i: a unit-stride transposed conv in class space.
static bool class_geom(const rsp_conv3d_desc* d, const int par[3], const void* dy, GatherGeom& g) {
  const int k[3] = {d->kt, d->kh, d->kw}, s[3] = {d->st, d->sh, d->sw}, pd[3] = {d->pt, d->ph, d->pw};
  const int in[3] = {d->Ti, d->Hi, d->Wi};
  int a0[3], na[3], off[3], cd[3];
  for (int i = 0; i < 3; ++i) {
    a0[i] = (par[i] + pd[i]) % s[i];
    na[i] = a0[i] < k[i] ? (k[i] - a0[i] + s[i] - 1) / s[i] : 0;
    off[i] = (par[i] + pd[i] - a0[i]) / s[i];
    cd[i] = in[i] > par[i] ? (in[i] - par[i] + s[i] - 1) / s[i] : 0;
    if (na[i] == 0 || cd[i] == 0) return false;  // no tap reaches this class: dx stays zero (memset by the caller)
  }
  g = GatherGeom{};
  g.N = d->N;
  g.Ts = (d->Ti + 2 * d->pt - d->kt) / d->st + 1;
  g.Hs = (d->Hi + 2 * d->ph - d->kh) / d->sh + 1;
  g.Ws = (d->Wi + 2 * d->pw - d->kw) / d->sw + 1;
  g.Cs = d->Co;
  g.Td = cd[0]; g.Hd = cd[1]; g.Wd = cd[2];
  g.kt = na[0]; g.kh = na[1]; g.kw = na[2];
  g.st = g.sh = g.sw = 1;
  g.pt = off[0]; g.ph = off[1]; g.pw = off[2];
  g.transposed = 1;
  g.pxs = 0;
  g.M = static_cast<long long>(g.N) * g.Td * g.Hd * g.Wd;
  g.numKb = na[0] * na[1] * na[2] * (g.Cs / 64);
  auto recip = [](int dd, unsigned long long& mul, int& sh) {
    int l = 0;
    while ((1ll << l) < dd) ++l;
    sh = 32 + l;
    mul = ((1ull << sh) + dd - 1) / dd;
  };
  recip(g.Wd, g.mulW, g.shW);
  recip(g.Hd, g.mulH, g.shH);
  recip(g.Td, g.mulT, g.shT);
  g.fast = 1;
  g.cls = 1;
  g.wa0 = a0[0]; g.was = s[0]; g.wb0 = a0[1]; g.wbs = s[1]; g.wc0 = a0[2]; g.wcs = s[2];
  g.wkh = d->kh; g.wkw = d->kw;
  g.oT = d->Ti; g.oH = d->Hi; g.oW = d->Wi;
  g.ost = s[0]; g.osh = s[1]; g.osw = s[2];
  g.oot = par[0]; g.ooh = par[1]; g.oow = par[2];
  g.src = static_cast<const __nv_bfloat16*>(dy);
  return g.M > 0;
}

int rsp_conv3d_dgrad(const rsp_conv3d_desc* d, const void* dy, const void* wd, void* dx, void* workspace,
                     void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ConvParams p{};
  RSP_REQUIRE(d->Ci % 64 == 0 && d->Co % 64 == 0, "conv3d dgrad: Ci=%d, Co=%d must be multiples of 64", d->Ci, d->Co);
  if ((d->st > 1 || d->sh > 1 || d->sw > 1) && d->kt <= 8 && d->kh <= 8 && d->kw <= 8) {
    const long long To = (d->Ti + 2 * d->pt - d->kt) / d->st + 1, Ho = (d->Hi + 2 * d->ph - d->kh) / d->sh + 1,
                    Wo = (d->Wi + 2 * d->pw - d->kw) / d->sw + 1;
    RSP_REQUIRE(To > 0 && Ho > 0 && Wo > 0, "conv3d dgrad: empty output");
    RSP_REQUIRE(static_cast<long long>(d->N) * To * Ho * Wo * d->Co < (1ll << 31) &&
                    static_cast<long long>(d->N) * d->Ti * d->Hi * d->Wi < (1ll << 31),
                "conv3d dgrad: tensor too large");
    if (d->kt < d->st || d->kh < d->sh || d->kw < d->sw) {  // some parity classes receive nothing
      cudaError_t e = rsp::zero_async(dx, static_cast<size_t>(d->N) * d->Ti * d->Hi * d->Wi * d->Ci * 2, stream);
      if (e != cudaSuccess) {
        set_error("dgrad memset: %s", cudaGetErrorString(e));
        return RSP_ERR_CUDA;
      }
    }
    // every parity class with at least one tap, kMaxClasses per launch
    ConvParamsMulti pm{};
    pm.base.wgt = static_cast<const __nv_bfloat16*>(wd);
    pm.base.out = static_cast<__nv_bfloat16*>(dx);
    pm.base.Nout = d->Ci;
    pm.base.wgtKb = d->kt * d->kh * d->kw * (d->Co / 64);  // filter rows carry the full K extent of the dgrad operand
    auto flush = [&]() -> int {
      if (pm.ncls == 0) return RSP_OK;
      pm.base.g = pm.gx[0];
      // short K loops (a class sees 1..8 of the 27 taps): two stages are enough and leave room for 4 CTAs per SM
      int rc = d->Ci % 128 == 0 ? launch_igemm_multi<128, 3>(pm, stream) : launch_igemm_multi<64, 2>(pm, stream);
      pm.ncls = 0;
      return rc;
    };
    // odd parities first: they see the most taps, so the longest CTAs start first
    for (int a = d->st - 1; a >= 0; --a)
      for (int b = d->sh - 1; b >= 0; --b)
        for (int c = d->sw - 1; c >= 0; --c) {
          const int par[3] = {a, b, c};
          if (!class_geom(d, par, dy, pm.gx[pm.ncls])) continue;
          const long long tiles = (pm.gx[pm.ncls].M + 127) / 128 + (pm.ncls ? pm.tileEnd[pm.ncls - 1] : 0);
          RSP_REQUIRE(tiles < (1ll << 31), "conv3d dgrad: too many tiles");
          pm.tileEnd[pm.ncls] = static_cast<int>(tiles);
          if (++pm.ncls == kMaxClasses) {
            int rc = flush();
            if (rc != RSP_OK) return rc;
          }
        }
    return flush();
  }
  if (direct_supported(d, 1)) return launch_direct(d, 1, dy, wd, nullptr, dx, nullptr, stream);
  int rc = fill_geom(p.g, d, MODE_GENERIC, 1);
  if (rc != RSP_OK) return rc;
  p.g.src = static_cast<const __nv_bfloat16*>(dy);
  p.wgt = static_cast<const __nv_bfloat16*>(wd);
  p.out = static_cast<__nv_bfloat16*>(dx);
  p.bias = nullptr;
  p.Nout = d->Ci;
  p.acc = static_cast<float*>(workspace);
  const int fullKb = skip_dead_frame_taps(p.g);
  return dispatch_igemm<MODE_GENERIC>(p, stream, fullKb);
}

// Timing / A-B switch (tools only, not in the public header): bit 0 = never take the direct wgrad kernel,
// bit 1 = take it whenever the geometry allows.
static int g_wgrad_debug = 0;
int rsp_debug_wgrad(int flags) {
  g_wgrad_debug = flags;
  return 0;
}

// The plane-run kernel pays a fixed 9 x 64 x 64 fp32 atomic epilogue per CTA and stages halo rows with every chunk:
// measured on R3D-18 at batch 64 it wins on layer1 (0.200 -> 0.094 ms), layer2 (0.077 -> 0.061) and layer3 (0.058 ->
// 0.053); tiny problems stay on the generic kernel.
static bool wgrad_direct_wanted(const rsp_conv3d_desc* d) {
  if (!wgrad_direct_supported(d)) return false;
  if (g_wgrad_debug & 2) return true;
  const long long pixels = static_cast<long long>(d->N) * d->Ti * d->Hi * d->Wi;
  return pixels >= 4096;
}

int rsp_conv3d_wgrad(const rsp_conv3d_desc* d, int Ci_logical, int Co_logical, const void* x, const void* dy,
                     float* dwt_workspace, float* dw, int accumulate, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int mode = conv_mode(d);
  WgradParams p{};
  int rc = fill_geom(p.g, d, mode, 0);
  if (rc != RSP_OK) return rc;
  RSP_REQUIRE(d->Co % 64 == 0, "conv3d wgrad: Co=%d must be a multiple of 64", d->Co);
  if (stem_wgrad_supported(d))
    return launch_stem_wgrad(d, Ci_logical, Co_logical, x, dy, dw, accumulate, device_sm_count(), stream);
  if (stem3_supported(d))
    return launch_stem3_wgrad(d, Ci_logical, Co_logical, x, dy, dwt_workspace, dw, accumulate, device_sm_count(), stream);
  p.g.src = static_cast<const __nv_bfloat16*>(x);
  p.dy = static_cast<const __nv_bfloat16*>(dy);
  p.dwt = dwt_workspace;
  p.Nout = d->Co;
  const int Kpad = p.g.numKb * 64;
  if (mode == MODE_GENERIC) skip_dead_frame_taps(p.g);   // their rows of dwt stay zero
  cudaError_t e = rsp::zero_async(dwt_workspace, static_cast<size_t>(Kpad) * d->Co * sizeof(float), stream);
  if (e != cudaSuccess) {
    set_error("wgrad memset: %s", cudaGetErrorString(e));
    return RSP_ERR_CUDA;
  }
  const int sms = device_sm_count();
  if (mode == MODE_GENERIC && !(g_wgrad_debug & 1) && wgrad_direct_wanted(d))
    rc = launch_wgrad_direct(d, x, dy, dwt_workspace, stream);
  else
    rc = mode == MODE_GENERIC ? dispatch_wgrad<MODE_GENERIC>(p, sms, stream) : dispatch_wgrad<MODE_SMALLC>(p, sms, stream);
  if (rc != RSP_OK) return rc;
  const int taps = d->kt * d->kh * d->kw;
  if (mode == MODE_GENERIC && taps <= 343) {
    int ciTile = taps == 27 ? 8 : 256 / taps;
    ciTile = ciTile < 1 ? 1 : (ciTile > 64 ? 64 : ciTile);
    dim3 grid(d->Co / 32, (Ci_logical + ciTile - 1) / ciTile);
    const size_t smem = static_cast<size_t>(32) * ((ciTile * taps) | 1) * sizeof(float);
    unpack_wgrad_tiled_kernel<<<grid, 256, smem, stream>>>(dwt_workspace, dw, Co_logical, d->Co, Ci_logical, d->Ci, taps,
                                                           ciTile, accumulate);
    return check_launch("unpack_wgrad_tiled");
  }
  size_t total = static_cast<size_t>(d->Co) * Kpad;
  unpack_wgrad_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      dwt_workspace, dw, mode, Co_logical, d->Co, Ci_logical, d->Ci, d->kt, d->kh, d->kw, p.g.pxs, Kpad, accumulate);
  return check_launch("unpack_wgrad");
}

}  // extern "C"
