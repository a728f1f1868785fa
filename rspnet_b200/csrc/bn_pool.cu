// Fused train-mode BatchNorm + ReLU + MaxPool3d on bf16 NDHWC tensors, forward.
//
// Reference call sites: models/resnet.py:203-206 (conv1 -> bn1 -> relu -> maxpool k3 s2 p1) and models/c3d.py:111-139
// (conv -> bn -> relu -> pool1..4).  Unfused, the post-activation tensor is written and read back once in the forward
// (bn_act_fwd + maxpool_fwd) and the pool gradient is materialised at full resolution in the backward (maxpool_bwd ->
// bn reduce -> bn apply).  Here:
//   forward : raw conv-output rows (bf16) --bulk copy--> smem; window scan on sign(scale)*x; affine + ReLU once per output
//             --> pooled y + uint8 argmax
//             --> pooled y (+ uint8 argmax and the raw value at the argmax in grad-enabled passes)
//   backward: bn_pool_bwd.cu (reductions from the pooled tensors, then one scatter + stream pass).
// HBM traffic per input element: forward 2 B read (+ pooled outputs).
#include "common.cuh"
#include "rspnet_b200.h"

namespace rsp {

int device_sm_count();
int bn_pool_bwd_row_bytes_ok(const rsp_pool3d_desc* d);

namespace {

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  v.x = pack_bf16x2(f[0], f[1]);
  v.y = pack_bf16x2(f[2], f[3]);
  v.z = pack_bf16x2(f[4], f[5]);
  v.w = pack_bf16x2(f[6], f[7]);
  return v;
}

struct FPGeom {
  int N, Ti, Hi, Wi, C, To, Ho, Wo;
  int kt, kh, kw, st, sh, sw, pt, ph, pw;
  int HB;       // forward: output rows per tile; backward: input rows per tile
  int bands;    // tiles per (n, frame)
  int rowsIn;   // forward: input rows staged per frame tap
  int numTiles;
};

__host__ __device__ __forceinline__ int floor_div(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }
__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return floor_div(a + b - 1, b); }

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// forward: one tile = HB output rows of one (n, to).  The kt x rowsIn raw input rows it needs are contiguous per frame in
// NDHWC, so one elected thread fetches them with bulk copies (UBLKCP) that complete on an mbarrier; 2-3 CTAs per SM keep
// the copy of one tile under the compute of another.  max/relu/affine commute per channel:
//   max_w relu(scale*x + shift) = relu(|scale| * max_w(sign(scale)*x) + shift)
// so the window scan works on sign-flipped raw bf16 pairs (HSETP2 / LOP3: ~5 instructions per pair and tap) and the
// affine + ReLU runs once per output.  KT/KH/KW > 0: compile-time window (fully unrolled scan); 0: runtime window.
// ---------------------------------------------------------------------------------------------------------------------
template <int KT, int KH, int KW>
__global__ void __launch_bounds__(256) bn_relu_maxpool_fwd_kernel(const uint4* __restrict__ x,
                                                                  const float* __restrict__ scale,
                                                                  const float* __restrict__ shift,
                                                                  uint4* __restrict__ y, uint2* __restrict__ idx,
                                                                  uint4* __restrict__ xmax, const FPGeom p) {
  extern __shared__ __align__(128) uint4 tile[];  // [kt][rowsIn][Wi][G] raw bf16
  __shared__ __align__(8) uint64_t bar;
  const int kt = KT ? KT : p.kt, kh = KH ? KH : p.kh, kw = KW ? KW : p.kw;
  const int G = p.C >> 3;
  const int g = threadIdx.x % G;   // 256 % G == 0: a thread keeps its channel group across strided loops
  float sc[8], sf[8];
  uint32_t flip[4];                // sign-bit masks per bf16 pair: compare sign(scale)*x
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sc[e] = scale[g * 8 + e];
    sf[e] = shift[g * 8 + e];
  }
#pragma unroll
  for (int e = 0; e < 4; ++e)
    flip[e] = (sc[2 * e] < 0.f ? 0x00008000u : 0u) | (sc[2 * e + 1] < 0.f ? 0x80000000u : 0u);
  const int rowVecs = p.Wi * G;
  const int frameVecs = p.rowsIn * rowVecs;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  uint32_t phase = 0;
  for (int tIdx = blockIdx.x; tIdx < p.numTiles; tIdx += gridDim.x, phase ^= 1) {
    const int band = tIdx % p.bands;
    const int q = tIdx / p.bands;
    const int to = q % p.To, n = q / p.To;
    const int ho0 = band * p.HB;
    const int hi0 = ho0 * p.sh - p.ph, ti0 = to * p.st - p.pt;
    const int h_lo = max(hi0, 0), h_hi = min(hi0 + p.rowsIn, p.Hi);   // valid input rows [h_lo, h_hi)
    __syncthreads();  // previous tile fully consumed
    if (threadIdx.x == 0) {
      fence_proxy_async_smem();
      uint32_t bytes = 0;
      const uint32_t chunk = static_cast<uint32_t>(h_hi - h_lo) * rowVecs * 16u;
      for (int a = 0; a < kt; ++a) {
        const int ti = ti0 + a;
        if (ti >= 0 && ti < p.Ti) bytes += chunk;
      }
      mbar_arrive_expect_tx(&bar, bytes);
      for (int a = 0; a < kt; ++a) {
        const int ti = ti0 + a;
        if (ti < 0 || ti >= p.Ti) continue;
        const uint4* src = x + ((static_cast<size_t>(n) * p.Ti + ti) * p.Hi + h_lo) * rowVecs;
        bulk_g2s(tile + a * frameVecs + (h_lo - hi0) * rowVecs, src, chunk, &bar);
      }
    }
    mbar_wait(&bar, phase);
    const int hbEff = min(p.HB, p.Ho - ho0);
    const int items = hbEff * p.Wo * G;
    for (int it = threadIdx.x; it < items; it += 256) {
      const int pix = it / G;
      const int hb = pix / p.Wo, wo = pix - hb * p.Wo;
      uint32_t best[4], bi[4];     // packed bf16 pairs / packed 16-bit tap indices
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        best[e] = 0xff80ff80u;     // (-inf, -inf)
        bi[e] = 0;
      }
      const int w0 = wo * p.sw - p.pw;
#pragma unroll(KT ? KT : 1)
      for (int a = 0; a < (KT ? KT : 8); ++a) {
        if (a >= kt) break;
        const int ti = ti0 + a;
        if (ti < 0 || ti >= p.Ti) continue;
#pragma unroll(KH ? KH : 1)
        for (int b = 0; b < (KH ? KH : 8); ++b) {
          if (b >= kh) break;
          const int r = hb * p.sh + b;
          const int hi = hi0 + r;
          if (hi < 0 || hi >= p.Hi) continue;
          const uint4* row = tile + a * frameVecs + r * rowVecs + w0 * G + g;
#pragma unroll(KW ? KW : 1)
          for (int c = 0; c < (KW ? KW : 8); ++c) {
            if (c >= kw) break;
            if (w0 + c < 0 || w0 + c >= p.Wi) continue;
            uint4 raw = row[c * G];
            const uint32_t lin = static_cast<uint32_t>((a * kh + b) * kw + c) * 0x00010001u;
            const uint32_t v[4] = {raw.x ^ flip[0], raw.y ^ flip[1], raw.z ^ flip[2], raw.w ^ flip[3]};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              // strict >: the first maximum in (kt, kh, kw) order wins, as in nn.MaxPool3d
              const uint32_t m = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&v[e]),
                                             *reinterpret_cast<const __nv_bfloat162*>(&best[e]));
              best[e] = (v[e] & m) | (best[e] & ~m);
              bi[e] = (lin & m) | (bi[e] & ~m);
            }
          }
        }
      }
      float f[8];
      uint4 bv;
      bv.x = best[0]; bv.y = best[1]; bv.z = best[2]; bv.w = best[3];
      unpack8(bv, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = fmaxf(fmaf(f[e], fabsf(sc[e]), sf[e]), 0.f);  // |s| * (sign*x) = s*x
      const size_t o = (((static_cast<size_t>(n) * p.To + to) * p.Ho + ho0 + hb) * p.Wo + wo) * G + g;
      y[o] = pack8(f);
      if (idx) {          // grad-enabled pass: the backward needs the argmax and the raw value it selected
        uint2 iv;
        iv.x = __byte_perm(bi[0], bi[1], 0x6420);
        iv.y = __byte_perm(bi[2], bi[3], 0x6420);
        idx[o] = iv;
        uint4 xm;
        xm.x = best[0] ^ flip[0]; xm.y = best[1] ^ flip[1]; xm.z = best[2] ^ flip[2]; xm.w = best[3] ^ flip[3];
        xmax[o] = xm;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// forward, 3x3x3 / stride 2 / padding 1 window (the R3D-18 stem pool), C = 64: frame sweep.
// The tiled kernel above stages the 3 frames x 3 rows of ONE output row per tile: every input element crosses the L2 -> SM
// path 2.25 times and the kernel sits at 0.35 of the HBM roofline.  A max over a 3x3x3 window is separable in time:
//   out[to] = max(sp[2to-1], sp[2to], sp[2to+1]),   sp[ti] = 3x3 spatial window max of input frame ti,
// and sp[2to+1] of one output frame is sp[2(to+1)-1] of the next.  A CTA therefore walks a CONTIGUOUS range of steps
// (n, band of 2 output rows, to) with `to` fastest: per step two new input frames (5 rows each) arrive in a 3-slot ring by
// bulk copy, each thread reduces its two (row, wo, 8-channel) items spatially and carries the odd frame's result (value +
// tap index) in registers to the next step.  Every input row is read 1.25 times (band halo), once per frame.
// Tie-breaking is unchanged: strict > in (frame, row, column) order = the first maximum wins, as in nn.MaxPool3d.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kSweepCompute = 224;            // 28 output columns x 8 channel octets
constexpr int kSweepThreads = kSweepCompute + 32;
constexpr int kSweepSlots = 3;
constexpr int kSweepHB = 2;
constexpr int kSweepRows = 2 * kSweepHB + 1;

struct SweepGeom {
  int N, Ti, Hi, Wi, To, Ho, Wo;
  int bands;
  long long numSteps;
};

// frames a step consumes, in order: [2to-1 when the carry is not valid], 2to, [2to+1 when inside the clip]
__device__ __forceinline__ int sweep_first_frame(int to, bool fresh) { return (fresh && to > 0) ? 2 * to - 1 : 2 * to; }

template <bool GRAD>
__global__ void __launch_bounds__(kSweepThreads, 2) bn_relu_maxpool_sweep_kernel(const uint4* __restrict__ x,
                                                                                 const float* __restrict__ scale,
                                                                                 const float* __restrict__ shift,
                                                                                 uint4* __restrict__ y, uint2* __restrict__ idx,
                                                                                 uint4* __restrict__ xmax, const SweepGeom p) {
  extern __shared__ __align__(128) uint4 ring[];   // [kSweepSlots][kSweepRows][Wi][8] raw bf16 octets
  __shared__ __align__(8) uint64_t full_bar[kSweepSlots], empty_bar[kSweepSlots];
  constexpr int G = 8;
  const int rowVecs = p.Wi * G;
  const int slotVecs = kSweepRows * rowVecs;
  const long long s0 = p.numSteps * blockIdx.x / gridDim.x, s1 = p.numSteps * (blockIdx.x + 1) / gridDim.x;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kSweepSlots; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kSweepCompute / 32);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (threadIdx.x >= kSweepCompute) {
    // ------------------------------------------------------------------ producer (one lane)
    if (threadIdx.x == kSweepCompute) {
      int slot = 0;
      uint32_t ph = 0;
      for (long long s = s0; s < s1; ++s) {
        const int to = static_cast<int>(s % p.To);
        const long long q = s / p.To;
        const int band = static_cast<int>(q % p.bands), n = static_cast<int>(q / p.bands);
        const int hi0 = band * kSweepHB * 2 - 1;
        const int h_lo = max(hi0, 0), h_hi = min(hi0 + kSweepRows, p.Hi);
        const uint32_t bytes = static_cast<uint32_t>(h_hi - h_lo) * rowVecs * 16u;
        const bool fresh = s == s0 || to == 0;
        for (int ti = sweep_first_frame(to, fresh); ti <= 2 * to + 1 && ti < p.Ti; ++ti) {
          mbar_wait(&empty_bar[slot], ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[slot], bytes);
          bulk_g2s(ring + slot * slotVecs + (h_lo - hi0) * rowVecs,
                   x + ((static_cast<size_t>(n) * p.Ti + ti) * p.Hi + h_lo) * rowVecs, bytes, &full_bar[slot]);
          if (++slot == kSweepSlots) {
            slot = 0;
            ph ^= 1;
          }
        }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumers: thread = (wo, channel octet), 2 rows
  const int g = threadIdx.x & 7, wo = threadIdx.x >> 3;
  const bool active = wo < p.Wo;
  float sc[8], sf[8];
  uint32_t flip[4];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sc[e] = scale[g * 8 + e];
    sf[e] = shift[g * 8 + e];
  }
#pragma unroll
  for (int e = 0; e < 4; ++e)
    flip[e] = (sc[2 * e] < 0.f ? 0x00008000u : 0u) | (sc[2 * e + 1] < 0.f ? 0x80000000u : 0u);
  const int w0 = 2 * wo - 1;
  uint32_t cval[kSweepHB][4], cidx[kSweepHB][4];   // carried spatial maximum of the last odd frame (+ its 3b+c index)
  int slot = 0;
  uint32_t ph = 0;

  // spatial 3x3 window maximum of the frame in `slot` for output row j of the band
  auto spatial = [&](const uint4* fr, int hi0, int j, uint32_t (&bv)[4], uint32_t (&bi)[4]) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      bv[e] = 0xff80ff80u;   // (-inf, -inf)
      bi[e] = 0;
    }
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const int r = 2 * j + b;
      const int hi = hi0 + r;
      if (hi < 0 || hi >= p.Hi) continue;
      const uint4* row = fr + r * rowVecs + w0 * G + g;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (w0 + c < 0 || w0 + c >= p.Wi) continue;
        const uint4 raw = row[c * G];
        const uint32_t lin = static_cast<uint32_t>(b * 3 + c) * 0x00010001u;
        const uint32_t v[4] = {raw.x ^ flip[0], raw.y ^ flip[1], raw.z ^ flip[2], raw.w ^ flip[3]};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t m = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&v[e]),
                                         *reinterpret_cast<const __nv_bfloat162*>(&bv[e]));
          bv[e] = (v[e] & m) | (bv[e] & ~m);
          if (GRAD) bi[e] = (lin & m) | (bi[e] & ~m);
        }
      }
    }
  };
  // wait for the next frame of the ring / hand its slot back
  auto acquire = [&]() {
    mbar_wait(&full_bar[slot], ph);
    return ring + slot * slotVecs;
  };
  auto release = [&]() {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&empty_bar[slot]);
    if (++slot == kSweepSlots) {
      slot = 0;
      ph ^= 1;
    }
  };

  for (long long s = s0; s < s1; ++s) {
    const int to = static_cast<int>(s % p.To);
    const long long q = s / p.To;
    const int band = static_cast<int>(q % p.bands), n = static_cast<int>(q / p.bands);
    const int ho0 = band * kSweepHB;
    const int hi0 = ho0 * 2 - 1;
    const bool fresh = s == s0 || to == 0;
    uint32_t best[kSweepHB][4], bidx[kSweepHB][4];
    if (fresh) {
      if (to > 0) {
        const uint4* fr = acquire();
        if (active) {
#pragma unroll
          for (int j = 0; j < kSweepHB; ++j) spatial(fr, hi0, j, cval[j], cidx[j]);
        }
        release();
      } else {
#pragma unroll
        for (int j = 0; j < kSweepHB; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            cval[j][e] = 0xff80ff80u;
            cidx[j][e] = 0;
          }
      }
    }
#pragma unroll
    for (int j = 0; j < kSweepHB; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        best[j][e] = cval[j][e];
        bidx[j][e] = cidx[j][e];      // frame tap 0: index 3b + c
      }
#pragma unroll
    for (int a = 1; a < 3; ++a) {
      const int ti = 2 * to - 1 + a;
      if (ti >= p.Ti) break;
      const uint4* fr = acquire();
      if (active) {
#pragma unroll
        for (int j = 0; j < kSweepHB; ++j) {
          uint32_t sv[4], si[4];
          spatial(fr, hi0, j, sv, si);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t m = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&sv[e]),
                                           *reinterpret_cast<const __nv_bfloat162*>(&best[j][e]));
            best[j][e] = (sv[e] & m) | (best[j][e] & ~m);
            if (GRAD) bidx[j][e] = ((si[e] + static_cast<uint32_t>(9 * a) * 0x00010001u) & m) | (bidx[j][e] & ~m);
            if (a == 2) {
              cval[j][e] = sv[e];
              cidx[j][e] = si[e];
            }
          }
        }
      }
      release();
    }
    if (active) {
#pragma unroll
      for (int j = 0; j < kSweepHB; ++j) {
        if (ho0 + j >= p.Ho) break;
        float f[8];
        uint4 bv;
        bv.x = best[j][0]; bv.y = best[j][1]; bv.z = best[j][2]; bv.w = best[j][3];
        unpack8(bv, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = fmaxf(fmaf(f[e], fabsf(sc[e]), sf[e]), 0.f);  // |s| * (sign*x) = s*x
        const size_t o = (((static_cast<size_t>(n) * p.To + to) * p.Ho + ho0 + j) * p.Wo + wo) * G + g;
        y[o] = pack8(f);
        if (GRAD) {
          uint2 iv;
          iv.x = __byte_perm(bidx[j][0], bidx[j][1], 0x6420);
          iv.y = __byte_perm(bidx[j][2], bidx[j][3], 0x6420);
          idx[o] = iv;
          uint4 xm;
          xm.x = best[j][0] ^ flip[0]; xm.y = best[j][1] ^ flip[1]; xm.z = best[j][2] ^ flip[2]; xm.w = best[j][3] ^ flip[3];
          xmax[o] = xm;
        }
      }
    }
  }
}

constexpr int kFusedSmemBudget = 72 * 1024;

int fill_fp(FPGeom& g, const rsp_pool3d_desc* d) {
  g.N = d->N; g.Ti = d->Ti; g.Hi = d->Hi; g.Wi = d->Wi; g.C = d->C;
  g.kt = d->kt; g.kh = d->kh; g.kw = d->kw;
  g.st = d->st; g.sh = d->sh; g.sw = d->sw;
  g.pt = d->pt; g.ph = d->ph; g.pw = d->pw;
  g.To = (d->Ti + 2 * d->pt - d->kt) / d->st + 1;
  g.Ho = (d->Hi + 2 * d->ph - d->kh) / d->sh + 1;
  g.Wo = (d->Wi + 2 * d->pw - d->kw) / d->sw + 1;
  RSP_REQUIRE(g.To > 0 && g.Ho > 0 && g.Wo > 0, "bn_relu_maxpool: empty output");
  RSP_REQUIRE(d->C % 8 == 0 && 256 % (d->C / 8) == 0, "bn_relu_maxpool: C/8 = %d must divide 256", d->C / 8);
  RSP_REQUIRE(d->kt * d->kh * d->kw <= 255, "bn_relu_maxpool: window too large for uint8 indices");
  return RSP_OK;
}

}  // namespace

}  // namespace rsp

using namespace rsp;

static int g_pool_debug = 0;   // bit 0: never take the frame-sweep forward kernel (A/B timing, tests)

extern "C" {

int rsp_debug_pool(int flags) {   // tools / tests only, not part of the public header
  g_pool_debug = flags;
  return 0;
}

int rsp_bn_relu_maxpool_supported(const rsp_pool3d_desc* d) {
  // C % 16: the argmax rows move with 16-byte-granular bulk copies
  if (d->C % 16 != 0 || 256 % (d->C / 8) != 0 || d->kt > 8 || d->kh > 8 || d->kw > 8) return 0;
  const size_t row = static_cast<size_t>(d->Wi) * d->C * 2;
  if (static_cast<size_t>(d->kt) * d->kh * row > static_cast<size_t>(kFusedSmemBudget)) return 0;
  FPGeom g;
  if (fill_fp(g, d) != RSP_OK) return 0;
  return bn_pool_bwd_row_bytes_ok(d);   // the backward keeps one fp32 row tile in shared memory (bn_pool_bwd.cu)
}

int rsp_bn_relu_maxpool_fwd(const rsp_pool3d_desc* d, const void* x, const float* scale, const float* shift, void* y,
                            uint8_t* idx, void* xmax, void* stream) {
  FPGeom g;
  int rc = fill_fp(g, d);
  if (rc != RSP_OK) return rc;
  RSP_REQUIRE(rsp_bn_relu_maxpool_supported(d), "bn_relu_maxpool_fwd: one window row set does not fit in shared memory");
  RSP_REQUIRE((idx == nullptr) == (xmax == nullptr), "bn_relu_maxpool_fwd: idx and xmax are written together or not at all");
  const size_t row = static_cast<size_t>(g.Wi) * g.C * 2;
  if (g.kt == 3 && g.kh == 3 && g.kw == 3 && g.st == 2 && g.sh == 2 && g.sw == 2 && g.pt == 1 && g.ph == 1 && g.pw == 1 &&
      g.C == 64 && g.Wo * 8 <= kSweepCompute && g.Wi >= 2 && !(g_pool_debug & 1)) {
    SweepGeom sg{};
    sg.N = g.N; sg.Ti = g.Ti; sg.Hi = g.Hi; sg.Wi = g.Wi; sg.To = g.To; sg.Ho = g.Ho; sg.Wo = g.Wo;
    sg.bands = (g.Ho + kSweepHB - 1) / kSweepHB;
    sg.numSteps = static_cast<long long>(g.N) * sg.bands * g.To;
    if (sg.numSteps == 0) return RSP_OK;
    const int smem = static_cast<int>(kSweepSlots * kSweepRows * row);
    if (smem <= 112 * 1024) {
      long long grid = 2ll * device_sm_count();
      if (grid > sg.numSteps) grid = sg.numSteps;
      cudaStream_t s = static_cast<cudaStream_t>(stream);
      cudaError_t e;
      if (idx) {
        e = cudaFuncSetAttribute(bn_relu_maxpool_sweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
        if (e == cudaSuccess)
          bn_relu_maxpool_sweep_kernel<true><<<static_cast<unsigned>(grid), kSweepThreads, smem, s>>>(
              static_cast<const uint4*>(x), scale, shift, static_cast<uint4*>(y), reinterpret_cast<uint2*>(idx),
              static_cast<uint4*>(xmax), sg);
      } else {
        e = cudaFuncSetAttribute(bn_relu_maxpool_sweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
        if (e == cudaSuccess)
          bn_relu_maxpool_sweep_kernel<false><<<static_cast<unsigned>(grid), kSweepThreads, smem, s>>>(
              static_cast<const uint4*>(x), scale, shift, static_cast<uint4*>(y), nullptr, nullptr, sg);
      }
      if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(bn_relu_maxpool_sweep): %s", cudaGetErrorString(e));
        return RSP_ERR_CUDA;
      }
      return check_launch("bn_relu_maxpool_sweep");
    }
  }
  // as many output rows per tile as the smem budget allows (fewer re-reads of rows shared by neighbouring windows)
  int hb = 1;
  while (hb < g.Ho && hb < 8 &&
         static_cast<size_t>(g.kt) * (hb * g.sh + g.kh) * row <= static_cast<size_t>(kFusedSmemBudget))
    ++hb;
  g.HB = hb;
  g.rowsIn = (hb - 1) * g.sh + g.kh;
  g.bands = (g.Ho + hb - 1) / hb;
  const long long tiles = static_cast<long long>(g.N) * g.To * g.bands;
  if (tiles == 0) return RSP_OK;
  RSP_REQUIRE(tiles < (1ll << 31), "bn_relu_maxpool_fwd: too many tiles");
  g.numTiles = static_cast<int>(tiles);
  const int smem = static_cast<int>(g.kt * g.rowsIn * row);
  const int per_sm = smem > 0 ? (200 * 1024) / (smem + 1024) : 8;
  long long grid = static_cast<long long>(device_sm_count()) * (per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm));
  if (grid > tiles) grid = tiles;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaSuccess;
#define RSP_LAUNCH_POOL_FWD(KT, KH, KW)                                                                                \
  do {                                                                                                                 \
    e = cudaFuncSetAttribute(bn_relu_maxpool_fwd_kernel<KT, KH, KW>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                             smem);                                                                                    \
    if (e == cudaSuccess)                                                                                              \
      bn_relu_maxpool_fwd_kernel<KT, KH, KW><<<static_cast<unsigned>(grid), 256, smem, s>>>(                           \
          static_cast<const uint4*>(x), scale, shift, static_cast<uint4*>(y), reinterpret_cast<uint2*>(idx),          \
          static_cast<uint4*>(xmax), g);                                                                               \
  } while (0)
  if (g.kt == 3 && g.kh == 3 && g.kw == 3) RSP_LAUNCH_POOL_FWD(3, 3, 3);
  else if (g.kt == 2 && g.kh == 2 && g.kw == 2) RSP_LAUNCH_POOL_FWD(2, 2, 2);
  else if (g.kt == 1 && g.kh == 2 && g.kw == 2) RSP_LAUNCH_POOL_FWD(1, 2, 2);
  else {
    RSP_REQUIRE(g.kt <= 8 && g.kh <= 8 && g.kw <= 8, "bn_relu_maxpool_fwd: window extents above 8 are not supported");
    RSP_LAUNCH_POOL_FWD(0, 0, 0);
  }
#undef RSP_LAUNCH_POOL_FWD
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(bn_relu_maxpool_fwd): %s", cudaGetErrorString(e));
    return RSP_ERR_CUDA;
  }
  return check_launch("bn_relu_maxpool_fwd");
}

}  // extern "C"
