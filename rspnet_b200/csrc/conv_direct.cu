// Direct (im2col-free) tcgen05 convolution for unit-stride filters on 64-channel-multiple tensors
// (reference: the 3x3x3 stride-1 nn.Conv3d of models/resnet.py:21-27 (BasicBlock), models/c3d.py:26-50, and their dgrad).
//
// Idea: flatten one padded (h, w) plane of the source with pitch Wp = Wi + 2*pw.  Output position q = ho*Wp + wo then
// reads source position q + b*Wp + c for filter tap (b, c): a pure OFFSET.  So a contiguous run of padded source pixels
// (one 128-byte row per pixel per 64-channel chunk, 128B-swizzled by address) sitting in shared memory is the K-major A
// operand of EVERY tap — the UMMA descriptor just starts b*Wp + c rows later (the swizzle is a function of the smem
// address, verified by tools/umma_shift_probe.cu).  A plane is therefore loaded once per (frame tap, channel chunk)
// instead of once per filter tap, and one filter tile feeds G=2..4 accumulators of 128 positions, which takes the kernel
// off the L2-bandwidth roofline that bounds the gather kernel for Co = 64.
//
// The plane run is one TMA box: the source is a 5-D tensor {C, W, H, T, N}; a box {64, Wp, rows, 1, 1} starting at
// (cc*64, -pw, row0 - ph, ts, n) lands as rows x Wp pixels of 128 B, 128B-swizzled, with the padding columns / rows
// zero-filled by the TMA unit (out-of-bounds coordinates) — exactly the padded-flattened layout above.  Filter tiles are
// 2-D TMA boxes {64, NT} of the packed filter.  No thread computes an address per element any more.
//
// CTA (persistent, 1/SM): warps 0-3 epilogue, warp 4 TMA producer (one elected lane), warp 5 MMA issuer.
// Rings: 2 plane buffers, 3-6 filter tiles, 2 TMEM accumulator sets (G * NT columns each).
#include <cstdlib>
#include "common.cuh"
#include "rspnet_b200.h"

namespace rsp {

int device_sm_count();

struct DirectParams {
  CUtensorMap tmapX;         // source as {Cs, Wi, Hi, Ti, N}, box {64, Wp, rowsBuf, 1, 1}
  CUtensorMap tmapW;         // filter as {taps*Cs, Nout}, box {64, NT}
  const __nv_bfloat16* x;    // source [N][Ti][Hi][Wi][Cs]
  const __nv_bfloat16* wgt;  // [Nout][taps*Cs]  (K index = tap*Cs + c)
  __nv_bfloat16* y;          // [N][To][Ho][Wo][Nout]
  const float* bias;
  float* stats;              // optional [2][Nout]
  int N, Ti, Hi, Wi, Cs;
  int To, Ho, Wo, Nout;
  int kt, kh, kw, pt, ph, pw;  // source = out - p + tap
  int flip;                    // 1: filter tap index is mirrored (dgrad)
  int Wp, Hp, P;               // padded pitch, padded rows, positions per plane (Ho*Wp)
  int G;                       // accumulators (128-position chunks) per work item
  int groups;                  // work items per plane
  int rowsBuf;                 // padded rows per plane buffer (TMA box height)
  int planeBytes;              // bytes reserved per plane buffer (multiple of 1024)
  int wStages;                 // filter-tile ring depth
  int numItems;                // ntiles * N * To * groups
  unsigned long long mulWp;    // reciprocal of Wp (shift 32 + shWp)
  int shWp;
};

// Diagnostic switches for timing experiments (tools/direct_probe.py); 0 in production.
//   1: the MMA lane does not wait for operands   2: the epilogue does not store   4: the producer loads nothing after
//   the first two planes / filter ring fill (barriers are still signalled)
__device__ int g_direct_debug = 0;

constexpr int kDirThreads = 192;
constexpr int kDirMaxWStages = 6;

__device__ __forceinline__ uint32_t fdiv64(uint32_t n, unsigned long long mul, int sh) {
  return static_cast<uint32_t>((static_cast<unsigned long long>(n) * mul) >> sh);
}

template <int NT>
__global__ void __launch_bounds__(kDirThreads, 1) conv_direct_kernel(const __grid_constant__ DirectParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int planeBytes = p.planeBytes;
  uint8_t* wring = smem + 2 * planeBytes;
  constexpr int W_BYTES = NT * 128;
  const int wStages = p.wStages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wring + wStages * W_BYTES);
  uint64_t* plane_full = bars;             // [2]
  uint64_t* plane_empty = bars + 2;        // [2]
  uint64_t* w_full = bars + 4;             // [wStages]
  uint64_t* w_empty = bars + 4 + kDirMaxWStages;
  uint64_t* acc_full = bars + 4 + 2 * kDirMaxWStages;   // [2]
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int t = threadIdx.x;
  const int warp = t >> 5;
  const int cch = p.Cs >> 6;
  const int accCols = p.G * NT;            // columns of one accumulator set
  const int dbg = g_direct_debug;

  if (t == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&plane_full[i], 1);
      mbar_init(&plane_empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 128);
    }
    for (int i = 0; i < wStages; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  // work item -> (group, to, ntile, n) with n fastest.  Items differ in cost (the last group of a plane has fewer position
  // chunks, the first / last frame skips a frame tap); a CTA takes every gridDim.x-th item, so the cost classes are the
  // slow indices: every CTA then gets its share of each class, and neighbouring CTAs still share the filter tile in L2.
  const int ntiles = p.Nout / NT;
  auto decode_item = [&](int it, int& nt, int& n, int& to, int& grp) {
    n = it % p.N;
    int q = it / p.N;
    nt = q % ntiles;
    q /= ntiles;
    to = q % p.To;
    grp = q / p.To;
  };

  if (warp == 4) {
    // ============================================================ TMA producer
    // one elected lane runs the whole loop: a single instruction stream without per-step warp re-convergence
    uint32_t pctr = 0, wctr = 0;
    uint32_t ws_ = 0, wph = 1;   // filter ring slot and the parity its "empty" barrier must have passed
    const uint32_t planeTx = static_cast<uint32_t>(p.rowsBuf) * p.Wp * 128u;
    const bool leader = elect_one();
    if (leader) {
      tma_prefetch_desc(&p.tmapX);
      tma_prefetch_desc(&p.tmapW);
    }
    for (int it = blockIdx.x; leader && it < p.numItems; it += gridDim.x) {
      int nt, n, to, grp;
      decode_item(it, nt, n, to, grp);
      const int q0 = grp * p.G * 128;
      const int row0 = static_cast<int>(fdiv64(static_cast<uint32_t>(q0), p.mulWp, p.shWp));  // first padded row
      for (int a = 0; a < p.kt; ++a) {
        const int ts = to - p.pt + a;
        if (ts < 0 || ts >= p.Ti) continue;  // temporal padding: the whole plane is zero, skip it (MMA warp agrees)
        for (int cc = 0; cc < cch; ++cc, ++pctr) {
          const int ps = pctr & 1;
          mbar_wait(&plane_empty[ps], ((pctr >> 1) & 1) ^ 1);
          if ((dbg & 4) && pctr >= 2) {
            mbar_arrive(&plane_full[ps]);
          } else {
            mbar_arrive_expect_tx(&plane_full[ps], planeTx);
            tma_load_5d(smem_u32(smem + ps * planeBytes), &p.tmapX, &plane_full[ps], cc * 64, -p.pw, row0 - p.ph, ts, n);
          }
          // the kh*kw filter tiles of this (frame tap, channel chunk)
          for (int b = 0; b < p.kh; ++b) {
            for (int c = 0; c < p.kw; ++c, ++wctr) {
              mbar_wait(&w_empty[ws_], wph);
              const int ta = p.flip ? p.kt - 1 - a : a, tb = p.flip ? p.kh - 1 - b : b, tc = p.flip ? p.kw - 1 - c : c;
              const int koff = ((ta * p.kh + tb) * p.kw + tc) * p.Cs + cc * 64;
              if ((dbg & 4) && wctr >= static_cast<uint32_t>(wStages)) {
                mbar_arrive(&w_full[ws_]);
              } else {
                mbar_arrive_expect_tx(&w_full[ws_], W_BYTES);
                tma_load_2d(smem_u32(wring + ws_ * W_BYTES), &p.tmapW, &w_full[ws_], koff, nt * NT);
              }
              if (++ws_ == static_cast<uint32_t>(wStages)) {
                ws_ = 0;
                wph ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp < 4) {
    // ============================================================ epilogue
    const int ew = warp, lane = t & 31;
    float ssum[NT / 32], ssq[NT / 32];
#pragma unroll
    for (int i = 0; i < NT / 32; ++i) ssum[i] = ssq[i] = 0.f;
    int cur_nt = -1;
    auto flush = [&]() {
      if (p.stats && cur_nt >= 0) {
#pragma unroll
        for (int i = 0; i < NT / 32; ++i) {
          atomicAdd(p.stats + cur_nt * NT + i * 32 + lane, ssum[i]);
          atomicAdd(p.stats + p.Nout + cur_nt * NT + i * 32 + lane, ssq[i]);
          ssum[i] = ssq[i] = 0.f;
        }
      }
    };
    uint32_t ictr = 0;
    for (int it = blockIdx.x; it < p.numItems; it += gridDim.x, ++ictr) {
      int nt, n, to, grp;
      decode_item(it, nt, n, to, grp);
      if (nt != cur_nt) {
        flush();
        cur_nt = nt;
      }
      const int buf = ictr & 1;
      const int q0 = grp * p.G * 128;
      int chunks = (p.P - q0 + 127) / 128;
      if (chunks > p.G) chunks = p.G;
      mbar_wait(&acc_full[buf], (ictr >> 1) & 1);
      tc_fence_after_sync();
      // columns outer, position chunks inner: BN statistics are reduced across lanes once per 32 columns and work item
#pragma unroll 1
      for (int c0 = 0; c0 < NT; c0 += 32) {
        float ra[32], qa[32];
#pragma unroll
        for (int jx = 0; jx < 32; ++jx) ra[jx] = qa[jx] = 0.f;
#pragma unroll 1
        for (int m = 0; m < chunks; ++m) {
          const uint32_t q = static_cast<uint32_t>(q0 + m * 128 + ew * 32 + lane);
          const uint32_t ho = fdiv64(q, p.mulWp, p.shWp);
          const int wo = static_cast<int>(q - ho * p.Wp);
          const bool ok = q < static_cast<uint32_t>(p.P) && wo < p.Wo;
          __nv_bfloat16* orow = p.y + ((((static_cast<size_t>(n) * p.To + to) * p.Ho + (ok ? ho : 0)) * p.Wo) +
                                       (ok ? wo : 0)) * p.Nout + nt * NT;
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + buf * accCols + m * NT + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int jx = 0; jx < 32; jx += 8) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              f[e] = __uint_as_float(v[jx + e]);
              if (p.bias) f[e] += __ldg(p.bias + nt * NT + c0 + jx + e);
            }
            uint4 o;
            o.x = pack_bf16x2(f[0], f[1]);
            o.y = pack_bf16x2(f[2], f[3]);
            o.z = pack_bf16x2(f[4], f[5]);
            o.w = pack_bf16x2(f[6], f[7]);
            if (ok && !(dbg & 2)) *reinterpret_cast<uint4*>(orow + c0 + jx) = o;
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
          // c0/32 is a runtime index into a tiny register array: resolve with a predicated unrolled loop
#pragma unroll
          for (int i = 0; i < NT / 32; ++i) {
            if (i == (c0 >> 5)) {
              ssum[i] += ra[0];
              ssq[i] += qa[0];
            }
          }
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&acc_empty[buf]);
    }
    flush();
  } else {
    // ============================================================ MMA issuer
    // one elected lane waits, issues and commits: the issue stream is ~2 instructions per MMA with no warp-level
    // re-convergence in between (measured with tools/umma_ts_probe: an elect/syncwarp per 4 MMAs costs ~25 clk per MMA)
    // The issue stream is kept to a few instructions per MMA: ring slots, parities and descriptors advance incrementally
    // (a single lane retires roughly one dependent instruction per 5-8 clocks, and an N=64 MMA lasts 48).
    constexpr uint32_t idesc = make_idesc_bf16(128, NT, 0, 0);
    constexpr int G = NT == 64 ? 4 : 2;                     // == p.G
    constexpr uint64_t kWStep = W_BYTES >> 4;
    uint32_t pctr = 0, ictr = 0;
    uint32_t ws_ = 0, wph = 0;
    const uint64_t bdesc_base = make_smem_desc_sw128(smem_u32(wring), 16, 1024);
    uint64_t bdesc = bdesc_base;
    const uint64_t rowStep = static_cast<uint64_t>((p.Wp - p.kw) * 8);   // descriptor units (16 B) from tap (b, kw) to (b+1, 0)
    const bool leader = elect_one();
    for (int it = blockIdx.x; leader && it < p.numItems; it += gridDim.x, ++ictr) {
      int nt, n, to, grp;
      decode_item(it, nt, n, to, grp);
      const int buf = ictr & 1;
      const int q0 = grp * G * 128;
      const int row0 = static_cast<int>(fdiv64(static_cast<uint32_t>(q0), p.mulWp, p.shWp));
      const int shift = q0 - row0 * p.Wp;   // the work item's first position inside the plane buffer
      int chunks = (p.P - q0 + 127) / 128;
      if (chunks > G) chunks = G;
      const uint32_t dacc = tmem_base + buf * accCols;
      mbar_wait(&acc_empty[buf], ((ictr >> 1) & 1) ^ 1);
      tc_fence_after_sync();
      uint32_t accflag = 0;
      for (int a = 0; a < p.kt; ++a) {
        const int ts = to - p.pt + a;
        if (ts < 0 || ts >= p.Ti) continue;
        for (int cc = 0; cc < cch; ++cc, ++pctr) {
          const int ps = pctr & 1;
          if (!(dbg & 1)) mbar_wait(&plane_full[ps], (pctr >> 1) & 1);
          uint64_t adesc0 = make_smem_desc_sw128(smem_u32(smem + ps * planeBytes) + shift * 128, 16, 1024);
          for (int b = 0; b < p.kh; ++b) {
            for (int c = 0; c < p.kw; ++c) {
              if (!(dbg & 1)) mbar_wait(&w_full[ws_], wph);
              tc_fence_after_sync();
              if (chunks == G) {
#pragma unroll
                for (int m = 0; m < G; ++m)
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_bf16(dacc + m * NT, adesc0 + static_cast<uint64_t>(m * 1024 + 2 * k), bdesc + 2 * k, idesc,
                              k == 0 ? accflag : 1u);
              } else {
                for (int m = 0; m < chunks; ++m)
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_bf16(dacc + m * NT, adesc0 + static_cast<uint64_t>(m * 1024 + 2 * k), bdesc + 2 * k, idesc,
                              k == 0 ? accflag : 1u);
              }
              umma_commit(&w_empty[ws_]);
              accflag = 1;
              adesc0 += 8;                                   // next tap along w: one pixel row = 128 B
              bdesc += kWStep;
              if (++ws_ == static_cast<uint32_t>(wStages)) {
                ws_ = 0;
                wph ^= 1;
                bdesc = bdesc_base;
              }
            }
            adesc0 += rowStep;
          }
          umma_commit(&plane_empty[ps]);
        }
      }
      umma_commit(&acc_full[buf]);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, 512);
}

// --------------------------------------------------------------------------------------------------------------------
static bool direct_geometry(const rsp_conv3d_desc* d, int transposed, DirectParams& p, int& NT, size_t& smem) {
  if (d->st != 1 || d->sh != 1 || d->sw != 1) return false;
  const int Cs = transposed ? d->Co : d->Ci, Nout = transposed ? d->Ci : d->Co;
  if (Cs % 64 != 0 || Nout % 64 != 0) return false;
  const int To = d->Ti + 2 * d->pt - d->kt + 1, Ho = d->Hi + 2 * d->ph - d->kh + 1, Wo = d->Wi + 2 * d->pw - d->kw + 1;
  if (To <= 0 || Ho <= 0 || Wo <= 0) return false;
  p.N = d->N;
  p.Cs = Cs;
  p.Nout = Nout;
  p.kt = d->kt; p.kh = d->kh; p.kw = d->kw;
  if (!transposed) {
    p.Ti = d->Ti; p.Hi = d->Hi; p.Wi = d->Wi;
    p.To = To; p.Ho = Ho; p.Wo = Wo;
    p.pt = d->pt; p.ph = d->ph; p.pw = d->pw;
    p.flip = 0;
  } else {  // dX[i] = sum_tap dY[i + p - tap] W[tap]: a unit-stride conv over dY with padding k-1-p and mirrored taps
    p.Ti = To; p.Hi = Ho; p.Wi = Wo;
    p.To = d->Ti; p.Ho = d->Hi; p.Wo = d->Wi;
    p.pt = d->kt - 1 - d->pt; p.ph = d->kh - 1 - d->ph; p.pw = d->kw - 1 - d->pw;
    p.flip = 1;
    if (p.pt < 0 || p.ph < 0 || p.pw < 0) return false;
  }
  p.Wp = p.Wi + 2 * p.pw;
  p.Hp = p.Hi + 2 * p.ph;
  if (p.Wo != p.Wp - p.kw + 1 || p.Ho != p.Hp - p.kh + 1) return false;
  p.P = p.Ho * p.Wp;
  NT = (Nout % 128 == 0) ? 128 : 64;
  p.G = NT == 64 ? 4 : 2;
  // small planes: the last 128-position chunk of a plane is partly empty; below 3/4 useful rows the gather kernel wastes less
  {
    const int chunksTotal = (p.P + 127) / 128;
    if (p.P < 128 || p.P * 4 < chunksTotal * 128 * 3) return false;
    // planes that do not fill whole groups (14 x 14: 224 of 256 rows): worth it only when the gather kernel's 128-row
    // tiles would not even fill three waves (R3D-18 layer2: 392 tiles) — measured both ways on R3D-18 / C3D shapes
    const long long gatherTiles = ((static_cast<long long>(p.N) * p.To * p.Ho * p.Wo + 127) / 128) * (Nout / NT);
    if (p.P < 128 * p.G && gatherTiles >= 6ll * device_sm_count()) return false;
  }
  if (p.kh * p.kw < 2) {
    // kt x 1 x 1 filters reuse nothing inside a plane, but the gather kernel pays for them twice: its im2col producer
    // issues one cp.async per 16 bytes, and with K = kt * Cs this short every 128-position tile is a CTA of its own with
    // its set-up cost.  The persistent TMA pipeline here has neither (R(2+1)D @ 16x56x56, batch 32: 3x1x1 fprop 128 -> 64
    // 0.385 -> 0.176 ms, 192 -> 64 0.43 -> 0.24 ms; dgrad 64 -> 192 1.00 -> 0.38 ms; S3D-G's unit-stride 1x1x1 branches: 26 -> 17 us at 28x28).
    static const int mode1x1 = [] {
      const char* e = getenv("RSP_DIRECT_1X1");   // 0: never, 1: temporal filters only, 2 (default): also unit-stride 1x1x1
      return e ? atoi(e) : 2;
    }();
    if (mode1x1 == 0 || (mode1x1 == 1 && p.kt < 2)) return false;
  }
  p.groups = (p.P + 128 * p.G - 1) / (128 * p.G);
  if (p.Wp > 256) return false;                                          // TMA box extent
  p.rowsBuf = (128 * p.G + p.kh * p.Wp + p.kw - 3) / p.Wp + 1;           // see the header: run + halo, row aligned
  if (p.rowsBuf > 256) return false;
  p.planeBytes = (p.rowsBuf * p.Wp * 128 + 1023) / 1024 * 1024;
  p.wStages = 0;
  for (int ws = kDirMaxWStages; ws >= 3; --ws) {
    smem = 2 * static_cast<size_t>(p.planeBytes) + static_cast<size_t>(ws) * NT * 128 + 1024 + 512;
    if (smem <= 227 * 1024) {
      p.wStages = ws;
      break;
    }
  }
  if (p.wStages == 0) return false;
  const long long items = static_cast<long long>(Nout / NT) * p.N * p.To * p.groups;
  if (items > (1ll << 30)) return false;
  p.numItems = static_cast<int>(items);
  int l = 0;
  while ((1 << l) < p.Wp) ++l;
  p.shWp = 32 + l;
  p.mulWp = ((1ull << p.shWp) + p.Wp - 1) / p.Wp;
  return true;
}

}  // namespace rsp

extern "C" int rsp_debug_direct(int flags) {   // timing experiments only (not part of the public header)
  return cudaMemcpyToSymbol(rsp::g_direct_debug, &flags, sizeof(int)) == cudaSuccess ? 0 : -2;
}

namespace rsp {

bool direct_supported(const rsp_conv3d_desc* d, int transposed) {
  DirectParams p{};
  int NT;
  size_t smem;
  return direct_geometry(d, transposed, p, NT, smem);
}

int launch_direct(const rsp_conv3d_desc* d, int transposed, const void* x, const void* wgt, const float* bias, void* y,
                  float* stats, cudaStream_t stream) {
  DirectParams p{};
  int NT;
  size_t smem;
  if (!direct_geometry(d, transposed, p, NT, smem)) {
    set_error("conv_direct: unsupported geometry");
    return RSP_ERR_INVALID;
  }
  p.x = static_cast<const __nv_bfloat16*>(x);
  p.wgt = static_cast<const __nv_bfloat16*>(wgt);
  p.y = static_cast<__nv_bfloat16*>(y);
  p.bias = bias;
  p.stats = stats;
  {
    const unsigned long long C = p.Cs, W = p.Wi, H = p.Hi, T = p.Ti, N = p.N;
    const unsigned long long dims[5] = {C, W, H, T, N};
    const unsigned long long strides[4] = {C * 2, W * C * 2, H * W * C * 2, T * H * W * C * 2};
    const unsigned box[5] = {64, static_cast<unsigned>(p.Wp), static_cast<unsigned>(p.rowsBuf), 1, 1};
    int rc = make_tmap_bf16(&p.tmapX, x, 5, dims, strides, box);
    if (rc != RSP_OK) return rc;
    const unsigned long long K = static_cast<unsigned long long>(p.kt) * p.kh * p.kw * p.Cs;
    const unsigned long long wdims[2] = {K, static_cast<unsigned long long>(p.Nout)};
    const unsigned long long wstrides[1] = {K * 2};
    const unsigned wbox[2] = {64, static_cast<unsigned>(NT)};
    rc = make_tmap_bf16(&p.tmapW, wgt, 2, wdims, wstrides, wbox);
    if (rc != RSP_OK) return rc;
  }
  const int sms = device_sm_count();
  const int grid = p.numItems < sms ? p.numItems : sms;
  cudaError_t e;
  if (NT == 64) {
    e = cudaFuncSetAttribute(conv_direct_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) conv_direct_kernel<64><<<grid, kDirThreads, smem, stream>>>(p);
  } else {
    e = cudaFuncSetAttribute(conv_direct_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) conv_direct_kernel<128><<<grid, kDirThreads, smem, stream>>>(p);
  }
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(conv_direct): %s", cudaGetErrorString(e));
    return RSP_ERR_CUDA;
  }
  return check_launch("conv_direct");
}

}  // namespace rsp
