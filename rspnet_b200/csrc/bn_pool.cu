// Fused train-mode BatchNorm + ReLU + MaxPool3d on bf16 NDHWC tensors, forward and backward.
//
// Reference call sites: models/resnet.py:203-206 (conv1 -> bn1 -> relu -> maxpool k3 s2 p1) and models/c3d.py:111-139
// (conv -> bn -> relu -> pool1..4).  Unfused, the post-activation tensor is written and read back once in the forward
// (bn_act_fwd + maxpool_fwd) and the pool gradient is materialised at full resolution in the backward (maxpool_bwd ->
// bn reduce -> bn apply).  Here:
//   forward : conv output x (bf16) --(scale, shift, relu, round to bf16)--> smem window rows --> pooled y + uint8 argmax
//   backward: dz(input position) = [bn(x) > 0] * sum of dy over the windows whose argmax is this position, rebuilt on the
//             fly from (dy, argmax) staged in smem; one pass reduces (sum dz, sum dz*xhat), one pass writes dx.
// HBM traffic per input element: forward 2 B read (+ pooled output), backward 2 x 2 B read + 2 B write.
#include "common.cuh"
#include "rspnet_b200.h"

namespace rsp {

int device_sm_count();

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

// ---------------------------------------------------------------------------------------------------------------------
// forward: one tile = HB output rows of one (n, to); smem holds the kt x rowsIn activated input rows it needs
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_relu_maxpool_fwd_kernel(const uint4* __restrict__ x,
                                                                  const float* __restrict__ scale,
                                                                  const float* __restrict__ shift,
                                                                  uint4* __restrict__ y, uint2* __restrict__ idx,
                                                                  const FPGeom p) {
  extern __shared__ uint4 tile[];  // [kt][rowsIn][Wi][G]
  const int G = p.C >> 3;
  const int g = threadIdx.x % G;   // 256 % G == 0: a thread keeps its channel group across strided loops
  float sc[8], sf[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sc[e] = scale[g * 8 + e];
    sf[e] = shift[g * 8 + e];
  }
  const int rowVecs = p.Wi * G;
  const int frameVecs = p.rowsIn * rowVecs;
  for (int tIdx = blockIdx.x; tIdx < p.numTiles; tIdx += gridDim.x) {
    const int band = tIdx % p.bands;
    const int q = tIdx / p.bands;
    const int to = q % p.To, n = q / p.To;
    const int ho0 = band * p.HB;
    const int hi0 = ho0 * p.sh - p.ph, ti0 = to * p.st - p.pt;
    __syncthreads();  // previous tile fully consumed
    for (int a = 0; a < p.kt; ++a) {
      const int ti = ti0 + a;
      if (ti < 0 || ti >= p.Ti) continue;
      const uint4* frame = x + (static_cast<size_t>(n) * p.Ti + ti) * p.Hi * rowVecs;
      for (int v = threadIdx.x; v < frameVecs; v += 256) {
        const int r = v / rowVecs;
        const int hi = hi0 + r;
        if (hi < 0 || hi >= p.Hi) continue;
        float f[8];
        unpack8(__ldg(frame + static_cast<size_t>(hi) * rowVecs + (v - r * rowVecs)), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = fmaxf(fmaf(f[e], sc[e], sf[e]), 0.f);
        tile[a * frameVecs + v] = pack8(f);
      }
    }
    __syncthreads();
    const int hbEff = min(p.HB, p.Ho - ho0);
    const int items = hbEff * p.Wo * G;
    for (int it = threadIdx.x; it < items; it += 256) {
      const int pix = it / G;
      const int hb = pix / p.Wo, wo = pix - hb * p.Wo;
      float best[8];
      unsigned bi[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        best[e] = -INFINITY;
        bi[e] = 0;
      }
      bool any = false;
      for (int a = 0; a < p.kt; ++a) {
        const int ti = ti0 + a;
        if (ti < 0 || ti >= p.Ti) continue;
        for (int b = 0; b < p.kh; ++b) {
          const int r = hb * p.sh + b;
          const int hi = hi0 + r;
          if (hi < 0 || hi >= p.Hi) continue;
          const uint4* row = tile + a * frameVecs + r * rowVecs + g;
          for (int c = 0; c < p.kw; ++c) {
            const int wi = wo * p.sw - p.pw + c;
            if (wi < 0 || wi >= p.Wi) continue;
            float v[8];
            unpack8(row[wi * G], v);
            const unsigned lin = (a * p.kh + b) * p.kw + c;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              if (!any || v[e] > best[e]) {
                best[e] = v[e];
                bi[e] = lin;
              }
            }
            any = true;
          }
        }
      }
      const size_t o = (((static_cast<size_t>(n) * p.To + to) * p.Ho + ho0 + hb) * p.Wo + wo) * G + g;
      y[o] = pack8(best);
      uint2 iv;
      iv.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
      iv.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
      idx[o] = iv;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward: one tile = HB input rows of one (n, ti); smem holds dy / argmax of every window that can select them
// MODE 0: per-channel sums (sum dz, sum dz*xhat) -> atomics.  MODE 1: dx = gamma*invstd*(dz - s1/M - xhat*s2/M).
// ---------------------------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) bn_relu_maxpool_bwd_kernel(
    const uint4* __restrict__ dy, const uint2* __restrict__ idx, const uint4* __restrict__ x,
    const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
    const float* __restrict__ invstd, const float* __restrict__ gamma, float* __restrict__ sum_dz,
    float* __restrict__ sum_dz_xhat, uint4* __restrict__ dx, const FPGeom p, int Cl, float inv_m, int maxWin) {
  extern __shared__ uint4 stage[];                       // dy vectors, then argmax vectors
  uint2* sidx = reinterpret_cast<uint2*>(stage + maxWin);
  __shared__ float red[MODE == 0 ? 2 * 256 * 8 : 1];
  const int G = p.C >> 3;
  const int g = threadIdx.x % G;
  float sc[8], sf[8], mu[8], is[8], k[8], s1[8], s2[8], a0[8], a1[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = g * 8 + e;
    sc[e] = scale[c];
    sf[e] = shift[c];
    mu[e] = mean[c];
    is[e] = invstd[c];
    a0[e] = a1[e] = 0.f;
    if (MODE == 1) {
      k[e] = c < Cl ? gamma[c] * is[e] : 0.f;
      s1[e] = sum_dz[c] * inv_m;
      s2[e] = sum_dz_xhat[c] * inv_m;
    }
  }
  const int rowVecs = p.Wi * G, orowVecs = p.Wo * G;
  for (int tIdx = blockIdx.x; tIdx < p.numTiles; tIdx += gridDim.x) {
    const int band = tIdx % p.bands;
    const int q = tIdx / p.bands;
    const int ti = q % p.Ti, n = q / p.Ti;
    const int hi0 = band * p.HB;
    const int hbEff = min(p.HB, p.Hi - hi0);
    const int to_lo = max(0, ceil_div(ti + p.pt - p.kt + 1, p.st)), to_hi = min(p.To - 1, floor_div(ti + p.pt, p.st));
    const int ho_lo = max(0, ceil_div(hi0 + p.ph - p.kh + 1, p.sh));
    const int ho_hi = min(p.Ho - 1, floor_div(hi0 + hbEff - 1 + p.ph, p.sh));
    const int nto = max(0, to_hi - to_lo + 1), nho = max(0, ho_hi - ho_lo + 1);
    __syncthreads();
    const int win = nto * nho * orowVecs;
    for (int v = threadIdx.x; v < win; v += 256) {
      const int r = v / orowVecs;                 // (to - to_lo) * nho + (ho - ho_lo)
      const int tt = r / nho, hh = r - tt * nho;
      const size_t o = ((static_cast<size_t>(n) * p.To + to_lo + tt) * p.Ho + ho_lo + hh) * orowVecs + (v - r * orowVecs);
      stage[v] = __ldg(dy + o);
      sidx[v] = __ldg(idx + o);
    }
    __syncthreads();
    const int items = hbEff * rowVecs;
    for (int it = threadIdx.x; it < items; it += 256) {
      const int pix = it / G;
      const int hb = pix / p.Wi, wi = pix - hb * p.Wi;
      const int hi = hi0 + hb;
      float acc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.f;
      const int h_lo = max(ho_lo, ceil_div(hi + p.ph - p.kh + 1, p.sh)), h_hi = min(ho_hi, floor_div(hi + p.ph, p.sh));
      const int w_lo = max(0, ceil_div(wi + p.pw - p.kw + 1, p.sw)), w_hi = min(p.Wo - 1, floor_div(wi + p.pw, p.sw));
      for (int to = to_lo; to <= to_hi; ++to) {
        const int a = ti + p.pt - to * p.st;
        for (int ho = h_lo; ho <= h_hi; ++ho) {
          const int b = hi + p.ph - ho * p.sh;
          const int base = ((to - to_lo) * nho + (ho - ho_lo)) * orowVecs + g;
          for (int wo = w_lo; wo <= w_hi; ++wo) {
            const int c = wi + p.pw - wo * p.sw;
            const unsigned lin = (a * p.kh + b) * p.kw + c;
            const uint2 iv = sidx[base + wo * G];
            float d[8];
            unpack8(stage[base + wo * G], d);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const unsigned sel = ((e < 4 ? iv.x : iv.y) >> (8 * (e & 3))) & 0xffu;
              if (sel == lin) acc[e] += d[e];
            }
          }
        }
      }
      const size_t o = ((static_cast<size_t>(n) * p.Ti + ti) * p.Hi + hi) * rowVecs + wi * G + g;
      float xv[8], out[8];
      unpack8(__ldg(x + o), xv);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float dz = fmaf(xv[e], sc[e], sf[e]) > 0.f ? acc[e] : 0.f;
        const float xh = (xv[e] - mu[e]) * is[e];
        if (MODE == 0) {
          a0[e] += dz;
          a1[e] = fmaf(dz, xh, a1[e]);
        } else {
          out[e] = k[e] * (dz - s1[e] - xh * s2[e]);
        }
      }
      if (MODE == 1) dx[o] = pack8(out);
    }
  }
  if (MODE == 0) {
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      red[threadIdx.x * 8 + e] = a0[e];
      red[2048 + threadIdx.x * 8 + e] = a1[e];
    }
    __syncthreads();
    const int reps = 256 / G;
    for (int c = threadIdx.x; c < p.C; c += 256) {
      const int gg = c >> 3, e = c & 7;
      float t0 = 0.f, t1 = 0.f;
      for (int r = 0; r < reps; ++r) {
        t0 += red[(r * G + gg) * 8 + e];
        t1 += red[2048 + (r * G + gg) * 8 + e];
      }
      atomicAdd(sum_dz + c, t0);
      atomicAdd(sum_dz_xhat + c, t1);
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

extern "C" {

int rsp_bn_relu_maxpool_supported(const rsp_pool3d_desc* d) {
  if (d->C % 8 != 0 || d->C < 8 || 256 % (d->C / 8) != 0 || d->kt * d->kh * d->kw > 255) return 0;
  const size_t row = static_cast<size_t>(d->Wi) * d->C * 2;
  return static_cast<size_t>(d->kt) * d->kh * row <= static_cast<size_t>(kFusedSmemBudget) ? 1 : 0;
}

int rsp_bn_relu_maxpool_fwd(const rsp_pool3d_desc* d, const void* x, const float* scale, const float* shift, void* y,
                            uint8_t* idx, void* stream) {
  FPGeom g;
  int rc = fill_fp(g, d);
  if (rc != RSP_OK) return rc;
  RSP_REQUIRE(rsp_bn_relu_maxpool_supported(d), "bn_relu_maxpool_fwd: one window row set does not fit in shared memory");
  const size_t row = static_cast<size_t>(g.Wi) * g.C * 2;
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
  cudaError_t e = cudaFuncSetAttribute(bn_relu_maxpool_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(bn_relu_maxpool_fwd): %s", cudaGetErrorString(e));
    return RSP_ERR_CUDA;
  }
  const int per_sm = smem > 0 ? (200 * 1024) / (smem + 1024) : 8;
  long long grid = static_cast<long long>(device_sm_count()) * (per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm));
  if (grid > tiles) grid = tiles;
  bn_relu_maxpool_fwd_kernel<<<static_cast<unsigned>(grid), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), scale, shift, static_cast<uint4*>(y), reinterpret_cast<uint2*>(idx), g);
  return check_launch("bn_relu_maxpool_fwd");
}

static int launch_bwd(int mode, const rsp_pool3d_desc* d, const void* dy, const uint8_t* idx, const void* x,
                      const float* scale, const float* shift, const float* mean, const float* invstd,
                      const float* gamma, float* sum_dz, float* sum_dz_xhat, void* dx, int C_logical, void* stream) {
  FPGeom g;
  int rc = fill_fp(g, d);
  if (rc != RSP_OK) return rc;
  g.HB = g.Hi < 4 ? g.Hi : 4;
  g.rowsIn = 0;
  g.bands = (g.Hi + g.HB - 1) / g.HB;
  const long long tiles = static_cast<long long>(g.N) * g.Ti * g.bands;
  if (tiles == 0) return RSP_OK;
  RSP_REQUIRE(tiles < (1ll << 31), "bn_relu_maxpool_bwd: too many tiles");
  g.numTiles = static_cast<int>(tiles);
  const int nto = (g.kt + g.st - 1) / g.st, nho = (g.HB + g.kh - 2) / g.sh + 1;
  const int maxWin = nto * nho * g.Wo * (g.C / 8);
  const int smem = maxWin * 24;
  RSP_REQUIRE(smem <= 160 * 1024, "bn_relu_maxpool_bwd: window staging (%d bytes) does not fit in shared memory", smem);
  const long long M = static_cast<long long>(g.N) * g.Ti * g.Hi * g.Wi;
  const float inv_m = 1.f / static_cast<float>(M);
  long long grid = static_cast<long long>(device_sm_count()) * 4;
  if (grid > tiles) grid = tiles;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  if (mode == 0) {
    e = cudaFuncSetAttribute(bn_relu_maxpool_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess)
      bn_relu_maxpool_bwd_kernel<0><<<static_cast<unsigned>(grid), 256, smem, s>>>(
          static_cast<const uint4*>(dy), reinterpret_cast<const uint2*>(idx), static_cast<const uint4*>(x), scale, shift,
          mean, invstd, gamma, sum_dz, sum_dz_xhat, nullptr, g, C_logical, inv_m, maxWin);
  } else {
    e = cudaFuncSetAttribute(bn_relu_maxpool_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess)
      bn_relu_maxpool_bwd_kernel<1><<<static_cast<unsigned>(grid), 256, smem, s>>>(
          static_cast<const uint4*>(dy), reinterpret_cast<const uint2*>(idx), static_cast<const uint4*>(x), scale, shift,
          mean, invstd, gamma, sum_dz, sum_dz_xhat, static_cast<uint4*>(dx), g, C_logical, inv_m, maxWin);
  }
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(bn_relu_maxpool_bwd): %s", cudaGetErrorString(e));
    return RSP_ERR_CUDA;
  }
  return check_launch("bn_relu_maxpool_bwd");
}

int rsp_bn_relu_maxpool_bwd_reduce(const rsp_pool3d_desc* d, const void* dy, const uint8_t* idx, const void* x,
                                   const float* scale, const float* shift, const float* mean, const float* invstd,
                                   float* sum_dz, float* sum_dz_xhat, void* stream) {
  return launch_bwd(0, d, dy, idx, x, scale, shift, mean, invstd, nullptr, sum_dz, sum_dz_xhat, nullptr, 0, stream);
}

int rsp_bn_relu_maxpool_bwd_apply(const rsp_pool3d_desc* d, const void* dy, const uint8_t* idx, const void* x,
                                  const float* scale, const float* shift, const float* mean, const float* invstd,
                                  const float* gamma, const float* sum_dz, const float* sum_dz_xhat, void* dx,
                                  int32_t C_logical, void* stream) {
  return launch_bwd(1, d, dy, idx, x, scale, shift, mean, invstd, gamma, const_cast<float*>(sum_dz),
                    const_cast<float*>(sum_dz_xhat), dx, C_logical, stream);
}

}  // extern "C"
