// Backward of the fused train-mode BatchNorm + ReLU + MaxPool3d block (bn_pool.cu holds the forward).
//
// Reference call sites: models/resnet.py:203-206 (conv1 -> bn1 -> relu -> maxpool k3 s2 p1), models/c3d.py:111-139.
//
// The pool gradient dz (gradient w.r.t. the BN output, after the ReLU mask) is SPARSE: at most one input position per
// window and channel is non-zero.  With x_max = the raw conv output at the argmax (saved by the forward next to the
// uint8 argmax) both BN-backward reductions only need the pooled tensors:
//     sum_dz      = sum_o  dy[o] * [scale*x_max[o] + shift > 0]
//     sum_dz_xhat = sum_o  dy[o] * [..] * (x_max[o] - mean) * invstd
// and the input gradient splits into a dense affine function of x and a sparse scatter of dy:
//     dx[p] = gamma*invstd*dz[p]  +  A*x[p] + B,   A = -gamma*invstd^2*c2,  B = -gamma*invstd*(c1 - mean*invstd*c2),
//     c1 = sum_dz / M,  c2 = sum_dz_xhat / M.
// Pass 1 (bn_pool_bwd_sums_kernel) streams dy and x_max once (1/8 of the input size for the R3D-18 stem pool);
// pass 2a (bn_pool_bwd_dense_kernel) streams x once and writes dx = A*x + B once; pass 2b (bn_pool_bwd_scatter_kernel)
// streams the pooled tensors once more and adds gamma*invstd*dy at each argmax with native bf16x2 atomics (the value is
// rounded to bf16 twice: once as A*x + B, once by the add — within the stated bf16 tolerance of the conv path).
// dz is never materialised at input resolution and nothing is tiled: no shared-memory staging, no window scan.
// HBM traffic per input element: 2 B read + 2 B written, plus twice the pooled tensors (5 B per OUTPUT element).
#include "common.cuh"
#include "rspnet_b200.h"

namespace rsp {

int device_sm_count();

namespace {

__device__ __forceinline__ void unpack8b(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8b(const float (&f)[8]) {
  uint4 v;
  v.x = pack_bf16x2(f[0], f[1]);
  v.y = pack_bf16x2(f[2], f[3]);
  v.z = pack_bf16x2(f[4], f[5]);
  v.w = pack_bf16x2(f[6], f[7]);
  return v;
}

struct BPGeom {
  int N, Ti, Hi, Wi, C, To, Ho, Wo;
  int kt, kh, kw, st, sh, sw, pt, ph, pw;
  int HB;        // input rows per tile
  int bands;     // tiles per (n, ti)
  int numTiles;
  long long outVecs;   // N*To*Ho*Wo*C/8
  float invM;    // 1 / (N*Ti*Hi*Wi)
};

__host__ __device__ __forceinline__ int fdiv(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }
__host__ __device__ __forceinline__ int cdiv(int a, int b) { return fdiv(a + b - 1, b); }

// ---------------------------------------------------------------------------------------------------------------------
// pass 1: per-channel (sum dz, sum dz*xhat) from the pooled tensors only
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_pool_bwd_sums_kernel(const uint4* __restrict__ dy,
                                                               const uint4* __restrict__ xmax,
                                                               const float* __restrict__ scale,
                                                               const float* __restrict__ shift,
                                                               const float* __restrict__ mean,
                                                               const float* __restrict__ invstd,
                                                               float* __restrict__ sum_dz,
                                                               float* __restrict__ sum_dz_xhat, int C, long long nvec) {
  __shared__ float red[2 * 256 * 8];
  const int G = C >> 3;
  const int g = threadIdx.x % G;        // 256 % G == 0 and the grid stride is a multiple of 256: g is fixed per thread
  float sc[8], sf[8], mu[8], a0[8], a1[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sc[e] = scale[g * 8 + e];
    sf[e] = shift[g * 8 + e];
    mu[e] = mean[g * 8 + e];
    a0[e] = a1[e] = 0.f;
  }
  const long long stride = static_cast<long long>(gridDim.x) * 256;
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < nvec; i += stride) {
    float d[8], xv[8];
    unpack8b(__ldg(dy + i), d);
    unpack8b(__ldg(xmax + i), xv);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float dz = fmaf(xv[e], sc[e], sf[e]) > 0.f ? d[e] : 0.f;
      a0[e] += dz;
      a1[e] = fmaf(dz, xv[e] - mu[e], a1[e]);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    red[threadIdx.x * 8 + e] = a0[e];
    red[2048 + threadIdx.x * 8 + e] = a1[e];
  }
  __syncthreads();
  const int reps = 256 / G;
  for (int c = threadIdx.x; c < C; c += 256) {
    const int gg = c >> 3, e = c & 7;
    float t0 = 0.f, t1 = 0.f;
    for (int r = 0; r < reps; ++r) {
      t0 += red[(r * G + gg) * 8 + e];
      t1 += red[2048 + (r * G + gg) * 8 + e];
    }
    atomicAdd(sum_dz + c, t0);
    atomicAdd(sum_dz_xhat + c, t1 * invstd[c]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// pass 2a: dense part, dx = A*x + B (pure streaming; A, B per channel from the two sums)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_pool_bwd_dense_kernel(const uint4* __restrict__ x,
                                                                const float* __restrict__ mean,
                                                                const float* __restrict__ invstd,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ sum_dz,
                                                                const float* __restrict__ sum_dz_xhat,
                                                                uint4* __restrict__ dx, int C, int C_logical, float invM,
                                                                long long nvec) {
  const int G = C >> 3;
  const int g = threadIdx.x % G;        // fixed per thread: 256 % G == 0 and the stride is a multiple of 256
  float A[8], Bc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = g * 8 + e;
    const float is = invstd[c];
    const float gm = c < C_logical ? gamma[c] : 0.f;     // padded channels: zero gradient
    const float c1 = sum_dz[c] * invM, c2 = sum_dz_xhat[c] * invM;
    A[e] = -gm * is * is * c2;
    Bc[e] = -gm * is * (c1 - mean[c] * is * c2);
  }
  const long long stride = static_cast<long long>(gridDim.x) * 256;
  long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  for (; i + 3 * stride < nvec; i += 4 * stride) {      // four independent 16-byte loads in flight per thread
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = __ldg(x + i + k * stride);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float f[8];
      unpack8b(v[k], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = fmaf(A[e], f[e], Bc[e]);
      dx[i + k * stride] = pack8b(f);
    }
  }
  for (; i < nvec; i += stride) {
    float f[8];
    unpack8b(__ldg(x + i), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = fmaf(A[e], f[e], Bc[e]);
    dx[i] = pack8b(f);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// pass 2b: sparse part, dx[argmax] += gamma*invstd*dy wherever bn(x_max) > 0 — one (output pixel, 8 channels) per thread,
// one native bf16x2 atomic add (ATOM.ADD.BF16x2, the other half adds zero) per live channel; two overlapping windows
// that selected the same input position simply add twice.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_pool_bwd_scatter_kernel(
    const uint4* __restrict__ dy, const uint2* __restrict__ idx, const uint4* __restrict__ xmax,
    const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ invstd,
    const float* __restrict__ gamma, __nv_bfloat16* __restrict__ dx, const BPGeom p, int C_logical) {
  __shared__ int taps[256];                             // window-local index -> a | b << 8 | c << 16
  const int G = p.C >> 3;
  const int g = threadIdx.x % G;
  for (int i = threadIdx.x; i < p.kt * p.kh * p.kw; i += 256) {
    const int c = i % p.kw, b = (i / p.kw) % p.kh, a = i / (p.kw * p.kh);
    taps[i] = a | (b << 8) | (c << 16);
  }
  float sc[8], sf[8], gi[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = g * 8 + e;
    sc[e] = scale[c];
    sf[e] = shift[c];
    gi[e] = (c < C_logical ? gamma[c] : 0.f) * invstd[c];
  }
  __syncthreads();
  const long long stride = static_cast<long long>(gridDim.x) * 256;
  for (long long o = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; o < p.outVecs; o += stride) {
    const uint2 iv = __ldg(idx + o);
    float d[8], xv[8];
    unpack8b(__ldg(dy + o), d);
    unpack8b(__ldg(xmax + o), xv);
    long long pix = o / G;                               // ((n*To + to)*Ho + ho)*Wo + wo
    const int wo = static_cast<int>(pix % p.Wo);
    pix /= p.Wo;
    const int ho = static_cast<int>(pix % p.Ho);
    pix /= p.Ho;
    const int to = static_cast<int>(pix % p.To);
    const long long n = pix / p.To;
    const int t0 = to * p.st - p.pt, h0 = ho * p.sh - p.ph, w0 = wo * p.sw - p.pw;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (!(fmaf(xv[e], sc[e], sf[e]) > 0.f)) continue;
      const uint32_t word = e < 4 ? iv.x : iv.y;
      const int tp = taps[(word >> (8 * (e & 3))) & 0xffu];
      const long long pos = ((n * p.Ti + t0 + (tp & 0xff)) * p.Hi + h0 + ((tp >> 8) & 0xff)) * p.Wi + w0 + (tp >> 16);
      __nv_bfloat16* target = dx + pos * p.C + g * 8 + e;
      const float v = gi[e] * d[e];
      // the bf16x2 word that holds the target; the partner half adds +0
      const uint32_t add = (e & 1) ? pack_bf16x2(0.f, v) : pack_bf16x2(v, 0.f);
      asm volatile("red.global.add.noftz.bf16x2 [%0], %1;" ::"l"(target - (e & 1)), "r"(add) : "memory");
    }
  }
}

int fill_bp(BPGeom& g, const rsp_pool3d_desc* d) {
  g.N = d->N; g.Ti = d->Ti; g.Hi = d->Hi; g.Wi = d->Wi; g.C = d->C;
  g.kt = d->kt; g.kh = d->kh; g.kw = d->kw;
  g.st = d->st; g.sh = d->sh; g.sw = d->sw;
  g.pt = d->pt; g.ph = d->ph; g.pw = d->pw;
  g.To = (d->Ti + 2 * d->pt - d->kt) / d->st + 1;
  g.Ho = (d->Hi + 2 * d->ph - d->kh) / d->sh + 1;
  g.Wo = (d->Wi + 2 * d->pw - d->kw) / d->sw + 1;
  RSP_REQUIRE(g.To > 0 && g.Ho > 0 && g.Wo > 0, "bn_relu_maxpool_bwd: empty output");
  RSP_REQUIRE(d->C % 8 == 0 && 256 % (d->C / 8) == 0, "bn_relu_maxpool_bwd: C/8 = %d must divide 256", d->C / 8);
  RSP_REQUIRE(d->kt * d->kh * d->kw <= 255, "bn_relu_maxpool_bwd: window too large for uint8 indices");
  g.outVecs = static_cast<long long>(g.N) * g.To * g.Ho * g.Wo * (g.C / 8);
  g.invM = 1.0f / (static_cast<float>(g.N) * g.Ti * g.Hi * g.Wi);
  return RSP_OK;
}

}  // namespace

int bn_pool_bwd_row_bytes_ok(const rsp_pool3d_desc* d) {
  (void)d;
  return 1;   // the backward has no shared-memory tile: any geometry the forward takes
}

}  // namespace rsp

using namespace rsp;

extern "C" {

int rsp_bn_relu_maxpool_bwd_sums(const rsp_pool3d_desc* d, const void* dy, const void* xmax, const float* scale,
                                 const float* shift, const float* mean, const float* invstd, float* sum_dz,
                                 float* sum_dz_xhat, void* stream) {
  BPGeom g;
  int rc = fill_bp(g, d);
  if (rc != RSP_OK) return rc;
  if (g.outVecs == 0) return RSP_OK;
  long long blocks = (g.outVecs + 256 * 8 - 1) / (256 * 8);       // ~8 vectors per thread
  const long long cap = static_cast<long long>(device_sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  bn_pool_bwd_sums_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(dy), static_cast<const uint4*>(xmax), scale, shift, mean, invstd, sum_dz, sum_dz_xhat,
      g.C, g.outVecs);
  return check_launch("bn_relu_maxpool_bwd_sums");
}

int rsp_bn_relu_maxpool_bwd_dx(const rsp_pool3d_desc* d, const void* dy, const uint8_t* idx, const void* xmax,
                               const void* x, const float* scale, const float* shift, const float* mean,
                               const float* invstd, const float* gamma, int32_t C_logical, const float* sum_dz,
                               const float* sum_dz_xhat, void* dx, void* stream) {
  BPGeom g;
  int rc = fill_bp(g, d);
  if (rc != RSP_OK) return rc;
  const long long inVecs = static_cast<long long>(g.N) * g.Ti * g.Hi * g.Wi * (g.C / 8);
  if (inVecs == 0) return RSP_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long cap = static_cast<long long>(device_sm_count()) * 8;
  long long blocks = (inVecs + 256 * 8 - 1) / (256 * 8);
  if (blocks > cap) blocks = cap;
  bn_pool_bwd_dense_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(
      static_cast<const uint4*>(x), mean, invstd, gamma, sum_dz, sum_dz_xhat, static_cast<uint4*>(dx), g.C, C_logical,
      g.invM, inVecs);
  rc = check_launch("bn_relu_maxpool_bwd_dx (dense)");
  if (rc != RSP_OK || g.outVecs == 0) return rc;
  blocks = (g.outVecs + 256 * 2 - 1) / (256 * 2);
  if (blocks > cap) blocks = cap;
  bn_pool_bwd_scatter_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(
      static_cast<const uint4*>(dy), reinterpret_cast<const uint2*>(idx), static_cast<const uint4*>(xmax), scale, shift,
      invstd, gamma, static_cast<__nv_bfloat16*>(dx), g, C_logical);
  return check_launch("bn_relu_maxpool_bwd_dx (scatter)");
}

}  // extern "C"
