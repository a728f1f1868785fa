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
// pass 2 (bn_pool_bwd_dx_kernel) takes one tile of input rows, scatters gamma*invstd*dy of every window that can select
// a position of the tile into an fp32 shared-memory tile (shared-memory atomics: two overlapping windows may pick the
// same position), then streams x once and writes dx once.  dz is never materialised at input resolution.
// HBM traffic per input element: 2 B read + 2 B written (+ the pooled dy / argmax / x_max reads, L2-resident re-reads).
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
// pass 2: dx for one tile of HB input rows of one (n, ti)
//   acc (shared, fp32) [HB][Wi][8][G]  <- scatter of gamma*invstd*dy*[bn(x_max) > 0] from every window whose argmax falls
//                                          into the tile (element order [e][g]: conflict-free for the dense phase)
//   dx = acc + A*x + B
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_pool_bwd_dx_kernel(
    const uint4* __restrict__ dy, const uint2* __restrict__ idx, const uint4* __restrict__ xmax,
    const uint4* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
    const float* __restrict__ sum_dz, const float* __restrict__ sum_dz_xhat, uint4* __restrict__ dx, const BPGeom p,
    int C_logical) {
  extern __shared__ __align__(16) float acc[];          // [HB * Wi][8][G]
  __shared__ int taps[256];                             // window-local index -> a | b << 8 | c << 16
  const int G = p.C >> 3;
  const int g = threadIdx.x % G;
  const int nTaps = p.kt * p.kh * p.kw;
  for (int i = threadIdx.x; i < nTaps; i += 256) {
    const int c = i % p.kw, b = (i / p.kw) % p.kh, a = i / (p.kw * p.kh);
    taps[i] = a | (b << 8) | (c << 16);
  }
  float sc[8], sf[8], gi[8], A[8], Bc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = g * 8 + e;
    sc[e] = scale[c];
    sf[e] = shift[c];
    const float is = invstd[c];
    const float gm = c < C_logical ? gamma[c] : 0.f;     // padded channels: zero gradient
    const float c1 = sum_dz[c] * p.invM, c2 = sum_dz_xhat[c] * p.invM;
    gi[e] = gm * is;
    A[e] = -gm * is * is * c2;
    Bc[e] = -gm * is * (c1 - mean[c] * is * c2);
  }
  const int rowVecs = p.Wi * G, orowVecs = p.Wo * G;
  for (int tIdx = blockIdx.x; tIdx < p.numTiles; tIdx += gridDim.x) {
    const int band = tIdx % p.bands;
    const int q = tIdx / p.bands;
    const int ti = q % p.Ti, n = q / p.Ti;
    const int hi0 = band * p.HB;
    const int hbEff = min(p.HB, p.Hi - hi0);
    const int to_lo = max(0, cdiv(ti + p.pt - p.kt + 1, p.st)), to_hi = min(p.To - 1, fdiv(ti + p.pt, p.st));
    const int ho_lo = max(0, cdiv(hi0 + p.ph - p.kh + 1, p.sh));
    const int ho_hi = min(p.Ho - 1, fdiv(hi0 + hbEff - 1 + p.ph, p.sh));
    const int nto = max(0, to_hi - to_lo + 1), nho = max(0, ho_hi - ho_lo + 1);
    __syncthreads();                                     // previous tile's dense phase is done with acc
    const int tileVecs4 = hbEff * rowVecs * 2;           // float4 count of the accumulator tile
    for (int i = threadIdx.x; i < tileVecs4; i += 256) reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    // ---- scatter phase: one (output pixel, channel group) per thread and iteration --------------------------------
    const int items = nto * nho * orowVecs;
    for (int it = threadIdx.x; it < items; it += 256) {
      const int r = it / orowVecs;
      const int v = it - r * orowVecs;                   // wo * G + g  (g == this thread's group: orowVecs % G == 0)
      const int to = to_lo + r / nho, ho = ho_lo + r % nho;
      const int wo = v / G;
      const size_t o = ((static_cast<size_t>(n) * p.To + to) * p.Ho + ho) * orowVecs + v;
      const uint2 iv = __ldg(idx + o);
      float d[8], xv[8];
      unpack8b(__ldg(dy + o), d);
      unpack8b(__ldg(xmax + o), xv);
      const int t0 = to * p.st - p.pt - ti, h0 = ho * p.sh - p.ph - hi0, w0 = wo * p.sw - p.pw;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const uint32_t word = e < 4 ? iv.x : iv.y;
        const int tp = taps[(word >> (8 * (e & 3))) & 0xffu];
        const int hh = h0 + ((tp >> 8) & 0xff);
        if (t0 + (tp & 0xff) != 0 || hh < 0 || hh >= hbEff) continue;
        if (!(fmaf(xv[e], sc[e], sf[e]) > 0.f)) continue;
        const int ww = w0 + (tp >> 16);
        atomicAdd(acc + (static_cast<size_t>(hh) * p.Wi + ww) * p.C + e * G + g, gi[e] * d[e]);
      }
    }
    __syncthreads();
    // ---- dense phase -------------------------------------------------------------------------------------------------
    const size_t base = ((static_cast<size_t>(n) * p.Ti + ti) * p.Hi + hi0) * rowVecs;
    const int tileVecs = hbEff * rowVecs;
    for (int i = threadIdx.x; i < tileVecs; i += 256) {  // i = pixel * G + g
      float xv[8], o[8];
      unpack8b(__ldg(x + base + i), xv);
      const float* a = acc + static_cast<size_t>(i / G) * p.C + g;
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = a[e * G] + fmaf(A[e], xv[e], Bc[e]);
      dx[base + i] = pack8b(o);
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

constexpr int kAccBudget = 64 * 1024;   // fp32 accumulator tile per CTA: three CTAs per SM

}  // namespace

int bn_pool_bwd_row_bytes_ok(const rsp_pool3d_desc* d) {
  return static_cast<size_t>(d->Wi) * d->C * 4 <= static_cast<size_t>(kAccBudget) ? 1 : 0;
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
  const size_t rowBytes = static_cast<size_t>(g.Wi) * g.C * 4;
  RSP_REQUIRE(rowBytes <= static_cast<size_t>(kAccBudget),
              "bn_relu_maxpool_bwd_dx: one input row (%zu bytes of fp32) does not fit the accumulator tile", rowBytes);
  int hb = static_cast<int>(kAccBudget / rowBytes);
  if (hb > g.Hi) hb = g.Hi;
  if (hb > 8) hb = 8;
  g.HB = hb;
  g.bands = (g.Hi + hb - 1) / hb;
  const long long tiles = static_cast<long long>(g.N) * g.Ti * g.bands;
  if (tiles == 0) return RSP_OK;
  RSP_REQUIRE(tiles < (1ll << 31), "bn_relu_maxpool_bwd_dx: too many tiles");
  g.numTiles = static_cast<int>(tiles);
  const int smem = static_cast<int>(hb * rowBytes);
  static int attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(bn_pool_bwd_dx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(bn_pool_bwd_dx): %s", cudaGetErrorString(e));
      return RSP_ERR_CUDA;
    }
    attr_smem = smem;
  }
  int per_sm = (220 * 1024) / (smem + 2048);
  per_sm = per_sm < 1 ? 1 : (per_sm > 6 ? 6 : per_sm);
  long long grid = static_cast<long long>(device_sm_count()) * per_sm;
  if (grid > tiles) grid = tiles;
  bn_pool_bwd_dx_kernel<<<static_cast<unsigned>(grid), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(dy), reinterpret_cast<const uint2*>(idx), static_cast<const uint4*>(xmax),
      static_cast<const uint4*>(x), scale, shift, mean, invstd, gamma, sum_dz, sum_dz_xhat, static_cast<uint4*>(dx), g,
      C_logical);
  return check_launch("bn_relu_maxpool_bwd_dx");
}

}  // extern "C"
