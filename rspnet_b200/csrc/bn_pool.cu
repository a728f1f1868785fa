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
