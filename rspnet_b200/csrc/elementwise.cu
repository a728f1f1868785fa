// HBM-bound activation kernels on bf16 NDHWC tensors: train-mode BatchNorm (+ReLU +residual), MaxPool3d,
// the projection heads, and layout conversion.  All of them move 16-byte vectors (8 channels) per thread.
// Reference call sites: models/resnet.py:54-75,137-139; models/c3d.py:22-50; moco/split_wrapper.py:128-169.
#include "common.cuh"
#include "rspnet_b200.h"

namespace rsp {

int device_sm_count();

struct bf8 {
  uint4 raw;
};
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

static inline unsigned ew_grid(size_t items, int block) {
  size_t need = (items + block - 1) / block;
  size_t cap = static_cast<size_t>(device_sm_count()) * 16;
  if (need < 1) need = 1;
  return static_cast<unsigned>(need < cap ? need : cap);
}

// ------------------------------------------------------------------------------------------------
// per-channel reductions: thread owns one 8-channel group, strides over rows
// MODE 0: sum x, sum x^2.   MODE 1: sum dz, sum dz*xhat (BN backward)
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) channel_reduce_kernel(const uint4* __restrict__ x, const uint4* __restrict__ out,
                                                             const uint4* __restrict__ dout,
                                                             const float* __restrict__ mean,
                                                             const float* __restrict__ invstd, int relu, size_t M,
                                                             int C, float* __restrict__ r0, float* __restrict__ r1) {
  const int G = C >> 3;             // channel groups per row (<= 256)
  const int rows_per_iter = 256 / G;
  const int g = threadIdx.x % G;
  const int rsub = threadIdx.x / G;
  const bool active = rsub < rows_per_iter;   // threads beyond rows_per_iter*G idle (G need not divide 256)
  float a[8], b[8], mu[8], is[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = b[i] = 0.f;
    if (MODE == 1) {
      mu[i] = mean[g * 8 + i];
      is[i] = invstd[g * 8 + i];
    }
  }
  const size_t rstride = static_cast<size_t>(gridDim.x) * rows_per_iter;
  size_t row = static_cast<size_t>(blockIdx.x) * rows_per_iter + rsub;
  // two rows per trip with all loads issued up front (the sums still run in row order: results are unchanged)
  for (; active && row + rstride < M; row += 2 * rstride) {
    const size_t o0 = row * G + g, o1 = (row + rstride) * G + g;
    const uint4 xa = __ldg(x + o0), xb = __ldg(x + o1);
    uint4 da = make_uint4(0, 0, 0, 0), db = da, oa = da, ob = da;
    if (MODE == 1) {
      da = __ldg(dout + o0);
      db = __ldg(dout + o1);
      if (relu) {
        oa = __ldg(out + o0);
        ob = __ldg(out + o1);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float xv[8];
      unpack8(u ? xb : xa, xv);
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          a[i] += xv[i];
          b[i] = fmaf(xv[i], xv[i], b[i]);
        }
      } else {
        float dv[8], ov[8];
        unpack8(u ? db : da, dv);
        if (relu) unpack8(u ? ob : oa, ov);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float dz = (relu && !(ov[i] > 0.f)) ? 0.f : dv[i];
          a[i] += dz;
          b[i] = fmaf(dz, (xv[i] - mu[i]) * is[i], b[i]);
        }
      }
    }
  }
  for (; active && row < M; row += rstride) {
    size_t o = row * G + g;
    float xv[8];
    unpack8(__ldg(x + o), xv);
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a[i] += xv[i];
        b[i] = fmaf(xv[i], xv[i], b[i]);
      }
    } else {
      float dv[8], ov[8];
      unpack8(__ldg(dout + o), dv);
      if (relu) unpack8(__ldg(out + o), ov);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float dz = (relu && !(ov[i] > 0.f)) ? 0.f : dv[i];
        a[i] += dz;
        b[i] = fmaf(dz, (xv[i] - mu[i]) * is[i], b[i]);
      }
    }
  }
  // block reduce across the rows_per_iter threads sharing a channel group
  __shared__ float sa[256 * 8];
  __shared__ float sb[256 * 8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sa[threadIdx.x * 8 + i] = a[i];
    sb[threadIdx.x * 8 + i] = b[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    int gg = c >> 3, i = c & 7;
    float s0 = 0.f, s1 = 0.f;
    for (int r = 0; r < rows_per_iter; ++r) {
      s0 += sa[(r * G + gg) * 8 + i];
      s1 += sb[(r * G + gg) * 8 + i];
    }
    atomicAdd(r0 + c, s0);
    atomicAdd(r1 + c, s1);
  }
}

__global__ void bn_finalize_kernel(float* __restrict__ sum, float* __restrict__ sumsq, int clear, float count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* __restrict__ rmean, float* __restrict__ rvar,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean,
                                   float* __restrict__ invstd, int C, int Cl) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float s1 = sum[c], s2 = sumsq[c];
  if (clear) {  // leave the accumulators zeroed for the next forward (persistent per-layer buffers, no memset launches)
    sum[c] = 0.f;
    sumsq[c] = 0.f;
  }
  if (c >= Cl) {
    scale[c] = 0.f;
    shift[c] = 0.f;
    mean[c] = 0.f;
    invstd[c] = 0.f;
    return;
  }
  float mu = s1 / count;
  float var = fmaxf(s2 / count - mu * mu, 0.f);
  float is = rsqrtf(var + eps);
  float sc = gamma[c] * is;
  scale[c] = sc;
  shift[c] = beta[c] - mu * sc;
  mean[c] = mu;
  invstd[c] = is;
  if (rmean) rmean[c] = (1.f - momentum) * rmean[c] + momentum * mu;
  if (rvar) {
    float unbiased = count > 1.f ? var * count / (count - 1.f) : var;
    rvar[c] = (1.f - momentum) * rvar[c] + momentum * unbiased;
  }
}

__global__ void __launch_bounds__(256) bn_act_fwd_kernel(const uint4* __restrict__ x, const float* __restrict__ scale,
                                                         const float* __restrict__ shift,
                                                         const uint4* __restrict__ res, int relu,
                                                         uint4* __restrict__ out, size_t nvec, int C) {
  extern __shared__ float sm[];
  float* ssc = sm;
  float* ssh = sm + C;
  for (int c = threadIdx.x; c < C; c += 256) {
    ssc[c] = scale[c];
    ssh[c] = shift[c];
  }
  __syncthreads();
  const int G = C >> 3;
  for (size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x; i < nvec;
       i += static_cast<size_t>(gridDim.x) * 256) {
    int g = static_cast<int>(i % G);
    float xv[8], rv[8];
    unpack8(__ldg(x + i), xv);
    if (res) unpack8(__ldg(res + i), rv);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float y = fmaf(xv[e], ssc[g * 8 + e], ssh[g * 8 + e]);
      if (res) y += rv[e];
      xv[e] = relu ? fmaxf(y, 0.f) : y;
    }
    out[i] = pack8(xv);
  }
}

// bn_finalize + bn_act_fwd in one launch: every block derives scale / shift of all channels from the batch sums (C rsqrt
// per block is noise next to the activation pass); block 0 also publishes (scale, shift, mean, invstd) for the backward,
// updates the running statistics and zeroes the OTHER parity's accumulators for the next forward of this layer (the
// conv epilogue of that forward is stream-ordered after this kernel; the sums read here are cleared one call later).
__global__ void __launch_bounds__(256) bn_finalize_act_fwd_kernel(
    const uint4* __restrict__ x, const float* __restrict__ sum, const float* __restrict__ sumsq,
    float* __restrict__ clear, float count, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
    float momentum, float* __restrict__ rmean, float* __restrict__ rvar, float* __restrict__ rows,
    const uint4* __restrict__ res, int relu, uint4* __restrict__ out, size_t nvec, int C, int Cl) {
  extern __shared__ float sm[];
  float* ssc = sm;
  float* ssh = sm + C;
  for (int c = threadIdx.x; c < C; c += 256) {
    float sc = 0.f, sh = 0.f, mu = 0.f, is = 0.f;
    if (c < Cl) {
      mu = sum[c] / count;
      const float var = fmaxf(sumsq[c] / count - mu * mu, 0.f);
      is = rsqrtf(var + eps);
      sc = gamma[c] * is;
      sh = beta[c] - mu * sc;
      if (blockIdx.x == 0) {
        if (rmean) rmean[c] = (1.f - momentum) * rmean[c] + momentum * mu;
        if (rvar) {
          const float unbiased = count > 1.f ? var * count / (count - 1.f) : var;
          rvar[c] = (1.f - momentum) * rvar[c] + momentum * unbiased;
        }
      }
    }
    ssc[c] = sc;
    ssh[c] = sh;
    if (blockIdx.x == 0) {
      rows[c] = sc;
      rows[C + c] = sh;
      rows[2 * C + c] = mu;
      rows[3 * C + c] = is;
      if (clear) {
        clear[c] = 0.f;
        clear[C + c] = 0.f;
      }
    }
  }
  __syncthreads();
  const int G = C >> 3;
  if (256 % G == 0) {
    // the block start and the grid stride are multiples of G: a thread keeps its channel group, so scale / shift live in
    // registers and the loop is four independent 16-byte loads (+ residuals) in flight per thread — no 64-bit modulo and
    // no shared-memory reads per element
    const int g = threadIdx.x % G;
    float sc[8], sh[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sc[e] = ssc[g * 8 + e];
      sh[e] = ssh[g * 8 + e];
    }
    const size_t stride = static_cast<size_t>(gridDim.x) * 256;
    size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
      uint4 xr[4], rr[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) xr[u] = __ldg(x + i + u * stride);
      if (res) {
#pragma unroll
        for (int u = 0; u < 4; ++u) rr[u] = __ldg(res + i + u * stride);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float xv[8], rv[8];
        unpack8(xr[u], xv);
        if (res) unpack8(rr[u], rv);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float y = fmaf(xv[e], sc[e], sh[e]);
          if (res) y += rv[e];
          xv[e] = relu ? fmaxf(y, 0.f) : y;
        }
        out[i + u * stride] = pack8(xv);
      }
    }
    for (; i < nvec; i += stride) {
      float xv[8], rv[8];
      unpack8(__ldg(x + i), xv);
      if (res) unpack8(__ldg(res + i), rv);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float y = fmaf(xv[e], sc[e], sh[e]);
        if (res) y += rv[e];
        xv[e] = relu ? fmaxf(y, 0.f) : y;
      }
      out[i] = pack8(xv);
    }
    return;
  }
  for (size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x; i < nvec;
       i += static_cast<size_t>(gridDim.x) * 256) {
    int g = static_cast<int>(i % G);
    float xv[8], rv[8];
    unpack8(__ldg(x + i), xv);
    if (res) unpack8(__ldg(res + i), rv);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float y = fmaf(xv[e], ssc[g * 8 + e], ssh[g * 8 + e]);
      if (res) y += rv[e];
      xv[e] = relu ? fmaxf(y, 0.f) : y;
    }
    out[i] = pack8(xv);
  }
}

__global__ void __launch_bounds__(256) bn_act_bwd_apply_kernel(
    const uint4* __restrict__ dout, const uint4* __restrict__ out, const uint4* __restrict__ x,
    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
    const float* __restrict__ sum_dz, const float* __restrict__ sum_dz_xhat, int relu, uint4* __restrict__ dx,
    uint4* __restrict__ dres, size_t nvec, int C, int Cl, float inv_m) {
  extern __shared__ float sm[];
  float* smu = sm;          // mean
  float* sis = sm + C;      // invstd
  float* sk = sm + 2 * C;   // gamma*invstd
  float* s1 = sm + 3 * C;   // sum_dz / M
  float* s2 = sm + 4 * C;   // sum_dz_xhat / M
  for (int c = threadIdx.x; c < C; c += 256) {
    bool live = c < Cl;
    smu[c] = mean[c];
    sis[c] = invstd[c];
    sk[c] = live ? gamma[c] * invstd[c] : 0.f;
    s1[c] = sum_dz[c] * inv_m;
    s2[c] = sum_dz_xhat[c] * inv_m;
  }
  __syncthreads();
  const int G = C >> 3;
  if (256 % G == 0) {   // see bn_finalize_act_fwd_kernel: per-thread channel group, parameters in registers, two rows in flight
    const int g = threadIdx.x % G;
    float mu[8], is[8], kk[8], m1[8], m2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      mu[e] = smu[g * 8 + e];
      is[e] = sis[g * 8 + e];
      kk[e] = sk[g * 8 + e];
      m1[e] = s1[g * 8 + e];
      m2[e] = s2[g * 8 + e];
    }
    const size_t stride = static_cast<size_t>(gridDim.x) * 256;
    size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
    auto one = [&](size_t at, const uint4& dr, const uint4& xr, const uint4& orr) {
      float dv[8], ov[8], xv[8], o[8];
      unpack8(dr, dv);
      unpack8(xr, xv);
      if (relu) unpack8(orr, ov);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float dz = (relu && !(ov[e] > 0.f)) ? 0.f : dv[e];
        dv[e] = dz;
        float xh = (xv[e] - mu[e]) * is[e];
        o[e] = kk[e] * (dz - m1[e] - xh * m2[e]);
      }
      dx[at] = pack8(o);
      if (dres) dres[at] = pack8(dv);
    };
    for (; i + stride < nvec; i += 2 * stride) {
      const uint4 d0 = __ldg(dout + i), d1 = __ldg(dout + i + stride);
      const uint4 x0 = __ldg(x + i), x1 = __ldg(x + i + stride);
      uint4 o0 = make_uint4(0, 0, 0, 0), o1 = o0;
      if (relu) {
        o0 = __ldg(out + i);
        o1 = __ldg(out + i + stride);
      }
      one(i, d0, x0, o0);
      one(i + stride, d1, x1, o1);
    }
    for (; i < nvec; i += stride) {
      const uint4 d0 = __ldg(dout + i), x0 = __ldg(x + i);
      const uint4 o0 = relu ? __ldg(out + i) : make_uint4(0, 0, 0, 0);
      one(i, d0, x0, o0);
    }
    return;
  }
  for (size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x; i < nvec;
       i += static_cast<size_t>(gridDim.x) * 256) {
    int g = static_cast<int>(i % G);
    float dv[8], ov[8], xv[8], o[8];
    unpack8(__ldg(dout + i), dv);
    unpack8(__ldg(x + i), xv);
    if (relu) unpack8(__ldg(out + i), ov);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      int c = g * 8 + e;
      float dz = (relu && !(ov[e] > 0.f)) ? 0.f : dv[e];
      dv[e] = dz;
      float xh = (xv[e] - smu[c]) * sis[c];
      o[e] = sk[c] * (dz - s1[c] - xh * s2[c]);
    }
    dx[i] = pack8(o);
    if (dres) dres[i] = pack8(dv);
  }
}

// ------------------------------------------------------------------------------------------------
// MaxPool3d
// ------------------------------------------------------------------------------------------------
struct PoolGeom {
  int N, Ti, Hi, Wi, C, To, Ho, Wo;
  int kt, kh, kw, st, sh, sw, pt, ph, pw;
};

__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y,
                                                          uint2* __restrict__ idx, PoolGeom p, size_t total) {
  const int G = p.C >> 3;
  for (size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * 256) {
    int g = static_cast<int>(i % G);
    size_t pix = i / G;
    int wo = static_cast<int>(pix % p.Wo);
    size_t q = pix / p.Wo;
    int ho = static_cast<int>(q % p.Ho);
    q /= p.Ho;
    int to = static_cast<int>(q % p.To);
    int n = static_cast<int>(q / p.To);
    float best[8];
    unsigned bi[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      best[e] = -INFINITY;
      bi[e] = 0;
    }
    bool any = false;
    for (int a = 0; a < p.kt; ++a) {
      int ti = to * p.st - p.pt + a;
      if (ti < 0 || ti >= p.Ti) continue;
      for (int b = 0; b < p.kh; ++b) {
        int hi = ho * p.sh - p.ph + b;
        if (hi < 0 || hi >= p.Hi) continue;
        for (int c = 0; c < p.kw; ++c) {
          int wi = wo * p.sw - p.pw + c;
          if (wi < 0 || wi >= p.Wi) continue;
          size_t o = (((static_cast<size_t>(n) * p.Ti + ti) * p.Hi + hi) * p.Wi + wi) * G + g;
          float v[8];
          unpack8(__ldg(x + o), v);
          unsigned lin = (a * p.kh + b) * p.kw + c;
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
    y[i] = pack8(best);
    uint2 iv;
    iv.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
    iv.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
    idx[i] = iv;
  }
}

// gather form: every input element sums dy over the windows that selected it (deterministic, no atomics)
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const uint4* __restrict__ dy, const uint2* __restrict__ idx,
                                                          uint4* __restrict__ dx, PoolGeom p, size_t total) {
  const int G = p.C >> 3;
  for (size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * 256) {
    int g = static_cast<int>(i % G);
    size_t pix = i / G;
    int wi = static_cast<int>(pix % p.Wi);
    size_t q = pix / p.Wi;
    int hi = static_cast<int>(q % p.Hi);
    q /= p.Hi;
    int ti = static_cast<int>(q % p.Ti);
    int n = static_cast<int>(q / p.Ti);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    // output positions o with o*s - p <= i <= o*s - p + k - 1
    int to_lo = max(0, (ti + p.pt - p.kt + p.st) / p.st), to_hi = min(p.To - 1, (ti + p.pt) / p.st);
    int ho_lo = max(0, (hi + p.ph - p.kh + p.sh) / p.sh), ho_hi = min(p.Ho - 1, (hi + p.ph) / p.sh);
    int wo_lo = max(0, (wi + p.pw - p.kw + p.sw) / p.sw), wo_hi = min(p.Wo - 1, (wi + p.pw) / p.sw);
    if (ti + p.pt - p.kt + 1 <= 0) to_lo = 0;
    if (hi + p.ph - p.kh + 1 <= 0) ho_lo = 0;
    if (wi + p.pw - p.kw + 1 <= 0) wo_lo = 0;
    for (int to = to_lo; to <= to_hi; ++to) {
      int a = ti + p.pt - to * p.st;
      if (a < 0 || a >= p.kt) continue;
      for (int ho = ho_lo; ho <= ho_hi; ++ho) {
        int b = hi + p.ph - ho * p.sh;
        if (b < 0 || b >= p.kh) continue;
        for (int wo = wo_lo; wo <= wo_hi; ++wo) {
          int c = wi + p.pw - wo * p.sw;
          if (c < 0 || c >= p.kw) continue;
          unsigned lin = (a * p.kh + b) * p.kw + c;
          size_t o = (((static_cast<size_t>(n) * p.To + to) * p.Ho + ho) * p.Wo + wo) * G + g;
          uint2 iv = __ldg(idx + o);
          float d[8];
          unpack8(__ldg(dy + o), d);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            unsigned sel = ((e < 4 ? iv.x : iv.y) >> (8 * (e & 3))) & 0xffu;
            if (sel == lin) acc[e] += d[e];
          }
        }
      }
    }
    dx[i] = pack8(acc);
  }
}

// ------------------------------------------------------------------------------------------------
// projection heads: avg-pool -> 2 x Linear -> L2 normalise. One block per sample.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) s += red[w];
  return s;
}

// forward: block per sample, 256 threads = 8 warps. Pooled features in smem; every output feature is a warp-level dot
// product with coalesced reads of its weight row.
__global__ void __launch_bounds__(256) head_fwd_kernel(const __nv_bfloat16* __restrict__ feat, int S, int C, int Cl,
                                                       int D, const float* __restrict__ w1,
                                                       const float* __restrict__ b1, const float* __restrict__ w2,
                                                       const float* __restrict__ b2, float* __restrict__ pooled,
                                                       float* __restrict__ raw1, float* __restrict__ raw2,
                                                       float* __restrict__ out1, float* __restrict__ out2) {
  extern __shared__ float sm[];
  float* sp = sm;             // [Cl]
  float* sraw = sm + Cl;      // [2][D]
  __shared__ float red[8];
  const int bidx = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __nv_bfloat16* f = feat + static_cast<size_t>(bidx) * S * C;
  for (int c = threadIdx.x; c < Cl; c += 256) {
    float s = 0.f;
    for (int i = 0; i < S; ++i) s += __bfloat162float(f[static_cast<size_t>(i) * C + c]);
    s /= S;
    sp[c] = s;
    pooled[static_cast<size_t>(bidx) * Cl + c] = s;
  }
  __syncthreads();
  for (int o = warp; o < 2 * D; o += 8) {
    const int head = o / D, j = o - head * D;
    const float* wr = (head ? w2 : w1) + static_cast<size_t>(j) * Cl;
    float acc = 0.f;
    for (int c = lane; c < Cl; c += 32) acc = fmaf(wr[c], sp[c], acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) {
      const float* bb = head ? b2 : b1;
      sraw[o] = acc + (bb ? bb[j] : 0.f);
    }
  }
  __syncthreads();
  for (int head = 0; head < 2; ++head) {
    float* raw = (head ? raw2 : raw1) + static_cast<size_t>(bidx) * D;
    float* out = (head ? out2 : out1) + static_cast<size_t>(bidx) * D;
    float sq = 0.f;
    for (int j = threadIdx.x; j < D; j += 256) sq += sraw[head * D + j] * sraw[head * D + j];
    float nrm = fmaxf(sqrtf(block_sum(sq, red)), 1e-12f);
    for (int j = threadIdx.x; j < D; j += 256) {
      raw[j] = sraw[head * D + j];
      out[j] = sraw[head * D + j] / nrm;
    }
  }
}

// backward, step 1 (block per sample): gradient w.r.t. the pre-normalisation outputs -> dr_ws [B][2][D];
// dfeat = (dr1 W1 + dr2 W2) / S broadcast over the S positions.
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ dout1, const float* __restrict__ dout2,
                                                       const float* __restrict__ raw1, const float* __restrict__ raw2,
                                                       int S, int C, int Cl, int D, const float* __restrict__ w1,
                                                       const float* __restrict__ w2, float* __restrict__ dr_ws,
                                                       __nv_bfloat16* __restrict__ dfeat) {
  extern __shared__ float sm[];
  float* dr = sm;  // [2][D]
  __shared__ float red[8];
  const int bidx = blockIdx.x;
  for (int head = 0; head < 2; ++head) {
    const float* raw = (head ? raw2 : raw1) + static_cast<size_t>(bidx) * D;
    const float* dout = (head ? dout2 : dout1) + static_cast<size_t>(bidx) * D;
    float sq = 0.f, dot = 0.f;
    for (int j = threadIdx.x; j < D; j += 256) {
      sq += raw[j] * raw[j];
      dot += raw[j] * dout[j];
    }
    const float ss = block_sum(sq, red);
    const float dd = block_sum(dot, red);
    const float nrm = sqrtf(ss);
    for (int j = threadIdx.x; j < D; j += 256) {
      // y = x/n: dx = dy/n - x * (x.dy) / n^3   (n clamped at eps like F.normalize)
      float g = nrm > 1e-12f ? dout[j] / nrm - raw[j] * dd / (nrm * nrm * nrm) : dout[j] / 1e-12f;
      dr[head * D + j] = g;
      dr_ws[(static_cast<size_t>(bidx) * 2 + head) * D + j] = g;
    }
  }
  __syncthreads();
  if (dfeat) {
    __nv_bfloat16* df = dfeat + static_cast<size_t>(bidx) * S * C;
    for (int c = threadIdx.x; c < C; c += 256) {   // coalesced over c: w[j][c]
      float acc = 0.f;
      if (c < Cl) {
        for (int j = 0; j < D; ++j)
          acc += dr[j] * w1[static_cast<size_t>(j) * Cl + c] + dr[D + j] * w2[static_cast<size_t>(j) * Cl + c];
        acc /= S;
      }
      const __nv_bfloat16 v = __float2bfloat16(acc);
      for (int i = 0; i < S; ++i) df[static_cast<size_t>(i) * C + c] = v;
    }
  }
}

// backward, step 2: dW_h[j][c] += sum_b dr[b][h][j] * pooled[b][c]; db_h[j] += sum_b dr[b][h][j].  No atomics.
__global__ void __launch_bounds__(256) head_wgrad_kernel(const float* __restrict__ dr_ws, const float* __restrict__ pooled,
                                                         int B, int Cl, int D, float* __restrict__ dw1,
                                                         float* __restrict__ db1, float* __restrict__ dw2,
                                                         float* __restrict__ db2) {
  const int head = blockIdx.z;
  const int j = blockIdx.y;
  float* dw = head ? dw2 : dw1;
  float* db = head ? db2 : db1;
  for (int c = blockIdx.x * 256 + threadIdx.x; c < Cl; c += gridDim.x * 256) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b)
      acc = fmaf(dr_ws[(static_cast<size_t>(b) * 2 + head) * D + j], pooled[static_cast<size_t>(b) * Cl + c], acc);
    dw[static_cast<size_t>(j) * Cl + c] += acc;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc += dr_ws[(static_cast<size_t>(b) * 2 + head) * D + j];
    db[j] += acc;
  }
}

// ------------------------------------------------------------------------------------------------
// layout conversion
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ncdhw_to_ndhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                             int C, int Cs, size_t THW, size_t total) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * 256) {
    size_t n = i / THW, p = i - n * THW;
    const float* s = x + n * C * THW + p;
    __nv_bfloat16* o = y + i * Cs;
    for (int c = 0; c < Cs; ++c) o[c] = __float2bfloat16(c < C ? s[c * THW] : 0.f);
  }
}
__global__ void __launch_bounds__(256) ndhwc_to_ncdhw_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y,
                                                             int C, int Cs, size_t THW, size_t total) {
  // total = N*C*THW, output-major so that writes are coalesced
  for (size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * 256) {
    size_t p = i % THW;
    size_t nc = i / THW;
    size_t n = nc / C;
    int c = static_cast<int>(nc - n * C);
    y[i] = __bfloat162float(x[(n * THW + p) * Cs + c]);
  }
}

static int check_c(int C, const char* what) {
  RSP_REQUIRE(C >= 8 && C % 8 == 0 && C <= 2048, "%s: channel count %d unsupported (need a multiple of 8, <= 2048)",
              what, C);
  return RSP_OK;
}

static int fill_pool(PoolGeom& g, const rsp_pool3d_desc* d) {
  g.N = d->N; g.Ti = d->Ti; g.Hi = d->Hi; g.Wi = d->Wi; g.C = d->C;
  g.kt = d->kt; g.kh = d->kh; g.kw = d->kw;
  g.st = d->st; g.sh = d->sh; g.sw = d->sw;
  g.pt = d->pt; g.ph = d->ph; g.pw = d->pw;
  g.To = (d->Ti + 2 * d->pt - d->kt) / d->st + 1;
  g.Ho = (d->Hi + 2 * d->ph - d->kh) / d->sh + 1;
  g.Wo = (d->Wi + 2 * d->pw - d->kw) / d->sw + 1;
  RSP_REQUIRE(g.To > 0 && g.Ho > 0 && g.Wo > 0, "maxpool3d: empty output");
  RSP_REQUIRE(d->C % 8 == 0, "maxpool3d: C %% 8 != 0");
  RSP_REQUIRE(d->kt * d->kh * d->kw <= 255, "maxpool3d: window too large for uint8 indices");
  return RSP_OK;
}

}  // namespace rsp

using namespace rsp;

extern "C" {

int rsp_bn_stats(const void* x, int64_t M, int32_t C, float* sum, float* sumsq, void* stream) {
  int rc = check_c(C, "bn_stats");
  if (rc != RSP_OK) return rc;
  if (M == 0) return RSP_OK;
  const int rows_per_iter = 256 / (C / 8);
  unsigned grid = ew_grid(static_cast<size_t>((M + rows_per_iter - 1) / rows_per_iter) * 256 / 8 + 1, 256);
  channel_reduce_kernel<0><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), nullptr, nullptr, nullptr, nullptr, 0, static_cast<size_t>(M), C, sum, sumsq);
  return check_launch("bn_stats");
}

int rsp_bn_finalize(float* sum, float* sumsq, int32_t clear_sums, int64_t count, const float* gamma,
                    const float* beta, float eps, float momentum, float* running_mean, float* running_var, float* scale,
                    float* shift, float* mean, float* invstd, int32_t C, int32_t C_logical, void* stream) {
  RSP_REQUIRE(count > 0 && C_logical <= C, "bn_finalize: bad count / channels");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      sum, sumsq, clear_sums, static_cast<float>(count), gamma, beta, eps, momentum, running_mean, running_var, scale, shift, mean,
      invstd, C, C_logical);
  return check_launch("bn_finalize");
}

int rsp_bn_act_fwd(const void* x, const float* scale, const float* shift, const void* residual, int relu, void* out,
                   int64_t M, int32_t C, void* stream) {
  RSP_REQUIRE(C % 8 == 0, "bn_act_fwd: C %% 8 != 0");
  if (M == 0) return RSP_OK;
  size_t nvec = static_cast<size_t>(M) * (C / 8);
  bn_act_fwd_kernel<<<ew_grid(nvec, 256), 256, 2 * C * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), scale, shift, static_cast<const uint4*>(residual), relu, static_cast<uint4*>(out),
      nvec, C);
  return check_launch("bn_act_fwd");
}

int rsp_bn_finalize_act_fwd(const void* x, const float* sum, const float* sumsq, float* clear_sums, int64_t count,
                            const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                            float* running_var, float* rows, const void* residual, int relu, void* out, int64_t M,
                            int32_t C, int32_t C_logical, void* stream) {
  RSP_REQUIRE(C % 8 == 0 && C <= 4096 && C_logical <= C && count > 0, "bn_finalize_act_fwd: bad channels / count");
  if (M == 0) return RSP_OK;
  size_t nvec = static_cast<size_t>(M) * (C / 8);
  bn_finalize_act_fwd_kernel<<<ew_grid(nvec, 256), 256, 2 * C * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), sum, sumsq, clear_sums, static_cast<float>(count), gamma, beta, eps, momentum,
      running_mean, running_var, rows, static_cast<const uint4*>(residual), relu, static_cast<uint4*>(out), nvec, C,
      C_logical);
  return check_launch("bn_finalize_act_fwd");
}

int rsp_bn_act_bwd_reduce(const void* dout, const void* out, const void* x, const float* mean, const float* invstd,
                          int relu, float* sum_dz, float* sum_dz_xhat, int64_t M, int32_t C, void* stream) {
  int rc = check_c(C, "bn_act_bwd_reduce");
  if (rc != RSP_OK) return rc;
  if (M == 0) return RSP_OK;
  const int rows_per_iter = 256 / (C / 8);
  unsigned grid = ew_grid(static_cast<size_t>((M + rows_per_iter - 1) / rows_per_iter) * 256 / 8 + 1, 256);
  channel_reduce_kernel<1><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), static_cast<const uint4*>(out), static_cast<const uint4*>(dout), mean, invstd, relu,
      static_cast<size_t>(M), C, sum_dz, sum_dz_xhat);
  return check_launch("bn_act_bwd_reduce");
}

int rsp_bn_act_bwd_apply(const void* dout, const void* out, const void* x, const float* mean, const float* invstd,
                         const float* gamma, const float* sum_dz, const float* sum_dz_xhat, int relu, void* dx,
                         void* dres, int64_t M, int32_t C, int32_t C_logical, void* stream) {
  RSP_REQUIRE(C % 8 == 0 && C_logical <= C, "bn_act_bwd_apply: bad channels");
  if (M == 0) return RSP_OK;
  size_t nvec = static_cast<size_t>(M) * (C / 8);
  bn_act_bwd_apply_kernel<<<ew_grid(nvec, 256), 256, 5 * C * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(dout), static_cast<const uint4*>(out), static_cast<const uint4*>(x), mean, invstd,
      gamma, sum_dz, sum_dz_xhat, relu, static_cast<uint4*>(dx), static_cast<uint4*>(dres), nvec, C, C_logical,
      1.f / static_cast<float>(M));
  return check_launch("bn_act_bwd_apply");
}

int rsp_maxpool3d_fwd(const rsp_pool3d_desc* d, const void* x, void* y, uint8_t* idx, void* stream) {
  PoolGeom g;
  int rc = fill_pool(g, d);
  if (rc != RSP_OK) return rc;
  size_t total = static_cast<size_t>(g.N) * g.To * g.Ho * g.Wo * (g.C / 8);
  if (total == 0) return RSP_OK;
  maxpool_fwd_kernel<<<ew_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(y), reinterpret_cast<uint2*>(idx), g, total);
  return check_launch("maxpool3d_fwd");
}

int rsp_maxpool3d_bwd(const rsp_pool3d_desc* d, const void* dy, const uint8_t* idx, void* dx, void* stream) {
  PoolGeom g;
  int rc = fill_pool(g, d);
  if (rc != RSP_OK) return rc;
  size_t total = static_cast<size_t>(g.N) * g.Ti * g.Hi * g.Wi * (g.C / 8);
  if (total == 0) return RSP_OK;
  maxpool_bwd_kernel<<<ew_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(dy), reinterpret_cast<const uint2*>(idx), static_cast<uint4*>(dx), g, total);
  return check_launch("maxpool3d_bwd");
}

int rsp_head_fwd(const void* feat, int32_t B, int32_t S, int32_t C, int32_t C_logical, int32_t D, const float* w1,
                 const float* b1, const float* w2, const float* b2, float* pooled, float* raw1, float* raw2,
                 float* out1, float* out2, void* stream) {
  RSP_REQUIRE(B > 0 && S > 0 && C_logical <= C && (C_logical + 2 * D) * sizeof(float) <= 48 * 1024,
              "head_fwd: bad sizes");
  head_fwd_kernel<<<B, 256, (C_logical + 2 * D) * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(feat), S, C, C_logical, D, w1, b1, w2, b2, pooled, raw1, raw2, out1, out2);
  return check_launch("head_fwd");
}

int rsp_head_bwd(const float* dout1, const float* dout2, const float* pooled, const float* raw1, const float* raw2,
                 int32_t B, int32_t S, int32_t C, int32_t C_logical, int32_t D, const float* w1, const float* w2,
                 float* dr_ws, float* dw1, float* db1, float* dw2, float* db2, void* dfeat, void* stream) {
  RSP_REQUIRE(B > 0 && S > 0 && C_logical <= C && 2 * D * sizeof(float) <= 48 * 1024, "head_bwd: bad sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  head_bwd_kernel<<<B, 256, 2 * D * sizeof(float), st>>>(dout1, dout2, raw1, raw2, S, C, C_logical, D, w1, w2, dr_ws,
                                                         static_cast<__nv_bfloat16*>(dfeat));
  dim3 grid((C_logical + 255) / 256, D, 2);
  head_wgrad_kernel<<<grid, 256, 0, st>>>(dr_ws, pooled, B, C_logical, D, dw1, db1, dw2, db2);
  return check_launch("head_bwd");
}

int rsp_ncdhw_to_ndhwc_bf16(const float* x, void* y, int32_t N, int32_t C, int32_t Cs, int64_t THW, void* stream) {
  RSP_REQUIRE(Cs >= C, "layout: Cs < C");
  size_t total = static_cast<size_t>(N) * THW;
  if (total == 0) return RSP_OK;
  ncdhw_to_ndhwc_kernel<<<ew_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<__nv_bfloat16*>(y), C, Cs, static_cast<size_t>(THW), total);
  return check_launch("ncdhw_to_ndhwc");
}

int rsp_ndhwc_bf16_to_ncdhw(const void* x, float* y, int32_t N, int32_t C, int32_t Cs, int64_t THW, void* stream) {
  RSP_REQUIRE(Cs >= C, "layout: Cs < C");
  size_t total = static_cast<size_t>(N) * C * THW;
  if (total == 0) return RSP_OK;
  ndhwc_to_ncdhw_kernel<<<ew_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), y, C, Cs, static_cast<size_t>(THW), total);
  return check_launch("ndhwc_to_ncdhw");
}

}  // extern "C"

// =====================================================================================================================
// On-GPU clip pipeline (reference: per-clip python loop of SequentialGPUCollateFn, datasets/transforms_video/
// transforms_tensor.py:214-233, applying ToTensorVideo -> Resize(bilinear, align_corners=False) -> RandomGrayScale ->
// RandomHorizontalFlip -> Normalize; crop boxes from RawVideoRandomCrop, frame indices from RandomStrideCrop).
// One launch for the whole batch: every output pixel gathers its 4 bilinear taps straight from the decoded uint8 frame
// pool through the per-clip descriptors.
// =====================================================================================================================
namespace rsp {

struct ClipGeom {
  int T, Hs, Ws, S;
  float mean[3], inv_std_unused[3], stdv[3];
};

// ColorJitter state of one clip (transforms_tensor.py:54-145): the four factors and the order in which the adjustments
// are applied (op ids: 0 brightness, 1 contrast, 2 saturation, 3 hue, 255 = none).
struct ClipJitter {
  float factor[4];
  uint8_t order[4];
};

__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }
__device__ __forceinline__ float luma(const float (&v)[3]) { return 0.2989f * v[0] + 0.5870f * v[1] + 0.1140f * v[2]; }

// functional_tensor.py:253-415: RGB -> HSV, h = (h + shift) mod 1, HSV -> RGB
__device__ __forceinline__ void hue_shift(float (&v)[3], float shift) {
  const float r = v[0], g = v[1], b = v[2];
  const float maxc = fmaxf(r, fmaxf(g, b)), minc = fminf(r, fminf(g, b));
  const float delta = maxc - minc;
  const float sat = maxc == 0.f ? 0.f : delta / maxc;
  float h;
  if (delta == 0.f) h = 0.f;
  else if (r == maxc) h = (g - b) / delta;           // torch.max returns the first maximal channel
  else if (g == maxc) h = (b - r) / delta + 2.0f;
  else h = (r - g) / delta + 4.0f;
  h = h / 6.0f;
  h = h - floorf(h);                                  // python-style mod 1
  h = h + shift;
  h = h - floorf(h);
  const float h6 = h * 6.f;
  const float hi = floorf(h6);
  const float f = h6 - hi;
  const float val = maxc;
  const float tt = val * (1.f - (1.f - f) * sat), pp = val * (1.f - sat), qq = val * (1.f - f * sat);
  int idx = static_cast<int>(hi) % 6;
  if (idx < 0) idx += 6;
  // channel_map of functional_tensor.py:293-297 as selects (a table indexed per pixel would live in local memory)
  v[0] = (idx == 0 || idx == 5) ? val : idx == 1 ? qq : idx == 4 ? tt : pp;
  v[1] = (idx == 1 || idx == 2) ? val : idx == 0 ? tt : idx == 3 ? qq : pp;
  v[2] = (idx == 3 || idx == 4) ? val : idx == 2 ? tt : idx == 5 ? qq : pp;
}

// ToTensorVideo's x / 255 for a byte, correctly rounded (== the IEEE quotient float(b) / 255.0f for all 256 values, checked
// exhaustively): one multiply by fl(1/255) and one FMA refinement step.  A 256-entry shared-memory table costs a 3-4-way
// bank-conflicted LDS per tap instead (the sampler was bound by its LSU wavefronts).
__device__ __forceinline__ float unit255(uint8_t b) {
  const float f = __uint_as_float(0x4B000000u | b) - 8388608.0f;   // exact u8 -> float without the conversion pipe
  const float k = __uint_as_float(0x3B808081u);                     // fl(1 / 255)
  const float q = f * k;
  const float r = fmaf(-q, 255.0f, f);
  return fmaf(r, k, q);
}

// Resized (bilinear, align_corners=False), optionally gray-scaled pixel of the cropped frame, in [0, 1].
__device__ __forceinline__ void clip_pixel(const uint8_t* __restrict__ frames, const int32_t* __restrict__ frame_idx,
                                           const int32_t* __restrict__ box, uint8_t fl, const ClipGeom& g, int clip, int t,
                                           int y, int xs, const float* __restrict__ lut, float (&v)[3]) {
  const int bi = box[clip * 4 + 0], bj = box[clip * 4 + 1], bh = box[clip * 4 + 2], bw = box[clip * 4 + 3];
  // torch upsample_bilinear2d, align_corners=False: src = scale*(dst+0.5)-0.5 clamped at 0
  const float sh = static_cast<float>(bh) / g.S, sw = static_cast<float>(bw) / g.S;
  float fy = sh * (y + 0.5f) - 0.5f, fx = sw * (xs + 0.5f) - 0.5f;
  fy = fy < 0.f ? 0.f : fy;
  fx = fx < 0.f ? 0.f : fx;
  const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
  const int y1 = y0 + (y0 < bh - 1 ? 1 : 0), x1 = x0 + (x0 < bw - 1 ? 1 : 0);
  const float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
  const uint8_t* f = frames + static_cast<size_t>(frame_idx[clip * g.T + t]) * g.Hs * g.Ws * 3;
  const uint8_t* p00 = f + (static_cast<size_t>(bi + y0) * g.Ws + bj + x0) * 3;
  const uint8_t* p01 = f + (static_cast<size_t>(bi + y0) * g.Ws + bj + x1) * 3;
  const uint8_t* p10 = f + (static_cast<size_t>(bi + y1) * g.Ws + bj + x0) * 3;
  const uint8_t* p11 = f + (static_cast<size_t>(bi + y1) * g.Ws + bj + x1) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float a = unit255(p00[c]), b = unit255(p01[c]), cc = unit255(p10[c]), d = unit255(p11[c]);
    v[c] = hy * (hx * a + lx * b) + ly * (hx * cc + lx * d);
  }
  if (fl & 2) {  // RandomGrayScale: ITU-R 601-2 luma, replicated on the three channels
    const float gray = luma(v);
    v[0] = v[1] = v[2] = gray;
  }
}

// Four horizontally adjacent output pixels x .. x+3 of row y (x % 4 == 0): the row terms (source rows, vertical weights,
// frame pointer) are computed once and the 48 byte loads of the quad are independent of each other — the one-pixel-per-
// thread form is bound by the latency of its dependent chain (index -> pointer -> bytes -> table), not by bandwidth.
// Per pixel the arithmetic is exactly clip_pixel's.
__device__ __forceinline__ void clip_pixels4(const uint8_t* __restrict__ frames, const int32_t* __restrict__ frame_idx,
                                             const int32_t* __restrict__ box, uint8_t fl, const ClipGeom& g, int clip, int t,
                                             int y, int x, const float* __restrict__ lut, float (&v)[4][3]) {
  const int bi = box[clip * 4 + 0], bj = box[clip * 4 + 1], bh = box[clip * 4 + 2], bw = box[clip * 4 + 3];
  const float sh = static_cast<float>(bh) / g.S, sw = static_cast<float>(bw) / g.S;
  float fy = sh * (y + 0.5f) - 0.5f;
  fy = fy < 0.f ? 0.f : fy;
  const int y0 = static_cast<int>(fy);
  const int y1 = y0 + (y0 < bh - 1 ? 1 : 0);
  const float ly = fy - y0, hy = 1.f - ly;
  const uint8_t* f = frames + static_cast<size_t>(frame_idx[clip * g.T + t]) * g.Hs * g.Ws * 3;
  const uint8_t* r0 = f + (static_cast<size_t>(bi + y0) * g.Ws + bj) * 3;
  const uint8_t* r1 = f + (static_cast<size_t>(bi + y1) * g.Ws + bj) * 3;
  uint8_t b00[4][3], b01[4][3], b10[4][3], b11[4][3];
  float lx[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int xs = (fl & 1) ? (g.S - 1 - (x + j)) : (x + j);   // horizontal flip acts on the resized clip
    float fx = sw * (xs + 0.5f) - 0.5f;
    fx = fx < 0.f ? 0.f : fx;
    const int x0 = static_cast<int>(fx);
    const int x1 = x0 + (x0 < bw - 1 ? 1 : 0);
    lx[j] = fx - x0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      b00[j][c] = r0[x0 * 3 + c];
      b01[j][c] = r0[x1 * 3 + c];
      b10[j][c] = r1[x0 * 3 + c];
      b11[j][c] = r1[x1 * 3 + c];
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float hx = 1.f - lx[j];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float a = unit255(b00[j][c]), b = unit255(b01[j][c]), cc = unit255(b10[j][c]), d = unit255(b11[j][c]);
      v[j][c] = hy * (hx * a + lx[j] * b) + ly * (hx * cc + lx[j] * d);
    }
    if (fl & 2) {
      const float gray = luma(v[j]);
      v[j][0] = v[j][1] = v[j][2] = gray;
    }
  }
}

// Applies the jitter ops order[first .. last) to one pixel; `mean` is the clip-wide gray mean the contrast op blends with.
__device__ __forceinline__ void jitter_ops(float (&v)[3], const ClipJitter& jt, int first, int last, float mean) {
  uint32_t order4;
  memcpy(&order4, jt.order, 4);
  const float f0 = jt.factor[0], f1 = jt.factor[1], f2 = jt.factor[2], f3 = jt.factor[3];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k < first || k >= last) continue;
    const int op = (order4 >> (8 * k)) & 0xff;
    const float fac = op == 0 ? f0 : op == 1 ? f1 : op == 2 ? f2 : f3;
    if (op == 0) {                     // adjust_brightness: blend with black
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = clamp01(fac * v[c] + (1.f - fac) * 0.f);
    } else if (op == 1) {              // adjust_contrast: blend with the mean gray level of the whole clip
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = clamp01(fac * v[c] + (1.f - fac) * mean);
    } else if (op == 2) {              // adjust_saturation: blend with the pixel's gray level
      const float gray = luma(v);
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = clamp01(fac * v[c] + (1.f - fac) * gray);
    } else if (op == 3) {
      hue_shift(v, fac);
    }
  }
}

// Pass 1 (only when a contrast op is present): sum over the clip of the gray level right before the contrast op.
__global__ void __launch_bounds__(256) clip_gray_sum_kernel(const uint8_t* __restrict__ frames,
                                                            const int32_t* __restrict__ frame_idx,
                                                            const int32_t* __restrict__ box,
                                                            const uint8_t* __restrict__ flags,
                                                            const ClipJitter* __restrict__ jitter, ClipGeom g,
                                                            float* __restrict__ sums, int per_clip_blocks) {
  const float* lut = nullptr;   // (the x / 255 table is gone: unit255)
  const int clip = blockIdx.x / per_clip_blocks, blk = blockIdx.x - clip * per_clip_blocks;
  const ClipJitter jt = jitter[clip];
  int cpos = -1;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (jt.order[k] == 1) cpos = k;
  float acc = 0.f;
  if (cpos >= 0) {
    const int per = g.T * g.S * g.S;
    const uint8_t fl = flags[clip];
    for (int i = blk * 256 + threadIdx.x; i < per; i += per_clip_blocks * 256) {
      const int x = i % g.S, y = (i / g.S) % g.S, t = i / (g.S * g.S);
      float v[3];
      clip_pixel(frames, frame_idx, box, fl, g, clip, t, y, x, lut, v);   // the mean does not depend on the flip
      jitter_ops(v, jt, 0, cpos, 0.f);
      acc += luma(v);
    }
  }
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && cpos >= 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    atomicAdd(sums + clip, s);
  }
}

// fp32 NCDHW output with colour jitter, two passes IN PLACE (the gray-sum pass above re-gathers every pixel — 12 byte loads,
// 12 table look-ups, bilinear blend — only to throw it away):
//   pass A: gather + the jitter ops in front of the contrast op, store the pixel into the output tensor, sum its gray level;
//   pass B: stream over the output tensor: remaining ops (they need the clip-wide gray mean) + normalise, in place.
// Same arithmetic on the same fp32 values as the single-pass kernel, so the pixels are bit-identical to it.
__global__ void __launch_bounds__(256) clip_sample_pre_kernel(const uint8_t* __restrict__ frames,
                                                              const int32_t* __restrict__ frame_idx,
                                                              const int32_t* __restrict__ box,
                                                              const uint8_t* __restrict__ flags,
                                                              const ClipJitter* __restrict__ jitter, ClipGeom g,
                                                              float* __restrict__ sums, float* __restrict__ out,
                                                              int per_clip_blocks) {
  const float* lut = nullptr;   // (the x / 255 table is gone: unit255)
  const int clip = blockIdx.x / per_clip_blocks, blk = blockIdx.x - clip * per_clip_blocks;
  const ClipJitter jt = jitter[clip];
  int split = 4;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (jt.order[k] == 1) split = k;
  const int per = g.T * g.S * g.S;
  const uint32_t plane2 = static_cast<uint32_t>(g.S) * g.S;
  const uint8_t fl = flags[clip];
  float* oc = out + static_cast<size_t>(clip) * 3 * per;
  float acc = 0.f;
  if ((g.S & 3) == 0) {
    for (int q = blk * 256 + threadIdx.x; q < (per >> 2); q += per_clip_blocks * 256) {
      const uint32_t i = static_cast<uint32_t>(q) * 4u;
      const uint32_t t = i / plane2, rem = i - t * plane2;
      const int y = static_cast<int>(rem / g.S), x = static_cast<int>(rem - (rem / g.S) * g.S);
      float v[4][3];
      clip_pixels4(frames, frame_idx, box, fl, g, clip, static_cast<int>(t), y, x, lut, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        jitter_ops(v[j], jt, 0, split, 0.f);
        acc += luma(v[j]);
      }
#pragma unroll
      for (int c = 0; c < 3; ++c)
        *reinterpret_cast<float4*>(oc + i + static_cast<size_t>(c) * per) = make_float4(v[0][c], v[1][c], v[2][c], v[3][c]);
    }
  } else {
    for (int i = blk * 256 + threadIdx.x; i < per; i += per_clip_blocks * 256) {
      const uint32_t t = static_cast<uint32_t>(i) / plane2, rem = static_cast<uint32_t>(i) - t * plane2;
      const int y = static_cast<int>(rem / g.S), x = static_cast<int>(rem - (rem / g.S) * g.S);
      const int xs = (fl & 1) ? (g.S - 1 - x) : x;                 // horizontal flip acts on the resized clip
      float v[3];
      clip_pixel(frames, frame_idx, box, fl, g, clip, static_cast<int>(t), y, xs, lut, v);
      jitter_ops(v, jt, 0, split, 0.f);
      acc += luma(v);
      oc[i] = v[0];
      oc[i + per] = v[1];
      oc[i + 2 * per] = v[2];
    }
  }
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && split < 4) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    atomicAdd(sums + clip, s);
  }
}

__global__ void __launch_bounds__(256) clip_sample_post_kernel(const ClipJitter* __restrict__ jitter,
                                                               const float* __restrict__ gray_sums, ClipGeom g,
                                                               float* __restrict__ out, int per_clip_blocks) {
  const int clip = blockIdx.x / per_clip_blocks, blk = blockIdx.x - clip * per_clip_blocks;
  const ClipJitter jt = jitter[clip];
  int split = 4;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (jt.order[k] == 1) split = k;
  const int per = g.T * g.S * g.S;
  const float mean = gray_sums[clip] / (static_cast<float>(g.T) * g.S * g.S);
  float* oc = out + static_cast<size_t>(clip) * 3 * per;
  if ((per & 3) == 0) {
    for (int q = blk * 256 + threadIdx.x; q < (per >> 2); q += per_clip_blocks * 256) {
      float4* p0 = reinterpret_cast<float4*>(oc) + q;
      float4* p1 = reinterpret_cast<float4*>(oc + per) + q;
      float4* p2 = reinterpret_cast<float4*>(oc + 2 * static_cast<size_t>(per)) + q;
      const float4 a = *p0, b = *p1, c4 = *p2;
      float v[4][3] = {{a.x, b.x, c4.x}, {a.y, b.y, c4.y}, {a.z, b.z, c4.z}, {a.w, b.w, c4.w}};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        jitter_ops(v[j], jt, split, 4, mean);
#pragma unroll
        for (int c = 0; c < 3; ++c) v[j][c] = (v[j][c] - g.mean[c]) / g.stdv[c];
      }
      *p0 = make_float4(v[0][0], v[1][0], v[2][0], v[3][0]);
      *p1 = make_float4(v[0][1], v[1][1], v[2][1], v[3][1]);
      *p2 = make_float4(v[0][2], v[1][2], v[2][2], v[3][2]);
    }
    return;
  }
  for (int i = blk * 256 + threadIdx.x; i < per; i += per_clip_blocks * 256) {
    float v[3] = {oc[i], oc[i + per], oc[i + 2 * per]};
    jitter_ops(v, jt, split, 4, mean);
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = (v[c] - g.mean[c]) / g.stdv[c];
    oc[i] = v[0];
    oc[i + per] = v[1];
    oc[i + 2 * per] = v[2];
  }
}

// fp32 NCDHW output without colour jitter, four pixels per thread (S % 4 == 0).
__global__ void __launch_bounds__(256) clip_sample4_kernel(const uint8_t* __restrict__ frames,
                                                           const int32_t* __restrict__ frame_idx,
                                                           const int32_t* __restrict__ box,
                                                           const uint8_t* __restrict__ flags, ClipGeom g,
                                                           float* __restrict__ out, size_t total4) {
  const float* lut = nullptr;   // (the x / 255 table is gone: unit255)
  const uint32_t plane2 = static_cast<uint32_t>(g.S) * g.S;
  const size_t per = static_cast<size_t>(g.T) * plane2;
  for (size_t q = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x; q < total4; q += static_cast<size_t>(gridDim.x) * 256) {
    const uint32_t i32 = static_cast<uint32_t>(q) * 4u;            // total < 2^32 is checked by the host
    const uint32_t ct = i32 / plane2;                              // clip * T + t
    const uint32_t rem = i32 - ct * plane2;
    const int y = static_cast<int>(rem / g.S), x = static_cast<int>(rem - (rem / g.S) * g.S);
    const int clip = static_cast<int>(ct / g.T), t = static_cast<int>(ct - (ct / g.T) * g.T);
    float v[4][3];
    clip_pixels4(frames, frame_idx, box, flags[clip], g, clip, t, y, x, lut, v);
    const size_t o = (static_cast<size_t>(clip) * 3 * g.T + t) * plane2 + rem;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      *reinterpret_cast<float4*>(out + o + c * per) =
          make_float4((v[0][c] - g.mean[c]) / g.stdv[c], (v[1][c] - g.mean[c]) / g.stdv[c],
                      (v[2][c] - g.mean[c]) / g.stdv[c], (v[3][c] - g.mean[c]) / g.stdv[c]);
  }
}

template <int LAYOUT>
__global__ void __launch_bounds__(256) clip_sample_kernel(const uint8_t* __restrict__ frames,
                                                          const int32_t* __restrict__ frame_idx,
                                                          const int32_t* __restrict__ box,
                                                          const uint8_t* __restrict__ flags,
                                                          const ClipJitter* __restrict__ jitter,
                                                          const float* __restrict__ gray_sums, ClipGeom g,
                                                          void* __restrict__ outv, size_t total) {
  const float* lut = nullptr;   // (the x / 255 table is gone: unit255)
  for (size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * 256) {
    // 32-bit index arithmetic (64-bit divisions cost ~100 instructions each)
    const uint32_t plane = static_cast<uint32_t>(g.S) * g.S;
    const uint32_t i32 = static_cast<uint32_t>(i);                // total < 2^32 is checked by the host
    const uint32_t ct = i32 / plane;                               // clip * T + t
    const uint32_t rem = i32 - ct * plane;
    const int y = static_cast<int>(rem / g.S), x = static_cast<int>(rem - (rem / g.S) * g.S);
    const int clip = static_cast<int>(ct / g.T), t = static_cast<int>(ct - (ct / g.T) * g.T);
    const uint8_t fl = flags[clip];
    const int xs = (fl & 1) ? (g.S - 1 - x) : x;                 // horizontal flip acts on the resized clip
    float v[3];
    clip_pixel(frames, frame_idx, box, fl, g, clip, t, y, xs, lut, v);
    if (jitter) {
      const float mean = gray_sums[clip] / (static_cast<float>(g.T) * g.S * g.S);
      jitter_ops(v, jitter[clip], 0, 4, mean);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = (v[c] - g.mean[c]) / g.stdv[c];
    if (LAYOUT == 0) {
      float* out = static_cast<float*>(outv);
      const size_t plane = static_cast<size_t>(g.T) * g.S * g.S;
      const size_t o = (static_cast<size_t>(clip) * 3 * g.T + t) * g.S * g.S + static_cast<size_t>(y) * g.S + x;
      out[o] = v[0];
      out[o + plane] = v[1];
      out[o + 2 * plane] = v[2];
    } else {
      uint2 w;
      w.x = pack_bf16x2(v[0], v[1]);
      w.y = pack_bf16x2(v[2], 0.f);
      static_cast<uint2*>(outv)[i] = w;
    }
  }
}

}  // namespace rsp

static int clip_sample_impl(const uint8_t* frames, const int32_t* frame_idx, const int32_t* box, const uint8_t* flags,
                            const void* jitter, float* gray_sums, const float* mean3, const float* std3, int32_t n_clips,
                            int32_t T, int32_t Hs, int32_t Ws, int32_t S, int32_t layout, void* out, void* stream) {
  RSP_REQUIRE(n_clips >= 0 && T > 0 && Hs > 0 && Ws > 0 && S > 0, "clip_sample: bad sizes");
  RSP_REQUIRE(layout == 0 || layout == 1, "clip_sample: layout must be 0 (fp32 NCDHW) or 1 (bf16 NDHWC4)");
  RSP_REQUIRE((jitter == nullptr) == (gray_sums == nullptr), "clip_sample: jitter and gray_sums go together");
  size_t total = static_cast<size_t>(n_clips) * T * S * S;
  if (total == 0) return rsp::RSP_OK;
  RSP_REQUIRE(total < (1ull << 32), "clip_sample: more than 2^32 output pixels in one call");
  rsp::ClipGeom g{};
  g.T = T; g.Hs = Hs; g.Ws = Ws; g.S = S;
  for (int c = 0; c < 3; ++c) {
    g.mean[c] = mean3[c];
    g.stdv[c] = std3[c];
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const rsp::ClipJitter* jt = static_cast<const rsp::ClipJitter*>(jitter);
  if (jt) {
    if (rsp::zero_async(gray_sums, sizeof(float) * n_clips, st) != cudaSuccess) {
      rsp::set_error("clip_sample: memset failed");
      return rsp::RSP_ERR_CUDA;
    }
    if (layout == 0) {   // fp32 output: two passes in place (see clip_sample_pre_kernel)
      int pcb = (T * S * S + 256 * 4 - 1) / (256 * 4);
      if (pcb < 1) pcb = 1;
      if (pcb > 512) pcb = 512;
      rsp::clip_sample_pre_kernel<<<n_clips * pcb, 256, 0, st>>>(frames, frame_idx, box, flags, jt, g, gray_sums,
                                                                 static_cast<float*>(out), pcb);
      int rc = rsp::check_launch("clip_sample_pre");
      if (rc != rsp::RSP_OK) return rc;
      rsp::clip_sample_post_kernel<<<n_clips * pcb, 256, 0, st>>>(jt, gray_sums, g, static_cast<float*>(out), pcb);
      return rsp::check_launch("clip_sample_post");
    }
    int per_clip_blocks = (T * S * S + 256 * 8 - 1) / (256 * 8);
    if (per_clip_blocks < 1) per_clip_blocks = 1;
    if (per_clip_blocks > 64) per_clip_blocks = 64;
    rsp::clip_gray_sum_kernel<<<n_clips * per_clip_blocks, 256, 0, st>>>(frames, frame_idx, box, flags, jt, g, gray_sums,
                                                                         per_clip_blocks);
    int rc = rsp::check_launch("clip_gray_sum");
    if (rc != rsp::RSP_OK) return rc;
  }
  unsigned grid = rsp::ew_grid(total, 256);
  if (layout == 0 && !jt && (S & 3) == 0)
    rsp::clip_sample4_kernel<<<rsp::ew_grid(total / 4, 256), 256, 0, st>>>(frames, frame_idx, box, flags, g,
                                                                          static_cast<float*>(out), total / 4);
  else if (layout == 0)
    rsp::clip_sample_kernel<0><<<grid, 256, 0, st>>>(frames, frame_idx, box, flags, jt, gray_sums, g, out, total);
  else
    rsp::clip_sample_kernel<1><<<grid, 256, 0, st>>>(frames, frame_idx, box, flags, jt, gray_sums, g, out, total);
  return rsp::check_launch("clip_sample");
}

extern "C" int rsp_clip_sample(const uint8_t* frames, const int32_t* frame_idx, const int32_t* box,
                               const uint8_t* flags, const float* mean3, const float* std3, int32_t n_clips, int32_t T,
                               int32_t Hs, int32_t Ws, int32_t S, int32_t layout, void* out, void* stream) {
  return clip_sample_impl(frames, frame_idx, box, flags, nullptr, nullptr, mean3, std3, n_clips, T, Hs, Ws, S, layout, out,
                          stream);
}

extern "C" int rsp_clip_sample_jitter(const uint8_t* frames, const int32_t* frame_idx, const int32_t* box,
                                      const uint8_t* flags, const void* jitter, float* gray_sums, const float* mean3,
                                      const float* std3, int32_t n_clips, int32_t T, int32_t Hs, int32_t Ws, int32_t S,
                                      int32_t layout, void* out, void* stream) {
  RSP_REQUIRE(jitter && gray_sums, "clip_sample_jitter: jitter table and gray_sums workspace are required");
  return clip_sample_impl(frames, frame_idx, box, flags, jitter, gray_sums, mean3, std3, n_clips, T, Hs, Ws, S, layout, out,
                          stream);
}

// =====================================================================================================================
// S3D-G pieces (reference: models/s3dg.py): self-gating of sep_conv (:54-72: AdaptiveAvgPool3d(1) -> 1x1x1 conv with
// bias -> sigmoid -> broadcast multiply) and the inception channel concat (:96, torch.cat(dim=1)).
// Activations bf16 [N][S][C] (S = T*H*W positions, C stored channels); gate statistics fp32 [N][C].
// =====================================================================================================================
namespace rsp {

// MODE 0: out[n][c] = sum_s x[n][s][c];  MODE 1: out[n][c] = sum_s x[n][s][c] * y[n][s][c]
// grid: (chunks over S, N); each thread owns one 8-channel group and strides over positions; fp32 atomics per block.
template <int MODE>
__global__ void __launch_bounds__(256) sample_channel_sum_kernel(const uint4* __restrict__ x,
                                                                 const uint4* __restrict__ y, int S, int C,
                                                                 float* __restrict__ out) {
  const int G = C >> 3;
  const int rows_per_iter = 256 / G;
  const int g = threadIdx.x % G, rsub = threadIdx.x / G;
  const bool active = rsub < rows_per_iter;
  const int n = blockIdx.y;
  const uint4* xb = x + static_cast<size_t>(n) * S * G;
  const uint4* yb = MODE ? y + static_cast<size_t>(n) * S * G : nullptr;
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 0.f;
  for (int s = blockIdx.x * rows_per_iter + rsub; active && s < S; s += gridDim.x * rows_per_iter) {
    float xv[8], yv[8];
    unpack8(__ldg(xb + static_cast<size_t>(s) * G + g), xv);
    if (MODE) unpack8(__ldg(yb + static_cast<size_t>(s) * G + g), yv);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] += MODE ? xv[i] * yv[i] : xv[i];
  }
  __shared__ float sa[256 * 8];
#pragma unroll
  for (int i = 0; i < 8; ++i) sa[threadIdx.x * 8 + i] = a[i];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float s0 = 0.f;
    for (int r = 0; r < rows_per_iter; ++r) s0 += sa[(r * G + (c >> 3)) * 8 + (c & 7)];
    atomicAdd(out + static_cast<size_t>(n) * C + c, s0);
  }
}

// gate[n][c] = sigmoid(b[c] + sum_k W[c][k] * mean[n][k]);  block per sample
__global__ void __launch_bounds__(256) gate_linear_fwd_kernel(const float* __restrict__ sums, float inv_s,
                                                              const float* __restrict__ w, const float* __restrict__ b,
                                                              int C, int Cl, float* __restrict__ pooled,
                                                              float* __restrict__ gate) {
  extern __shared__ float sp[];
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < Cl; c += 256) {
    float m = sums[static_cast<size_t>(n) * C + c] * inv_s;
    sp[c] = m;
    pooled[static_cast<size_t>(n) * Cl + c] = m;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float g = 0.f;
    if (c < Cl) {
      float z = b[c];
      const float* wr = w + static_cast<size_t>(c) * Cl;
      for (int k = 0; k < Cl; ++k) z = fmaf(wr[k], sp[k], z);
      g = 1.f / (1.f + __expf(-z));
    }
    gate[static_cast<size_t>(n) * C + c] = g;
  }
}

// y = x * gate[n][c]   (MODE 0)        dx = dy * gate[n][c] + add[n][c]   (MODE 1)
template <int MODE>
__global__ void __launch_bounds__(256) gate_scale_kernel(const uint4* __restrict__ x, const float* __restrict__ gate,
                                                         const float* __restrict__ add, uint4* __restrict__ out,
                                                         size_t nvec, int S, int C) {
  const int G = C >> 3;
  for (size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x; i < nvec;
       i += static_cast<size_t>(gridDim.x) * 256) {
    const int g = static_cast<int>(i % G);
    const size_t n = i / (static_cast<size_t>(S) * G);
    const float* gr = gate + n * C + g * 8;
    float v[8];
    unpack8(__ldg(x + i), v);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = MODE ? fmaf(v[e], gr[e], add[n * C + g * 8 + e]) : v[e] * gr[e];
    out[i] = pack8(v);
  }
}

// backward of the gate linear: dz = dgate * g * (1-g); dW += dz^T pooled; db += dz; dmean_over_S[n][k] = (dz W)[k] / S
__global__ void __launch_bounds__(256) gate_linear_bwd_kernel(const float* __restrict__ dgate,
                                                              const float* __restrict__ gate,
                                                              const float* __restrict__ pooled,
                                                              const float* __restrict__ w, int C, int Cl, float inv_s,
                                                              float* __restrict__ dw, float* __restrict__ db,
                                                              float* __restrict__ dadd) {
  extern __shared__ float sm[];
  float* dz = sm;        // [Cl]
  float* sp = sm + Cl;   // [Cl]
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < Cl; c += 256) {
    float g = gate[static_cast<size_t>(n) * C + c];
    dz[c] = dgate[static_cast<size_t>(n) * C + c] * g * (1.f - g);
    sp[c] = pooled[static_cast<size_t>(n) * Cl + c];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Cl; c += 256) atomicAdd(db + c, dz[c]);
  for (int i = threadIdx.x; i < Cl * Cl; i += 256) {
    int c = i / Cl, k = i - c * Cl;
    atomicAdd(dw + i, dz[c] * sp[k]);
  }
  for (int k = threadIdx.x; k < C; k += 256) {
    float a = 0.f;
    if (k < Cl)
      for (int c = 0; c < Cl; ++c) a = fmaf(dz[c], w[static_cast<size_t>(c) * Cl + k], a);
    dadd[static_cast<size_t>(n) * C + k] = a * inv_s;
  }
}

// dst[m][dst_off + j] = src[m][src_off + j], j < n_ch, 8-channel (16 B) granularity
__global__ void __launch_bounds__(256) copy_channels_kernel(const uint4* __restrict__ src, int src_g, int src_off_g,
                                                            uint4* __restrict__ dst, int dst_g, int dst_off_g,
                                                            int n_g, size_t total) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * 256) {
    const size_t m = i / n_g;
    const int j = static_cast<int>(i - m * n_g);
    dst[m * dst_g + dst_off_g + j] = __ldg(src + m * src_g + src_off_g + j);
  }
}

}  // namespace rsp

extern "C" {

int rsp_gate_fwd(const void* x, int32_t N, int32_t S, int32_t C, int32_t C_logical, const float* w, const float* b,
                 float* sums_ws, float* pooled, float* gate, void* y, void* stream_) {
  using namespace rsp;
  int rc = check_c(C, "gate_fwd");
  if (rc != RSP_OK) return rc;
  RSP_REQUIRE(N > 0 && S > 0 && C_logical <= C && C_logical * sizeof(float) <= 48 * 1024, "gate_fwd: bad sizes");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  cudaError_t e = rsp::zero_async(sums_ws, static_cast<size_t>(N) * C * sizeof(float), stream);
  if (e != cudaSuccess) {
    set_error("gate_fwd memset: %s", cudaGetErrorString(e));
    return RSP_ERR_CUDA;
  }
  const int rows_per_iter = 256 / (C / 8);
  int gx = (S + rows_per_iter * 8 - 1) / (rows_per_iter * 8);
  if (gx < 1) gx = 1;
  if (gx > 64) gx = 64;
  sample_channel_sum_kernel<0><<<dim3(gx, N), 256, 0, stream>>>(static_cast<const uint4*>(x), nullptr, S, C, sums_ws);
  gate_linear_fwd_kernel<<<N, 256, C_logical * sizeof(float), stream>>>(sums_ws, 1.f / S, w, b, C, C_logical, pooled,
                                                                        gate);
  size_t nvec = static_cast<size_t>(N) * S * (C / 8);
  gate_scale_kernel<0><<<ew_grid(nvec, 256), 256, 0, stream>>>(static_cast<const uint4*>(x), gate, nullptr,
                                                               static_cast<uint4*>(y), nvec, S, C);
  return check_launch("gate_fwd");
}

int rsp_gate_bwd(const void* dy, const void* x, int32_t N, int32_t S, int32_t C, int32_t C_logical, const float* w,
                 const float* pooled, const float* gate, float* dgate_ws, float* dadd_ws, float* dw, float* db,
                 void* dx, void* stream_) {
  using namespace rsp;
  int rc = check_c(C, "gate_bwd");
  if (rc != RSP_OK) return rc;
  RSP_REQUIRE(N > 0 && S > 0 && C_logical <= C && 2 * C_logical * sizeof(float) <= 48 * 1024, "gate_bwd: bad sizes");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  cudaError_t e = rsp::zero_async(dgate_ws, static_cast<size_t>(N) * C * sizeof(float), stream);
  if (e != cudaSuccess) {
    set_error("gate_bwd memset: %s", cudaGetErrorString(e));
    return RSP_ERR_CUDA;
  }
  const int rows_per_iter = 256 / (C / 8);
  int gx = (S + rows_per_iter * 8 - 1) / (rows_per_iter * 8);
  if (gx < 1) gx = 1;
  if (gx > 64) gx = 64;
  sample_channel_sum_kernel<1><<<dim3(gx, N), 256, 0, stream>>>(static_cast<const uint4*>(dy),
                                                                 static_cast<const uint4*>(x), S, C, dgate_ws);
  gate_linear_bwd_kernel<<<N, 256, 2 * C_logical * sizeof(float), stream>>>(dgate_ws, gate, pooled, w, C, C_logical,
                                                                            1.f / S, dw, db, dadd_ws);
  size_t nvec = static_cast<size_t>(N) * S * (C / 8);
  gate_scale_kernel<1><<<ew_grid(nvec, 256), 256, 0, stream>>>(static_cast<const uint4*>(dy), gate, dadd_ws,
                                                               static_cast<uint4*>(dx), nvec, S, C);
  return check_launch("gate_bwd");
}

int rsp_copy_channels(const void* src, int32_t c_src, int32_t src_off, void* dst, int32_t c_dst, int32_t dst_off,
                      int32_t n_ch, int64_t M, void* stream) {
  using namespace rsp;
  RSP_REQUIRE(c_src % 8 == 0 && c_dst % 8 == 0 && src_off % 8 == 0 && dst_off % 8 == 0 && n_ch % 8 == 0 &&
                  src_off + n_ch <= c_src && dst_off + n_ch <= c_dst,
              "copy_channels: channel ranges must be multiples of 8 and in range");
  size_t total = static_cast<size_t>(M) * (n_ch / 8);
  if (total == 0) return RSP_OK;
  copy_channels_kernel<<<ew_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(src), c_src / 8, src_off / 8, static_cast<uint4*>(dst), c_dst / 8, dst_off / 8,
      n_ch / 8, total);
  return check_launch("copy_channels");
}

}  // extern "C"
