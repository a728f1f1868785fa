// HBM-bound kernels of the MoCo / relative-speed objective
// (reference: moco/builder_diffspeed_diffloss.py; line numbers cited per kernel).
#include "common.cuh"
#include "rspnet_b200.h"

namespace rsp {

int device_sm_count();

static inline unsigned grid_for(size_t work_items, int block, int per_sm = 8) {
  size_t need = (work_items + block - 1) / block;
  size_t cap = static_cast<size_t>(device_sm_count()) * per_sm;
  if (need < 1) need = 1;
  return static_cast<unsigned>(need < cap ? need : cap);
}

// ------------------------------------------------------------------------------------------------
// _momentum_update_key_encoder (:337-343): k = k*m + q*(1-m), one pass over the flat parameter buffer.
// Rounding follows the reference expression: two fp32 products, then one fp32 add (no FMA contraction).
// 12 algorithmic bytes per parameter.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ema_kernel(float* __restrict__ k, const float* __restrict__ q, size_t n4,
                                                  size_t n, float m, float om) {
  size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  float4* k4 = reinterpret_cast<float4*>(k);
  const float4* q4 = reinterpret_cast<const float4*>(q);
  for (size_t j = i; j < n4; j += stride) {
    float4 a = k4[j];
    float4 b = __ldg(q4 + j);
    a.x = __fadd_rn(__fmul_rn(a.x, m), __fmul_rn(b.x, om));
    a.y = __fadd_rn(__fmul_rn(a.y, m), __fmul_rn(b.y, om));
    a.z = __fadd_rn(__fmul_rn(a.z, m), __fmul_rn(b.z, om));
    a.w = __fadd_rn(__fmul_rn(a.w, m), __fmul_rn(b.w, om));
    k4[j] = a;
  }
  for (size_t j = n4 * 4 + i; j < n; j += stride) k[j] = __fadd_rn(__fmul_rn(k[j], m), __fmul_rn(q[j], om));
}

// ------------------------------------------------------------------------------------------------
// SGD with momentum and weight decay over flat buffers (pretrain.py:65-72,163-165; torch.optim.SGD math):
//   d = g*grad_scale + wd*p ; buf = first ? d : mom*buf + d ; p -= lr*buf.     16 B read + 8 B written / param
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sgd_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                  float* __restrict__ buf, size_t n4, size_t n, float lr, float mom,
                                                  float wd, float gs, int first) {
  size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* b4 = reinterpret_cast<float4*>(buf);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  auto upd = [&](float& pv, float gv, float& bv) {
    float d = gv * gs + wd * pv;
    bv = first ? d : mom * bv + d;
    pv = pv - lr * bv;
  };
  for (size_t j = i; j < n4; j += stride) {
    float4 pv = p4[j], gv = __ldg(g4 + j), bv = first ? make_float4(0, 0, 0, 0) : b4[j];
    upd(pv.x, gv.x, bv.x);
    upd(pv.y, gv.y, bv.y);
    upd(pv.z, gv.z, bv.z);
    upd(pv.w, gv.w, bv.w);
    p4[j] = pv;
    b4[j] = bv;
  }
  for (size_t j = n4 * 4 + i; j < n; j += stride) {
    float bv = first ? 0.f : buf[j];
    upd(p[j], g[j], bv);
    buf[j] = bv;
  }
}

// ------------------------------------------------------------------------------------------------
// _diff_speed (:421-447): temporal re-sampling of (im_q, im_k) into (q, k, k_neg).
// grid: x = chunks of one frame plane, y = j*Tr + t (j indexes the permutation), z = which output.
// ------------------------------------------------------------------------------------------------
template <int LAYOUT>
__global__ void __launch_bounds__(256) speed_gather_kernel(const float* __restrict__ im_q,
                                                           const float* __restrict__ im_k,
                                                           const int64_t* __restrict__ perm, int B, int C, int T,
                                                           int HW, int n_s1, int d, int Tr, void* __restrict__ out_q,
                                                           void* __restrict__ out_k, void* __restrict__ out_kneg) {
  const int j = blockIdx.y / Tr, t = blockIdx.y % Tr;
  const int which = blockIdx.z;
  const int b = static_cast<int>(perm[j]);
  const bool s1 = j < n_s1;
  // q and k: s1 rows keep speed 1, the rest speed d; k_neg swaps the two
  const int speed = (which == 2) ? (s1 ? d : 1) : (s1 ? 1 : d);
  const float* src = (which == 0 ? im_q : im_k) + (static_cast<size_t>(b) * C * T + static_cast<size_t>(t) * speed) * HW;
  void* outv = which == 0 ? out_q : (which == 1 ? out_k : out_kneg);
  const size_t plane = static_cast<size_t>(T) * HW;  // channel stride in the source
  if (LAYOUT == 0) {
    float* out = static_cast<float*>(outv) + (static_cast<size_t>(b) * C * Tr + t) * HW;
    const size_t oplane = static_cast<size_t>(Tr) * HW;
    for (int c = 0; c < C; ++c) {
      const float* s = src + c * plane;
      float* o = out + c * oplane;
      if ((HW & 3) == 0) {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW / 4; i += gridDim.x * blockDim.x)
          reinterpret_cast<float4*>(o)[i] = __ldg(reinterpret_cast<const float4*>(s) + i);
      } else {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) o[i] = s[i];
      }
    }
  } else {
    uint2* out = static_cast<uint2*>(outv) + (static_cast<size_t>(b) * Tr + t) * HW;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
      float r = __ldg(src + i), g = __ldg(src + plane + i), bl = __ldg(src + 2 * plane + i);
      uint2 v;
      v.x = pack_bf16x2(r, g);
      v.y = pack_bf16x2(bl, 0.f);
      out[i] = v;
    }
  }
}

// dst[i] = src[index[i]], rows of row_bytes (multiple of 16)
__global__ void __launch_bounds__(256) gather_rows_kernel(const uint4* __restrict__ src,
                                                          const int64_t* __restrict__ index, uint4* __restrict__ dst,
                                                          size_t vec_per_row) {
  const size_t row = blockIdx.y;
  const uint4* s = src + static_cast<size_t>(index[row]) * vec_per_row;
  uint4* o = dst + row * vec_per_row;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < vec_per_row;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    o[i] = __ldg(s + i);
}

// ------------------------------------------------------------------------------------------------
// _dequeue_and_enqueue (:345-359): ring-buffer write of n key columns; pointer lives on the device.
// ------------------------------------------------------------------------------------------------
__global__ void enqueue_kernel(float* __restrict__ queue, const float* __restrict__ keys,
                               const int64_t* __restrict__ ptr, int D, int K, int n) {
  __shared__ float tile[32][33];
  const int p = static_cast<int>(*ptr);
  const int i0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int i = i0 + r, dd = d0 + threadIdx.x;
    tile[r][threadIdx.x] = (i < n && dd < D) ? keys[static_cast<size_t>(i) * D + dd] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int dd = d0 + r, i = i0 + threadIdx.x;
    if (dd < D && i < n) queue[static_cast<size_t>(dd) * K + p + i] = tile[threadIdx.x][r];
  }
}
__global__ void advance_ptr_kernel(int64_t* ptr, int K, int n) { *ptr = (*ptr + n) % K; }

// ------------------------------------------------------------------------------------------------
// logits (:521-536) + logsumexp for the two cross-entropies
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// one warp per row: the four positive/ranking dot products, divided by T like the reference (`/= self.T`)
__global__ void rowdots_kernel(const float* __restrict__ q_a, const float* __restrict__ q_m,
                               const float* __restrict__ k_a, const float* __restrict__ k_m,
                               const float* __restrict__ kn_a, const float* __restrict__ kn_m, int N, int D, int K,
                               float T, float* __restrict__ logits1, float* __restrict__ logits2,
                               float* __restrict__ lpos_m, float* __restrict__ lneg_m, float* __restrict__ pos1,
                               float* __restrict__ pos2) {
  int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (n >= N) return;
  float a1 = 0, a2 = 0, m1 = 0, m2 = 0;
  for (int dd = lane; dd < D; dd += 32) {
    size_t o = static_cast<size_t>(n) * D + dd;
    float qa = q_a[o], qm = q_m[o];
    a1 += qa * k_a[o];
    a2 += qa * kn_a[o];
    m1 += qm * k_m[o];
    m2 += qm * kn_m[o];
  }
  a1 = warp_sum(a1) / T;
  a2 = warp_sum(a2) / T;
  m1 = warp_sum(m1) / T;
  m2 = warp_sum(m2) / T;
  if (lane == 0) {
    pos1[n] = a1;
    pos2[n] = a2;
    lpos_m[n] = m1;
    lneg_m[n] = m2;
    if (logits1) logits1[static_cast<size_t>(n) * (K + 1)] = a1;
    if (logits2) logits2[static_cast<size_t>(n) * (K + 1)] = a2;
  }
}

constexpr int kNegCols = 256;  // queue columns per block
constexpr int kNegRows = 16;   // query rows per block

// l_neg tile: 256 queue columns x 16 rows; writes logits and per-(row, tile) (max, sum exp(l - max))
__global__ void __launch_bounds__(kNegCols) neg_logits_kernel(const float* __restrict__ q_a,
                                                              const float* __restrict__ queue, int N, int D, int K,
                                                              float T, float* __restrict__ logits1,
                                                              float* __restrict__ logits2, float* __restrict__ ws,
                                                              int tiles, const float* __restrict__ pos1,
                                                              const float* __restrict__ pos2,
                                                              int32_t* __restrict__ ranks) {
  extern __shared__ float sm[];  // [kNegRows][D] query rows, then [8][kNegRows][2] reduction scratch
  float* qs = sm;
  float* red = sm + kNegRows * D;
  const int n0 = blockIdx.y * kNegRows;
  const int k = blockIdx.x * kNegCols + threadIdx.x;
  for (int i = threadIdx.x; i < kNegRows * D; i += blockDim.x) {
    int r = i / D, dd = i - r * D;
    qs[i] = (n0 + r < N) ? q_a[static_cast<size_t>(n0 + r) * D + dd] : 0.f;
  }
  __syncthreads();
  float acc[kNegRows];
#pragma unroll
  for (int r = 0; r < kNegRows; ++r) acc[r] = 0.f;
  if (k < K) {
    for (int dd = 0; dd < D; ++dd) {
      float qv = __ldg(queue + static_cast<size_t>(dd) * K + k);
#pragma unroll
      for (int r = 0; r < kNegRows; ++r) acc[r] = fmaf(qs[r * D + dd], qv, acc[r]);
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < kNegRows; ++r) {
    float l = acc[r] / T;
    if (k < K && n0 + r < N) {
      size_t o = static_cast<size_t>(n0 + r) * (K + 1) + 1 + k;
      if (logits1) logits1[o] = l;
      if (logits2) logits2[o] = l;
    }
    if (ranks) {   // contrastive accuracy without top-k: how many negatives beat each positive (uniform branch)
      const bool rv = k < K && n0 + r < N;
      const unsigned b1 = __ballot_sync(0xffffffffu, rv && l > pos1[n0 + r]);
      const unsigned b2 = __ballot_sync(0xffffffffu, rv && l > pos2[n0 + r]);
      if (lane == 0 && b1) atomicAdd(ranks + n0 + r, __popc(b1));
      if (lane == 0 && b2) atomicAdd(ranks + N + n0 + r, __popc(b2));
    }
    float lm = k < K ? l : -INFINITY;
    float mx = warp_max(lm);
    float se = warp_sum(k < K ? __expf(l - mx) : 0.f);
    if (lane == 0) {
      red[(warp * kNegRows + r) * 2] = mx;
      red[(warp * kNegRows + r) * 2 + 1] = se;
    }
  }
  __syncthreads();
  if (threadIdx.x < kNegRows) {
    int r = threadIdx.x;
    float mx = -INFINITY;
    for (int w = 0; w < kNegCols / 32; ++w) mx = fmaxf(mx, red[(w * kNegRows + r) * 2]);
    float se = 0.f;
    for (int w = 0; w < kNegCols / 32; ++w) {
      float wm = red[(w * kNegRows + r) * 2];
      if (wm > -INFINITY) se += red[(w * kNegRows + r) * 2 + 1] * __expf(wm - mx);
    }
    if (n0 + r < N) {
      ws[(static_cast<size_t>(n0 + r) * tiles + blockIdx.x) * 2] = mx;
      ws[(static_cast<size_t>(n0 + r) * tiles + blockIdx.x) * 2 + 1] = se;
    }
  }
}

// one warp per row: merge tile partials with each positive logit
__global__ void lse_finalize_kernel(const float* __restrict__ ws, const float* __restrict__ pos1,
                                    const float* __restrict__ pos2, int N, int tiles, float* __restrict__ lse1,
                                    float* __restrict__ lse2) {
  int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (n >= N) return;
  float mx = -INFINITY;
  for (int i = lane; i < tiles; i += 32) mx = fmaxf(mx, ws[(static_cast<size_t>(n) * tiles + i) * 2]);
  mx = warp_max(mx);
  float se = 0.f;
  for (int i = lane; i < tiles; i += 32) {
    float tm = ws[(static_cast<size_t>(n) * tiles + i) * 2];
    se += ws[(static_cast<size_t>(n) * tiles + i) * 2 + 1] * expf(tm - mx);
  }
  se = warp_sum(se);
  if (lane == 0) {
    float p1 = pos1[n], p2 = pos2[n];
    float m1 = fmaxf(mx, p1), m2 = fmaxf(mx, p2);
    lse1[n] = m1 + logf(se * expf(mx - m1) + expf(p1 - m1));
    lse2[n] = m2 + logf(se * expf(mx - m2) + expf(p2 - m2));
  }
}

// positive-key part of dq (overwrites dq_a / dq_m); one warp per row
__global__ void rowdots_bwd_kernel(const float* __restrict__ k_a, const float* __restrict__ k_m,
                                   const float* __restrict__ kn_a, const float* __restrict__ kn_m, int N, int D,
                                   int K, float T, const float* __restrict__ pos1, const float* __restrict__ pos2,
                                   const float* __restrict__ lse1, const float* __restrict__ lse2,
                                   const float* __restrict__ g_lse1, const float* __restrict__ g_lse2,
                                   const float* __restrict__ g_pos1, const float* __restrict__ g_pos2,
                                   const float* __restrict__ g_lpos_m, const float* __restrict__ g_lneg_m,
                                   const float* __restrict__ g_logits1, const float* __restrict__ g_logits2,
                                   float* __restrict__ dq_a, float* __restrict__ dq_m) {
  int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (n >= N) return;
  float c1 = g_lse1[n] * expf(pos1[n] - lse1[n]) + g_pos1[n];
  float c2 = g_lse2[n] * expf(pos2[n] - lse2[n]) + g_pos2[n];
  if (g_logits1) c1 += g_logits1[static_cast<size_t>(n) * (K + 1)];
  if (g_logits2) c2 += g_logits2[static_cast<size_t>(n) * (K + 1)];
  float cm1 = g_lpos_m[n], cm2 = g_lneg_m[n];
  for (int dd = lane; dd < D; dd += 32) {
    size_t o = static_cast<size_t>(n) * D + dd;
    dq_a[o] = (c1 * k_a[o] + c2 * kn_a[o]) / T;
    dq_m[o] = (cm1 * k_m[o] + cm2 * kn_m[o]) / T;
  }
}

constexpr int kBwdK = 64;       // queue columns per chunk
constexpr int kBwdChunks = 4;   // chunks per block
// queue part of dq_a: dq_a[n] += (1/T) sum_k w[n,k] queue[:,k] with
// w = g_lse1*softmax1 + g_lse2*softmax2 (+ dense logits grads). D is 64, 128 or 256 (moco.dim).
template <int D>
__global__ void __launch_bounds__(256) neg_logits_bwd_kernel(
    const float* __restrict__ q_a, const float* __restrict__ queue, int N, int K, float T,
    const float* __restrict__ lse1, const float* __restrict__ lse2, const float* __restrict__ g_lse1,
    const float* __restrict__ g_lse2, const float* __restrict__ g_logits1, const float* __restrict__ g_logits2,
    float* __restrict__ dq_a) {
  constexpr int DJ = D / 16;          // feature columns per thread
  extern __shared__ float sm[];
  float* qs = sm;                     // [16][D]
  float* Qs = qs + 16 * D;            // [D][kBwdK + 1]
  float* wsm = Qs + D * (kBwdK + 1);  // [16][kBwdK + 1]
  const int n0 = blockIdx.y * 16;
  const int t = threadIdx.x;
  for (int i = t; i < 16 * D; i += 256) {
    int r = i / D;
    qs[i] = (n0 + r < N) ? q_a[static_cast<size_t>(n0 + r) * D + (i - r * D)] : 0.f;
  }
  const int r = t >> 4, c = t & 15;
  const int n = n0 + r;
  const bool nok = n < N;
  const float l1 = nok ? lse1[n] : 0.f, l2 = nok ? lse2[n] : 0.f;
  const float g1 = nok ? g_lse1[n] : 0.f, g2 = nok ? g_lse2[n] : 0.f;
  float acc[DJ];
#pragma unroll
  for (int j = 0; j < DJ; ++j) acc[j] = 0.f;
  for (int ch = 0; ch < kBwdChunks; ++ch) {
    const int k0 = (blockIdx.x * kBwdChunks + ch) * kBwdK;
    __syncthreads();
    for (int i = t; i < D * kBwdK; i += 256) {
      int dd = i / kBwdK, kk = i - dd * kBwdK;
      Qs[dd * (kBwdK + 1) + kk] = (k0 + kk < K) ? __ldg(queue + static_cast<size_t>(dd) * K + k0 + kk) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kBwdK / 16; ++j) {
      int kk = c + 16 * j;
      float l = 0.f;
      for (int dd = 0; dd < D; ++dd) l = fmaf(qs[r * D + dd], Qs[dd * (kBwdK + 1) + kk], l);
      l /= T;
      float w = 0.f;
      if (nok && k0 + kk < K) {
        w = g1 * expf(l - l1) + g2 * expf(l - l2);
        size_t o = static_cast<size_t>(n) * (K + 1) + 1 + k0 + kk;
        if (g_logits1) w += g_logits1[o];
        if (g_logits2) w += g_logits2[o];
      }
      wsm[r * (kBwdK + 1) + kk] = w;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < DJ; ++j) {
      int dd = c + 16 * j;
      float a = 0.f;
      for (int kk = 0; kk < kBwdK; ++kk) a = fmaf(wsm[r * (kBwdK + 1) + kk], Qs[dd * (kBwdK + 1) + kk], a);
      acc[j] += a;
    }
  }
  if (nok) {
#pragma unroll
    for (int j = 0; j < DJ; ++j) atomicAdd(dq_a + static_cast<size_t>(n) * D + c + 16 * j, acc[j] / T);
  }
}

// Loss.forward (:272-283); single block
__global__ void moco_loss_kernel(const float* __restrict__ lse1, const float* __restrict__ lse2,
                                 const float* __restrict__ pos1, const float* __restrict__ pos2,
                                 const float* __restrict__ lpos_m, const float* __restrict__ lneg_m, int N,
                                 float margin, float A, float M, float* __restrict__ out3) {
  __shared__ float red[3][32];
  float c1 = 0, c2 = 0, rk = 0;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    c1 += lse1[n] - pos1[n];
    c2 += lse2[n] - pos2[n];
    rk += fmaxf(0.f, -(lpos_m[n] - lneg_m[n]) + margin);
  }
  c1 = warp_sum(c1);
  c2 = warp_sum(c2);
  rk = warp_sum(rk);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[0][warp] = c1;
    red[1][warp] = c2;
    red[2][warp] = rk;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s1 = 0, s2 = 0, s3 = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) {
      s1 += red[0][w];
      s2 += red[1][w];
      s3 += red[2][w];
    }
    float ce1 = s1 / N, ce2 = s2 / N, rank = s3 / N;
    out3[0] = A * (ce1 + ce2) + M * rank;
    out3[1] = ce1 + ce2;
    out3[2] = rank;
  }
}

__global__ void moco_loss_bwd_kernel(const float* __restrict__ lpos_m, const float* __restrict__ lneg_m, int N,
                                     float margin, float A, float M, const float* __restrict__ g3,
                                     float* __restrict__ g_lse1, float* __restrict__ g_lse2,
                                     float* __restrict__ g_pos1, float* __restrict__ g_pos2,
                                     float* __restrict__ g_lpos_m, float* __restrict__ g_lneg_m) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float g_ce = (g3[0] * A + g3[1]) / N;
  float g_rank = (g3[0] * M + g3[2]) / N;
  g_lse1[n] = g_ce;
  g_lse2[n] = g_ce;
  g_pos1[n] = -g_ce;
  g_pos2[n] = -g_ce;
  bool active = (-(lpos_m[n] - lneg_m[n]) + margin) > 0.f;
  g_lpos_m[n] = active ? -g_rank : 0.f;
  g_lneg_m[n] = active ? g_rank : 0.f;
}

// dense CE with target 0: one block per row
__global__ void __launch_bounds__(256) ce0_fwd_kernel(const float* __restrict__ logits, int L,
                                                      float* __restrict__ lse) {
  __shared__ float red[8];
  __shared__ float bmx;
  const float* row = logits + static_cast<size_t>(blockIdx.x) * L;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < L; i += 256) mx = fmaxf(mx, row[i]);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = red[0];
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
    bmx = m;
  }
  __syncthreads();
  mx = bmx;
  float se = 0.f;
  for (int i = threadIdx.x; i < L; i += 256) se += expf(row[i] - mx);
  se = warp_sum(se);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = se;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0;
    for (int w = 0; w < 8; ++w) s += red[w];
    lse[blockIdx.x] = mx + logf(s);
  }
}
__global__ void __launch_bounds__(256) ce0_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ lse,
                                                      int N, int L, const float* __restrict__ g,
                                                      float* __restrict__ dlogits) {
  const size_t base = static_cast<size_t>(blockIdx.x) * L;
  const float l = lse[blockIdx.x];
  const float s = g[0] / N;
  for (int i = threadIdx.x; i < L; i += 256) {
    float p = expf(logits[base + i] - l);
    dlogits[base + i] = s * (p - (i == 0 ? 1.f : 0.f));
  }
}

}  // namespace rsp

using namespace rsp;

__global__ void __launch_bounds__(256) metrics_update_kernel(const float* __restrict__ loss3,
                                                             const int32_t* __restrict__ ranks,
                                                             const float* __restrict__ lpos_m,
                                                             const float* __restrict__ lneg_m, int N,
                                                             float* __restrict__ meters) {
  __shared__ int cnt[5];
  if (threadIdx.x < 5) cnt[threadIdx.x] = 0;
  __syncthreads();
  int c[5] = {0, 0, 0, 0, 0};
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const int r1 = ranks[n], r2 = ranks[N + n];
    c[0] += r1 == 0;                     // top-1 of logits1: nothing in the queue beats the positive
    c[1] += r1 < 5;                      // top-5
    c[2] += r2 == 0;
    c[3] += r2 < 5;
    c[4] += lpos_m[n] >= lneg_m[n];      // top-1 of cat(l_pos_M, l_neg_M) against class 0 (ties go to the lower index)
  }
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    int v = c[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&cnt[i], v);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float scale = 100.0f / N;
    const float val[8] = {loss3[0],       loss3[1],       cnt[0] * scale, cnt[1] * scale,
                          cnt[2] * scale, cnt[3] * scale, loss3[2],       cnt[4] * scale};
    for (int i = 0; i < 8; ++i) {
      meters[i] = val[i];
      meters[8 + i] += val[i] * N;
    }
    reinterpret_cast<int32_t*>(meters)[16] += N;
  }
}

extern "C" {

int rsp_ema_update(float* k, const float* q, int64_t n, float m, float one_minus_m, void* stream) {
  RSP_REQUIRE(n >= 0, "ema: negative size");
  if (n == 0) return RSP_OK;
  RSP_REQUIRE((reinterpret_cast<uintptr_t>(k) & 15) == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0,
              "ema: buffers must be 16-byte aligned");
  size_t n4 = static_cast<size_t>(n) / 4;
  ema_kernel<<<grid_for(n4 + 1, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(k, q, n4, n, m, one_minus_m);
  return check_launch("ema_update");
}

int rsp_sgd_step(float* p, const float* grad, float* mom, int64_t n, float lr, float momentum, float weight_decay,
                 float grad_scale, int first_step, void* stream) {
  if (n == 0) return RSP_OK;
  RSP_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(mom)) &
               15) == 0,
              "sgd: buffers must be 16-byte aligned");
  size_t n4 = static_cast<size_t>(n) / 4;
  sgd_kernel<<<grid_for(n4 + 1, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p, grad, mom, n4, n, lr, momentum, weight_decay, grad_scale, first_step);
  return check_launch("sgd_step");
}

int rsp_speed_gather(const float* im_q, const float* im_k, const int64_t* perm, int32_t B, int32_t C, int32_t T,
                     int32_t H, int32_t W, int32_t n_s1, int32_t d, int32_t layout, void* out_q, void* out_k,
                     void* out_kneg, void* stream) {
  RSP_REQUIRE(d >= 1 && T / d >= 1, "speed_gather: bad speed %d for T=%d", d, T);
  RSP_REQUIRE(layout == 0 || (layout == 1 && C == 3), "speed_gather: layout 1 needs C == 3");
  RSP_REQUIRE(n_s1 >= 0 && n_s1 <= B, "speed_gather: n_s1 out of range");
  if (B == 0) return RSP_OK;
  const int Tr = T / d, HW = H * W;
  RSP_REQUIRE(static_cast<long long>(B) * Tr <= 65535, "speed_gather: B*T too large");
  int per = layout == 0 ? (HW / 4 > 0 ? HW / 4 : HW) : HW;
  dim3 grid((per + 255) / 256 > 64 ? 64 : (per + 255) / 256, B * Tr, 3);
  if (layout == 0)
    speed_gather_kernel<0><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(im_q, im_k, perm, B, C, T, HW, n_s1, d,
                                                                                 Tr, out_q, out_k, out_kneg);
  else
    speed_gather_kernel<1><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(im_q, im_k, perm, B, C, T, HW, n_s1, d,
                                                                                 Tr, out_q, out_k, out_kneg);
  return check_launch("speed_gather");
}

int rsp_gather_rows(const void* src, const int64_t* index, void* dst, int64_t n_rows, int64_t row_bytes,
                    void* stream) {
  if (n_rows == 0) return RSP_OK;
  RSP_REQUIRE(row_bytes % 16 == 0 && n_rows <= 65535, "gather_rows: row_bytes %% 16 != 0 or too many rows");
  size_t vec = static_cast<size_t>(row_bytes) / 16;
  unsigned gx = static_cast<unsigned>((vec + 255) / 256);
  if (gx > 128) gx = 128;
  dim3 grid(gx, static_cast<unsigned>(n_rows));
  gather_rows_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(src), index, static_cast<uint4*>(dst), vec);
  return check_launch("gather_rows");
}

int rsp_queue_enqueue(float* queue, const float* keys, int64_t* queue_ptr, int32_t D, int32_t K, int32_t n,
                      void* stream) {
  RSP_REQUIRE(n > 0 && K % n == 0, "enqueue: K=%d must be a multiple of the gathered batch %d", K, n);
  dim3 grid((n + 31) / 32, (D + 31) / 32), block(32, 8);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  enqueue_kernel<<<grid, block, 0, s>>>(queue, keys, queue_ptr, D, K, n);
  advance_ptr_kernel<<<1, 1, 0, s>>>(queue_ptr, K, n);
  return check_launch("queue_enqueue");
}

int64_t rsp_moco_logits_workspace(int32_t N, int32_t K) {
  return static_cast<int64_t>(N) * ((K + kNegCols - 1) / kNegCols) * 2 * sizeof(float);
}

static int moco_logits_fwd_impl(const float* q_a, const float* q_m, const float* k_a, const float* k_m, const float* kn_a,
                                const float* kn_m, const float* queue, int32_t N, int32_t D, int32_t K, float temperature,
                                float* logits1, float* logits2, float* lpos_m, float* lneg_m, float* lse1, float* lse2,
                                float* pos1, float* pos2, float* workspace, int32_t* ranks, void* stream) {
  RSP_REQUIRE(N > 0 && D > 0 && K > 0 && D <= 512, "moco_logits: bad sizes N=%d D=%d K=%d", N, D, K);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (ranks && rsp::zero_async(ranks, sizeof(int32_t) * 2 * N, s) != cudaSuccess) {
    set_error("moco_logits: memset failed");
    return RSP_ERR_CUDA;
  }
  rowdots_kernel<<<(N + 7) / 8, 256, 0, s>>>(q_a, q_m, k_a, k_m, kn_a, kn_m, N, D, K, temperature, logits1, logits2,
                                             lpos_m, lneg_m, pos1, pos2);
  const int tiles = (K + kNegCols - 1) / kNegCols;
  dim3 grid(tiles, (N + kNegRows - 1) / kNegRows);
  size_t smem = (static_cast<size_t>(kNegRows) * D + (kNegCols / 32) * kNegRows * 2) * sizeof(float);
  neg_logits_kernel<<<grid, kNegCols, smem, s>>>(q_a, queue, N, D, K, temperature, logits1, logits2, workspace, tiles,
                                                 pos1, pos2, ranks);
  lse_finalize_kernel<<<(N + 7) / 8, 256, 0, s>>>(workspace, pos1, pos2, N, tiles, lse1, lse2);
  return check_launch("moco_logits_fwd");
}

int rsp_moco_logits_fwd(const float* q_a, const float* q_m, const float* k_a, const float* k_m, const float* kn_a,
                        const float* kn_m, const float* queue, int32_t N, int32_t D, int32_t K, float temperature,
                        float* logits1, float* logits2, float* lpos_m, float* lneg_m, float* lse1, float* lse2,
                        float* pos1, float* pos2, float* workspace, void* stream) {
  return moco_logits_fwd_impl(q_a, q_m, k_a, k_m, kn_a, kn_m, queue, N, D, K, temperature, logits1, logits2, lpos_m, lneg_m,
                              lse1, lse2, pos1, pos2, workspace, nullptr, stream);
}

int rsp_moco_logits_fwd_ranked(const float* q_a, const float* q_m, const float* k_a, const float* k_m, const float* kn_a,
                               const float* kn_m, const float* queue, int32_t N, int32_t D, int32_t K, float temperature,
                               float* logits1, float* logits2, float* lpos_m, float* lneg_m, float* lse1, float* lse2,
                               float* pos1, float* pos2, float* workspace, int32_t* ranks, void* stream) {
  RSP_REQUIRE(ranks != nullptr, "moco_logits_fwd_ranked: ranks must not be null");
  return moco_logits_fwd_impl(q_a, q_m, k_a, k_m, kn_a, kn_m, queue, N, D, K, temperature, logits1, logits2, lpos_m, lneg_m,
                              lse1, lse2, pos1, pos2, workspace, ranks, stream);
}

// Device-side AverageMeters of the training loop (pretrain.py:97-106,169-196; framework/meters/average.py:4-44;
// framework/metrics/classification.py:6-20): one thread block turns the step's loss triple, the rank counters and the
// ranking pair into the eight logged values and folds them into val / sum / count without a host round trip.
// meters: float val[8], float sum[8], then int32 count — order Loss, Loss_A, Acc@1_A, Acc@5_A, Acc@1_A_n, Acc@5_A_n,
// Loss_M, Acc@1_M.
int rsp_metrics_update(const float* loss3, const int32_t* ranks, const float* lpos_m, const float* lneg_m, int32_t N,
                       float* meters, void* stream) {
  RSP_REQUIRE(N > 0, "metrics_update: empty batch");
  metrics_update_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(loss3, ranks, lpos_m, lneg_m, N, meters);
  return check_launch("metrics_update");
}

int rsp_moco_logits_bwd(const float* q_a, const float* q_m, const float* k_a, const float* k_m, const float* kn_a,
                        const float* kn_m, const float* queue, int32_t N, int32_t D, int32_t K, float temperature,
                        const float* pos1, const float* pos2, const float* lse1, const float* lse2,
                        const float* g_lse1, const float* g_lse2, const float* g_pos1, const float* g_pos2,
                        const float* g_lpos_m, const float* g_lneg_m, const float* g_logits1, const float* g_logits2,
                        float* dq_a, float* dq_m, void* stream) {
  (void)q_m;
  RSP_REQUIRE(D == 64 || D == 128 || D == 256, "moco_logits_bwd: feature dimension must be 64, 128 or 256 (got %d)", D);
  RSP_REQUIRE(N > 0 && K > 0, "moco_logits_bwd: bad sizes N=%d K=%d", N, K);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // positive keys first (overwrites dq), then the queue part accumulates with atomics
  rowdots_bwd_kernel<<<(N + 7) / 8, 256, 0, s>>>(k_a, k_m, kn_a, kn_m, N, D, K, temperature, pos1, pos2, lse1, lse2,
                                                 g_lse1, g_lse2, g_pos1, g_pos2, g_lpos_m, g_lneg_m, g_logits1,
                                                 g_logits2, dq_a, dq_m);
  const int per_block = kBwdK * kBwdChunks;
  dim3 grid((K + per_block - 1) / per_block, (N + 15) / 16);
  size_t smem = (16 * D + D * (kBwdK + 1) + 16 * (kBwdK + 1)) * sizeof(float);
#define RSP_LAUNCH_NEG_BWD(DD)                                                                                          \
  do {                                                                                                                  \
    static bool attr_set = false;                                                                                       \
    if (!attr_set) {                                                                                                    \
      cudaFuncSetAttribute(neg_logits_bwd_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
      attr_set = true;                                                                                                  \
    }                                                                                                                   \
    neg_logits_bwd_kernel<DD><<<grid, 256, smem, s>>>(q_a, queue, N, K, temperature, lse1, lse2, g_lse1, g_lse2,       \
                                                      g_logits1, g_logits2, dq_a);                                      \
  } while (0)
  if (D == 64) RSP_LAUNCH_NEG_BWD(64);
  else if (D == 128) RSP_LAUNCH_NEG_BWD(128);
  else RSP_LAUNCH_NEG_BWD(256);
#undef RSP_LAUNCH_NEG_BWD
  return check_launch("moco_logits_bwd");
}

int rsp_moco_loss_fwd(const float* lse1, const float* lse2, const float* pos1, const float* pos2, const float* lpos_m,
                      const float* lneg_m, int32_t N, float margin, float A, float M, float* out3, void* stream) {
  RSP_REQUIRE(N > 0, "moco_loss: empty batch");
  moco_loss_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(lse1, lse2, pos1, pos2, lpos_m, lneg_m, N, margin,
                                                                      A, M, out3);
  return check_launch("moco_loss_fwd");
}

int rsp_moco_loss_bwd(const float* lpos_m, const float* lneg_m, int32_t N, float margin, float A, float M,
                      const float* g_out3, float* g_lse1, float* g_lse2, float* g_pos1, float* g_pos2,
                      float* g_lpos_m, float* g_lneg_m, void* stream) {
  moco_loss_bwd_kernel<<<(N + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      lpos_m, lneg_m, N, margin, A, M, g_out3, g_lse1, g_lse2, g_pos1, g_pos2, g_lpos_m, g_lneg_m);
  return check_launch("moco_loss_bwd");
}

int rsp_ce0_fwd(const float* logits, int32_t N, int32_t L, float* lse, void* stream) {
  RSP_REQUIRE(N > 0 && L > 0, "ce0: empty input");
  ce0_fwd_kernel<<<N, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, L, lse);
  return check_launch("ce0_fwd");
}

int rsp_ce0_bwd(const float* logits, const float* lse, int32_t N, int32_t L, const float* g_scalar, float* dlogits,
                void* stream) {
  ce0_bwd_kernel<<<N, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, lse, N, L, g_scalar, dlogits);
  return check_launch("ce0_bwd");
}

}  // extern "C"
