// Direct (im2col-free) tcgen05 filter gradient for unit-stride kt x 3 x 3 convolutions on 64-channel-multiple tensors
// (reference: autograd of the 3x3x3 stride-1 nn.Conv3d in models/resnet.py:21-27 (BasicBlock) — convolution_backward's
// grad_weight).  Companion of conv_direct.cu: the same padded-plane flattening, now with the PIXELS as the K axis.
//
//   dW[a,b,c][ci][co] = sum_{n,to} sum_q  X_{n, to-pt+a}[q + b*Wp + c][ci] * dY_{n,to}[q][co]
//
// q = ho*Wp + wo runs over one output plane flattened with the PADDED pitch Wp = Wi + 2 (the Wp - Wo surplus columns of
// dY are zero-filled by the TMA unit, as are the halo rows / columns of X), so a filter tap is again a pure offset:
//   * X run in shared memory (one 128-byte swizzled row per pixel, 64 channels) = MN-major A operand of EVERY tap (b, c):
//     the descriptor starts b*Wp + c rows later.  An M = 128 tile holds two taps: the second 64-wide panel is the same
//     buffer `LBO` rows further (overlapping panels, verified by tools/umma_mn_shift_probe.cu).
//   * dY run = MN-major B operand.  With Co chunks of 64 an N = 64 MMA would hold the pipe 48 clocks for 32 clocks of
//     math, so the second 64 columns are the SAME dY run one pixel earlier: sum_q X[q+t] dY[q-1] = dW[t+1] — a shifted
//     B panel is another tap.  (The K range of a plane is [0, P+1) so that the shifted panel also sees every pixel; what
//     lies outside the plane is zero-filled.)  One M128 x N128 tile = four taps from one A fetch.
// The generic wgrad re-gathers x once per tap pair and dy once per tile (24 KB staged per MFLOP, L2-bound at 445 TF/s on
// R3D-18 layer1); here a plane chunk is staged once per frame tap for nine taps (4.9 KB per MFLOP).
//
// Work split: a CTA group = (frame tap a, ci chunk, co chunk) keeps its nine 64 x 64 gradients in TMEM for the whole
// kernel (3 tiles: 128 + 128 + 64 columns) while its CTAs stream the planes (n, to); at the end every CTA adds its
// partial sums into dwt[(tap*Cs + ci)][co] (fp32 red.global.add.v4), the layout the generic kernel produces.
//   tile 0: M = {(0,0), (1,0)}  N = {+1, +0}  -> (0,1) (1,1) | (0,0) (1,0)
//   tile 1: M = {(0,2), (2,0)}  N = {+1, +0}  -> ( -  ) (2,1) | (0,2) (2,0)
//   tile 2: M = {(1,2), (2,2)}  N = {+0}      -> (1,2) (2,2)
// CTA: warps 0-3 epilogue, warp 4 TMA producer (one lane), warp 5 MMA issuer (one lane).
#include "common.cuh"
#include "rspnet_b200.h"

namespace rsp {

int device_sm_count();

constexpr int kWDMaxGroups = 64;

struct WDirectParams {
  CUtensorMap tmapX;    // x  as {Cs, Wi, Hi, Ti, N}, box {64, Wp, rowsX, 1, 1}
  CUtensorMap tmapDy;   // dy as {Co, Wo, Ho, To, N}, box {64, Wp, rowsDy, 1, 1}
  float* dwt;           // [kt*9*Cs][Co] fp32, zeroed by the caller
  int N, Ti, To, Cs, Co;
  int kt, pt;
  int Wp, P;            // padded pitch, positions per output plane (Ho * Wp)
  int Kp;               // K extent per plane: P + 1 rounded up to 16
  int R;                // positions per chunk (multiple of 16)
  int chunks;           // ceil(Kp / R)
  int rowsX, rowsDy;    // TMA box heights
  int xBytes, stageBytes;
  int stages;
  int groups;           // kt * (Cs/64) * (Co/64)
  int groupStart[kWDMaxGroups + 1];   // first CTA of each group
  unsigned long long mulWp;
  int shWp;
};

constexpr int kWDThreads = 192;

__global__ void __launch_bounds__(kWDThreads, 1) conv_wgrad_direct_kernel(const __grid_constant__ WDirectParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * p.stageBytes);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* accum_bar = empty_bar + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int t = threadIdx.x;
  const int warp = t >> 5;

  // group of this CTA and its rank inside the group
  int grp = 0;
  while (grp + 1 < p.groups && static_cast<int>(blockIdx.x) >= p.groupStart[grp + 1]) ++grp;
  const int worker = blockIdx.x - p.groupStart[grp];
  const int workers = p.groupStart[grp + 1] - p.groupStart[grp];
  const int cchunks = p.Cs >> 6, ochunks = p.Co >> 6;
  const int oc = grp % ochunks;
  const int cc = (grp / ochunks) % cchunks;
  const int a = grp / (ochunks * cchunks);

  // planes (n, to) of this CTA: item = to * N + n, every workers-th one; planes whose source frame is outside the clip
  // contribute nothing
  const int numPlanes = p.N * p.To;
  auto plane_valid = [&](int item, int& n, int& to, int& ts) {
    n = item % p.N;
    to = item / p.N;
    ts = to - p.pt + a;
    return ts >= 0 && ts < p.Ti;
  };
  int myPlanes = 0;
  for (int item = worker; item < numPlanes; item += workers) {
    int n, to, ts;
    if (plane_valid(item, n, to, ts)) ++myPlanes;
  }

  if (t == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (myPlanes > 0) {
    if (warp == 4) {
      // ============================================================ TMA producer
      if (elect_one()) {
        tma_prefetch_desc(&p.tmapX);
        tma_prefetch_desc(&p.tmapDy);
        const uint32_t tx = static_cast<uint32_t>(p.rowsX + p.rowsDy) * p.Wp * 128u;
        int s = 0;
        uint32_t ph = 0;
        for (int item = worker; item < numPlanes; item += workers) {
          int n, to, ts;
          if (!plane_valid(item, n, to, ts)) continue;
          for (int ch = 0; ch < p.chunks; ++ch) {
            const int k0 = ch * p.R;
            const int rx0 = static_cast<int>((static_cast<unsigned long long>(k0) * p.mulWp) >> p.shWp);      // k0 / Wp
            const int rd0 = k0 == 0 ? -1 : static_cast<int>((static_cast<unsigned long long>(k0 - 1) * p.mulWp) >> p.shWp);
            mbar_wait(&empty_bar[s], ph ^ 1);
            const uint32_t stage = smem_u32(smem + s * p.stageBytes);
            mbar_arrive_expect_tx(&full_bar[s], tx);
            tma_load_5d(stage, &p.tmapX, &full_bar[s], cc * 64, -1, rx0 - 1, ts, n);
            tma_load_5d(stage + p.xBytes, &p.tmapDy, &full_bar[s], oc * 64, 0, rd0, to, n);
            if (++s == p.stages) {
              s = 0;
              ph ^= 1;
            }
          }
        }
      }
    } else if (warp == 5) {
      // ============================================================ MMA issuer
      if (elect_one()) {
        constexpr uint32_t idesc128 = make_idesc_bf16(128, 128, 1, 1);
        constexpr uint32_t idesc64 = make_idesc_bf16(128, 64, 1, 1);
        const uint32_t Wp = static_cast<uint32_t>(p.Wp);
        // tap offsets (pixel rows) of the first panel and distance to the second panel of each tile
        const uint32_t t1[3] = {0u, 2u, Wp + 2u};
        const uint32_t lbo[3] = {Wp, 2u * Wp - 2u, Wp};
        int s = 0;
        uint32_t ph = 0, acc = 0;
        for (int pl = 0; pl < myPlanes; ++pl) {
          for (int ch = 0; ch < p.chunks; ++ch) {
            const int k0 = ch * p.R;
            const int rx0 = static_cast<int>((static_cast<unsigned long long>(k0) * p.mulWp) >> p.shWp);
            const int rd0 = k0 == 0 ? -1 : static_cast<int>((static_cast<unsigned long long>(k0 - 1) * p.mulWp) >> p.shWp);
            const uint32_t shiftX = static_cast<uint32_t>(k0 - rx0 * p.Wp);          // position k0 inside the X box
            const uint32_t shiftDy = static_cast<uint32_t>(k0 - rd0 * p.Wp);         // position k0 inside the dY box (>= 1)
            int ksteps = (p.Kp - k0) >> 4;
            if (ksteps > (p.R >> 4)) ksteps = p.R >> 4;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after_sync();
            const uint32_t stage = smem_u32(smem + s * p.stageBytes);
            // MN-major, 128-byte swizzle: 8-pixel groups 1024 B apart (SBO), second 64-wide panel LBO bytes further
            const uint64_t a0 = make_smem_desc_sw128(stage + (shiftX + t1[0]) * 128u, lbo[0] * 128u, 1024);
            const uint64_t a1 = make_smem_desc_sw128(stage + (shiftX + t1[1]) * 128u, lbo[1] * 128u, 1024);
            const uint64_t a2 = make_smem_desc_sw128(stage + (shiftX + t1[2]) * 128u, lbo[2] * 128u, 1024);
            const uint64_t bsh = make_smem_desc_sw128(stage + p.xBytes + (shiftDy - 1u) * 128u, 128u, 1024);  // {dY[q-1] | dY[q]}
            const uint64_t b0 = make_smem_desc_sw128(stage + p.xBytes + shiftDy * 128u, 128u, 1024);           // dY[q]
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint64_t step = static_cast<uint64_t>(ks) * 128u;   // 16 pixels = 2048 B
              umma_bf16(tmem_base, a0 + step, bsh + step, idesc128, acc);
              umma_bf16(tmem_base + 128, a1 + step, bsh + step, idesc128, acc);
              umma_bf16(tmem_base + 256, a2 + step, b0 + step, idesc64, acc);
              acc = 1;
            }
            umma_commit(&empty_bar[s]);
            if (++s == p.stages) {
              s = 0;
              ph ^= 1;
            }
          }
        }
        umma_commit(accum_bar);
      }
    } else {
      // ============================================================ epilogue: TMEM lane = (panel, ci), columns = (tap shift, co)
      mbar_wait(accum_bar, 0);
      tc_fence_after_sync();
      const int lane = t & 31;
      const int l = warp * 32 + lane;
      const int half = l >> 6, ci = l & 63;
      // (b, c) of the 64-column blocks: tile 0 {+1 | +0}, tile 1 {+1 | +0}, tile 2 {+0}; -1 = unused slot
      const int tb[5] = {half ? 1 : 0, half ? 1 : 0, half ? 2 : -1, half ? 2 : 0, half ? 2 : 1};
      const int tcx[5] = {1, 0, 1, half ? 0 : 2, 2};
#pragma unroll 1
      for (int blk = 0; blk < 5; ++blk) {
        if (tb[blk] < 0) continue;   // uniform per warp (half is a function of the warp)
        const int tap = (a * 3 + tb[blk]) * 3 + tcx[blk];
        float* drow = p.dwt + (static_cast<size_t>(tap) * p.Cs + cc * 64 + ci) * p.Co + oc * 64;
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + blk * 64 + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + c0 + j),
                         "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])),
                         "f"(__uint_as_float(v[j + 3]))
                         : "memory");
          }
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, 512);
}

// --------------------------------------------------------------------------------------------------------------------
static bool wdirect_geometry(const rsp_conv3d_desc* d, int sm_count, WDirectParams& p, size_t& smem) {
  if (d->st != 1 || d->sh != 1 || d->sw != 1) return false;
  if (d->kh != 3 || d->kw != 3 || d->ph != 1 || d->pw != 1) return false;
  if (d->Ci % 64 != 0 || d->Co % 64 != 0) return false;
  const int To = d->Ti + 2 * d->pt - d->kt + 1, Ho = d->Hi, Wo = d->Wi;
  if (To <= 0) return false;
  p.N = d->N; p.Ti = d->Ti; p.To = To; p.Cs = d->Ci; p.Co = d->Co;
  p.kt = d->kt; p.pt = d->pt;
  p.Wp = d->Wi + 2;
  if (p.Wp > 256) return false;
  p.P = Ho * p.Wp;
  p.Kp = (p.P + 1 + 15) / 16 * 16;
  p.groups = d->kt * (d->Ci / 64) * (d->Co / 64);
  if (p.groups > kWDMaxGroups || p.groups > sm_count) return false;
  // chunking: the largest chunk that still leaves three pipeline stages fixes the chunk COUNT; the chunk length is then
  // evened out over that count (a short last chunk would still stage a full box)
  p.stages = 0;
  const int tmax = 2 * p.Wp + 2;
  auto try_chunk = [&](int R) {
    const int rowsX = (R + tmax - 1) / p.Wp + 2, rowsDy = R / p.Wp + 2;
    if (rowsX > 256 || rowsDy > 256) return false;
    const int xBytes = (rowsX * p.Wp * 128 + 1023) / 1024 * 1024;
    const int stageBytes = xBytes + (rowsDy * p.Wp * 128 + 1023) / 1024 * 1024;
    if (3ll * stageBytes + 1024 + 512 > 227 * 1024) return false;
    p.R = R;
    p.rowsX = rowsX;
    p.rowsDy = rowsDy;
    p.xBytes = xBytes;
    p.stageBytes = stageBytes;
    p.stages = static_cast<int>((227 * 1024 - 1024 - 512) / stageBytes);
    if (p.stages > 6) p.stages = 6;
    return true;
  };
  for (int R = p.Kp < 512 ? p.Kp : 512; R >= 64; R -= 16) {
    if (!try_chunk(R)) continue;
    const int chunks = (p.Kp + R - 1) / R;
    const int even = ((p.Kp + chunks - 1) / chunks + 15) / 16 * 16;
    if (even < R) try_chunk(even);
    break;
  }
  if (p.stages < 3) return false;
  p.chunks = (p.Kp + p.R - 1) / p.R;
  smem = static_cast<size_t>(p.stages) * p.stageBytes + 1024 + 512;
  // CTAs per group in proportion to the planes whose source frame exists for the group's frame tap
  {
    long long valid[16];
    long long total = 0;
    if (d->kt > 16) return false;
    for (int a = 0; a < d->kt; ++a) {
      int cnt = 0;
      for (int to = 0; to < To; ++to) cnt += (to - d->pt + a >= 0 && to - d->pt + a < d->Ti) ? 1 : 0;
      valid[a] = static_cast<long long>(cnt) * d->N;
      total += valid[a];
    }
    if (total == 0) return false;
    const int per = (d->Ci / 64) * (d->Co / 64);
    int start = 0;
    int left = sm_count;
    long long remaining = total * per;
    for (int g = 0; g < p.groups; ++g) {
      const int a = g / per;
      int w = remaining > 0 ? static_cast<int>((static_cast<long long>(left) * valid[a] + remaining / 2) / remaining) : 1;
      const int groupsLeft = p.groups - g - 1;
      if (w < 1) w = 1;
      if (w > left - groupsLeft) w = left - groupsLeft;
      p.groupStart[g] = start;
      start += w;
      left -= w;
      remaining -= valid[a];
    }
    p.groupStart[p.groups] = start;
  }
  int l = 0;
  while ((1 << l) < p.Wp) ++l;
  p.shWp = 32 + l;
  p.mulWp = ((1ull << p.shWp) + p.Wp - 1) / p.Wp;
  return true;
}

bool wgrad_direct_supported(const rsp_conv3d_desc* d) {
  WDirectParams p{};
  size_t smem;
  return wdirect_geometry(d, device_sm_count(), p, smem);
}

// dwt [kt*9*Ci][Co] fp32 must be zero on entry (the caller zeroes it and unpacks it afterwards, as for the generic kernel).
int launch_wgrad_direct(const rsp_conv3d_desc* d, const void* x, const void* dy, float* dwt, cudaStream_t stream) {
  WDirectParams p{};
  size_t smem;
  if (!wdirect_geometry(d, device_sm_count(), p, smem)) {
    set_error("conv_wgrad_direct: unsupported geometry");
    return RSP_ERR_INVALID;
  }
  p.dwt = dwt;
  {
    const unsigned long long C = d->Ci, W = d->Wi, H = d->Hi, T = d->Ti, N = d->N;
    const unsigned long long dims[5] = {C, W, H, T, N};
    const unsigned long long strides[4] = {C * 2, W * C * 2, H * W * C * 2, T * H * W * C * 2};
    const unsigned box[5] = {64, static_cast<unsigned>(p.Wp), static_cast<unsigned>(p.rowsX), 1, 1};
    int rc = make_tmap_bf16(&p.tmapX, x, 5, dims, strides, box);
    if (rc != RSP_OK) return rc;
    const unsigned long long Co = d->Co, To = p.To;
    const unsigned long long ddims[5] = {Co, W, H, To, N};
    const unsigned long long dstrides[4] = {Co * 2, W * Co * 2, H * W * Co * 2, To * H * W * Co * 2};
    const unsigned dbox[5] = {64, static_cast<unsigned>(p.Wp), static_cast<unsigned>(p.rowsDy), 1, 1};
    rc = make_tmap_bf16(&p.tmapDy, dy, 5, ddims, dstrides, dbox);
    if (rc != RSP_OK) return rc;
  }
  // always the device maximum: the attribute is per function, not per launch, and a tool that re-launches a captured
  // graph node on its own (ncu --graph-profiling node) would otherwise see the value of the LAST eager launch
  cudaError_t e = cudaFuncSetAttribute(conv_wgrad_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(conv_wgrad_direct): %s", cudaGetErrorString(e));
    return RSP_ERR_CUDA;
  }
  conv_wgrad_direct_kernel<<<p.groupStart[p.groups], kWDThreads, smem, stream>>>(p);
  return check_launch("conv_wgrad_direct");
}

}  // namespace rsp
