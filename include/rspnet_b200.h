/* rspnet_b200 — C ABI of the B200 (sm_100a) RSPNet pretraining hot path.
 *
 * The reference (PeihaoChen/RSPNet) has no native layer: every operation below is reached there through
 * torch.nn / ATen / cuDNN / cuBLAS from Python.  Each entry point names the reference call site it
 * replaces (file:line relative to the reference root).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (PyTorch caching allocator); nothing is
 *    allocated, freed or retained by the library (sole exception: the rsp_peer_* IPC buffers);
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *  - return value 0 = ok, negative = error (see rsp_last_error(), thread-local);
 *  - activations on the conv path are bf16, channels-last NDHWC ([N][T][H][W][C]); "C" is the STORED
 *    channel count (RGB input padded 3->4, everything else a multiple of 64);
 *  - there is no CPU path and no other architecture: rsp_init() fails unless the device is sm_100.
 */
#ifndef RSPNET_B200_H_
#define RSPNET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSP_ABI_VERSION 2

int rsp_abi_version(void);
/* Checks that the current device is compute capability 10.x; caches the SM count. */
int rsp_init(void);
const char* rsp_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Conv3d as implicit GEMM on tcgen05 (reference: nn.Conv3d at models/resnet.py:21-27,130-136,170-175,
 * models/c3d.py:21-50, models/r2plus1d_vcop.py:56-67, models/s3dg.py:20-21 and their autograd).
 * ------------------------------------------------------------------------------------------------ */
typedef struct rsp_conv3d_desc {
  int32_t N, Ti, Hi, Wi;   /* input pixel grid                                   */
  int32_t Ci, Co;          /* stored channels: Ci == 4 (RGB+pad) or Ci % 64 == 0; Co % 64 == 0 */
  int32_t kt, kh, kw;      /* filter                                             */
  int32_t st, sh, sw;      /* stride                                             */
  int32_t pt, ph, pw;      /* zero padding                                       */
} rsp_conv3d_desc;

/* which = 0: K extent (elements per output channel) of the packed fprop filter;
 * which = 1: K extent of the packed dgrad filter. Negative on error. */
int rsp_conv3d_kpad(const rsp_conv3d_desc* d, int which);
/* Number of bf16 elements of the packed operand (which = 0 may carry extra direct-conv slabs for the RGB stem). */
int64_t rsp_conv3d_packed_elems(const rsp_conv3d_desc* d, int which);
/* w: fp32 [Co_logical][Ci_logical][kt][kh][kw] (the nn.Conv3d parameter).
 * which = 0 -> wp bf16 [Co][kpad(0)] (+ stem slabs)  (fprop / wgrad K order);  which = 1 -> wd bf16 [Ci][kpad(1)]
 * (dgrad).  wp must hold rsp_conv3d_packed_elems(d, which) elements. */
int rsp_conv3d_pack_weight(const rsp_conv3d_desc* d, int Ci_logical, int Co_logical, const float* w, void* wp,
                           int which, void* stream);
/* The which = 0 packing of n filters in one launch per 24 filters (host arrays of descriptors / pointers): what the
 * training loop calls after every optimizer and momentum-encoder update. */
int rsp_conv3d_pack_weights(int32_t n, const rsp_conv3d_desc* descs, const int32_t* ci_logical,
                            const int32_t* co_logical, const float* const* w, void* const* wp, void* stream);
/* Bytes of the fp32 split-K accumulation buffer fprop (which = 0) / dgrad (which = 1) want for this geometry
 * (0: the K loop is not split). Passing NULL as workspace is always legal and disables split-K. */
int64_t rsp_conv3d_workspace_bytes(const rsp_conv3d_desc* d, int which);
/* y[N,To,Ho,Wo,Co] = conv(x[N,Ti,Hi,Wi,Ci], wp) (+ bias[Co] fp32, may be NULL).
 * stats (may be NULL): fp32 [2][Co], the epilogue ADDS the per-channel sum and sum of squares of the stored bf16
 * output (the BatchNorm batch statistics); ignored (left untouched) when split-K is active, i.e. when a non-NULL
 * workspace is passed for a geometry with rsp_conv3d_workspace_bytes() > 0 — use rsp_bn_stats then. */
int rsp_conv3d_fprop(const rsp_conv3d_desc* d, const void* x, const void* wp, const float* bias, void* y,
                     void* workspace, float* stats, void* stream);
/* dx[N,Ti,Hi,Wi,Ci] = conv_transpose(dy[N,To,Ho,Wo,Co], wd). Strides must be powers of two. */
int rsp_conv3d_dgrad(const rsp_conv3d_desc* d, const void* dy, const void* wd, void* dx, void* workspace,
                     void* stream);
/* dw fp32 [Co_logical][Ci_logical][kt][kh][kw] (=, or += when accumulate) from x and dy.
 * dwt_workspace: fp32 [kpad(0)][Co], overwritten. */
int rsp_conv3d_wgrad(const rsp_conv3d_desc* d, int Ci_logical, int Co_logical, const void* x, const void* dy,
                     float* dwt_workspace, float* dw, int accumulate, void* stream);

/* ------------------------------------------------------------------------------------------------
 * BatchNorm3d (train mode) + ReLU + residual, NDHWC bf16 (reference: nn.BatchNorm3d / nn.ReLU /
 * `out += residual` at models/resnet.py:54-75,137-138; models/c3d.py:22-50).
 * ------------------------------------------------------------------------------------------------ */
/* sum[c] += sum_m x[m][c], sumsq[c] += sum_m x^2 (fp32; caller zeroes them). */
int rsp_bn_stats(const void* x, int64_t M, int32_t C, float* sum, float* sumsq, void* stream);
/* mean/var from the sums; writes scale = gamma*invstd, shift = beta - mean*scale, mean, invstd, and updates
 * running_mean / running_var (unbiased) with `momentum` exactly like F.batch_norm(training=True).
 * running_* may be NULL. C_logical <= C: padded channels get scale = shift = 0.
 * clear_sums != 0: sum / sumsq are reset to zero after being read (persistent accumulators, no memset needed). */
int rsp_bn_finalize(float* sum, float* sumsq, int32_t clear_sums, int64_t count, const float* gamma,
                    const float* beta, float eps, float momentum, float* running_mean, float* running_var, float* scale,
                    float* shift, float* mean, float* invstd, int32_t C, int32_t C_logical, void* stream);
/* out = act(x*scale + shift (+ residual)); relu != 0 applies max(0,.). residual may be NULL. */
int rsp_bn_act_fwd(const void* x, const float* scale, const float* shift, const void* residual, int relu, void* out,
                   int64_t M, int32_t C, void* stream);
/* rsp_bn_finalize + rsp_bn_act_fwd in one launch.  sum / sumsq: the batch sums of x (read only); clear_sums: 2*C floats
 * zeroed for the layer's next forward (the other half of a double-buffered accumulator), may be NULL; rows: [4][C] =
 * scale, shift, mean, invstd (written for the backward). */
int rsp_bn_finalize_act_fwd(const void* x, const float* sum, const float* sumsq, float* clear_sums, int64_t count,
                            const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                            float* running_var, float* rows, const void* residual, int relu, void* out, int64_t M,
                            int32_t C, int32_t C_logical, void* stream);
/* dz = dout * (out > 0 if relu); sum_dz[c] += sum dz, sum_dz_xhat[c] += sum dz * (x-mean)*invstd (caller zeroes). */
int rsp_bn_act_bwd_reduce(const void* dout, const void* out, const void* x, const float* mean, const float* invstd,
                          int relu, float* sum_dz, float* sum_dz_xhat, int64_t M, int32_t C, void* stream);
/* dx = gamma*invstd*(dz - sum_dz/M - xhat*sum_dz_xhat/M); dres = dz (may be NULL). */
int rsp_bn_act_bwd_apply(const void* dout, const void* out, const void* x, const float* mean, const float* invstd,
                         const float* gamma, const float* sum_dz, const float* sum_dz_xhat, int relu, void* dx,
                         void* dres, int64_t M, int32_t C, int32_t C_logical, void* stream);

/* ------------------------------------------------------------------------------------------------
 * MaxPool3d NDHWC bf16 (reference: models/resnet.py:139, models/c3d.py:24,29,37,45).
 * idx: uint8 [N,To,Ho,Wo,C] window-local argmax (first maximum in (kt,kh,kw) scan order, as ATen).
 * ------------------------------------------------------------------------------------------------ */
typedef struct rsp_pool3d_desc {
  int32_t N, Ti, Hi, Wi, C;
  int32_t kt, kh, kw, st, sh, sw, pt, ph, pw;
} rsp_pool3d_desc;
int rsp_maxpool3d_fwd(const rsp_pool3d_desc* d, const void* x, void* y, uint8_t* idx, void* stream);
int rsp_maxpool3d_bwd(const rsp_pool3d_desc* d, const void* dy, const uint8_t* idx, void* dx, void* stream);

/* Fused BatchNorm(train) -> ReLU -> MaxPool3d (reference: models/resnet.py:203-206 bn1/relu/maxpool after conv1,
 * models/c3d.py:111-139 bn/relu/pool1..4).  x is the conv output (bf16 NDHWC, the pool's input geometry); scale/shift/
 * mean/invstd come from rsp_bn_finalize.  The post-activation tensor is never materialised, and neither is the pool
 * gradient at input resolution.
 *   fwd      : y bf16 [N][To][Ho][Wo][C]; idx uint8 argmax inside the window (first maximum in (kt,kh,kw) order) and
 *              xmax bf16 = x at that argmax, both [N][To][Ho][Wo][C] — pass both or, for no-grad passes, NULL for both
 *   bwd_sums : sum_dz[c] += sum_o dy*[bn(xmax) > 0], sum_dz_xhat[c] += sum_o dy*[..]*(xmax-mean)*invstd (zero on entry):
 *              the two BatchNorm-backward reductions from the POOLED tensors only
 *   bwd_dx   : dx = gamma*invstd*(dz - sum_dz/M - xhat*sum_dz_xhat/M), M = N*Ti*Hi*Wi, with dz scattered from (dy, idx)
 *              into a shared-memory tile of input rows and x streamed once; gamma has C_logical entries.
 * rsp_bn_relu_maxpool_supported: 1 when the geometry fits the kernels' shared-memory tiles (else use the unfused calls). */
int rsp_bn_relu_maxpool_supported(const rsp_pool3d_desc* d);
int rsp_bn_relu_maxpool_fwd(const rsp_pool3d_desc* d, const void* x, const float* scale, const float* shift, void* y,
                            uint8_t* idx, void* xmax, void* stream);
int rsp_bn_relu_maxpool_bwd_sums(const rsp_pool3d_desc* d, const void* dy, const void* xmax, const float* scale,
                                 const float* shift, const float* mean, const float* invstd, float* sum_dz,
                                 float* sum_dz_xhat, void* stream);
int rsp_bn_relu_maxpool_bwd_dx(const rsp_pool3d_desc* d, const void* dy, const uint8_t* idx, const void* xmax,
                               const void* x, const float* scale, const float* shift, const float* mean,
                               const float* invstd, const float* gamma, int32_t C_logical, const float* sum_dz,
                               const float* sum_dz_xhat, void* dx, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Projection heads (reference: moco/split_wrapper.py:128-152,164-169): global average pool over the
 * S = t*h*w positions, two Linear(C -> D) heads, L2 normalise each (F.normalize, eps 1e-12).
 * feat bf16 [B][S][C]; w fp32 [D][C_logical]; outputs fp32.
 * ------------------------------------------------------------------------------------------------ */
int rsp_head_fwd(const void* feat, int32_t B, int32_t S, int32_t C, int32_t C_logical, int32_t D, const float* w1,
                 const float* b1, const float* w2, const float* b2, float* pooled, float* raw1, float* raw2,
                 float* out1, float* out2, void* stream);
/* dw*, db*: += (caller owns zeroing); dfeat bf16 [B][S][C] overwritten (may be NULL); dr_ws: fp32 [B][2][D] scratch. */
int rsp_head_bwd(const float* dout1, const float* dout2, const float* pooled, const float* raw1, const float* raw2,
                 int32_t B, int32_t S, int32_t C, int32_t C_logical, int32_t D, const float* w1, const float* w2,
                 float* dr_ws, float* dw1, float* db1, float* dw2, float* db2, void* dfeat, void* stream);

/* ------------------------------------------------------------------------------------------------
 * S3D-G (reference: models/s3dg.py). Self-gating of sep_conv (:54-72): gate = sigmoid(W * mean_S(x) + b), y = x*gate.
 * x, y, dy, dx bf16 [N][S][C]; w fp32 [C_logical][C_logical] (the 1x1x1 excitation conv), b fp32 [C_logical];
 * sums_ws / dgate_ws / dadd_ws fp32 [N][C] scratch; pooled fp32 [N][C_logical]; gate fp32 [N][C]; dw, db: +=.
 * rsp_copy_channels moves a channel range between NDHWC tensors (inception concat :96 and its backward).
 * ------------------------------------------------------------------------------------------------ */
int rsp_gate_fwd(const void* x, int32_t N, int32_t S, int32_t C, int32_t C_logical, const float* w, const float* b,
                 float* sums_ws, float* pooled, float* gate, void* y, void* stream);
int rsp_gate_bwd(const void* dy, const void* x, int32_t N, int32_t S, int32_t C, int32_t C_logical, const float* w,
                 const float* pooled, const float* gate, float* dgate_ws, float* dadd_ws, float* dw, float* db,
                 void* dx, void* stream);
int rsp_copy_channels(const void* src, int32_t c_src, int32_t src_off, void* dst, int32_t c_dst, int32_t dst_off,
                      int32_t n_ch, int64_t M, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Layout: fp32 NCDHW [N][C][T][H][W] <-> bf16 NDHWC [N][T][H][W][Cs] (zero padded channels).
 * ------------------------------------------------------------------------------------------------ */
int rsp_ncdhw_to_ndhwc_bf16(const float* x, void* y, int32_t N, int32_t C, int32_t Cs, int64_t THW, void* stream);
int rsp_ndhwc_bf16_to_ncdhw(const void* x, float* y, int32_t N, int32_t C, int32_t Cs, int64_t THW, void* stream);

/* ------------------------------------------------------------------------------------------------
 * On-GPU clip pipeline (reference: datasets/transforms_video/transforms_tensor.py:214-233 per-clip loop applying
 * ToTensorVideo, Resize (transforms_spatial.py:16-25, bilinear, align_corners=False), RandomGrayScale
 * (transforms_tensor.py:13-31), RandomHorizontalFlipVideo, NormalizeVideo; crop boxes from RawVideoRandomCrop
 * (transforms_spatial.py:42-83), frame indices from RandomStrideCrop (transforms_temporal.py:25-50)).
 * frames: uint8 [F][Hs][Ws][3] decoded frame pool; frame_idx int32 [n_clips][T] rows of the pool;
 * box int32 [n_clips][4] = (i, j, h, w); flags uint8 [n_clips]: bit0 = horizontal flip, bit1 = grayscale;
 * mean3/std3: HOST pointers to 3 floats. layout 0: out fp32 [n_clips][3][T][S][S]; 1: bf16 [n_clips][T][S][S][4].
 * ------------------------------------------------------------------------------------------------ */
int rsp_clip_sample(const uint8_t* frames, const int32_t* frame_idx, const int32_t* box, const uint8_t* flags,
                    const float* mean3, const float* std3, int32_t n_clips, int32_t T, int32_t Hs, int32_t Ws,
                    int32_t S, int32_t layout, void* out, void* stream);
/* The same with ColorJitter (transforms_tensor.py:54-145; colour kernels functional_tensor.py:88-162,253-415) between
 * RandomGrayScale and the flip: jitter is a device table of n_clips records {float factor[4]; uint8 order[4]} — factors
 * of brightness / contrast / saturation / hue and the order of application (op ids 0..3 in that sequence, 255 = skip);
 * gray_sums: fp32 [n_clips] workspace (the contrast op blends with the clip-wide mean gray level, taken by a first pass). */
int rsp_clip_sample_jitter(const uint8_t* frames, const int32_t* frame_idx, const int32_t* box, const uint8_t* flags,
                           const void* jitter, float* gray_sums, const float* mean3, const float* std3, int32_t n_clips,
                           int32_t T, int32_t Hs, int32_t Ws, int32_t S, int32_t layout, void* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * MoCo / RSP objective (reference: moco/builder_diffspeed_diffloss.py)
 * ------------------------------------------------------------------------------------------------ */
/* _momentum_update_key_encoder (:337-343) over a flat parameter buffer: k = k*m + q*(1-m), with the
 * reference's rounding (two fp32 products, one fp32 add; `one_minus_m` is float(1.0 - m) from double). */
int rsp_ema_update(float* k, const float* q, int64_t n, float m, float one_minus_m, void* stream);
/* SGD(momentum, dampening 0, no nesterov) + weight decay over flat buffers (pretrain.py:65-72,163-165).
 * first_step != 0: momentum buffer is initialised with the gradient (torch.optim.SGD semantics).
 * grad_scale multiplies the gradient first (1/world_size after a sum all-reduce). */
int rsp_sgd_step(float* p, const float* grad, float* mom, int64_t n, float lr, float momentum, float weight_decay,
                 float grad_scale, int first_step, void* stream);
/* _diff_speed (:421-447). im_q/im_k fp32 NCDHW [B][C][T][H][W]; perm int64 [B] (the randperm);
 * rows perm[:n_s1] keep speed 1 for q/k and speed d for k_neg, the rest the other way round.
 * layout 0: outputs fp32 NCDHW [B][C][T/d][H][W] (reference tensors);
 * layout 1: outputs bf16 NDHWC [B][T/d][H][W][4] (conv-ready, C must be 3). */
int rsp_speed_gather(const float* im_q, const float* im_k, const int64_t* perm, int32_t B, int32_t C, int32_t T,
                     int32_t H, int32_t W, int32_t n_s1, int32_t d, int32_t layout, void* out_q, void* out_k,
                     void* out_kneg, void* stream);
/* Batched row gather: dst[i] = src[index[i]] for rows of `row_bytes` (multiple of 16) — the local half of
 * _batch_shuffle_ddp / _batch_unshuffle_ddp (:361-406) and the pack step of the permutation exchange. */
int rsp_gather_rows(const void* src, const int64_t* index, void* dst, int64_t n_rows, int64_t row_bytes, void* stream);
/* Shuffle-BN over NVLink peer memory (_batch_shuffle_ddp, :361-387: all_gather + x_gather[idx_this]).  Each rank keeps
 * its key clips in a buffer its peers have mapped; one kernel per rank pulls the B rows the permutation assigns to it
 * out of the owners' memory.  These five calls are the only ones that allocate: a peer buffer must be its own
 * cudaMalloc allocation for the IPC handle, so the library owns it (rsp_peer_alloc / rsp_peer_free).
 *   rsp_peer_export : 64-byte cudaIpcMemHandle_t of a buffer from rsp_peer_alloc (exchanged between ranks by the caller)
 *   rsp_peer_open   : maps a peer's buffer into this process (cudaIpcMemLazyEnablePeerAccess); rsp_peer_close unmaps
 *   rsp_gather_rows_peer : dst[i] = peer_bases[index[i] / rows_per_peer] row (index[i] % rows_per_peer);
 *                          peer_bases is a DEVICE array of W base pointers (own buffer at [rank]), index int64 on
 *                          the device (the permutation slice idx_shuffle.view(W,-1)[rank]), rows of row_bytes (% 16).
 *   rsp_invert_permutation : inv[perm[i]] = i — idx_unshuffle = argsort(idx_shuffle) (:381) without a sort. */
int rsp_peer_alloc(int64_t bytes, void** dev_ptr);
int rsp_peer_free(void* dev_ptr);
int rsp_peer_export(void* dev_ptr, uint8_t* handle64);
int rsp_peer_open(const uint8_t* handle64, void** mapped);
int rsp_peer_close(void* mapped);
int rsp_gather_rows_peer(const void* const* peer_bases, const int64_t* index, void* dst, int64_t n_rows,
                         int32_t rows_per_peer, int64_t row_bytes, void* stream);
int rsp_invert_permutation(const int64_t* perm, int64_t* inv, int32_t n, void* stream);
/* _dequeue_and_enqueue (:345-359): queue[:, ptr:ptr+n] = keys.T; ptr = (ptr+n) % K, ptr read/written on device.
 * queue fp32 [D][K]; keys fp32 [n][D]; queue_ptr int64[1]. K % n must be 0. */
int rsp_queue_enqueue(float* queue, const float* keys, int64_t* queue_ptr, int32_t D, int32_t K, int32_t n,
                      void* stream);
/* Logits of forward() (:521-536) fused with the row-wise logsumexp needed by the two cross-entropies.
 * q/k fp32 [N][D]; queue fp32 [D][K]; logits1/2 fp32 [N][1+K] (may be NULL: not materialised);
 * lpos_m/lneg_m/pos1/pos2 fp32 [N] (pos_i = logits_i[:,0]); lse1/lse2 fp32 [N]; every dot product is divided by
 * `temperature` like the reference (`/= self.T`); workspace fp32 [N][ceil(K/256)][2] (rsp_moco_logits_workspace). */
int64_t rsp_moco_logits_workspace(int32_t N, int32_t K);
int rsp_moco_logits_fwd(const float* q_a, const float* q_m, const float* k_a, const float* k_m, const float* kn_a,
                        const float* kn_m, const float* queue, int32_t N, int32_t D, int32_t K, float temperature,
                        float* logits1, float* logits2, float* lpos_m, float* lneg_m, float* lse1, float* lse2,
                        float* pos1, float* pos2, float* workspace, void* stream);
/* The same, also counting per row how many queue negatives beat each positive: ranks int32 [2][N] (ranks[0][n] =
 * #{k : logits1[n][1+k] > logits1[n][0]}, ranks[1] for logits2).  accuracy(output, target, topk=(1,5)) of
 * pretrain.py:169-175 (framework/metrics/classification.py:6-20) is then rank == 0 / rank < 5 — no top-k pass over
 * [N][1+K] and no materialised logits. */
int rsp_moco_logits_fwd_ranked(const float* q_a, const float* q_m, const float* k_a, const float* k_m, const float* kn_a,
                               const float* kn_m, const float* queue, int32_t N, int32_t D, int32_t K, float temperature,
                               float* logits1, float* logits2, float* lpos_m, float* lneg_m, float* lse1, float* lse2,
                               float* pos1, float* pos2, float* workspace, int32_t* ranks, void* stream);
/* Device-side AverageMeters (framework/meters/average.py:4-44) of the eight values pretrain.py:169-196 logs, in the
 * order Loss, Loss_A, Acc@1_A, Acc@5_A, Acc@1_A_n, Acc@5_A_n, Loss_M, Acc@1_M: meters = float val[8], float sum[8],
 * int32 count.  loss3 = Loss.forward's triple; val = this step's value, sum += val * N, count += N.  No host sync. */
int rsp_metrics_update(const float* loss3, const int32_t* ranks, const float* lpos_m, const float* lneg_m, int32_t N,
                       float* meters, void* stream);
/* Gradient w.r.t. q_a / q_m given per-row gradients of lse1, lse2, pos1, pos2, lpos_m, lneg_m and (optionally,
 * may be NULL) dense gradients of the materialised logits. dq_a/dq_m overwritten. */
int rsp_moco_logits_bwd(const float* q_a, const float* q_m, const float* k_a, const float* k_m, const float* kn_a,
                        const float* kn_m, const float* queue, int32_t N, int32_t D, int32_t K, float temperature,
                        const float* pos1, const float* pos2, const float* lse1, const float* lse2,
                        const float* g_lse1, const float* g_lse2,
                        const float* g_pos1, const float* g_pos2, const float* g_lpos_m, const float* g_lneg_m,
                        const float* g_logits1, const float* g_logits2, float* dq_a, float* dq_m, void* stream);
/* Loss.forward (:263-283): out[0] = A*(ce1+ce2) + M*rank, out[1] = ce1+ce2, out[2] = rank with
 * ce_i = mean(lse_i - pos_i), rank = mean(max(0, -(lpos_m - lneg_m) + margin)). */
int rsp_moco_loss_fwd(const float* lse1, const float* lse2, const float* pos1, const float* pos2, const float* lpos_m,
                      const float* lneg_m, int32_t N, float margin, float A, float M, float* out3, void* stream);
/* g_out3: gradients of the three outputs. Writes the six per-row gradient vectors. */
int rsp_moco_loss_bwd(const float* lpos_m, const float* lneg_m, int32_t N, float margin, float A, float M,
                      const float* g_out3, float* g_lse1, float* g_lse2, float* g_pos1, float* g_pos2,
                      float* g_lpos_m, float* g_lneg_m, void* stream);
/* Dense cross entropy with target class 0 over logits [N][L] (for callers that hand Loss plain tensors):
 * lse[n] and, for bwd, dlogits = g * (softmax - onehot0) / N. */
int rsp_ce0_fwd(const float* logits, int32_t N, int32_t L, float* lse, void* stream);
int rsp_ce0_bwd(const float* logits, const float* lse, int32_t N, int32_t L, const float* g_scalar, float* dlogits,
                void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RSPNET_B200_H_ */
