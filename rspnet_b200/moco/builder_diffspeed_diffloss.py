"""MoCo-style relative-speed contrastive builder — the interface of the reference's
``moco/builder_diffspeed_diffloss.py`` (``MoCoDiffLossTwoFc`` :286-547, ``MoCoDiffLoss`` :11-245,
``concat_all_gather`` :249-260, ``Loss`` :263-283) on B200 kernels:

* momentum update  -> one coalesced EMA pass over the flattened parameters          (ref :337-343)
* ``_diff_speed``   -> one gather kernel that writes q / k / k_neg conv-ready        (ref :421-447)
* shuffle-BN        -> permutation all-to-all instead of all_gather + index          (ref :361-406)
* logits + CE       -> q.queue tiles fused with the row logsumexp                    (ref :521-538, :272-283)
* enqueue           -> ring-buffer transpose write, pointer kept on the device       (ref :345-359)

Random draws are made with the same generators in the same order as the reference, so
permutations, queue pointers and gathered keys are bit-identical.
"""
import logging
import os
import random
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor, nn

from .. import nn as rnn
from .. import ops
from . import exchange

logger = logging.getLogger(__name__)


@torch.no_grad()
def concat_all_gather(tensor: Tensor) -> Tensor:
    """all_gather along dim 0 in rank order (no gradient), as the reference's helper of the same name."""
    return exchange.all_gather_rows(tensor)


# ----------------------------------------------------------------------------------------------------------------
# autograd functions of the objective
# ----------------------------------------------------------------------------------------------------------------
class _LogitsFn(torch.autograd.Function):
    """(q_A, q_M, keys, queue) -> logits1, logits2, l_pos_M, l_neg_M and the row statistics
    rows = [l_pos_M, l_neg_M, lse1, lse2, pos1, pos2] that the fused loss consumes."""

    @staticmethod
    def forward(ctx, q_a, q_m, k_a, k_m, kn_a, kn_m, queue, temperature, materialize):
        args = [t.contiguous().float() for t in (q_a, q_m, k_a, k_m, kn_a, kn_m)]
        if any(ctx.needs_input_grad[:2]):
            # the queue is overwritten in place by _dequeue_and_enqueue before backward runs; backward recomputes the
            # negative logits, so it needs the queue as it was (the reference clones it for the same reason, :529)
            queue = queue.clone()
        logits, rows, ranks = ops.moco_logits_fwd(*args, queue, temperature, materialize)
        ctx.temperature = temperature
        ctx.save_for_backward(*args, queue, rows)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(ranks)
        l1 = logits[0] if materialize else None
        l2 = logits[1] if materialize else None
        return l1, l2, rows[0].unsqueeze(-1), rows[1].unsqueeze(-1), rows, ranks

    @staticmethod
    def backward(ctx, g_l1, g_l2, g_lpm, g_lnm, g_rows, _g_ranks):
        *args, queue, rows = ctx.saved_tensors
        g = torch.zeros_like(rows) if g_rows is None else g_rows.contiguous().clone()
        if g_lpm is not None:
            g[0] += g_lpm.flatten()
        if g_lnm is not None:
            g[1] += g_lnm.flatten()
        g_l1 = g_l1.contiguous() if g_l1 is not None else None
        g_l2 = g_l2.contiguous() if g_l2 is not None else None
        dq_a, dq_m = ops.moco_logits_bwd(*args, queue, ctx.temperature, rows, g, g_l1, g_l2)
        return dq_a, dq_m, None, None, None, None, None, None, None


class _FusedLossFn(torch.autograd.Function):
    """rows -> (A*(ce1+ce2) + M*rank, ce1+ce2, rank)."""

    @staticmethod
    def forward(ctx, rows, margin, a, m):
        ctx.cfg = (margin, a, m)
        ctx.save_for_backward(rows)
        return ops.moco_loss_fwd(rows, margin, a, m)

    @staticmethod
    def backward(ctx, g3):
        (rows,) = ctx.saved_tensors
        return ops.moco_loss_bwd(rows, *ctx.cfg, g3.contiguous()), None, None, None


class _DenseCE0Fn(torch.autograd.Function):
    """mean cross entropy against class 0 over arbitrary dense logits."""

    @staticmethod
    def forward(ctx, logits):
        logits = logits.contiguous().float()
        lse = ops.ce0_fwd(logits)
        ctx.save_for_backward(logits, lse)
        return (lse - logits[:, 0]).mean()

    @staticmethod
    def backward(ctx, g):
        logits, lse = ctx.saved_tensors
        return ops.ce0_bwd(logits, lse, g.reshape(1).contiguous().float())


class Loss(nn.Module):
    """``A * (CE(logits1) + CE(logits2)) + M * MarginRanking(l_pos_M, l_neg_M)``; returns (loss, ce_sum, ranking)."""

    def __init__(self, margin=1.0, A: float = 1.0, M: float = 1.0):
        super().__init__()
        self.A = A
        self.M = M
        self.margin = margin

    def forward(self, output: Tuple[Tensor, Tensor], target: Tensor, ranking_logits: Tuple[Tensor, Tensor],
                ranking_target: Tensor):
        rows = getattr(output[0], "_rsp_rows", None)
        fused = (rows is not None and getattr(output[1], "_rsp_rows", None) is rows and
                 getattr(ranking_logits[0], "_rsp_rows", None) is rows and
                 getattr(ranking_logits[1], "_rsp_rows", None) is rows)
        if fused:
            # target == 0 and ranking_target == 1 by construction of MoCoDiffLoss*.forward
            out3 = _FusedLossFn.apply(rows, self.margin, self.A, self.M)
            return out3[0], out3[1], out3[2]
        # plain tensors: dense cross-entropy kernels + the ranking term on an assembled statistics block
        if not bool((target == 0).all()) or not bool((ranking_target == 1).all()):
            raise NotImplementedError("rspnet_b200.Loss: only the MoCo targets (class 0 / ranking +1) are supported")
        ce = _DenseCE0Fn.apply(output[0]) + _DenseCE0Fn.apply(output[1])
        lp, ln = ranking_logits[0].flatten().float(), ranking_logits[1].flatten().float()
        z = torch.zeros_like(lp)
        rank = _FusedLossFn.apply(torch.stack([lp, ln, z, z, z, z]), self.margin, 0.0, 1.0)[2]
        return self.A * ce + self.M * rank, ce, rank


# ----------------------------------------------------------------------------------------------------------------
# flat parameter storage (one buffer per encoder so EMA / SGD / all-reduce are single passes)
# ----------------------------------------------------------------------------------------------------------------
def flatten_parameters(module: nn.Module) -> Tensor:
    """Moves every parameter of ``module`` into one contiguous fp32 buffer (16-byte aligned segments) and returns it.
    Parameters keep their names, shapes and values; ``p.data`` becomes a view."""
    params = list(module.parameters())
    sizes = [(p.numel() + 3) // 4 * 4 for p in params]
    flat = torch.zeros(sum(sizes), dtype=torch.float32, device=params[0].device)
    off = 0
    for p, n in zip(params, sizes):
        view = flat[off:off + p.numel()].view_as(p)
        view.copy_(p.data)
        p.data = view
        off += n
    return flat


class _MoCoBase(nn.Module):
    """State and kernels shared by the two builder variants."""

    def _init_common(self, base_encoder, dim, K, m, T, diff_speed, mlp=False):
        if dim not in (64, 128, 256):
            # the logits backward kernel is instantiated for these feature dimensions (csrc/moco.cu); fail at
            # construction, not inside loss.backward()
            raise ValueError(f"rspnet_b200: moco.dim must be 64, 128 or 256 (got {dim})")
        self.K = K
        self.m = m
        self.T = T
        self.diff_speed = diff_speed
        logger.warning('Using diffspeed: %s', self.diff_speed)
        self.encoder_q = base_encoder(num_classes=dim)
        self.encoder_k = base_encoder(num_classes=dim)
        if mlp:  # ref :45-48 / :318-321 — the reference's brute-force replacement of the backbone's ``fc``
            dim_mlp = self.encoder_q.fc.weight.shape[1]
            self.encoder_q.fc = nn.Sequential(nn.Linear(dim_mlp, dim_mlp), nn.ReLU(), self.encoder_q.fc)
            self.encoder_k.fc = nn.Sequential(nn.Linear(dim_mlp, dim_mlp), nn.ReLU(), self.encoder_k.fc)
        for param_q, param_k in zip(self.encoder_q.parameters(), self.encoder_k.parameters()):
            param_k.data.copy_(param_q.data)
            param_k.requires_grad = False
        self.register_buffer("queue", torch.randn(dim, K))
        self.queue = nn.functional.normalize(self.queue, dim=0)
        self.register_buffer("queue_ptr", torch.zeros(1, dtype=torch.long))
        self.alpha = 0.5
        self.materialize_logits = True   # forward() returns [N, 1+K] logits like the reference
        # key-encoder passes on a side stream next to the query pass (forward())
        self.overlap_key_passes = os.environ.get("RSP_KEY_OVERLAP", "1") != "0"
        self._side_stream = None
        self._pull_stream = None
        self._exchanges = {}             # (rows, row shape, dtype) -> exchange.ShuffleExchange (key-clip buffers + transport)
        self._flat_q: Optional[Tensor] = None
        self._flat_k: Optional[Tensor] = None
        assert self.diff_speed is not None, "This branch is for diff speed"

    # -- flat buffers ------------------------------------------------------------------------------------------
    def _ensure_flat(self):
        p0 = next(self.encoder_q.parameters())
        stale = (self._flat_q is None or self._flat_q.device != p0.device or
                 p0.data_ptr() != self._flat_q.data_ptr())
        if stale:
            self._flat_q = flatten_parameters(self.encoder_q)
            self._flat_k = flatten_parameters(self.encoder_k)
            for enc in (self.encoder_q, self.encoder_k):
                enc._rsp_bn_counters = rnn.batch_bn_counters(enc)
            rnn.bump_weight_epoch()

    def flat_parameters(self) -> Tuple[Tensor, Tensor]:
        """(flat encoder_q parameters, flat encoder_k parameters)."""
        self._ensure_flat()
        return self._flat_q, self._flat_k

    @torch.no_grad()
    def _momentum_update_key_encoder(self):
        self._ensure_flat()
        ops.ema_update_(self._flat_k, self._flat_q, self.m)
        rnn.bump_weight_epoch()
        rnn.refresh_packed_weights()   # both encoders are final for this step: one batched re-pack

    @torch.no_grad()
    def _dequeue_and_enqueue(self, keys, gathered: bool = False):
        if not gathered:
            keys = concat_all_gather(keys)
        assert self.K % keys.shape[0] == 0  # for simplicity
        ops.queue_enqueue_(self.queue, keys.float(), self.queue_ptr)

    # -- shuffle-BN --------------------------------------------------------------------------------------------
    def _exchange_for(self, shape, dtype, device) -> "exchange.ShuffleExchange":
        """The exchange (key-clip buffers + transport) for per-rank batches of this shape ([B, ...])."""
        key = (tuple(shape), dtype, device)
        ex = self._exchanges.get(key)
        if ex is None:
            ex = self._exchanges[key] = exchange.ShuffleExchange(shape[0], shape[1:], dtype, device)
        return ex

    @torch.no_grad()
    def _batch_shuffle_ddp(self, x, idx_shuffle=None, idx_host=None, slot=exchange.ShuffleExchange.SLOT_KNEG):
        """ref :361-387.  ``x``: this rank's batch; returns (rows ``concat_all_gather(x)[idx_shuffle.view(W,-1)[rank]]``,
        idx_unshuffle).  forward() pre-draws both permutations of the step (same generator, same order as the reference)
        and passes ``idx_shuffle`` (int64 [B*W], on the device, already published); called without it this draws and
        publishes its own permutation like the reference method does."""
        ex = self._exchange_for(x.shape, x.dtype, x.device)
        if idx_shuffle is None:
            if not ex.holds(x, slot):
                ex.begin_step()[slot].copy_(x)
            idx_all, idx_hosts = ex.draw(x.shape[0] * ex.world, count=1)
            ex.publish(idx_all)
            idx_shuffle, idx_host = idx_all[0], (idx_hosts[0] if idx_hosts is not None else None)
        elif not ex.holds(x, slot):
            raise RuntimeError("rspnet_b200: pre-published permutations need the clips in the exchange buffers")
        idx_unshuffle = ops.invert_permutation(idx_shuffle) if idx_shuffle.is_cuda else torch.argsort(idx_shuffle)
        return ex.pull(slot, idx_shuffle, idx_host, ops.gather_rows), idx_unshuffle

    @torch.no_grad()
    def _batch_unshuffle_ddp(self, x, idx_unshuffle, return_all: bool = False):
        rank, world = exchange.world_info()
        x_gather = concat_all_gather(x)
        restored = ops.gather_rows(x_gather, idx_unshuffle.to(x.device, non_blocking=True))  # global original order
        b = x.shape[0]
        mine = restored[rank * b:(rank + 1) * b]
        return (mine, restored) if return_all else mine

    @torch.no_grad()
    def _forward_encoder_k(self, im_k, return_all: bool = False, idx_shuffle=None, idx_host=None,
                           slot=exchange.ShuffleExchange.SLOT_KNEG, shuffled=None):
        if shuffled is None:
            im_k, idx_unshuffle = self._batch_shuffle_ddp(rnn.as_ndhwc(im_k), idx_shuffle, idx_host, slot)
        else:
            im_k, idx_unshuffle = shuffled
        k_a, k_m = self.encoder_k(im_k)
        both = torch.cat([k_a, k_m], dim=1)  # one gather for both heads
        res = self._batch_unshuffle_ddp(both, idx_unshuffle, return_all)
        d = k_a.shape[1]
        if return_all:
            mine, everyone = res
            return mine[:, :d].contiguous(), mine[:, d:].contiguous(), everyone[:, :d].contiguous()
        return res[:, :d].contiguous(), res[:, d:].contiguous()

    @torch.no_grad()
    def draw_step(self, im_q: Tensor, static: Optional[dict] = None):
        """Every random draw of one forward(), in the reference's order and from the same generators: randperm(B) on the
        input's device (ref :424), random.choice(diff_speed) (:426), then the two CPU randperm(B*W) of the k_neg and k
        shuffles (:375; nothing else touches those generators in between in the reference either).  Returns
        (random_indices, diff_speed, idx_all [2, B*W] on the device, host copy or None).  ``static``: persistent device
        tensors {"random_indices", "idx_all"} to fill (what a captured CUDA graph of the step reads)."""
        B = im_q.shape[0]
        random_indices = torch.randperm(B, device=im_q.device)
        diff_speed = int(random.choice(self.diff_speed))
        shape = (B, im_q.shape[2] // diff_speed, im_q.shape[3], im_q.shape[4], 4)
        ex = self._exchange_for(shape, torch.bfloat16, im_q.device)
        if static is not None:
            static["random_indices"].copy_(random_indices)
            random_indices = static["random_indices"]
        idx_all, idx_host = ex.draw(B * ex.world, count=2, out=None if static is None else static["idx_all"])
        return random_indices, diff_speed, idx_all, idx_host

    @torch.no_grad()
    def _speed_views(self, im_q: Tensor, im_k: Tensor, into_exchange: bool = False, random_indices=None,
                     diff_speed=None):
        """ref :421-443: draws randperm(B) on the input's device and random.choice(diff_speed), then re-samples.
        into_exchange: the k / k_neg clips are written straight into this step's exchange buffers (returned as such)."""
        B = im_q.shape[0]
        if random_indices is None:
            random_indices = torch.randperm(B, device=im_q.device)
            diff_speed = int(random.choice(self.diff_speed))
        outs = None
        if into_exchange:
            shape = (B, im_q.shape[2] // diff_speed, im_q.shape[3], im_q.shape[4], 4)   # conv-ready bf16 NDHWC(4)
            kneg_buf, k_buf = self._exchange_for(shape, torch.bfloat16, im_q.device).begin_step()
            outs = (None, k_buf, kneg_buf)
        return ops.speed_gather(im_q.float(), im_k.float(), random_indices, int(B * self.alpha), diff_speed, 1, outs)

    def _logits(self, q_a, q_m, k_a, k_m, kn_a, kn_m):
        l1, l2, lpm, lnm, rows, ranks = _LogitsFn.apply(q_a, q_m, k_a, k_m, kn_a, kn_m, self.queue, self.T,
                                                        self.materialize_logits)
        if l1 is None:  # statistics-only mode: hand the Loss something to hold on to
            l1, l2 = rows[2].unsqueeze(-1), rows[3].unsqueeze(-1)
        for t in (l1, l2, lpm, lnm):
            t._rsp_rows = rows
            t._rsp_ranks = ranks     # [2][N] negatives beating each positive: accuracy without top-k (meters.py)
        l1._rsp_slot, l2._rsp_slot = 0, 1
        return (l1, l2), (lpm, lnm)


class MoCoDiffLossTwoFc(_MoCoBase):
    def __init__(self, base_encoder, dim=128, K=65536, m=0.999, T=0.07, mlp=False,
                 diff_speed: Optional[List[int]] = None):
        """
        dim: feature dimension (default: 128); K: queue size; m: momentum of the key encoder; T: softmax temperature
        """
        super().__init__()
        # mlp=True rewires ``encoder.fc``; with the two-head MultiTaskWrapper (no ``fc``) it raises AttributeError in
        # the reference (:318-321) and does the same here
        self._init_common(base_encoder, dim, K, m, T, diff_speed, mlp)
        self.gap = nn.AdaptiveAvgPool3d((1, 1, 1))

    @torch.no_grad()
    def _diff_speed(self, im_q: Tensor, im_k: Tensor):
        q, k, k_neg = self._speed_views(im_q, im_k)
        kn_a, kn_m, kn_a_all = self._forward_encoder_k(k_neg, return_all=True, slot=exchange.ShuffleExchange.SLOT_KNEG)
        self._enqueue_payload = kn_a_all
        return q, k, kn_a, kn_m

    def forward(self, im_q, im_k, draws=None):
        """
        Input: im_q, im_k: [B, 3, T, H, W] clips (T = diff_speed * clip length).
        Output: (logits1, logits2), labels_A (zeros), (l_pos_M, l_neg_M), labels_M (ones) — as the reference.
        ``draws``: the result of ``draw_step`` when the caller made the step's random draws itself (CUDA-graph replays).
        """
        if not (self.overlap_key_passes and im_q.is_cuda):
            with torch.no_grad():
                self._momentum_update_key_encoder()
                im_q, im_k, k_neg_A, k_neg_M = self._diff_speed(im_q, im_k)
                k_A, k_M = self._forward_encoder_k(im_k, slot=exchange.ShuffleExchange.SLOT_K)
            q_A, q_M = self.encoder_q(im_q)
        else:
            # The two key-encoder passes (no grad) and the query pass are independent: the key passes go to a side
            # stream so that the small layer3 / layer4 kernels of one pass fill the SMs the other leaves idle.  The
            # random draws are made up front in the reference's order (speed perm, diff_speed choice, shuffle of k_neg,
            # shuffle of k — nothing else touches those generators in between in the reference either), so that both
            # permutations travel in ONE small all-reduce that also orders every peer's pull after every rank's clips.
            main = torch.cuda.current_stream()
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream()
                self._pull_stream = torch.cuda.Stream()
            side, pull = self._side_stream, self._pull_stream
            SL = exchange.ShuffleExchange
            with torch.no_grad():
                self._momentum_update_key_encoder()
                random_indices, diff_speed, idx_all, idx_host = draws if draws is not None else self.draw_step(im_q)
                im_q, im_k, k_neg = self._speed_views(im_q, im_k, into_exchange=True, random_indices=random_indices,
                                                      diff_speed=diff_speed)
                ex = self._exchange_for(k_neg.shape, k_neg.dtype, k_neg.device)
            # No Tensor.record_stream here: the exchange buffers are persistent, and what the side streams allocate is
            # only reused by them after their next wait_stream(main).  (record_stream defers block reuse to event
            # polling; the caching allocator then grows with cudaMalloc in the middle of training steps.)
            side.wait_stream(main)
            with torch.cuda.stream(side), torch.no_grad():
                ex.publish(idx_all)
                pull.wait_stream(side)
                with torch.cuda.stream(pull):   # the k rows cross NVLink while the k_neg pass runs
                    k_shuffled = self._batch_shuffle_ddp(im_k, idx_all[1], None if idx_host is None else idx_host[1],
                                                         SL.SLOT_K)
                k_neg_A, k_neg_M, self._enqueue_payload = self._forward_encoder_k(
                    k_neg, return_all=True, idx_shuffle=idx_all[0], idx_host=None if idx_host is None else idx_host[0],
                    slot=SL.SLOT_KNEG)
                side.wait_stream(pull)
                k_A, k_M = self._forward_encoder_k(None, shuffled=k_shuffled)
            q_A, q_M = self.encoder_q(im_q)
            main.wait_stream(side)
        logits_A, logits_M = self._logits(q_A, q_M, k_A, k_M, k_neg_A, k_neg_M)
        labels_A = torch.zeros(q_A.shape[0], dtype=torch.long, device=q_A.device)
        labels_M = torch.ones_like(labels_A)
        # the unshuffle gather of the k_neg pass already holds every rank's keys in rank order
        self._dequeue_and_enqueue(self._enqueue_payload, gathered=True)
        self._enqueue_payload = None
        return logits_A, labels_A, logits_M, labels_M

    @torch.no_grad()
    def cam_visualize(self, im_q, im_k):
        """CAM maps (ref :449-490). Inference-only tooling: the contractions are tiny and use torch.einsum."""
        im_q, im_k, _, _ = self._diff_speed(im_q, im_k)
        self._forward_encoder_k(im_k)
        k_F = self.encoder_k._get_last_feature()
        self.encoder_q(im_q)
        q_F = self.encoder_q._get_last_feature()
        q_X = q_F.mean(dim=(2, 3, 4))
        k_X = k_F.mean(dim=(2, 3, 4))
        q_wA, q_wM = self.encoder_q._get_fc_weight()
        k_wA, k_wM = self.encoder_k._get_fc_weight()

        def cam(w_other, x_other, w_self, f_self):
            return torch.einsum('bc,bcthw->bthw', torch.einsum('bn,nc->bc', torch.einsum('nc,bc->bn', w_other, x_other),
                                                                w_self), f_self)
        return (cam(k_wA, k_X, q_wA, q_F), cam(k_wM, k_X, q_wM, q_F), cam(q_wA, q_X, k_wA, k_F),
                cam(q_wM, q_X, k_wM, k_F))


class MoCoDiffLoss(_MoCoBase):
    """Single-head variant (ref :11-245): the encoder returns one embedding, normalised here; the ranking pair is
    (q.k, q.speed_k) and the enqueued key is k."""

    def __init__(self, base_encoder, dim=128, K=65536, m=0.999, T=0.07, mlp=False,
                 diff_speed: Optional[List[int]] = None):
        super().__init__()
        self._init_common(base_encoder, dim, K, m, T, diff_speed, mlp)

    @torch.no_grad()
    def _forward_encoder_k(self, im_k, return_all: bool = False):
        """ref :171-183: shuffle, encoder_k, L2 normalise, unshuffle (one embedding per clip)."""
        im_k, idx_unshuffle = self._batch_shuffle_ddp(rnn.as_ndhwc(im_k))
        k = nn.functional.normalize(self.encoder_k(im_k).float(), dim=1)
        return self._batch_unshuffle_ddp(k, idx_unshuffle, return_all)

    @torch.no_grad()
    def _diff_speed(self, im_q: Tensor, im_k: Tensor):
        """ref :136-162: re-sampled q / k clips and the embedding of the speed-swapped key clip."""
        q, k, k_neg = self._speed_views(im_q, im_k)
        return q, k, self._forward_encoder_k(k_neg)

    def forward(self, im_q, im_k):
        """
        Input: im_q, im_k: [B, 3, T, H, W] clips (T = diff_speed * clip length).
        Output (ref :184-245): (logits1, logits2), labels (zeros), (l_pos, l_neg_speed), ranking_target (ones) with
        logits1 = [q.k | q.queue] / T, logits2 = [q.speed_k | q.queue] / T; the key ``k`` is enqueued.
        """
        with torch.no_grad():
            self._momentum_update_key_encoder()
            im_q, im_k, speed_k = self._diff_speed(im_q, im_k)
            k, k_all = self._forward_encoder_k(im_k, return_all=True)
        q = nn.functional.normalize(self.encoder_q(im_q).float(), dim=1)
        # the two-head kernel with both heads fed by the same embedding: logits1 = [q.k | q.queue], logits2 =
        # [q.speed_k | q.queue], ranking pair (q.k, q.speed_k); autograd sums the two gradient paths into q
        logits, ranking_logits = self._logits(q, q, k, k, speed_k, speed_k)
        labels = torch.zeros(q.shape[0], dtype=torch.long, device=q.device)
        ranking_target = torch.ones_like(labels)
        self._dequeue_and_enqueue(k_all, gathered=True)
        return logits, labels, ranking_logits, ranking_target
