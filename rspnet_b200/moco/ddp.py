"""Data-parallel wrapper with the surface the reference's training loop uses from
``nn.parallel.DistributedDataParallel`` (moco/__init__.py:49-53; pretrain.py:154-165,249-260): ``__call__``,
``.module``, ``parameters()``, ``train()``.

Differences by design:
* gradients of encoder_q are staged in ONE flat fp32 buffer and summed with bucketed NCCL all-reduces launched from
  autograd hooks on a side stream, so communication overlaps the remaining wgrad/dgrad kernels;
* the per-forward buffer broadcast is dropped — queue / queue_ptr are bit-identical on all ranks by construction and
  BN running statistics do not influence train-mode outputs (rank 0's are what a checkpoint stores, as in the
  reference);
* parameters that receive no gradient (``encoder.fc``; hence find_unused_parameters=True upstream) keep
  ``grad = None`` so that torch.optim.SGD skips them exactly as it does in the reference.
"""
import os
from typing import List

import torch
import torch.distributed as dist
from torch import nn

from . import exchange


class FlatDDP(nn.Module):
    def __init__(self, module: nn.Module, bucket_bytes: int = 32 << 20):
        super().__init__()
        bucket_bytes = int(float(os.environ.get("RSP_DDP_BUCKET_MB", bucket_bytes / (1 << 20))) * (1 << 20))
        self.module = module
        self.rank, self.world = exchange.world_info()
        flat_q, flat_k = module.flat_parameters()
        if self.world > 1:
            # what DDP's constructor does: rank 0's parameters and buffers everywhere
            dist.broadcast(flat_q, src=0)
            dist.broadcast(flat_k, src=0)
            for b in module.buffers():
                dist.broadcast(b, src=0)
        self.flat_grad = torch.zeros_like(flat_q)
        self._params = [p for p in module.encoder_q.parameters()]
        self._views, self._bounds = [], []
        off = 0
        for p in self._params:
            self._views.append(self.flat_grad[off:off + p.numel()].view_as(p))
            self._bounds.append((off, off + (p.numel() + 3) // 4 * 4))
            off = self._bounds[-1][1]
        # buckets over the flat buffer, filled from the back (backward produces the last layers first)
        self._buckets: List[List[int]] = []
        cur, cur_bytes = [], 0
        for i in reversed(range(len(self._params))):
            cur.append(i)
            cur_bytes += self._params[i].numel() * 4
            if cur_bytes >= bucket_bytes:
                self._buckets.append(cur)
                cur, cur_bytes = [], 0
        if cur:
            self._buckets.append(cur)
        self._bucket_of = {i: b for b, idxs in enumerate(self._buckets) for i in idxs}
        self._pending = [0] * len(self._buckets)
        self._fired = set()
        self._armed = False
        self._comm_stream = torch.cuda.Stream() if (self.world > 1 and flat_q.is_cuda) else None
        self.used_parameter_ids = None  # indices that received gradients in the last backward
        # Optimizer-in-backward: ``bucket_hook(segments)`` is called once per bucket, on the stream on which that
        # bucket's gradient becomes final (after its all-reduce), with the [lo, hi) ranges of the flat buffers whose
        # parameters received gradients.  ``defer_average`` leaves the SUM in the flat gradient (the hook's optimizer
        # multiplies by 1/world itself) instead of spending one elementwise pass per bucket on the mean.
        self.bucket_hook = None
        self.defer_average = False
        self._hook_stream = None
        for i, p in enumerate(self._params):
            p.register_post_accumulate_grad_hook(self._make_hook(i))
        from .. import nn as rnn
        rnn.register_grad_slots(self._params, self._views)

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    # ------------------------------------------------------------------------------------------------------
    def _arm(self):
        self._armed = True
        self._fired = set()
        self._pending = [len(b) for b in self._buckets]
        torch.autograd.Variable._execution_engine.queue_callback(self._finalize)

    def _make_hook(self, i):
        def hook(p):
            if not self._armed:
                self._arm()
            view = self._views[i]
            if p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
                p.grad = view
            self._fired.add(i)
            b = self._bucket_of[i]
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._reduce_bucket(b)
        return hook

    def _bucket_segments(self, b):
        """Contiguous [lo, hi) ranges of bucket ``b`` whose parameters have received their gradient."""
        segs = []
        for i in sorted(self._buckets[b]):
            if i not in self._fired:
                continue
            lo, hi = self._bounds[i]
            if segs and segs[-1][1] == lo:
                segs[-1][1] = hi
            else:
                segs.append([lo, hi])
        return [tuple(s) for s in segs]

    def _reduce_bucket(self, b):
        from .. import nn as rnn
        idxs = self._buckets[b]
        lo = min(self._bounds[i][0] for i in idxs)
        hi = max(self._bounds[i][1] for i in idxs)
        seg = self.flat_grad[lo:hi]
        if self.world > 1 and self._comm_stream is not None:
            stream = self._comm_stream
        elif self.bucket_hook is not None and seg.is_cuda:
            if self._hook_stream is None:
                self._hook_stream = torch.cuda.Stream()
            stream = self._hook_stream
        else:
            stream = None
        if stream is None:
            if self.world > 1:
                dist.all_reduce(seg)
                if not self.defer_average:
                    seg.mul_(1.0 / self.world)
            if self.bucket_hook is not None:
                self.bucket_hook(self._bucket_segments(b))
            return
        stream.wait_stream(torch.cuda.current_stream())
        if rnn.wgrad_stream() is not None:      # filter gradients are produced on their own stream
            stream.wait_stream(rnn.wgrad_stream())
        with torch.cuda.stream(stream):
            if self.world > 1:
                dist.all_reduce(seg)
                if not self.defer_average:
                    seg.mul_(1.0 / self.world)
            if self.bucket_hook is not None:
                self.bucket_hook(self._bucket_segments(b))

    def _finalize(self):
        # buckets holding parameters that never fired (unused parameters): zero their slots, reduce anyway so that
        # every rank issues the same collectives
        for b, left in enumerate(self._pending):
            if left > 0:
                for i in self._buckets[b]:
                    if i not in self._fired:
                        self._views[i].zero_()
                self._reduce_bucket(b)
        for st in (self._comm_stream, self._hook_stream):
            if st is not None:
                torch.cuda.current_stream().wait_stream(st)
        for i, p in enumerate(self._params):
            if i not in self._fired:
                p.grad = None
        self.used_parameter_ids = sorted(self._fired)
        self._armed = False

    def used_segments(self):
        """Contiguous [lo, hi) ranges of the flat buffers whose parameters received gradients."""
        segs = []
        for i in (self.used_parameter_ids or []):
            lo, hi = self._bounds[i]
            if segs and segs[-1][1] == lo:
                segs[-1][1] = hi
            else:
                segs.append([lo, hi])
        return [tuple(s) for s in segs]
