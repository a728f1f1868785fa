"""Training-step driver for the pretraining hot path (the loop body of the reference's
``pretrain.Engine.train_epoch``, pretrain.py:154-165): forward, 3-term loss, backward, gradient all-reduce and
SGD(momentum, weight decay) — with the optimizer running as one fused pass over the flat parameter buffer.
"""
import math
from typing import Optional, Tuple

import torch

from . import nn as rnn
from . import ops
from .moco import FlatDDP, Loss
from .moco import exchange


def scale_learning_rate(lr: float, world_size: int, batch_size: int, base_batch: int = 64) -> float:
    """framework/utils/environment.py:13-16."""
    return lr * world_size * batch_size / base_batch


class PretrainEngine:
    def __init__(self, model, criterion: Loss, lr: float, momentum: float = 0.9, weight_decay: float = 1e-4,
                 num_epochs: int = 200, track_metrics: bool = False):
        self.ddp = model if isinstance(model, FlatDDP) else FlatDDP(model)
        self.model = self.ddp.module
        self.criterion = criterion
        self.base_lr = lr
        self.lr = lr
        self.momentum = momentum
        self.weight_decay = weight_decay
        self.num_epochs = num_epochs
        self.flat_q, _ = self.model.flat_parameters()
        self.momentum_buf = torch.zeros_like(self.flat_q)
        self._first = True
        # the eight AverageMeters of pretrain.py:97-106, kept on the device (meters.py); off by default like any logging
        self.meters = None
        if track_metrics:
            from .meters import ContrastiveMeters
            self.meters = ContrastiveMeters(self.flat_q.device)
        self._bind_grads()

    def _bind_grads(self):
        for p, v in zip(self.ddp._params, self.ddp._views):
            p.grad = v

    def set_epoch(self, epoch: int):
        """CosineAnnealingLR(T_max=num_epochs, eta_min=lr/1000) evaluated per epoch (pretrain.py:74-79)."""
        eta_min = self.base_lr / 1000
        self.lr = eta_min + (self.base_lr - eta_min) * (1 + math.cos(math.pi * epoch / self.num_epochs)) / 2

    def step(self, clip_q: torch.Tensor, clip_k: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """One optimisation step; returns (loss, loss_A, loss_M) as device scalars (no host sync)."""
        # gradients are produced straight into their slots of the flat buffer (rnn.register_grad_slots); slots of
        # parameters that receive nothing are zeroed by FlatDDP._finalize
        for p in self.ddp._params:
            p.grad = None
        rnn.begin_grad_epoch()
        output, target, ranking_logits, ranking_target = self.ddp(clip_q, clip_k)
        loss, loss_a, loss_m = self.criterion(output, target, ranking_logits, ranking_target)
        loss.backward()
        for lo, hi in self.ddp.used_segments():
            ops.sgd_step_(self.flat_q[lo:hi], self.ddp.flat_grad[lo:hi], self.momentum_buf[lo:hi], self.lr,
                          self.momentum, self.weight_decay, 1.0, self._first)
        self._first = False
        rnn.bump_weight_epoch()
        if self.meters is not None:
            self.meters.update((loss, loss_a, loss_m), output, ranking_logits)
        self.last_output = (output, ranking_logits)
        return loss.detach(), loss_a.detach(), loss_m.detach()
