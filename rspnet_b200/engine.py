"""Training-step driver for the pretraining hot path (the loop body of the reference's
``pretrain.Engine.train_epoch``, pretrain.py:154-165): forward, 3-term loss, backward, gradient all-reduce and
SGD(momentum, weight decay) — with the optimizer running as one fused pass over the flat parameter buffer.

Checkpoint hand-off: ``checkpoint_state`` / ``load_checkpoint`` speak the dictionary the
reference writes and reads (pretrain.py:118-125,249-259) — ``model`` with the reference's state_dict names, ``optimizer``
and ``scheduler`` in ``torch.optim.SGD`` / ``CosineAnnealingLR`` format — so a run can move between the two code bases
and ``finetune.py:273-310`` / ``retrieval.py:84-101`` load the encoder unchanged.
"""
import math
import warnings
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
                 num_epochs: int = 200, track_metrics: bool = False, cuda_graph: bool = False):
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
        # cuda_graph=True: after three eager steps the whole step (forward x3, loss, backward, gradient all-reduce, SGD,
        # enqueue — side streams and NCCL collectives included) is captured into CUDA graphs and replayed; only the
        # random draws (made eagerly, same generators / same order as the reference) and the input hand-off stay on the
        # host.  Two graphs alternate (even / odd steps): per-layer double buffers (BN statistics, exchange buffers) flip
        # every step, and a 2-deep input ring then maps one input pair to each graph.
        self.cuda_graph = cuda_graph
        self._graph_capable = cuda_graph      # stays True after a failed capture: the dedicated stream is kept
        self._stream = None
        self._graphs = [None, None]
        self._graph_steps = 0
        self._eager_steps = 0
        self._graph_pool = None
        self.graph_error = None
        self.phase_log = None     # set to a list to collect CUDA events at the phase boundaries of every step (bench.py)
        # the eight AverageMeters of pretrain.py:97-106, kept on the device (meters.py); off by default like any logging
        self.meters = None
        if track_metrics:
            from .meters import ContrastiveMeters
            self.meters = ContrastiveMeters(self.flat_q.device)
        # SGD runs bucket by bucket inside backward, on the stream where each bucket's all-reduced gradient becomes final
        # (FlatDDP.bucket_hook): only the last small bucket's reduce + update is left after the final wgrad kernel.  The
        # 1/world of DDP's mean is folded into the update (grad_scale), not spent on a pass over the gradient.
        self.ddp.bucket_hook = self._sgd_segments
        self.ddp.defer_average = True
        self._bind_grads()

    def _bind_grads(self):
        for p, v in zip(self.ddp._params, self.ddp._views):
            p.grad = v

    def _sgd_segments(self, segments):
        for lo, hi in segments:
            ops.sgd_step_(self.flat_q[lo:hi], self.ddp.flat_grad[lo:hi], self.momentum_buf[lo:hi], self.lr,
                          self.momentum, self.weight_decay, 1.0 / self.ddp.world, self._first)

    def set_epoch(self, epoch: int):
        """CosineAnnealingLR(T_max=num_epochs, eta_min=lr/1000) evaluated per epoch (pretrain.py:74-79)."""
        eta_min = self.base_lr / 1000
        self.lr = eta_min + (self.base_lr - eta_min) * (1 + math.cos(math.pi * epoch / self.num_epochs)) / 2

    def step(self, clip_q: torch.Tensor, clip_k: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """One optimisation step; returns (loss, loss_A, loss_M) as device scalars (no host sync)."""
        if not (self._graph_capable and clip_q.is_cuda):
            return self._step_eager(clip_q, clip_k)
        # Graph mode runs EVERY step (eager warm-up, capture, replay) on one dedicated stream: autograd pins each
        # parameter's AccumulateGrad node to the stream it was created on, and a capture must not touch the legacy
        # default stream.  The caller's stream is ordered before and after.
        caller = torch.cuda.current_stream()
        if self._stream is None:
            self._stream = torch.cuda.Stream()
        self._stream.wait_stream(caller)
        with torch.cuda.stream(self._stream):
            if (self.cuda_graph and self._eager_steps >= 3 and not self._first and self.meters is None and
                    self.phase_log is None):
                out = self._graph_step(clip_q, clip_k)
            else:
                out = self._step_eager(clip_q, clip_k)
        caller.wait_stream(self._stream)
        return out

    def _step_eager(self, clip_q, clip_k):
        self._eager_steps += 1
        if self._graphs[0] is not None or self._graphs[1] is not None:
            self._graphs = [None, None]      # an eager step flips the per-step double buffers: captured graphs are stale
        return self._step_body(clip_q, clip_k)

    def _step_body(self, clip_q, clip_k, draws=None):
        # gradients are produced straight into their slots of the flat buffer (rnn.register_grad_slots); slots of
        # parameters that receive nothing are zeroed by FlatDDP._finalize
        marks = [] if self.phase_log is not None else None

        def mark():
            if marks is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append(ev)

        for p in self.ddp._params:
            p.grad = None
        mark()
        if draws is None:
            output, target, ranking_logits, ranking_target = self.ddp(clip_q, clip_k)
        else:
            output, target, ranking_logits, ranking_target = self.ddp(clip_q, clip_k, draws=draws)
        mark()
        loss, loss_a, loss_m = self.criterion(output, target, ranking_logits, ranking_target)
        loss.backward()          # includes the gradient all-reduce and the SGD update of every bucket (_sgd_segments)
        mark()
        mark()
        if marks is not None:
            self.phase_log.append(marks)
        self._first = False
        rnn.bump_weight_epoch()
        if self.meters is not None:
            self.meters.update((loss, loss_a, loss_m), output, ranking_logits)
        # detached: holding the attached outputs would keep the whole autograd graph (every saved activation and the
        # parameters' AccumulateGrad nodes, with the stream they were created on) alive until the next step
        self.last_output = (tuple(o.detach() for o in output), tuple(r.detach() for r in ranking_logits))
        return loss.detach(), loss_a.detach(), loss_m.detach()

    # ------------------------------------------------------------------------------------------------ CUDA graphs
    def release_graphs(self):
        """Drops the captured graphs (they re-capture on the next steps).  Call before tearing down the process group:
        the graphs hold NCCL kernels."""
        self._graphs = [None, None]
        self._graph_pool = None

    def next_input_pair(self, ahead: int = 0) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
        """The ``(clip_q, clip_k)`` buffers that the next graph-mode step (``ahead=0``; ``ahead=1``: the one after it)
        reads in place, or None while that step would still run eagerly / be captured.  A loader that fills them
        (``GPUClipSampler(..., out=pair)``) and passes them to ``step`` saves the device copy of both clips into the
        captured graph's inputs.  Two graphs alternate, so the pair also belongs to the step two steps earlier: fill it
        only after that step has finished (stream-order the fill behind it)."""
        if not (self.cuda_graph and self._graph_capable and self._eager_steps >= 3):
            return None
        G = self._graphs[(self._graph_steps + ahead) & 1]
        return None if G is None else (G["q"], G["k"])

    def _graph_step(self, clip_q, clip_k):
        slot = self._graph_steps & 1
        key = (tuple(clip_q.shape), tuple(clip_k.shape), clip_q.dtype, self.lr, self.momentum, self.weight_decay)
        G = self._graphs[slot]
        if G is not None and G["key"] != key:
            self._graphs = [None, None]
            G = None
        if G is None:
            G = self._capture(clip_q, clip_k, key)
            if isinstance(G, tuple):
                return G                     # capture refused: the step ran eagerly with the draws already made
            self._graphs[slot] = G
        else:
            if clip_q.data_ptr() != G["q"].data_ptr():
                G["q"].copy_(clip_q)
            if clip_k.data_ptr() != G["k"].data_ptr():
                G["k"].copy_(clip_k)
            self.model.draw_step(G["q"], static=G["static"])      # eager: reference-order draws into the graph's inputs
        G["graph"].replay()
        self._graph_steps += 1
        rnn.bump_weight_epoch()
        self.last_output = G["out"]
        return G["loss"]

    def _capture(self, clip_q, clip_k, key):
        """Records one full step into a CUDA graph whose inputs are (clip_q, clip_k) themselves and two small persistent
        tensors holding the step's random draws.  When the step cannot be captured (e.g. the all_to_all exchange, whose
        split sizes live on the host) graphs are switched off and the eager result of this step is returned instead."""
        dev = clip_q.device
        B = clip_q.shape[0]
        static = {"random_indices": torch.empty(B, dtype=torch.int64, device=dev),
                  "idx_all": torch.empty((2, B * self.ddp.world), dtype=torch.int64, device=dev)}
        draws = self.model.draw_step(clip_q, static=static)
        if draws[3] is not None:
            self.cuda_graph, self.graph_error = False, "the all_to_all shuffle exchange plans on the host"
            return self._fallback_after_draws(clip_q, clip_k, draws)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(graph, pool=self._graph_pool, stream=self._stream, capture_error_mode="thread_local"):
                loss = self._step_body(clip_q, clip_k, draws=draws)
        except Exception as e:   # capture is an optimisation: the eager path is always valid
            self.cuda_graph, self.graph_error = False, f"{type(e).__name__}: {e}"[:300]
            try:
                torch.cuda.synchronize()
            except Exception:
                pass
            try:   # an aborted capture leaves the CUDA generator in capture mode: a trivial complete capture resets it
                with torch.cuda.graph(torch.cuda.CUDAGraph()):
                    torch.zeros(1, device=dev)
            except Exception:
                pass
            self.ddp._armed = False
            return self._fallback_after_draws(clip_q, clip_k, draws)
        if self._graph_pool is None:
            self._graph_pool = graph.pool()
        return dict(graph=graph, q=clip_q, k=clip_k, static=static, loss=loss, out=self.last_output, key=key)

    def _fallback_after_draws(self, clip_q, clip_k, draws):
        """Capture is off, but this step's draws were already consumed from the generators: run it eagerly with them."""
        self._eager_steps += 1
        self._graphs = [None, None]
        return self._step_body(clip_q, clip_k, draws=draws)

    # ------------------------------------------------------------------------------------------------ checkpoints
    def _momentum_views(self):
        return [self.momentum_buf[lo:lo + p.numel()].view_as(p)
                for p, (lo, _) in zip(self.ddp._params, self.ddp._bounds)]

    def _torch_optimizer(self, epoch: int):
        """A ``torch.optim.SGD`` / ``CosineAnnealingLR`` pair over ``model.parameters()`` (what pretrain.py:64-79
        builds) holding this engine's state; used for (de)serialisation only, never stepped."""
        opt = torch.optim.SGD(self.model.parameters(), lr=self.base_lr, momentum=self.momentum, dampening=0.0,
                              weight_decay=self.weight_decay, nesterov=False)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=self.num_epochs, eta_min=self.base_lr / 1000)
        return opt, sched

    def checkpoint_state(self, epoch: int, arch: str, best_loss: float = float("inf")) -> dict:
        """The dictionary of pretrain.py:249-259 (``model`` is ``self.model.state_dict()``, i.e. DDP's ``.module``)."""
        opt, sched = self._torch_optimizer(epoch)
        used = self.ddp.used_parameter_ids if not self._first else []
        views = self._momentum_views()
        for i in (used or []):          # torch.optim.SGD only holds a buffer for parameters that received a gradient
            opt.state[self.ddp._params[i]]["momentum_buffer"] = views[i].detach().clone()
        self.set_epoch(epoch)
        opt.param_groups[0]["lr"] = self.lr
        sched.last_epoch = epoch
        sched._step_count = epoch + 1
        sched._last_lr = [self.lr]
        return {"epoch": epoch, "arch": arch, "model": self.model.state_dict(), "best_loss": best_loss,
                "optimizer": opt.state_dict(), "scheduler": sched.state_dict()}

    def load_checkpoint(self, states: dict, arch: Optional[str] = None) -> int:
        """pretrain.py:112-125.  Returns the epoch to continue from."""
        if arch is not None and states["arch"] != arch:
            raise ValueError(f'Loading checkpoint arch {states["arch"]} does not match current arch {arch}')
        self.model.load_state_dict(states["model"])
        rnn.bump_weight_epoch()                      # packed bf16 filters are stale now
        opt, sched = self._torch_optimizer(0)
        opt.load_state_dict(states["optimizer"])
        sched.load_state_dict(states["scheduler"])
        self.momentum_buf.zero_()
        any_state = False
        for p, view in zip(self.ddp._params, self._momentum_views()):
            buf = opt.state.get(p, {}).get("momentum_buffer")
            if buf is not None:
                view.copy_(buf)
                any_state = True
        # without buffers the next step is torch's first step (buf = grad); with them, parameters that never had a
        # gradient keep a zero buffer, which gives the same update as torch's lazy initialisation (momentum * 0 + grad)
        self._first = not any_state
        group = opt.param_groups[0]
        self.momentum, self.weight_decay = group["momentum"], group["weight_decay"]
        self.base_lr = group.get("initial_lr", self.base_lr)
        self.num_epochs = sched.T_max
        self.set_epoch(states["epoch"])
        return states["epoch"]
