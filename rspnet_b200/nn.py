"""Autograd glue between torch modules and the sm_100a kernels (rspnet_b200.ops).

The conv path works on bf16 NDHWC tensors ``[N, T, H, W, C]``.  Each fused block is one ``autograd.Function``:

* ``ConvBNAct``   — Conv3d (+bias) -> train-mode BatchNorm3d -> (+residual) -> (ReLU)
                    (reference: the conv/bn/relu triplets of models/resnet.py:59-77,203-213, models/c3d.py:111-150)
* ``MaxPool3dFn`` — nn.MaxPool3d
* ``HeadsFn``     — AdaptiveAvgPool3d + Flatten + Linear (x2) + F.normalize (moco/split_wrapper.py:128-152)
* ``ToNCDHW`` / ``ToNDHWC`` — layout conversion at the public module boundary

Packed bf16 filter operands are cached per parameter and invalidated by ``bump_weight_epoch()`` (called by the
EMA / SGD kernels, which update parameters through raw pointers) or by torch's own version counter.
"""
import os
from typing import Optional

import torch

from . import ops

_weight_epoch = 0


def bump_weight_epoch():
    """Invalidate every cached packed filter (parameters were changed behind torch's back)."""
    global _weight_epoch
    _weight_epoch += 1


# Gradient slots: when a data-parallel wrapper owns one flat gradient buffer it registers the slot of every parameter
# here; the wgrad kernels then write straight into the slot and autograd adopts that tensor as ``param.grad``
# (no temporary, no accumulate kernel).  Valid once per backward and parameter: a second use in the same backward, or
# a parameter that already carries a gradient (zero_grad(set_to_none=False), gradient accumulation), falls back to a
# temporary that autograd accumulates.  "Same backward" is the autograd graph-task id, so nothing depends on the
# caller announcing a new step.
_grad_slots = {}
_grad_written = {}


def register_grad_slots(params, views):
    for p, v in zip(params, views):
        _grad_slots[p.data_ptr()] = v


def begin_grad_epoch():
    """Kept for callers of the round-1 API; slot bookkeeping is keyed by the autograd graph task now."""


def _grad_slot(weight, task):
    key = weight.data_ptr()
    v = _grad_slots.get(key)
    if v is None or weight.grad is not None or task == -1 or _grad_written.get(key) == task or v.shape != weight.shape:
        return None
    _grad_written[key] = task
    return v


# Filter gradients are off the critical path of backward (only the optimizer / the all-reduce read them), so they run
# on a side stream while the main stream carries on with dgrad and the next layer; a callback queued on the autograd
# graph task joins the streams when backward ends (FlatDDP waits on the same stream before it reduces a bucket).
wgrad_overlap = os.environ.get("RSP_WGRAD_OVERLAP", "1") != "0"
_wgrad_stream = None
_wgrad_task = -1          # autograd graph task that already has the join callback queued


def wgrad_stream():
    return _wgrad_stream


_wgrad_keepalive = []     # (x, dy) of every wgrad in flight: released after the join instead of Tensor.record_stream


def _join_wgrad():
    global _wgrad_task
    _wgrad_task = -1
    torch.cuda.current_stream().wait_stream(_wgrad_stream)
    _wgrad_keepalive.clear()   # later main-stream work is ordered after the side stream's reads


def _wgrad(desc, x, dy, weight):
    global _wgrad_stream, _wgrad_task
    task = torch._C._current_graph_task_id()
    slot = _grad_slot(weight, task)
    if slot is None:
        # No slot: the result is a temporary that autograd's AccumulateGrad (and any post-accumulate hook) reads on the
        # CURRENT stream right after this function returns, so it must be produced on the current stream.
        return ops.conv3d_wgrad(desc, x, dy, weight.shape)
    if not (wgrad_overlap and x.is_cuda):
        ops.conv3d_wgrad(desc, x, dy, weight.shape, out=slot)
        return slot.view(slot.shape)   # a fresh alias: autograd takes it over as param.grad without copying
    # Slot path: nobody reads the slot before the end of backward except FlatDDP's bucket reduce, which waits on the
    # side stream; the callback queued on the graph task joins the streams before backward() returns.
    if _wgrad_stream is None:
        _wgrad_stream = torch.cuda.Stream()
    if task != _wgrad_task:
        torch.autograd.Variable._execution_engine.queue_callback(_join_wgrad)
        _wgrad_task = task
    side = _wgrad_stream
    side.wait_stream(torch.cuda.current_stream())
    _wgrad_keepalive.append((x, dy))
    with torch.cuda.stream(side):
        ops.conv3d_wgrad(desc, x, dy, weight.shape, out=slot)
    return slot.view(slot.shape)


def refresh_packed_weights():
    """Batch re-pack of all stale filter operands (see _PackCache.refresh)."""
    _pack_cache.refresh()


class _PackCache:
    """Per-parameter cache of packed filter operands, keyed by (which, geometry, weight version)."""

    def __init__(self):
        self._store = {}

    @staticmethod
    def _tag(weight):
        return (_weight_epoch, weight._version, weight.data_ptr())

    def get(self, weight: torch.Tensor, desc, which: int):
        key = (id(weight), which, desc.Ci, desc.Co, desc.kt, desc.kh, desc.kw)
        hit = self._store.get(key)
        tag = self._tag(weight)
        if hit is not None and hit[0] == tag:
            return hit[1]
        packed = ops.conv3d_pack_weight(desc, weight, which)
        self._store[key] = (tag, packed, weight, desc)
        return packed

    def refresh(self):
        """Re-pack every fprop/wgrad operand seen so far whose parameter changed, in place, a whole encoder per launch
        (called once per step after the SGD / EMA updates instead of one small launch per layer)."""
        stale = [(k, e) for k, e in self._store.items() if k[1] == 0 and e[0] != self._tag(e[2]) and
                 e[2].dtype == torch.float32 and e[2].is_contiguous()]
        if not stale:
            return
        ops.conv3d_pack_weights([e[3] for _, e in stale], [e[2] for _, e in stale], [e[1] for _, e in stale])
        for k, e in stale:
            self._store[k] = (self._tag(e[2]), e[1], e[2], e[3])


_pack_cache = _PackCache()
_stats_acc = {}


def _pad_vec(v: Optional[torch.Tensor], n: int):
    if v is None or v.numel() == n:
        return v
    out = torch.zeros(n, dtype=v.dtype, device=v.device)
    out[:v.numel()] = v
    return out


def _conv_stats(x, weight, bias, gamma, kernel, stride, padding):
    """Conv fprop with the BN batch statistics taken in its epilogue.  The per-layer accumulator is double buffered:
    this forward adds into one half (zero at rest), the BN kernel that consumes it zeroes the other half for the next
    forward.  Returns desc, y, the sums of this forward and the half to clear."""
    co = ops.pad_channels(weight.shape[0])
    desc = ops.conv_desc(x.shape, co, kernel, stride, padding)
    wp = _pack_cache.get(weight, desc, 0)
    key = (id(gamma), co)
    st = _stats_acc.get(key)
    if st is None or st[0].device != x.device:
        st = _stats_acc[key] = [torch.zeros((2, 2, co), dtype=torch.float32, device=x.device), 0]
    acc, parity = st
    st[1] = parity ^ 1
    sums, other = acc[parity], acc[parity ^ 1]
    y = ops.conv3d_fprop(desc, x, wp, _pad_vec(bias, co), stats=sums)   # statistics fused into the conv epilogue
    if not ops.conv3d_fprop.stats_done:
        ops.bn_stats(y, out=sums)                                        # split-K layers: separate reduction
    return desc, y, sums, other


def _conv_bn_stats(x, weight, bias, gamma, beta, running_mean, running_var, kernel, stride, padding, eps, momentum):
    """_conv_stats followed by the stand-alone bn_finalize (used where the consumer is not the BN-apply kernel).
    Returns desc, y and the (scale, shift, mean, invstd) rows."""
    desc, y, sums, other = _conv_stats(x, weight, bias, gamma, kernel, stride, padding)
    co = y.shape[-1]
    # stand-alone finalize clears the half it read, so both halves are zero at rest on this path
    rows = ops.bn_finalize(sums[0], sums[1], y.numel() // co, gamma, beta, eps, momentum, running_mean, running_var, co,
                           clear_sums=True)
    return desc, y, rows


def _conv_bn_act_fwd(x, weight, bias, gamma, beta, running_mean, running_var, residual, kernel, stride, padding, eps,
                     momentum, relu):
    desc, y, sums, other = _conv_stats(x, weight, bias, gamma, kernel, stride, padding)
    co = y.shape[-1]
    out, rows = ops.bn_finalize_act_fwd(y, sums, other, y.numel() // co, gamma, beta, eps, momentum, running_mean,
                                        running_var, residual, relu)
    return desc, y, out, rows


class ConvBNAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, running_mean, running_var, residual, kernel, stride, padding, eps,
                momentum, relu):
        desc, y, out, (scale, shift, mean, invstd) = _conv_bn_act_fwd(x, weight, bias, gamma, beta, running_mean,
                                                                      running_var, residual, kernel, stride, padding,
                                                                      eps, momentum, relu)
        ctx.desc = desc
        ctx.relu = relu
        ctx.has_bias = bias is not None
        ctx.has_res = residual is not None
        ctx.save_for_backward(x, weight, y, out if relu else None, mean, invstd, gamma)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, weight, y, out, mean, invstd, gamma = ctx.saved_tensors
        desc = ctx.desc
        dout = dout.contiguous()
        dy, dres, dgamma, dbeta = ops.bn_act_bwd(dout, out, y, mean, invstd, gamma, ctx.relu, ctx.has_res)
        dw = _wgrad(desc, x, dy, weight)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.conv3d_dgrad(desc, dy, _pack_cache.get(weight, desc, 1))
        # a bias feeding straight into train-mode BN has an exactly zero gradient
        dbias = torch.zeros(weight.shape[0], dtype=torch.float32, device=weight.device) if ctx.has_bias else None
        return dx, dw, dbias, dgamma, dbeta, None, None, dres, None, None, None, None, None, None


class ConvBNReLUPool(torch.autograd.Function):
    """Conv3d (+bias) -> train-mode BatchNorm3d -> ReLU -> MaxPool3d with the BN/ReLU/pool part fused into one kernel
    each way (reference: models/resnet.py:203-206, models/c3d.py:111-139)."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, running_mean, running_var, kernel, stride, padding, eps, momentum,
                pool_k, pool_s, pool_p):
        desc, y, (scale, shift, mean, invstd) = _conv_bn_stats(x, weight, bias, gamma, beta, running_mean, running_var,
                                                               kernel, stride, padding, eps, momentum)
        pdesc = ops.pool_desc(y.shape, pool_k, pool_s, pool_p)
        out, idx, xmax = ops.bn_relu_maxpool_fwd(pdesc, y, scale, shift)
        ctx.desc, ctx.pdesc = desc, pdesc
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, weight, y, idx, xmax, scale, shift, mean, invstd, gamma)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, weight, y, idx, xmax, scale, shift, mean, invstd, gamma = ctx.saved_tensors
        desc = ctx.desc
        dy, dgamma, dbeta = ops.bn_relu_maxpool_bwd(ctx.pdesc, dout.contiguous(), idx, xmax, y, scale, shift, mean,
                                                    invstd, gamma)
        dw = _wgrad(desc, x, dy, weight)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.conv3d_dgrad(desc, dy, _pack_cache.get(weight, desc, 1))
        dbias = torch.zeros(weight.shape[0], dtype=torch.float32, device=weight.device) if ctx.has_bias else None
        return (dx, dw, dbias, dgamma, dbeta) + (None,) * 10


def conv_bn_relu_pool(x, conv: torch.nn.Conv3d, bn: torch.nn.BatchNorm3d, pool: torch.nn.MaxPool3d):
    """pool(relu(bn(conv(x)))) — fused when the pool geometry fits the kernel, else the two-step path."""
    if pool.ceil_mode or pool.dilation not in (1, (1, 1, 1)):
        raise NotImplementedError("rspnet_b200: ceil_mode / dilated MaxPool3d is not on the pretraining path")
    pk = ops._triple(pool.kernel_size)
    ps = ops._triple(pool.stride if pool.stride is not None else pool.kernel_size)
    pp = ops._triple(pool.padding)
    k, s, p = tuple(conv.kernel_size), tuple(conv.stride), tuple(conv.padding)
    co = ops.pad_channels(conv.weight.shape[0])
    cd = ops.conv_desc(x.shape, co, k, s, p)
    to, ho, wo = cd.out_dims()
    if not bn.training or not ops.bn_relu_maxpool_supported(ops.pool_desc((x.shape[0], to, ho, wo, co), pk, ps, pp)):
        return max_pool3d(conv_bn_act(x, conv, bn, relu=True), pool)
    if conv.groups != 1 or tuple(conv.dilation) != (1, 1, 1):
        raise NotImplementedError("rspnet_b200: grouped / dilated Conv3d is not on the pretraining path")
    momentum = bn.momentum if bn.momentum is not None else 0.0
    if not torch.is_grad_enabled():
        _, y, (scale, shift, _, _) = _conv_bn_stats(x, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean,
                                                    bn.running_var, k, s, p, bn.eps, momentum)
        out = ops.bn_relu_maxpool_fwd(ops.pool_desc(y.shape, pk, ps, pp), y, scale, shift, aux=False)[0]
    else:
        out = ConvBNReLUPool.apply(x, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, k, s,
                                   p, bn.eps, momentum, pk, ps, pp)
    if bn.track_running_stats and bn.num_batches_tracked is not None and not getattr(bn, "_rsp_counter_batched", False):
        bn.num_batches_tracked += 1
    return out


def conv_bn_act(x, conv: torch.nn.Conv3d, bn: torch.nn.BatchNorm3d, relu: bool = True, residual=None):
    """Fused Conv3d -> BatchNorm3d -> (+residual) -> (ReLU) on NDHWC bf16 activations.  Train-mode BN is the pretraining
    path; eval-mode BN (``model.eval()``: validation, retrieval, feature extraction after the checkpoint hand-off) is an
    inference-only forward with scale / shift taken from the running statistics."""
    if conv.groups != 1 or tuple(conv.dilation) != (1, 1, 1):
        raise NotImplementedError("rspnet_b200: grouped / dilated Conv3d is not on the pretraining path")
    if not bn.training:
        return _conv_bn_act_eval(x, conv, bn, relu, residual)
    momentum = bn.momentum if bn.momentum is not None else 0.0
    if not torch.is_grad_enabled():
        # key-encoder passes: same kernels without the autograd.Function round trip
        out = _conv_bn_act_fwd(x, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, residual,
                               conv.kernel_size, conv.stride, conv.padding, bn.eps, momentum, relu)[2]
    else:
        out = ConvBNAct.apply(x, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, residual,
                              tuple(conv.kernel_size), tuple(conv.stride), tuple(conv.padding), bn.eps, momentum, relu)
    if bn.track_running_stats and bn.num_batches_tracked is not None and not getattr(bn, "_rsp_counter_batched", False):
        bn.num_batches_tracked += 1
    return out


def _conv_bn_act_eval(x, conv, bn, relu, residual):
    """Eval-mode BatchNorm (F.batch_norm(training=False)): y = (conv(x) - running_mean) / sqrt(running_var + eps) * gamma
    + beta.  Inference only — the backward of frozen-statistics BN is not on the pretraining path."""
    if torch.is_grad_enabled() and (x.requires_grad or conv.weight.requires_grad):
        raise NotImplementedError("rspnet_b200: eval-mode BatchNorm is forward-only; wrap the call in torch.no_grad()")
    co = ops.pad_channels(conv.weight.shape[0])
    desc = ops.conv_desc(x.shape, co, conv.kernel_size, conv.stride, conv.padding)
    y = ops.conv3d_fprop(desc, x, _pack_cache.get(conv.weight, desc, 0), _pad_vec(conv.bias, co))
    with torch.no_grad():   # [C]-sized parameter preparation; padded channels get scale = shift = 0
        if bn.track_running_stats and bn.running_mean is not None:
            invstd = torch.rsqrt(bn.running_var.float() + bn.eps)
            mean = bn.running_mean.float()
        else:
            raise NotImplementedError("rspnet_b200: eval-mode BatchNorm needs running statistics")
        gamma = bn.weight.float() if bn.weight is not None else torch.ones_like(mean)
        beta = bn.bias.float() if bn.bias is not None else torch.zeros_like(mean)
        scale = _pad_vec(gamma * invstd, co)
        shift = _pad_vec(beta - mean * gamma * invstd, co)
    return ops.bn_act_fwd(y, scale.contiguous(), shift.contiguous(), residual, relu)


class ConvBiasAct(torch.autograd.Function):
    """Conv3d (+bias) -> optional ReLU on NDHWC bf16 without a BatchNorm behind it (the 'conv' projection heads,
    reference moco/split_wrapper.py:17-40).  The convolutions run on the tcgen05 kernels; the ReLU mask and the bias
    gradient of this non-hot variant are plain elementwise / reduction ops."""

    @staticmethod
    def forward(ctx, x, weight, bias, kernel, stride, padding, relu):
        co = ops.pad_channels(weight.shape[0])
        desc = ops.conv_desc(x.shape, co, kernel, stride, padding)
        y = ops.conv3d_fprop(desc, x, _pack_cache.get(weight, desc, 0), _pad_vec(bias, co))
        out = torch.relu_(y) if relu else y
        ctx.desc, ctx.relu, ctx.has_bias = desc, relu, bias is not None
        ctx.save_for_backward(x, weight, out if relu else None)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, weight, out = ctx.saved_tensors
        dy = dout.contiguous()
        if ctx.relu:
            dy = dy * (out > 0)
        dw = _wgrad(ctx.desc, x, dy, weight)
        dx = ops.conv3d_dgrad(ctx.desc, dy, _pack_cache.get(weight, ctx.desc, 1)) if ctx.needs_input_grad[0] else None
        dbias = dy.float().sum(dim=(0, 1, 2, 3))[:weight.shape[0]] if ctx.has_bias else None
        return dx, dw, dbias, None, None, None, None


def conv_bias_act(x, conv: torch.nn.Conv3d, relu: bool = False):
    if conv.groups != 1 or tuple(conv.dilation) != (1, 1, 1):
        raise NotImplementedError("rspnet_b200: grouped / dilated Conv3d is not on the pretraining path")
    return ConvBiasAct.apply(x, conv.weight, conv.bias, tuple(conv.kernel_size), tuple(conv.stride), tuple(conv.padding),
                             relu)


def batch_bn_counters(module: torch.nn.Module):
    """Re-homes every BatchNorm ``num_batches_tracked`` of ``module`` into one int64 buffer (the per-layer buffers become
    views, names/values unchanged) so that one increment per forward replaces one tiny kernel per layer.
    Returns the flat counter tensor (or None when there is no BatchNorm)."""
    bns = [m for m in module.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm) and
           m.track_running_stats and m.num_batches_tracked is not None]
    if not bns:
        return None
    flat = torch.stack([m.num_batches_tracked.detach().reshape(()) for m in bns]).clone()
    for i, m in enumerate(bns):
        m._buffers["num_batches_tracked"] = flat[i]
        m._rsp_counter_batched = True
    return flat


class MaxPool3dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kernel, stride, padding):
        desc = ops.pool_desc(x.shape, kernel, stride, padding)
        y, idx = ops.maxpool3d_fwd(desc, x)
        ctx.desc = desc
        ctx.save_for_backward(idx)
        return y

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        return ops.maxpool3d_bwd(ctx.desc, dy.contiguous(), idx), None, None, None


def max_pool3d(x, pool: torch.nn.MaxPool3d):
    if pool.ceil_mode or pool.dilation not in (1, (1, 1, 1)):
        raise NotImplementedError("rspnet_b200: ceil_mode / dilated MaxPool3d is not on the pretraining path")
    stride = pool.stride if pool.stride is not None else pool.kernel_size
    return MaxPool3dFn.apply(x, pool.kernel_size, stride, pool.padding)


class HeadsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, w1, b1, w2, b2, c_logical):
        o1, o2, pooled, raw = ops.head_fwd(feat, c_logical, w1, b1, w2, b2)
        ctx.feat_shape = tuple(feat.shape)
        ctx.save_for_backward(pooled, raw, w1, w2)
        return o1, o2

    @staticmethod
    def backward(ctx, g1, g2):
        pooled, raw, w1, w2 = ctx.saved_tensors
        if g1 is None:
            g1 = torch.zeros_like(raw[0])
        if g2 is None:
            g2 = torch.zeros_like(raw[1])
        dw1, db1, dw2, db2, dfeat = ops.head_bwd(g1, g2, pooled, raw, ctx.feat_shape, w1, w2,
                                                 want_dfeat=ctx.needs_input_grad[0])
        return dfeat, dw1, db1, dw2, db2, None


class ToNCDHW(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, c_logical):
        ctx.cs = x.shape[-1]
        return ops.to_ncdhw_f32(x, c_logical)

    @staticmethod
    def backward(ctx, g):
        return ops.to_ndhwc_bf16(g.contiguous(), ctx.cs), None


class ToNDHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, c_stored):
        ctx.c = x.shape[1]
        return ops.to_ndhwc_bf16(x, c_stored)

    @staticmethod
    def backward(ctx, g):
        return ops.to_ncdhw_f32(g.contiguous(), ctx.c), None


def is_ndhwc(x: torch.Tensor) -> bool:
    """Internal activations are tagged by dtype: bf16 5-D tensors are NDHWC, anything else is NCDHW."""
    return x.dim() == 5 and x.dtype == torch.bfloat16


def as_ndhwc(x: torch.Tensor) -> torch.Tensor:
    if is_ndhwc(x):
        return x
    if not x.is_cuda:
        raise RuntimeError("rspnet_b200: forward needs CUDA tensors (there is no CPU path)")
    return ToNDHWC.apply(x, ops.pad_channels(x.shape[1]))


class GateFn(torch.autograd.Function):
    """S3D-G self-gating: y = x * sigmoid(excitation(mean_{T,H,W} x))  (models/s3dg.py:64-72)."""

    @staticmethod
    def forward(ctx, x, weight, bias, c_logical):
        y, pooled, gate = ops.gate_fwd(x, c_logical, weight, bias)
        ctx.c_logical = c_logical
        ctx.wshape = tuple(weight.shape)
        ctx.save_for_backward(x, weight, pooled, gate)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, pooled, gate = ctx.saved_tensors
        dx, dw, db = ops.gate_bwd(dy.contiguous(), x, ctx.c_logical, weight, pooled, gate)
        return dx, dw.view(ctx.wshape), db, None


class ConcatChannelsFn(torch.autograd.Function):
    """torch.cat(dim=1) of NDHWC tensors with zero-padded channels: logical channels are packed back to back."""

    @staticmethod
    def forward(ctx, logical, *xs):
        total = sum(logical)
        cs = ops.pad_channels(total)
        out = torch.zeros(tuple(xs[0].shape[:-1]) + (cs,), dtype=torch.bfloat16, device=xs[0].device)
        off = 0
        for x, cl in zip(xs, logical):
            ops.copy_channels(x, 0, out, off, cl)
            off += cl
        ctx.logical = logical
        ctx.stored = [x.shape[-1] for x in xs]
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = dout.contiguous()
        grads, off = [], 0
        for cl, cs in zip(ctx.logical, ctx.stored):
            g = torch.zeros(tuple(dout.shape[:-1]) + (cs,), dtype=torch.bfloat16, device=dout.device) if cs != cl \
                else torch.empty(tuple(dout.shape[:-1]) + (cs,), dtype=torch.bfloat16, device=dout.device)
            ops.copy_channels(dout, off, g, 0, cl)
            grads.append(g)
            off += cl
        return (None, *grads)


def concat_channels(xs, logical):
    for cl in logical:
        if cl % 8:
            raise NotImplementedError("rspnet_b200: channel concat needs multiples of 8 channels per branch")
    return ConcatChannelsFn.apply(tuple(logical), *xs)
