"""Thin, allocation-owning wrappers over the C ABI (one python function per kernel family).

Tensors on the conv path are ``torch.bfloat16`` in NDHWC layout, stored as 5-D tensors ``[N, T, H, W, C]``.
Nothing here falls back to ATen compute; torch only allocates outputs.
"""
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import ConvDesc, PoolDesc, call, ptr, stream_ptr
import ctypes as C


def _triple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v, v)


def pad_channels(c: int) -> int:
    """Stored channel count of an activation with ``c`` logical channels."""
    if c <= 4:
        return 4
    return (c + 63) // 64 * 64


def conv_desc(x_shape, co: int, kernel, stride, padding) -> ConvDesc:
    n, t, h, w, ci = x_shape
    k, s, p = _triple(kernel), _triple(stride), _triple(padding)
    return ConvDesc(n, t, h, w, ci, co, k[0], k[1], k[2], s[0], s[1], s[2], p[0], p[1], p[2])


# ------------------------------------------------------------------------------------------------ conv
def conv3d_pack_weight(desc: ConvDesc, weight: torch.Tensor, which: int = 0) -> torch.Tensor:
    """fp32 ``[Co, Ci, kt, kh, kw]`` parameter -> packed bf16 operand (which=0 fprop/wgrad, 1 dgrad)."""
    co_l, ci_l = weight.shape[0], weight.shape[1]
    kpad = _lib.load().rsp_conv3d_kpad(C.byref(desc), which)
    if kpad <= 0:
        raise RuntimeError("rsp_conv3d_kpad failed: " + _lib.load().rsp_last_error().decode())
    n_elems = _lib.load().rsp_conv3d_packed_elems(C.byref(desc), which)
    out = torch.empty((n_elems,), dtype=torch.bfloat16, device=weight.device)
    w = weight.detach().contiguous().float()
    call("rsp_conv3d_pack_weight", C.byref(desc), ci_l, co_l, ptr(w), ptr(out), which, stream_ptr())
    return out


def conv3d_pack_weights(descs, weights, outs):
    """Batched ``conv3d_pack_weight(which=0)`` into existing packed buffers."""
    n = len(descs)
    d_arr = (ConvDesc * n)(*descs)
    ci = (C.c_int32 * n)(*[w.shape[1] for w in weights])
    co = (C.c_int32 * n)(*[w.shape[0] for w in weights])
    w_arr = (C.c_void_p * n)(*[w.data_ptr() for w in weights])
    o_arr = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
    call("rsp_conv3d_pack_weights", n, d_arr, ci, co, w_arr, o_arr, stream_ptr())


def _splitk_workspace(desc: ConvDesc, which: int, device):
    """fp32 accumulation buffer for split-K on small-M layers (None when the library does not want one)."""
    _lib._ensure_device()
    n = _lib.load().rsp_conv3d_workspace_bytes(C.byref(desc), which)
    return torch.empty((n // 4,), dtype=torch.float32, device=device) if n > 0 else None


def conv3d_fprop(desc: ConvDesc, x: torch.Tensor, wp: torch.Tensor, bias: Optional[torch.Tensor] = None,
                 stats: Optional[torch.Tensor] = None):
    """Returns y; when ``stats`` ([2, Co] fp32, zeroed) is given the epilogue accumulates the BN batch statistics into
    it — unless the layer runs split-K, in which case ``conv3d_fprop.stats_done`` is False and the caller must run
    ``bn_stats``."""
    assert x.dtype == torch.bfloat16 and x.is_contiguous()
    to, ho, wo = desc.out_dims()
    y = torch.empty((desc.N, to, ho, wo, desc.Co), dtype=torch.bfloat16, device=x.device)
    ws = _splitk_workspace(desc, 0, x.device)
    conv3d_fprop.stats_done = stats is not None and ws is None
    call("rsp_conv3d_fprop", C.byref(desc), ptr(x), ptr(wp), ptr(bias), ptr(y), ptr(ws),
         ptr(stats) if conv3d_fprop.stats_done else None, stream_ptr())
    return y


conv3d_fprop.stats_done = False


def conv3d_dgrad(desc: ConvDesc, dy: torch.Tensor, wd: torch.Tensor):
    assert dy.dtype == torch.bfloat16 and dy.is_contiguous()
    dx = torch.empty((desc.N, desc.Ti, desc.Hi, desc.Wi, desc.Ci), dtype=torch.bfloat16, device=dy.device)
    ws = _splitk_workspace(desc, 1, dy.device)
    call("rsp_conv3d_dgrad", C.byref(desc), ptr(dy), ptr(wd), ptr(dx), ptr(ws), stream_ptr())
    return dx


def conv3d_wgrad(desc: ConvDesc, x: torch.Tensor, dy: torch.Tensor, weight_shape, out: Optional[torch.Tensor] = None,
                 accumulate: bool = False):
    """Gradient of the fp32 parameter ``[Co, Ci, kt, kh, kw]``."""
    co_l, ci_l = weight_shape[0], weight_shape[1]
    kpad = _lib.load().rsp_conv3d_kpad(C.byref(desc), 0)
    ws = torch.empty((kpad, desc.Co), dtype=torch.float32, device=x.device)
    if out is None:
        out = torch.empty(tuple(weight_shape), dtype=torch.float32, device=x.device)
        accumulate = False
    call("rsp_conv3d_wgrad", C.byref(desc), ci_l, co_l, ptr(x), ptr(dy), ptr(ws), ptr(out), int(accumulate),
         stream_ptr())
    return out


# ------------------------------------------------------------------------------------------------ batch norm
def bn_stats(x: torch.Tensor, out: Optional[torch.Tensor] = None):
    """Per-channel sum and sum of squares; accumulates into ``out`` ([2, C], must be zero) when given."""
    c = x.shape[-1]
    m = x.numel() // c
    s = out if out is not None else torch.zeros((2, c), dtype=torch.float32, device=x.device)
    call("rsp_bn_stats", ptr(x), m, c, ptr(s[0]), ptr(s[1]), stream_ptr())
    return s[0], s[1]


def bn_finalize(s, ss, count, gamma, beta, eps, momentum, running_mean, running_var, c_stored, clear_sums=False):
    out = torch.empty((4, c_stored), dtype=torch.float32, device=s.device)
    base, row = out.data_ptr(), 4 * c_stored
    call("rsp_bn_finalize", ptr(s), ptr(ss), int(clear_sums), count, ptr(gamma), ptr(beta), eps, momentum,
         ptr(running_mean), ptr(running_var), base, base + row, base + 2 * row, base + 3 * row, c_stored,
         gamma.numel(), stream_ptr())
    return out.unbind(0)  # scale, shift, mean, invstd


def bn_finalize_act_fwd(x, sums, clear, count, gamma, beta, eps, momentum, running_mean, running_var, residual,
                        relu: bool):
    """bn_finalize + bn_act_fwd in one launch; ``sums`` [2, C] are read, ``clear`` [2, C] (or None) is zeroed.
    Returns out and the (scale, shift, mean, invstd) rows."""
    c = x.shape[-1]
    out = torch.empty_like(x)
    rows = torch.empty((4, c), dtype=torch.float32, device=x.device)
    base = sums.data_ptr()
    call("rsp_bn_finalize_act_fwd", ptr(x), base, base + 4 * c, ptr(clear), count, ptr(gamma), ptr(beta), eps, momentum,
         ptr(running_mean), ptr(running_var), ptr(rows), ptr(residual), int(relu), ptr(out), x.numel() // c, c,
         gamma.numel(), stream_ptr())
    return out, rows.unbind(0)


def bn_act_fwd(x, scale, shift, residual, relu: bool):
    c = x.shape[-1]
    out = torch.empty_like(x)
    call("rsp_bn_act_fwd", ptr(x), ptr(scale), ptr(shift), ptr(residual), int(relu), ptr(out), x.numel() // c, c,
         stream_ptr())
    return out


def bn_act_bwd(dout, out, x, mean, invstd, gamma, relu: bool, want_dres: bool):
    c = x.shape[-1]
    m = x.numel() // c
    sums = torch.zeros((2, c), dtype=torch.float32, device=x.device)
    call("rsp_bn_act_bwd_reduce", ptr(dout), ptr(out), ptr(x), ptr(mean), ptr(invstd), int(relu), ptr(sums[0]),
         ptr(sums[1]), m, c, stream_ptr())
    dx = torch.empty_like(x)
    dres = torch.empty_like(x) if want_dres else None
    call("rsp_bn_act_bwd_apply", ptr(dout), ptr(out), ptr(x), ptr(mean), ptr(invstd), ptr(gamma), ptr(sums[0]),
         ptr(sums[1]), int(relu), ptr(dx), ptr(dres), m, c, gamma.numel(), stream_ptr())
    cl = gamma.numel()
    return dx, dres, sums[1][:cl], sums[0][:cl]  # dx, dres, dgamma, dbeta


# ------------------------------------------------------------------------------------------------ pooling
def pool_desc(x_shape, kernel, stride, padding) -> PoolDesc:
    n, t, h, w, c = x_shape
    k, s, p = _triple(kernel), _triple(stride), _triple(padding)
    return PoolDesc(n, t, h, w, c, k[0], k[1], k[2], s[0], s[1], s[2], p[0], p[1], p[2])


def maxpool3d_fwd(desc: PoolDesc, x):
    to, ho, wo = desc.out_dims()
    y = torch.empty((desc.N, to, ho, wo, desc.C), dtype=torch.bfloat16, device=x.device)
    idx = torch.empty((desc.N, to, ho, wo, desc.C), dtype=torch.uint8, device=x.device)
    call("rsp_maxpool3d_fwd", C.byref(desc), ptr(x), ptr(y), ptr(idx), stream_ptr())
    return y, idx


def maxpool3d_bwd(desc: PoolDesc, dy, idx):
    dx = torch.empty((desc.N, desc.Ti, desc.Hi, desc.Wi, desc.C), dtype=torch.bfloat16, device=dy.device)
    call("rsp_maxpool3d_bwd", C.byref(desc), ptr(dy), ptr(idx), ptr(dx), stream_ptr())
    return dx


def bn_relu_maxpool_supported(desc: PoolDesc) -> bool:
    return bool(_lib.load().rsp_bn_relu_maxpool_supported(C.byref(desc)))


def bn_relu_maxpool_fwd(desc: PoolDesc, x, scale, shift, aux: bool = True):
    """maxpool(relu(x*scale + shift)) without materialising the activation.  Returns (y, argmax, x_max); with
    ``aux=False`` (no-grad passes) the argmax and the raw value it selected are neither tracked nor written."""
    to, ho, wo = desc.out_dims()
    y = torch.empty((desc.N, to, ho, wo, desc.C), dtype=torch.bfloat16, device=x.device)
    idx = torch.empty((desc.N, to, ho, wo, desc.C), dtype=torch.uint8, device=x.device) if aux else None
    xmax = torch.empty_like(y) if aux else None
    call("rsp_bn_relu_maxpool_fwd", C.byref(desc), ptr(x), ptr(scale), ptr(shift), ptr(y), ptr(idx), ptr(xmax),
         stream_ptr())
    return y, idx, xmax


def bn_relu_maxpool_bwd(desc: PoolDesc, dy, idx, xmax, x, scale, shift, mean, invstd, gamma):
    """Gradient w.r.t. the conv output x plus (dgamma, dbeta) of the fused BN -> ReLU -> MaxPool block: the two BN
    reductions from the pooled tensors (dy, x_max), then one pass that scatters dy and streams x (csrc/bn_pool_bwd.cu)."""
    c = x.shape[-1]
    sums = torch.zeros((2, c), dtype=torch.float32, device=x.device)
    call("rsp_bn_relu_maxpool_bwd_sums", C.byref(desc), ptr(dy), ptr(xmax), ptr(scale), ptr(shift), ptr(mean),
         ptr(invstd), ptr(sums[0]), ptr(sums[1]), stream_ptr())
    dx = torch.empty_like(x)
    call("rsp_bn_relu_maxpool_bwd_dx", C.byref(desc), ptr(dy), ptr(idx), ptr(xmax), ptr(x), ptr(scale), ptr(shift),
         ptr(mean), ptr(invstd), ptr(gamma), gamma.numel(), ptr(sums[0]), ptr(sums[1]), ptr(dx), stream_ptr())
    cl = gamma.numel()
    return dx, sums[1][:cl], sums[0][:cl]


# ------------------------------------------------------------------------------------------------ heads
def head_fwd(feat, c_logical, w1, b1, w2, b2):
    b = feat.shape[0]
    c = feat.shape[-1]
    s = feat.numel() // (b * c)
    d = w1.shape[0]
    dev = feat.device
    pooled = torch.empty((b, c_logical), dtype=torch.float32, device=dev)
    raw = torch.empty((2, b, d), dtype=torch.float32, device=dev)
    out = torch.empty((2, b, d), dtype=torch.float32, device=dev)
    call("rsp_head_fwd", ptr(feat), b, s, c, c_logical, d, ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(pooled),
         ptr(raw[0]), ptr(raw[1]), ptr(out[0]), ptr(out[1]), stream_ptr())
    return out[0], out[1], pooled, raw


def head_bwd(dout1, dout2, pooled, raw, feat_shape, w1, w2, want_dfeat=True):
    b = feat_shape[0]
    c = feat_shape[-1]
    s = 1
    for v in feat_shape[1:-1]:
        s *= v
    d, cl = w1.shape
    dev = w1.device
    dw = torch.zeros((2, d, cl), dtype=torch.float32, device=dev)
    db = torch.zeros((2, d), dtype=torch.float32, device=dev)
    dfeat = torch.empty(tuple(feat_shape), dtype=torch.bfloat16, device=dev) if want_dfeat else None
    dr_ws = torch.empty((b, 2, d), dtype=torch.float32, device=dev)
    call("rsp_head_bwd", ptr(dout1.contiguous()), ptr(dout2.contiguous()), ptr(pooled), ptr(raw[0]), ptr(raw[1]), b, s,
         c, cl, d, ptr(w1), ptr(w2), ptr(dr_ws), ptr(dw[0]), ptr(db[0]), ptr(dw[1]), ptr(db[1]), ptr(dfeat),
         stream_ptr())
    return dw[0], db[0], dw[1], db[1], dfeat


# ------------------------------------------------------------------------------------------------ layout
def to_ndhwc_bf16(x: torch.Tensor, c_stored: Optional[int] = None) -> torch.Tensor:
    """fp32 NCDHW -> bf16 NDHWC with zero-padded channels."""
    n, c, t, h, w = x.shape
    cs = c_stored or pad_channels(c)
    y = torch.empty((n, t, h, w, cs), dtype=torch.bfloat16, device=x.device)
    xc = x.contiguous().float()
    call("rsp_ncdhw_to_ndhwc_bf16", ptr(xc), ptr(y), n, c, cs, t * h * w, stream_ptr())
    return y


def to_ncdhw_f32(x: torch.Tensor, c_logical: int) -> torch.Tensor:
    n, t, h, w, cs = x.shape
    y = torch.empty((n, c_logical, t, h, w), dtype=torch.float32, device=x.device)
    call("rsp_ndhwc_bf16_to_ncdhw", ptr(x), ptr(y), n, c_logical, cs, t * h * w, stream_ptr())
    return y


# ------------------------------------------------------------------------------------------------ MoCo
def ema_update_(k_flat: torch.Tensor, q_flat: torch.Tensor, m: float):
    """In place k = k*m + q*(1-m) (builder_diffspeed_diffloss.py:337-343)."""
    call("rsp_ema_update", ptr(k_flat), ptr(q_flat), k_flat.numel(), float(m), float(1.0 - m), stream_ptr())


def sgd_step_(p, grad, mom, lr, momentum, weight_decay, grad_scale=1.0, first_step=False):
    call("rsp_sgd_step", ptr(p), ptr(grad), ptr(mom), p.numel(), float(lr), float(momentum), float(weight_decay),
         float(grad_scale), int(first_step), stream_ptr())


def speed_gather(im_q, im_k, perm, n_s1: int, d: int, layout: int, outs=None):
    """_diff_speed re-sampling. layout 0 -> fp32 NCDHW, 1 -> bf16 NDHWC(4).  ``outs`` = (q, k, k_neg) destination
    tensors (entries may be None): the key clips are written straight into the peer-visible exchange buffers."""
    b, c, t, h, w = im_q.shape
    tr = t // d
    dev = im_q.device
    shape, dtype = (((b, c, tr, h, w), torch.float32) if layout == 0 else ((b, tr, h, w, 4), torch.bfloat16))
    outs = list(outs) if outs is not None else [None, None, None]
    for i in range(3):
        if outs[i] is None:
            outs[i] = torch.empty(shape, dtype=dtype, device=dev)
        assert tuple(outs[i].shape) == shape and outs[i].dtype == dtype and outs[i].is_contiguous()
    call("rsp_speed_gather", ptr(im_q.contiguous()), ptr(im_k.contiguous()), ptr(perm), b, c, t, h, w, n_s1, d, layout,
         ptr(outs[0]), ptr(outs[1]), ptr(outs[2]), stream_ptr())
    return outs


def gather_rows(src: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """dst[i] = src[index[i]] along dim 0."""
    src = src.contiguous()
    row_bytes = src[0].numel() * src.element_size() if src.shape[0] else 0
    out = torch.empty((index.numel(),) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    if index.numel():
        call("rsp_gather_rows", ptr(src), ptr(index), ptr(out), index.numel(), row_bytes, stream_ptr())
    return out


def gather_rows_peer(peer_table: torch.Tensor, index: torch.Tensor, rows_per_peer: int, row_shape, dtype) -> torch.Tensor:
    """dst[i] = peer[index[i] // rows_per_peer][index[i] % rows_per_peer]; ``peer_table`` int64 [W] of device base
    pointers (NVLink-mapped peer buffers, own buffer at [rank]), ``index`` int64 on the device."""
    n = index.numel()
    out = torch.empty((n,) + tuple(row_shape), dtype=dtype, device=index.device)
    row_bytes = out[0].numel() * out.element_size()
    call("rsp_gather_rows_peer", ptr(peer_table), ptr(index), ptr(out), n, int(rows_per_peer), row_bytes, stream_ptr())
    return out


def invert_permutation(perm: torch.Tensor) -> torch.Tensor:
    """argsort of a permutation (int64, device): inv[perm[i]] = i."""
    inv = torch.empty_like(perm)
    call("rsp_invert_permutation", ptr(perm), ptr(inv), perm.numel(), stream_ptr())
    return inv


def queue_enqueue_(queue, keys, queue_ptr):
    d, k = queue.shape
    call("rsp_queue_enqueue", ptr(queue), ptr(keys.contiguous()), ptr(queue_ptr), d, k, keys.shape[0], stream_ptr())


def moco_logits_fwd(q_a, q_m, k_a, k_m, kn_a, kn_m, queue, temperature: float, materialize: bool = True):
    n, d = q_a.shape
    k = queue.shape[1]
    dev = q_a.device
    logits = torch.empty((2, n, k + 1), dtype=torch.float32, device=dev) if materialize else None
    rows = torch.empty((6, n), dtype=torch.float32, device=dev)  # lpos_m, lneg_m, lse1, lse2, pos1, pos2
    ws = torch.empty((_lib.load().rsp_moco_logits_workspace(n, k) // 4,), dtype=torch.float32, device=dev)
    ranks = torch.empty((2, n), dtype=torch.int32, device=dev)   # negatives beating each positive (zeroed by the call)
    call("rsp_moco_logits_fwd_ranked", ptr(q_a), ptr(q_m), ptr(k_a), ptr(k_m), ptr(kn_a), ptr(kn_m), ptr(queue), n, d, k,
         float(temperature), ptr(logits[0]) if materialize else None, ptr(logits[1]) if materialize else None,
         ptr(rows[0]), ptr(rows[1]), ptr(rows[2]), ptr(rows[3]), ptr(rows[4]), ptr(rows[5]), ptr(ws), ptr(ranks),
         stream_ptr())
    return logits, rows, ranks


def metrics_update(loss3, ranks, lpos_m, lneg_m, meters):
    """meters: fp32 [17] device buffer (val[8], sum[8], int32 count) — see rsp_metrics_update."""
    call("rsp_metrics_update", ptr(loss3), ptr(ranks), ptr(lpos_m), ptr(lneg_m), ranks.shape[1], ptr(meters),
         stream_ptr())


def moco_logits_bwd(q_a, q_m, k_a, k_m, kn_a, kn_m, queue, temperature, rows, g_rows, g_logits1, g_logits2):
    """rows = (lpos_m, lneg_m, lse1, lse2, pos1, pos2); g_rows same order."""
    n, d = q_a.shape
    k = queue.shape[1]
    dq = torch.empty((2, n, d), dtype=torch.float32, device=q_a.device)
    call("rsp_moco_logits_bwd", ptr(q_a), ptr(q_m), ptr(k_a), ptr(k_m), ptr(kn_a), ptr(kn_m), ptr(queue), n, d, k,
         float(temperature), ptr(rows[4]), ptr(rows[5]), ptr(rows[2]), ptr(rows[3]), ptr(g_rows[2]), ptr(g_rows[3]),
         ptr(g_rows[4]), ptr(g_rows[5]), ptr(g_rows[0]), ptr(g_rows[1]), ptr(g_logits1), ptr(g_logits2), ptr(dq[0]),
         ptr(dq[1]), stream_ptr())
    return dq[0], dq[1]


def moco_loss_fwd(rows, margin, a, m):
    out = torch.empty((3,), dtype=torch.float32, device=rows.device)
    call("rsp_moco_loss_fwd", ptr(rows[2]), ptr(rows[3]), ptr(rows[4]), ptr(rows[5]), ptr(rows[0]), ptr(rows[1]),
         rows.shape[1], float(margin), float(a), float(m), ptr(out), stream_ptr())
    return out


def moco_loss_bwd(rows, margin, a, m, g_out3):
    g = torch.empty_like(rows)
    call("rsp_moco_loss_bwd", ptr(rows[0]), ptr(rows[1]), rows.shape[1], float(margin), float(a), float(m),
         ptr(g_out3), ptr(g[2]), ptr(g[3]), ptr(g[4]), ptr(g[5]), ptr(g[0]), ptr(g[1]), stream_ptr())
    return g


def ce0_fwd(logits):
    n, l = logits.shape
    lse = torch.empty((n,), dtype=torch.float32, device=logits.device)
    call("rsp_ce0_fwd", ptr(logits), n, l, ptr(lse), stream_ptr())
    return lse


def ce0_bwd(logits, lse, g_scalar):
    n, l = logits.shape
    out = torch.empty_like(logits)
    call("rsp_ce0_bwd", ptr(logits), ptr(lse), n, l, ptr(g_scalar), ptr(out), stream_ptr())
    return out


# ------------------------------------------------------------------------------------------------ S3D-G
def gate_fwd(x: torch.Tensor, c_logical: int, w: torch.Tensor, b: torch.Tensor):
    """Self-gating (models/s3dg.py:64-72): y = x * sigmoid(W mean_S(x) + b). Returns y, pooled, gate."""
    n, c = x.shape[0], x.shape[-1]
    s = x.numel() // (n * c)
    dev = x.device
    sums = torch.empty((n, c), dtype=torch.float32, device=dev)
    pooled = torch.empty((n, c_logical), dtype=torch.float32, device=dev)
    gate = torch.empty((n, c), dtype=torch.float32, device=dev)
    y = torch.empty_like(x)
    w2 = w.detach().reshape(c_logical, c_logical).contiguous().float()
    call("rsp_gate_fwd", ptr(x), n, s, c, c_logical, ptr(w2), ptr(b), ptr(sums), ptr(pooled), ptr(gate), ptr(y),
         stream_ptr())
    return y, pooled, gate


def gate_bwd(dy, x, c_logical, w, pooled, gate):
    n, c = x.shape[0], x.shape[-1]
    s = x.numel() // (n * c)
    dev = x.device
    ws = torch.empty((2, n, c), dtype=torch.float32, device=dev)
    dw = torch.zeros((c_logical, c_logical), dtype=torch.float32, device=dev)
    db = torch.zeros((c_logical,), dtype=torch.float32, device=dev)
    dx = torch.empty_like(x)
    w2 = w.detach().reshape(c_logical, c_logical).contiguous().float()
    call("rsp_gate_bwd", ptr(dy), ptr(x), n, s, c, c_logical, ptr(w2), ptr(pooled), ptr(gate), ptr(ws[0]), ptr(ws[1]),
         ptr(dw), ptr(db), ptr(dx), stream_ptr())
    return dx, dw, db


def copy_channels(src, src_off, dst, dst_off, n_ch):
    m = src.numel() // src.shape[-1]
    call("rsp_copy_channels", ptr(src), src.shape[-1], src_off, ptr(dst), dst.shape[-1], dst_off, n_ch, m, stream_ptr())
