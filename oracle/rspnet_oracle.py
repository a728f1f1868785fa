"""TEST INFRASTRUCTURE ONLY — CPU restatement (plain torch fp32 ops) of the RSPNet pretraining step.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file;
the product package (rspnet_b200/) never does.  Every function cites the reference lines it restates
(paths relative to the reference root).  Pinning: tests/test_oracle_cpu.py checks this file against the golden
vectors in tests/golden/ (generated from the unmodified reference by oracle/make_golden.py) and, where
/root/reference is present, against the live reference modules.

The restatement is functional: all state lives in a flat ``dict`` keyed exactly like
``MoCoDiffLossTwoFc.state_dict()`` ('queue', 'queue_ptr', 'encoder_q.encoder.conv1.weight', ...), and W data-parallel
ranks are simulated in one process (all_gather == concatenation in rank order), so multi-rank semantics
(shuffle-BN, key gather, gradient averaging) are covered without a process group.
"""
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

State = Dict[str, torch.Tensor]

# When True the restatement rounds to bf16 at exactly the points where the B200 path stores bf16 (network input,
# conv operands, conv output, output of every fused BN(+residual)(+ReLU)); accumulation stays fp32.  This separates
# "precision of the stated bf16 conv path" from "logic" in the GPU parity tests.  Default: exact fp32 reference math.
EMULATE_BF16 = False


def _r(x):
    return x.bfloat16().float() if EMULATE_BF16 else x


# --------------------------------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------------------------------
def _bn(x, sd: State, name: str, train: bool, eps=1e-5, momentum=0.1):
    """nn.BatchNorm3d forward (train mode updates running stats in ``sd`` in place, like the module does)."""
    rm, rv = sd[name + ".running_mean"], sd[name + ".running_var"]
    y = F.batch_norm(x, rm, rv, sd[name + ".weight"], sd[name + ".bias"], train, momentum, eps)
    if train and (name + ".num_batches_tracked") in sd:
        sd[name + ".num_batches_tracked"] += 1
    return y


def _conv(x, sd: State, name: str, stride, padding):
    return _r(F.conv3d(_r(x), _r(sd[name + ".weight"]), sd.get(name + ".bias"), stride, padding))


def resnet18_feature(x, sd: State, p: str, train=True):
    """models/resnet.py:203-213 (get_feature) with BasicBlock (:59-77), layers [2,2,2,2], shortcut B (:170-175)."""
    x = _conv(x, sd, p + "conv1", (1, 2, 2), (3, 3, 3))
    x = _r(F.relu(_bn(x, sd, p + "bn1", train)))
    x = F.max_pool3d(x, 3, 2, 1)
    for li, planes in enumerate((64, 128, 256, 512), start=1):
        for bi in range(2):
            b = f"{p}layer{li}.{bi}."
            stride = 2 if (li > 1 and bi == 0) else 1
            residual = x
            out = _conv(x, sd, b + "conv1", stride, 1)
            out = _r(F.relu(_bn(out, sd, b + "bn1", train)))
            out = _conv(out, sd, b + "conv2", 1, 1)
            out = _bn(out, sd, b + "bn2", train)
            if (b + "downsample.0.weight") in sd:
                residual = _conv(x, sd, b + "downsample.0", stride, 0)
                residual = _r(_bn(residual, sd, b + "downsample.1", train))
            x = _r(F.relu(out + residual))
    return x


def c3d_feature(x, sd: State, p: str, train=True):
    """models/c3d.py:111-150 (get_feature, return_conv=False): 8 conv(bias)+BN+ReLU, pools 1-4, no pool5."""
    def cbr(x, c, b):
        return _r(F.relu(_bn(_conv(x, sd, p + c, 1, 1), sd, p + b, train)))
    x = F.max_pool3d(cbr(x, "conv1", "bn1"), (1, 2, 2), (1, 2, 2))
    x = F.max_pool3d(cbr(x, "conv2", "bn2"), 2, 2)
    x = F.max_pool3d(cbr(cbr(x, "conv3a", "bn3a"), "conv3b", "bn3b"), 2, 2)
    x = F.max_pool3d(cbr(cbr(x, "conv4a", "bn4a"), "conv4b", "bn4b"), 2, 2)
    x = cbr(cbr(x, "conv5a", "bn5a"), "conv5b", "bn5b")
    return x


def r2plus1d_feature(x, sd: State, p: str, train=True):
    """models/r2plus1d_vcop.py:218-224 (get_feature) with R2Plus1DNet((1,1,1,1)): SpatioTemporalConv (:13-72) =
    (1,k,k) conv -> BN -> ReLU -> (k,1,1) conv; SpatioTemporalResBlock (:75-123)."""
    def stconv(x, name, k, stride, pad):
        x = _conv(x, sd, name + ".spatial_conv", (1, stride[1], stride[2]), (0, pad[1], pad[2]))
        x = _r(F.relu(_bn(x, sd, name + ".bn", train)))
        return _conv(x, sd, name + ".temporal_conv", (stride[0], 1, 1), (pad[0], 0, 0))

    x = stconv(x, p + "conv1", (3, 7, 7), (1, 2, 2), (1, 3, 3))
    x = _r(F.relu(_bn(x, sd, p + "bn1", train)))
    for li, down in ((2, False), (3, True), (4, True), (5, True)):
        b = f"{p}conv{li}.block1."
        s = (2, 2, 2) if down else (1, 1, 1)
        res = stconv(x, b + "conv1", (3, 3, 3), s, (1, 1, 1))
        res = _r(F.relu(_bn(res, sd, b + "bn1", train)))
        res = stconv(res, b + "conv2", (3, 3, 3), (1, 1, 1), (1, 1, 1))
        res = _bn(res, sd, b + "bn2", train)
        if down:
            x = stconv(x, b + "downsampleconv", (1, 1, 1), (2, 2, 2), (0, 0, 0))
            x = _r(_bn(x, sd, b + "downsamplebn", train))
        x = _r(F.relu(x + res))
    return x


_S3DG_INC = ["sepInc_3b", "sepInc_3c", "sepInc_4b", "sepInc_4c", "sepInc_4d", "sepInc_4e", "sepInc_4f", "sepInc_5b",
             "sepInc_5c"]
# models/s3dg.py:105-121, in order: (name, kind, arguments)
S3DG_STAGES = [("sepConv1", "sep", (7, 2, 3)), ("maxPool1", "pool", ((1, 3, 3), (1, 2, 2), (0, 1, 1))),
               ("basicConv3d", "basic", ()), ("sep_conv2", "sep", (3, 1, 1)),
               ("maxPool2", "pool", ((1, 3, 3), (1, 2, 2), (0, 1, 1))), ("sepInc_3b", "inc", ()), ("sepInc_3c", "inc", ()),
               ("maxPool3", "pool", (3, 2, 1)), ("sepInc_4b", "inc", ()), ("sepInc_4c", "inc", ()), ("sepInc_4d", "inc", ()),
               ("sepInc_4e", "inc", ()), ("sepInc_4f", "inc", ()), ("maxpool4", "pool", (2, 2, 0)), ("sepInc_5b", "inc", ()),
               ("sepInc_5c", "inc", ())]


def _s3dg_basic(x, sd, name, train, stride=(1, 1, 1), pad=(0, 0, 0)):
    """BasicConv3d (models/s3dg.py:6-33): conv (no bias) -> BN(eps 1e-3, momentum 0.001) -> ReLU."""
    y = _conv(x, sd, name + ".conv3d", stride, pad)
    return _r(F.relu(_bn(y, sd, name + ".bn", train, eps=1e-3, momentum=0.001)))


def _s3dg_sep(x, sd, name, train, k, stride, pad):
    """sep_conv with self-gating (models/s3dg.py:36-72)."""
    x = _s3dg_basic(x, sd, name + ".sep_conv.0", train, (stride, stride, stride), (0, pad, pad))
    x = _s3dg_basic(x, sd, name + ".sep_conv.1", train, (1, 1, 1), (pad, 0, 0))
    w = x.mean(dim=(2, 3, 4), keepdim=True)
    w = torch.sigmoid(F.conv3d(w, sd[name + ".excitation.weight"], sd[name + ".excitation.bias"]))
    return _r(w * x)


def _s3dg_inc(x, sd, name, train):
    """sep_inc (models/s3dg.py:74-99)."""
    o0 = _s3dg_basic(x, sd, name + ".branch0", train)
    o1 = _s3dg_sep(_s3dg_basic(x, sd, name + ".branch1.0", train), sd, name + ".branch1.1", train, 3, 1, 1)
    o2 = _s3dg_sep(_s3dg_basic(x, sd, name + ".branch2.0", train), sd, name + ".branch2.1", train, 3, 1, 1)
    o3 = _s3dg_basic(F.max_pool3d(x, 3, 1, 1), sd, name + ".branch3.1", train)
    return torch.cat((o0, o1, o2, o3), 1)


def s3dg_stage(x, sd: State, p: str, stage: str, train=True):
    """One entry of S3D_G.feature (models/s3dg.py:105-121) on its own; ``p`` is the prefix of the backbone's keys."""
    kind, args = next((k, a) for n, k, a in S3DG_STAGES if n == stage)
    name = p + "feature." + stage
    if kind == "pool":
        return F.max_pool3d(x, *args)
    if kind == "basic":
        return _s3dg_basic(x, sd, name, train)
    if kind == "sep":
        return _s3dg_sep(x, sd, name, train, *args)
    return _s3dg_inc(x, sd, name, train)


def s3dg_feature(x, sd: State, p: str, train=True, upto: Optional[str] = None, taps: Optional[list] = None):
    """models/s3dg.py:151-153 (get_feature): the 16 stages in order.  ``taps`` (a list) receives every stage input."""
    for stage, _, _ in S3DG_STAGES:
        if taps is not None:
            taps.append((stage, x))
        x = s3dg_stage(x, sd, p, stage, train)
        if upto == stage:
            break
    return x


FEATURES = {"resnet18": resnet18_feature, "c3d": c3d_feature, "r2plus1d-vcop": r2plus1d_feature,
            "s3dg": s3dg_feature}


def wrapper_forward(arch: str, x, sd: State, p: str, train=True):
    """moco/split_wrapper.py:128-152 with fc_type='linear' (:164-169), groups=1: two pooled linear heads, L2-normalised."""
    feat = FEATURES[arch](_r(x), sd, p + "encoder.", train)
    pooled = feat.mean(dim=(2, 3, 4))
    x1 = F.linear(pooled, sd[p + "fc1.2.weight"], sd[p + "fc1.2.bias"])
    x2 = F.linear(pooled, sd[p + "fc2.2.weight"], sd[p + "fc2.2.bias"])
    return F.normalize(x1, dim=1), F.normalize(x2, dim=1)


def param_names(sd: State, prefix: str) -> List[str]:
    """Parameter (not buffer) keys under ``prefix`` in state_dict order."""
    skip = ("running_mean", "running_var", "num_batches_tracked")
    return [k for k in sd if k.startswith(prefix) and not k.endswith(skip)]


# --------------------------------------------------------------------------------------------------------------
# MoCo pieces
# --------------------------------------------------------------------------------------------------------------
def momentum_update(sd: State, m: float):
    """moco/builder_diffspeed_diffloss.py:337-343: k = k*m + q*(1-m) for every parameter (buffers untouched)."""
    for kq in param_names(sd, "encoder_q."):
        kk = "encoder_k." + kq[len("encoder_q."):]
        sd[kk] = sd[kk] * m + sd[kq] * (1. - m)


def diff_speed(im_q, im_k, perm, d: int, alpha=0.5):
    """builder:421-443 (the re-sampling part of _diff_speed; ``perm`` is the randperm(B) drawn at :424)."""
    B, C, T, H, W = im_q.shape
    s1, s2 = perm[:int(B * alpha)], perm[int(B * alpha):]
    t_real = T // d
    speed1 = torch.arange(0, T, 1)[:t_real]
    speed2 = torch.arange(0, T, d)[:t_real]
    q = torch.empty(B, C, t_real, H, W)
    k = torch.empty_like(q)
    kneg = torch.empty_like(q)
    q[s1] = im_q.index_select(0, s1).index_select(2, speed1)
    q[s2] = im_q.index_select(0, s2).index_select(2, speed2)
    k[s1] = im_k.index_select(0, s1).index_select(2, speed1)
    k[s2] = im_k.index_select(0, s2).index_select(2, speed2)
    kneg[s1] = im_k.index_select(0, s1).index_select(2, speed2)
    kneg[s2] = im_k.index_select(0, s2).index_select(2, speed1)
    return q, k, kneg


def forward_encoder_k(arch, xs: Sequence[torch.Tensor], sds: Sequence[State], idx_shuffle):
    """builder:408-419 over W simulated ranks: shuffle (:361-387), encoder_k per rank, unshuffle (:389-406).

    xs[r] is rank r's batch, sds[r] its state (BN running stats of encoder_k are per rank).
    Returns per-rank (k_A, k_M) and the per-rank shuffled inputs (for exchange tests)."""
    W = len(xs)
    B = xs[0].shape[0]
    x_gather = torch.cat(list(xs), 0)
    idx_unshuffle = torch.argsort(idx_shuffle)
    shuffled = [x_gather[idx_shuffle.view(W, -1)[r]] for r in range(W)]
    outs = [wrapper_forward(arch, shuffled[r], sds[r], "encoder_k.", train=True) for r in range(W)]
    ka_all = torch.cat([o[0] for o in outs], 0)
    km_all = torch.cat([o[1] for o in outs], 0)
    res = []
    for r in range(W):
        idx_this = idx_unshuffle.view(W, -1)[r]
        res.append((ka_all[idx_this], km_all[idx_this]))
    assert all(t[0].shape[0] == B for t in res)
    return res, shuffled


def logits(q_a, q_m, k_a, k_m, kn_a, kn_m, queue, T: float):
    """builder:521-538."""
    l_pos_a1 = torch.einsum('nc,nc->n', [q_a, k_a]).unsqueeze(-1)
    l_pos_a2 = torch.einsum('nc,nc->n', [q_a, kn_a]).unsqueeze(-1)
    l_pos_m = torch.einsum('nc,nc->n', [q_m, k_m]).unsqueeze(-1)
    l_neg_a = torch.einsum('nc,ck->nk', [q_a, queue.clone().detach()])
    l_neg_m = torch.einsum('nc,nc->n', [q_m, kn_m]).unsqueeze(-1)
    l_pos_a1, l_pos_a2, l_neg_a, l_pos_m, l_neg_m = (v / T for v in (l_pos_a1, l_pos_a2, l_neg_a, l_pos_m, l_neg_m))
    return (torch.cat([l_pos_a1, l_neg_a], 1), torch.cat([l_pos_a2, l_neg_a], 1)), (l_pos_m, l_neg_m)


def loss(logits_a, logits_m, margin=2.0, A=1.0, M=1.0):
    """builder:263-283 (Loss.forward) with the torch-1.6 broadcasting of MarginRankingLoss written out."""
    target = torch.zeros(logits_a[0].shape[0], dtype=torch.long)
    ce1 = F.cross_entropy(logits_a[0], target)
    ce2 = F.cross_entropy(logits_a[1], target)
    ranking = torch.clamp(-(logits_m[0] - logits_m[1]) + margin, min=0).mean()
    return A * (ce1 + ce2) + M * ranking, ce1 + ce2, ranking


def enqueue(sd: State, keys_all):
    """builder:345-359 with ``keys_all`` already gathered in rank order."""
    n = keys_all.shape[0]
    K = sd["queue"].shape[1]
    ptr = int(sd["queue_ptr"])
    assert K % n == 0
    sd["queue"][:, ptr:ptr + n] = keys_all.T
    sd["queue_ptr"][0] = (ptr + n) % K


def single_head_encoder(x, sd: State, p: str, train=True):
    """models/resnet.py:215-223 (get_output_and_feature / forward) for resnet18: AvgPool3d over the (1, 4, 4) feature map of
    a 16 x 112 x 112 clip, then ``fc``; L2-normalised as builder:177,207 do."""
    feat = resnet18_feature(_r(x), sd, p, train)
    pooled = F.avg_pool3d(feat, (1, 4, 4), stride=1).flatten(1)
    return F.normalize(F.linear(pooled, sd[p + "fc.weight"], sd[p + "fc.bias"]), dim=1)


def single_head_forward(sd: State, im_q, im_k, perm, idx_shuffles, *, d=2, m=0.999, T=0.07):
    """MoCoDiffLoss.forward (builder:184-245), one process (the shuffles permute within the batch): returns
    (logits1, logits2), (l_pos, l_neg_speed), q (with grad) and enqueues ``k``.  ``sd`` holds leaves for encoder_q."""
    momentum_update(sd, m)
    q_in, k_in, kneg_in = diff_speed(im_q, im_k, perm, d)

    def enc_k(x, idx):                                       # builder:171-183
        with torch.no_grad():
            k = single_head_encoder(x[idx], sd, "encoder_k.", True)
        return k[torch.argsort(idx)]

    speed_k = enc_k(kneg_in, idx_shuffles[0])
    k = enc_k(k_in, idx_shuffles[1])
    q = single_head_encoder(q_in, sd, "encoder_q.", True)
    l_pos = torch.einsum('nc,nc->n', [q, k]).unsqueeze(-1) / T
    l_neg = torch.einsum('nc,ck->nk', [q, sd["queue"].clone().detach()]) / T
    l_neg_speed = torch.einsum('nc,nc->n', [q, speed_k]).unsqueeze(-1) / T
    out = (torch.cat([l_pos, l_neg], 1), torch.cat([l_neg_speed, l_neg], 1)), (l_pos, l_neg_speed)
    enqueue(sd, k)
    return out


# --------------------------------------------------------------------------------------------------------------
# one full training step over W simulated ranks
# --------------------------------------------------------------------------------------------------------------
def train_step(arch: str, sds: List[State], im_q: List[torch.Tensor], im_k: List[torch.Tensor],
               perms: List[torch.Tensor], idx_shuffles: Tuple[torch.Tensor, torch.Tensor], *, d=2, m=0.999, T=0.07,
               margin=2.0, A=1.0, M=1.0, lr=0.1, momentum=0.9, weight_decay=1e-4,
               mom_bufs: Optional[Dict[str, torch.Tensor]] = None, do_update=True):
    """MoCoDiffLossTwoFc.forward (builder:492-547) + Loss + backward + SGD (pretrain.py:154-165) for W ranks.

    sds[r]: rank r's state (identical parameters / queue on all ranks; BN buffers may differ).
    perms[r]: rank r's randperm(B) (builder:424). idx_shuffles: rank 0's two randperm(B*W) (builder:375) for the
    k_neg pass and the k pass, in that order.  Returns a dict of per-rank outputs and the averaged gradients.
    """
    W = len(sds)
    for sd in sds:
        momentum_update(sd, m)
    views = [diff_speed(im_q[r], im_k[r], perms[r], d) for r in range(W)]
    kneg, shuf_neg = forward_encoder_k(arch, [v[2] for v in views], sds, idx_shuffles[0])
    kpos, shuf_pos = forward_encoder_k(arch, [v[1] for v in views], sds, idx_shuffles[1])
    kneg = [(a.detach(), b.detach()) for a, b in kneg]
    kpos = [(a.detach(), b.detach()) for a, b in kpos]
    names = param_names(sds[0], "encoder_q.")
    out = {"logits_a": [], "logits_m": [], "loss": [], "q": [], "k": kpos, "kneg": kneg, "views": views,
           "shuffled": (shuf_neg, shuf_pos)}
    grads = None
    for r in range(W):
        sd = dict(sds[r])
        leaves = {}
        for n in names:
            leaves[n] = sds[r][n].detach().clone().requires_grad_(True)
            sd[n] = leaves[n]
        q_a, q_m = wrapper_forward(arch, views[r][0], sd, "encoder_q.", train=True)
        for key in sds[r]:  # BN buffers updated by the functional forward
            if key.startswith("encoder_q.") and key not in leaves:
                sds[r][key] = sd[key]
        la, lm = logits(q_a, q_m, kpos[r][0], kpos[r][1], kneg[r][0], kneg[r][1], sds[r]["queue"], T)
        total, ce, rank = loss(la, lm, margin, A, M)
        used = [n for n in names if not n.startswith("encoder_q.encoder.fc.") and
                not n.startswith("encoder_q.encoder.linear.")]  # classifier heads are never on the path
        gs = torch.autograd.grad(total, [leaves[n] for n in used], allow_unused=True)
        g = {n: (t if t is not None else torch.zeros_like(leaves[n])) for n, t in zip(used, gs)}
        grads = g if grads is None else {n: grads[n] + g[n] for n in g}
        out["logits_a"].append(tuple(t.detach() for t in la))
        out["logits_m"].append(tuple(t.detach() for t in lm))
        out["loss"].append((total.detach(), ce.detach(), rank.detach()))
        out["q"].append((q_a.detach(), q_m.detach()))
    grads = {n: g / W for n, g in grads.items()}  # DDP averages (moco/__init__.py:49-53)
    out["grads"] = grads
    keys_all = torch.cat([kneg[r][0] for r in range(W)], 0)
    out["keys_all"] = keys_all
    for sd in sds:
        enqueue(sd, keys_all)
    if do_update:
        # torch.optim.SGD (pretrain.py:65-72): params without grad (encoder.fc / encoder.linear) are skipped
        for n, g in grads.items():
            for sd in sds:
                p = sd[n]
                dp = g + weight_decay * p
                if mom_bufs is not None:
                    key = n
                    if key not in mom_bufs:
                        mom_bufs[key] = dp.clone()
                    elif sd is sds[0]:
                        mom_bufs[key] = momentum * mom_bufs[key] + dp
                    dp = mom_bufs[key]
                sd[n] = p - lr * dp
    return out
