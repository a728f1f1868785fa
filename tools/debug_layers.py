"""Diagnostic: layer-by-layer forward and per-parameter gradient comparison between the B200 path and the
bf16-emulating oracle on the r3d18_w1 / c3d_w1 fixture inputs (GPU box only)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import build_product_moco, load_golden, make_inputs  # noqa: E402
from oracle import rspnet_oracle as oracle  # noqa: E402
from rspnet_b200 import nn as rnn, ops  # noqa: E402
from rspnet_b200.moco import Loss  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "r3d18_w1"
g = load_golden(name)
cfg, hyper = g["config"], g["hyper"]
rec = g["ranks"][0]["steps"][0]
draws = [rec["perm"], rec["idx_shuffle_neg"], rec["idx_shuffle_pos"]]
im_q, im_k = make_inputs(cfg, 0, 0)

# ---------------- oracle side (records conv inputs by parameter name)
sd = {k: v.clone() for k, v in build_product_moco(cfg, hyper).state_dict().items()}
o_in = {}
orig_conv = oracle._conv


def rec_conv(x, sdd, nm, stride, padding):
    o_in.setdefault(nm, []).append(oracle._r(x).detach().clone())
    return orig_conv(x, sdd, nm, stride, padding)


oracle._conv = rec_conv
oracle.EMULATE_BF16 = True
emu = oracle.train_step(cfg["arch"], [sd], [im_q], [im_k], [draws[0]], (draws[1], draws[2]), d=2, m=hyper["m"],
                        T=hyper["T"], margin=hyper["margin"], do_update=False)
oracle.EMULATE_BF16 = False
oracle._conv = orig_conv

# ---------------- product side
model = build_product_moco(cfg, hyper).cuda()
names = {id(m): n for n, m in model.named_modules()}
p_in = {}
orig_cba = rnn.conv_bn_act


def rec_cba(x, conv, bn, relu=True, residual=None):
    ci = conv.weight.shape[1]
    p_in.setdefault(names[id(conv)], []).append(ops.to_ncdhw_f32(x.detach(), ci).cpu())
    return orig_cba(x, conv, bn, relu, residual)


rnn.conv_bn_act = rec_cba
import rspnet_b200.models.resnet as R, rspnet_b200.models.c3d as C3  # noqa
_mlb, _hb = ops.moco_logits_bwd, ops.head_bwd


def dbg_mlb(q_a, q_m, k_a, k_m, kn_a, kn_m, queue, T, rows, g_rows, gl1, gl2):
    r = _mlb(q_a, q_m, k_a, k_m, kn_a, kn_m, queue, T, rows, g_rows, gl1, gl2)
    print("LOGITS_BWD rows", rows.cpu().tolist())
    print("LOGITS_BWD g_rows", g_rows.cpu().tolist(), "gl1", gl1 is not None, "gl2", gl2 is not None)
    print("LOGITS_BWD dq_a norm", r[0].norm().item(), "dq_m norm", r[1].norm().item(),
          "cos(q_a,k_a)", (q_a * k_a).sum(1).cpu().tolist())
    return r


def dbg_hb(d1, d2, pooled, raw, shape, w1, w2, want_dfeat=True):
    print("HEAD_BWD dout1 norm", d1.norm().item(), "dout2 norm", d2.norm().item(), "shape", shape)
    r = _hb(d1, d2, pooled, raw, shape, w1, w2, want_dfeat)
    print("HEAD_BWD dw1 norm", r[0].norm().item(), "dw2 norm", r[2].norm().item(), "dfeat norm", r[4].float().norm().item())
    return r


it = iter(draws)
orig_rp = torch.randperm
torch.randperm = lambda n, *a, **k: (lambda r: r.to(k["device"]) if "device" in k else r.clone())(next(it))
out, tgt, rl, rt = model(im_q.cuda(), im_k.cuda())
torch.randperm = orig_rp
loss, ce, rank = Loss(hyper["margin"], 1.0, 1.0)(out, tgt, rl, rt)
loss.backward()
torch.cuda.synchronize()
print("loss product", [float(loss), float(ce), float(rank)], "emu", [float(t) for t in emu["loss"][0]], "ref", rec["loss"].tolist())
print("logits max diff vs emu", (out[0].detach().cpu() - emu["logits_a"][0][0]).abs().max().item(),
      "vs ref", (out[0].detach().cpu() - rec["logits1"]).abs().max().item())
print("ORACLE q_a.k_a", (emu["q"][0][0] * emu["k"][0][0]).sum(1).tolist(), "logits_m", [t.flatten().tolist() for t in emu["logits_m"][0]])
# forward: encoder_k pass 1 (k_neg), pass 2 (k), then encoder_q
for nm in p_in:
    for i, (a, b) in enumerate(zip(p_in[nm], o_in.get(nm, []))):
        d = (a - b).abs().max().item()
        print(f"fwd in {nm:45s} pass{i} maxdiff {d:.4f} absmax {b.abs().max().item():.3f}")
# gradients
named = dict(model.named_parameters())
for k, gref in emu["grads"].items():
    got = named[k].grad
    if got is None:
        print(f"grad {k:50s} MISSING (ref norm {gref.norm().item():.4e})")
        continue
    got = got.cpu()
    cos = float((got.flatten().double() @ gref.flatten().double()) / (got.norm().double() * gref.norm().double() + 1e-30))
    print(f"grad {k:50s} cos {cos:+.4f} norm ratio {got.norm().item() / (gref.norm().item() + 1e-30):.4f} refnorm {gref.norm().item():.3e}")
