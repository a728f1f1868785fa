import sys
from pathlib import Path
import torch
import torch.nn.functional as F
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rspnet_b200 import ops
from rspnet_b200.moco.builder_diffspeed_diffloss import _LogitsFn, Loss

torch.manual_seed(0)
for (n, k, degenerate) in [(4, 64, False), (4, 64, True), (64, 16384, True), (4, 512, True)]:
    d, T = 128, 0.07
    base = F.normalize(torch.randn(1, d), dim=1)
    def mk():
        if degenerate:
            return F.normalize(base + 0.02 * torch.randn(n, d), dim=1)
        return F.normalize(torch.randn(n, d), dim=1)
    f = [mk() for _ in range(6)]
    queue = F.normalize(torch.randn(d, k), dim=0)
    qa, qm = f[0].clone().requires_grad_(True), f[1].clone().requires_grad_(True)
    l_pos1 = (qa * f[2]).sum(1, keepdim=True) / T
    l_pos2 = (qa * f[4]).sum(1, keepdim=True) / T
    l_neg = qa @ queue / T
    tgt = torch.zeros(n, dtype=torch.long)
    ce = F.cross_entropy(torch.cat([l_pos1, l_neg], 1), tgt) + F.cross_entropy(torch.cat([l_pos2, l_neg], 1), tgt)
    lpm = (qm * f[3]).sum(1) / T
    lnm = (qm * f[5]).sum(1) / T
    rank = torch.clamp(-(lpm - lnm) + 2.0, min=0).mean()
    (ce + rank).backward()
    gq_a, gq_m = f[0].cuda().requires_grad_(True), f[1].cuda().requires_grad_(True)
    l1, l2, a, b, rows, _ranks = _LogitsFn.apply(gq_a, gq_m, *[t.cuda() for t in f[2:]], queue.cuda(), T, True)
    for t in (l1, l2, a, b):
        t._rsp_rows = rows
    out = Loss(2.0, 1.0, 1.0)((l1, l2), tgt.cuda(), (a, b), torch.ones(n, dtype=torch.long).cuda())
    out[0].backward()
    print(n, k, degenerate, "loss", float(out[0]), float(ce + rank), "dq_a err", (gq_a.grad.cpu() - qa.grad).abs().max().item(),
          "ref max", qa.grad.abs().max().item(), "dq_m err", (gq_m.grad.cpu() - qm.grad).abs().max().item())
    # direct ops call
    logits, rows2, _ = ops.moco_logits_fwd(*[t.cuda() for t in f], queue.cuda(), T, True)
    g3 = torch.tensor([1.0, 0, 0], device="cuda")
    g_rows = ops.moco_loss_bwd(rows2, 2.0, 1.0, 1.0, g3)
    dqa, dqm = ops.moco_logits_bwd(*[t.cuda() for t in f], queue.cuda(), T, rows2, g_rows, None, None)
    print("   direct dq_a err", (dqa.cpu() - qa.grad).abs().max().item(), "g_rows", g_rows[:, 0].tolist())
