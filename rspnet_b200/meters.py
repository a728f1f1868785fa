"""Contrastive training metrics without top-k and without host syncs.

The reference's loop (pretrain.py:169-196) runs ``accuracy(output, target, topk=(1, 5))`` — a ``topk`` over
``[N, 1+K]`` — three times per step and feeds eight ``AverageMeter``s (framework/meters/average.py:4-44).  Here the
logits kernel already counts, per row, how many queue negatives beat the positive (``rsp_moco_logits_fwd_ranked``), so
top-1 is ``rank == 0`` and top-5 ``rank < 5``; ``rsp_metrics_update`` folds a step into device-side val / sum / count
and the host reads them only when it logs.
"""
from typing import Dict, List, Sequence

import torch
import torch.distributed as dist
from torch import Tensor

from . import ops

NAMES = ("Loss", "Loss_A", "Acc@1_A", "Acc@5_A", "Acc@1_A_n", "Acc@5_A_n", "Loss_M", "Acc@1_M")
_FMT = {"Loss": ":f", "Loss_A": ":f", "Loss_M": ":f"}


def accuracy(output: Tensor, target: Tensor, topk: Sequence[int] = (1,)) -> List[Tensor]:
    """framework/metrics/classification.py:6-20.  Logits produced by ``MoCoDiffLoss*.forward`` carry their rank
    counters, which answer the question exactly (target is class 0 by construction); any other tensor takes the
    reference's topk route."""
    ranks = getattr(output, "_rsp_ranks", None)
    with torch.no_grad():
        batch_size = target.size(0)
        if ranks is not None and hasattr(output, "_rsp_slot"):
            r = ranks[output._rsp_slot]          # logits1 -> slot 0, logits2 -> slot 1
            return [(r < k).sum(dtype=torch.float) * (100.0 / batch_size) for k in topk]
        maxk = max(topk)
        _, pred = output.topk(maxk, 1, True, True)
        correct = pred.t().eq(target[None])
        return [correct[:k].flatten().sum(dtype=torch.float) * (100.0 / batch_size) for k in topk]


class ContrastiveMeters:
    """The eight meters of pretrain.Engine (pretrain.py:97-106) as one device buffer."""

    def __init__(self, device):
        self.buf = torch.zeros(17, dtype=torch.float32, device=device)   # val[8], sum[8], int32 count

    def reset(self):
        self.buf.zero_()

    @torch.no_grad()
    def update(self, loss3, output, ranking_logits):
        """loss3: (loss, loss_A, loss_M) as returned by ``Loss.forward`` / ``PretrainEngine.step``; output /
        ranking_logits: what ``model(clip_q, clip_k)`` returned.  Enqueues one tiny kernel; never synchronises."""
        ranks = getattr(output[0], "_rsp_ranks", None)
        if ranks is None:
            raise RuntimeError("ContrastiveMeters.update needs the logits returned by rspnet_b200's MoCoDiffLoss* forward")
        if torch.is_tensor(loss3):
            l3 = loss3.detach()
        elif (loss3[1].data_ptr() == loss3[0].data_ptr() + 4 and loss3[2].data_ptr() == loss3[0].data_ptr() + 8
              and loss3[0].dtype == torch.float32):
            l3 = loss3[0].detach().as_strided((3,), (1,))    # the fused loss kernel wrote the triple side by side
        else:
            l3 = torch.stack([t.detach().reshape(()) for t in loss3]).float()
        ops.metrics_update(l3.contiguous(), ranks, ranking_logits[0].detach().reshape(-1).contiguous(),
                           ranking_logits[1].detach().reshape(-1).contiguous(), self.buf)

    def sync_distributed(self):
        """framework/meters/average.py:40-44: sums and counts added over ranks."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.buf[8:16])                          # fp32 sums
            dist.all_reduce(self.buf[16:17].view(torch.int32))       # the count is an int32 in the last slot

    def summary(self) -> Dict[str, Dict[str, float]]:
        """One device -> host read: {name: {"val": last value, "avg": sum / count}}."""
        host = self.buf.cpu()
        count = int(host[16:17].view(torch.int32)[0])
        return {n: {"val": float(host[i]), "avg": float(host[8 + i]) / max(count, 1)} for i, n in enumerate(NAMES)}

    def __str__(self):
        s = self.summary()
        parts = []
        for n in NAMES:
            f = "{:f}" if n in _FMT else "{:6.2f}"
            parts.append(f"{n} {f.format(s[n]['val'])} ({f.format(s[n]['avg'])})")
        return "\t".join(parts)
