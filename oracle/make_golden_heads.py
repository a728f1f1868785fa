"""TEST INFRASTRUCTURE ONLY — golden vectors for the projection-head variants of ``MultiTaskWrapper``
(reference moco/split_wrapper.py:17-64,108-126): fc_type 'conv' (ConvFc) and 'convbn' (ConvBnFc), plus the
``finetune=True`` classifier branch (:104-106,131-135), from the UNMODIFIED reference on CPU.  R3D-18 backbone, 4 clips of
8x64x64, a fixed linear functional of the outputs as the loss.

Run in the build container only:  python oracle/make_golden_heads.py  ->  tests/golden/r3d18_heads.pt
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import ref_loader  # noqa: E402
from oracle.make_golden import initialize_seed, pack, summarize  # noqa: E402

CASES = [dict(fc_type="conv", finetune=False, groups=1), dict(fc_type="convbn", finetune=False, groups=1),
         dict(fc_type="convbn", finetune=False, groups=2), dict(fc_type="linear", finetune=True, groups=1)]
SHAPE = (4, 3, 8, 64, 64)
SEED = 3


def inputs():
    g = torch.Generator().manual_seed(77)
    x = torch.randn(SHAPE, generator=g)
    w1 = torch.randn(SHAPE[0], 128, generator=g)
    w2 = torch.randn(SHAPE[0], 128, generator=g)
    return x, w1, w2


def main():
    assert ref_loader.available(), "/root/reference is required to (re)generate goldens"
    torch.set_num_threads(8)
    mods = ref_loader.modules()
    out = {"shape": SHAPE, "seed": SEED, "cases": []}
    x, w1, w2 = inputs()
    for case in CASES:
        initialize_seed(SEED)
        model = mods["wrapper"].MultiTaskWrapper(ref_loader.backbone_ctor("resnet18"), num_classes=128, **case)
        model.train()
        init = {k: summarize(v.float()) for k, v in model.state_dict().items()}
        y = model(x)
        if case["finetune"]:
            loss = (y * w1).sum()
            outs = [y.detach().clone()]
        else:
            loss = (y[0] * w1).sum() + (y[1] * w2).sum()
            outs = [y[0].detach().clone(), y[1].detach().clone()]
        loss.backward()
        # the stem filter gradient (the end of the longest backward chain) is kept in full so that its direction is
        # gated over all 65,856 values, not over a 32-value prefix; the 7 M-value layer4 filter stays a summary
        grads = {k: (p.grad.detach().clone() if k == "encoder.conv1.weight" else pack(p.grad))
                 for k, p in model.named_parameters()
                 if p.grad is not None and (not k.startswith("encoder.") or k in ("encoder.conv1.weight",
                                                                                  "encoder.layer4.1.conv2.weight"))}
        out["cases"].append(dict(case=case, keys=list(model.state_dict().keys()), init=init, outs=outs,
                                 loss=float(loss), grads=grads))
        print(case, "loss", float(loss), "n grads", len(grads))
    path = ROOT / "tests" / "golden" / "r3d18_heads.pt"
    torch.save(out, path)
    print("written", path, path.stat().st_size)


if __name__ == "__main__":
    main()
