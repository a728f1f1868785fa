"""TEST INFRASTRUCTURE ONLY — golden vectors for the clip pipeline (subsystem 4) from the UNMODIFIED reference
transforms (datasets/transforms_video/*): random decisions under a fixed python-random seed, and the per-clip GPU
transform chain (ToTensorVideo -> Resize -> gray -> flip -> Normalize) evaluated on CPU.
Run in the build container only:  python oracle/make_golden_sampler.py"""
import json
import random
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, "/root/reference")
from datasets.transforms_video import transforms_spatial, transforms_temporal  # noqa: E402
from datasets.transforms_video import functional_tensor as FT  # noqa: E402

OUT = ROOT / "tests" / "golden"


def draws():
    rec = []
    random.seed(1234)
    for size, strides, n in [(32, [{'stride': 1, 'weight': 1}], 300), (32, [{'stride': 1, 'weight': 1}], 20),
                             (32, [{'stride': 2, 'weight': 1}], 50), (16, [{'stride': 1, 'weight': 8}, {'stride': 2, 'weight': 1},
                                                                          {'stride': 4, 'weight': 1}], 120),
                             (8, [{'stride': 4, 'weight': 1}], 9)]:
        tc = transforms_temporal.RandomStrideCrop(size=size, strides=[dict(s) for s in strides])
        for _ in range(3):
            rec.append(dict(kind="temporal", size=size, strides=strides, n=n,
                            out=[int(v) for v in tc(np.arange(n))]))
    crop = transforms_spatial.RawVideoRandomCrop(scale=(0.4, 1.0))
    for (h, w) in [(128, 171), (240, 320), (50, 400), (400, 50)]:
        for _ in range(4):
            clip = torch.zeros(2, h, w, 3, dtype=torch.uint8)
            rec.append(dict(kind="crop", h=h, w=w, out=[int(v) for v in crop.get_params(clip)]))
    rec.append(dict(kind="state", value=random.random()))
    return rec


def pipeline():
    g = torch.Generator().manual_seed(7)
    frames = torch.randint(0, 256, (10, 40, 52, 3), generator=g, dtype=torch.uint8)
    idx = torch.tensor([[0, 1, 2, 3], [9, 7, 5, 3], [2, 2, 4, 4]], dtype=torch.int32)
    box = torch.tensor([[3, 5, 30, 40], [0, 0, 40, 52], [10, 20, 17, 13]], dtype=torch.int32)
    flags = torch.tensor([0, 1 | 2, 1], dtype=torch.uint8)
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    size = 16
    outs = []
    for c in range(idx.shape[0]):
        clip = frames[idx[c].long()]                                   # [T,H,W,3] uint8 (vr.get_batch)
        i, j, h, w = box[c].tolist()
        clip = clip[..., i:i + h, j:j + w, :].contiguous()             # RawVideoCrop.__call__
        x = transforms_spatial.ToTensor()(clip)
        x = transforms_spatial.Resize(size)(x)
        if int(flags[c]) & 2:
            x = FT.rgb_to_grayscale(x)
        if int(flags[c]) & 1:
            x = x.flip(-1)                                             # RandomHorizontalFlipVideo when it fires
        x = transforms_spatial.Normalize(mean, std, inplace=True)(x)
        outs.append(x)
    return dict(frames=frames, idx=idx, box=box, flags=flags, mean=mean, std=std, size=size, out=torch.stack(outs))


if __name__ == "__main__":
    (OUT / "sampler_draws.json").write_text(json.dumps(draws()))
    torch.save(pipeline(), OUT / "sampler_clip.pt")
    print("written", (OUT / "sampler_draws.json").stat().st_size, (OUT / "sampler_clip.pt").stat().st_size)
