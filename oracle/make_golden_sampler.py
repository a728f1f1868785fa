"""TEST INFRASTRUCTURE ONLY — golden vectors for the clip pipeline (subsystem 4) from the UNMODIFIED reference
transforms (datasets/transforms_video/*): random decisions under a fixed python-random seed, and the per-clip GPU
transform chain (ToTensorVideo -> Resize -> gray -> [ColorJitter] -> flip -> Normalize) evaluated on CPU.
Run in the build container only:  python oracle/make_golden_sampler.py"""
import json
import random
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, "/root/reference")
from datasets.transforms_video import transforms_spatial, transforms_temporal, transforms_tensor  # noqa: E402
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


def jitter_pipeline():
    """The non-aug_plus chain of datasets/classification/__init__.py:188-202 with the reference's own ColorJitter:
    the python-random state is seeded, the chain runs, and the decisions it took are recovered by replaying the same
    seed through the documented draw order (gray, 4 uniforms, shuffle, flip) — the CPU test checks the product's
    host draws against these records, the GPU test the kernel against ``out``."""
    g = torch.Generator().manual_seed(11)
    n, t, size = 12, 4, 16
    frames = torch.randint(0, 256, (10, 40, 52, 3), generator=g, dtype=torch.uint8)
    # smooth content too, so that hue sectors / saturation are not only noise
    yy, xx = torch.meshgrid(torch.arange(40), torch.arange(52), indexing="ij")
    frames[5:, :, :, 0] = (xx * 4).clamp(0, 255).to(torch.uint8)
    frames[5:, :, :, 1] = (yy * 6).clamp(0, 255).to(torch.uint8)
    frames[5:, :, :, 2] = ((xx + yy) * 2).clamp(0, 255).to(torch.uint8)
    frames[9] = 0                                                      # black frame: v == 0 branch of rgb_to_hsv
    idx = torch.randint(0, 10, (n, t), generator=g, dtype=torch.int32)
    box = torch.tensor([[3, 5, 30, 40], [0, 0, 40, 52], [10, 20, 17, 13]] * 4, dtype=torch.int32)
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    chain = transforms_tensor.Compose([
        transforms_spatial.ToTensor(), transforms_spatial.Resize(size), transforms_spatial.RandomGrayScale(p=0.2),
        transforms_spatial.ColorJitter(brightness=0.4, contrast=0.4, saturation=0.4, hue=0.4),
        transforms_spatial.RandomHorizontalFlip(), transforms_spatial.Normalize(mean, std, inplace=True)])
    random.seed(4321)
    outs = []
    for c in range(n):
        clip = frames[idx[c].long()]
        i, j, h, w = box[c].tolist()
        outs.append(chain(clip[..., i:i + h, j:j + w, :].contiguous()))
    end_state = random.random()
    # replay of the decisions
    random.seed(4321)
    flags, factors, orders = [], [], []
    for c in range(n):
        gray = random.random() < 0.2
        f = [random.uniform(0.6, 1.4), random.uniform(0.6, 1.4), random.uniform(0.6, 1.4), random.uniform(-0.4, 0.4)]
        ops = [0, 1, 2, 3]
        random.shuffle(ops)
        flip = random.random() < 0.5
        flags.append((1 if flip else 0) | (2 if gray else 0))
        factors.append(f)
        orders.append(ops)
    assert random.random() == end_state, "replay consumed a different number of draws than the reference chain"
    return dict(frames=frames, idx=idx, box=box, flags=torch.tensor(flags, dtype=torch.uint8),
                factors=torch.tensor(factors, dtype=torch.float64), orders=torch.tensor(orders, dtype=torch.uint8),
                mean=mean, std=std, size=size, seed=4321, end_state=end_state, out=torch.stack(outs))


if __name__ == "__main__":
    torch.save(jitter_pipeline(), OUT / "sampler_jitter.pt")
    (OUT / "sampler_draws.json").write_text(json.dumps(draws()))
    torch.save(pipeline(), OUT / "sampler_clip.pt")
    print("written", (OUT / "sampler_draws.json").stat().st_size, (OUT / "sampler_clip.pt").stat().st_size)
