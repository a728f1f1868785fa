"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's per-clip GPU transform chain (the non-aug_plus chain,
datasets/classification/__init__.py:188-202) from explicit random decisions.  Pinned by tests/test_oracle_cpu.py /
tests/test_sampler.py against tests/golden/sampler_clip.pt and sampler_jitter.pt, which were produced by the
unmodified reference transforms (oracle/make_golden_sampler.py).  Never imported by the product."""
import torch
import torch.nn.functional as F


def _gray(x):                       # functional_tensor.py:89-100
    return 0.2989 * x[0] + 0.5870 * x[1] + 0.1140 * x[2]


def _blend(a, b, ratio):            # functional_tensor.py:103-106
    return (ratio * a + (1 - ratio) * b).clamp(0, 1)


def _hue(x, shift):                 # functional_tensor.py:254-345, 376-415
    flat = x.reshape(3, -1)
    r, g, b = flat
    maxc, which = flat.max(0)
    minc = flat.min(0).values
    delta = maxc - minc
    sat = torch.where(maxc == 0, torch.zeros(()), delta / maxc)
    cand = torch.stack([(g - b) / delta, (b - r) / delta + 2.0, (r - g) / delta + 4.0])
    h = cand.gather(0, which[None])[0]
    h[delta == 0] = 0.0
    h = (h / 6.0) % 1.0
    h = (h + shift) % 1.0
    hi = torch.floor(h * 6)
    f = h * 6 - hi
    vtpq = torch.stack([maxc, maxc * (1 - (1 - f) * sat), maxc * (1 - sat), maxc * (1 - f * sat)])
    sector = hi.long() % 6
    cmap = torch.tensor([[0, 3, 2, 2, 1, 0], [1, 0, 0, 3, 2, 2], [2, 2, 1, 0, 0, 3]])
    return vtpq.gather(0, cmap[:, sector]).reshape(x.shape)


def color_jitter(x, factor, order):
    """x [3,T,H,W] in [0,1]; factor[4] = brightness, contrast, saturation, hue; order = op ids in application order."""
    for op in order:
        op = int(op)
        if op == 0:
            x = _blend(x, torch.zeros_like(x), float(factor[0]))
        elif op == 1:
            x = _blend(x, _gray(x).mean(), float(factor[1]))         # mean over the whole clip
        elif op == 2:
            x = _blend(x, _gray(x)[None].expand_as(x), float(factor[2]))
        elif op == 3:
            x = _hue(x, float(factor[3]))
    return x


def clip_chain(frames, idx, box, flag, mean, std, size, factor=None, order=None):
    """ToTensorVideo -> Resize(bilinear) -> gray? -> jitter? -> flip? -> Normalize for one clip; returns [3,T,S,S]."""
    i, j, h, w = [int(v) for v in box]
    clip = frames[idx.long()][:, i:i + h, j:j + w, :]
    x = clip.permute(3, 0, 1, 2).float() / 255.0
    x = F.interpolate(x, size=(size, size), mode="bilinear", align_corners=False)
    if int(flag) & 2:
        x = _gray(x)[None].expand_as(x).contiguous()
    if factor is not None:
        x = color_jitter(x, factor, order)
    if int(flag) & 1:
        x = x.flip(-1)
    m = torch.tensor(mean, dtype=torch.float32)[:, None, None, None]
    s = torch.tensor(std, dtype=torch.float32)[:, None, None, None]
    return (x - m) / s
