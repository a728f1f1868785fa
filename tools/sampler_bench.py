"""Times the clip sampler alone on the bench's e2e shapes (diagnostic, GPU box only)."""
import random
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rspnet_b200.sampler import GPUClipSampler  # noqa: E402

B, FV, HS, WS = 64, 64, 128, 171
frames = torch.randint(0, 256, (B * FV, HS, WS, 3), dtype=torch.uint8, device="cuda")
offsets, lengths = [v * FV for v in range(B)], [FV] * B
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, jit in (("no jitter", None), ("jitter", dict(brightness=0.4, contrast=0.4, saturation=0.4, hue=0.4))):
    for layout in (0, 1):
        s = GPUClipSampler(size=112, temporal_size=32, color_jitter=jit)
        random.seed(0)
        for _ in range(3):
            s(frames, offsets, lengths, layout=layout)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            s(frames, offsets, lengths, layout=layout)
        e1.record()
        torch.cuda.synchronize()
        print(f"{name:10s} layout {layout}: {e0.elapsed_time(e1) / 10:.3f} ms per batch of {2 * B} clips x 32 x 112 x 112")
