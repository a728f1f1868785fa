"""Where does an end-to-end step from uint8 host frames spend its time?  (diagnostic, GPU box only)"""
import random
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from rspnet_b200.engine import PretrainEngine  # noqa: E402
from rspnet_b200.models import get_model_class  # noqa: E402
from rspnet_b200.moco import Loss, MoCoDiffLossTwoFc, MultiTaskWrapper  # noqa: E402
from rspnet_b200.sampler import GPUClipSampler  # noqa: E402

H = bench.HYPER
dev = torch.device("cuda", 0)
torch.manual_seed(0)
base = get_model_class(arch="resnet18")
model = MoCoDiffLossTwoFc(lambda num_classes=128: MultiTaskWrapper(base, num_classes=num_classes, fc_type="linear"),
                          dim=H["dim"], K=H["K"], m=H["m"], T=H["T"], diff_speed=H["diff_speed"]).to(dev)
engine = PretrainEngine(model, Loss(H["margin"], H["A"], H["M"]), H["lr"], H["momentum"], H["weight_decay"])
B, FV, HS, WS = 64, 64, 128, 171
random.seed(0)
sampler = GPUClipSampler(size=112, temporal_size=32, color_jitter=dict(brightness=.4, contrast=.4, saturation=.4, hue=.4))
frames = torch.randint(0, 256, (B * FV, HS, WS, 3), dtype=torch.uint8, device=dev)
offsets, lengths = [v * FV for v in range(B)], [FV] * B
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]


def run(n, with_sampler, sync_each):
    th_s = th_e = 0.0
    g_s = g_e = 0.0
    torch.cuda.synchronize()
    t_all = time.perf_counter()
    q = k = None
    if not with_sampler:
        (q, k), _ = sampler(frames, offsets, lengths)
    for _ in range(n):
        t0 = time.perf_counter()
        ev[0].record()
        if with_sampler:
            (q, k), _ = sampler(frames, offsets, lengths)
        ev[1].record()
        t1 = time.perf_counter()
        engine.step(q, k)
        ev[2].record()
        t2 = time.perf_counter()
        th_s += t1 - t0
        th_e += t2 - t1
        if sync_each:
            torch.cuda.synchronize()
            g_s += ev[0].elapsed_time(ev[1])
            g_e += ev[1].elapsed_time(ev[2])
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t_all) * 1e3 / n
    print(f"sampler={with_sampler} sync_each={sync_each}: wall {wall:.2f} ms/step | host sampler {th_s / n * 1e3:.2f} "
          f"step {th_e / n * 1e3:.2f} | gpu sampler {g_s / n:.2f} step {g_e / n:.2f}")


for _ in range(2):
    run(5, True, False)
run(10, False, False)
run(10, True, False)
run(10, True, True)
run(10, False, True)

# ---- the same with the frames crossing PCIe each step (double-buffered on a copy stream), as bench.py does
host = [torch.randint(0, 256, (B * FV, HS, WS, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
stage = [torch.empty_like(frames) for _ in range(2)]
copy_stream = torch.cuda.Stream()
ready = [torch.cuda.Event() for _ in range(2)]
consumed = [torch.cuda.Event() for _ in range(2)]
c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(copy_stream):
    c0.record()
    stage[0].copy_(host[0], non_blocking=True)
    c1.record()
torch.cuda.synchronize()
print(f"one H2D copy of {host[0].numel() / 1e6:.0f} MB alone: {c0.elapsed_time(c1):.2f} ms")


def feed(n, consume_early):
    for s in range(2):
        consumed[s].record()

    def prefetch(i):
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[s])
            stage[s].copy_(host[s], non_blocking=True)
            ready[s].record(copy_stream)

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    prefetch(0)
    for i in range(n):
        s = i % 2
        prefetch(i + 1)
        torch.cuda.current_stream().wait_event(ready[s])
        (q, k), _ = sampler(stage[s], offsets, lengths)
        if consume_early:
            consumed[s].record()
        engine.step(q, k)
        if not consume_early:
            consumed[s].record()
    torch.cuda.synchronize()
    print(f"H2D feed, consumed recorded {'after the sampler' if consume_early else 'after the step'}: "
          f"{(time.perf_counter() - t0) * 1e3 / n:.2f} ms/step")


feed(5, False)
feed(20, False)
feed(20, True)
