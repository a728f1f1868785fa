"""Per-stream device timeline of the pretraining step (nsys is not in the image: torch.profiler / CUPTI instead).

    python tools/step_timeline.py                                   # 1 GPU
    torchrun --nproc-per-node N tools/step_timeline.py              # N GPUs, rank 0 writes the file
    -> gpurun_out/r02_timeline_n{N}.json : {"kernels": [[stream, start_us, dur_us, name], ...], "step_marks": [...]}
    -> prints a summary: per stream busy time, NCCL kernels, gaps on the main stream, what overlaps what

Same model / batch as bench.py's default line (R3D-18, batch 64, 2x16x112x112, K=16384)."""
import json
import os
import sys
from collections import defaultdict
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from rspnet_b200.engine import PretrainEngine, scale_learning_rate  # noqa: E402
from rspnet_b200.models import get_model_class  # noqa: E402
from rspnet_b200.moco import Loss, MoCoDiffLossTwoFc, MultiTaskWrapper  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    arch = sys.argv[1] if len(sys.argv) > 1 else "resnet18"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.manual_seed(rank)
    base = get_model_class(arch=arch)
    model = MoCoDiffLossTwoFc(lambda num_classes=128: MultiTaskWrapper(base, num_classes=num_classes), dim=128, K=16384,
                              m=0.999, T=0.07, diff_speed=[2]).to(dev)
    engine = PretrainEngine(model, Loss(2.0, 1.0, 1.0), scale_learning_rate(0.1, world, B))
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    shape = (B, 3, 32, 112, 112)
    ring = [(torch.randn(shape, device=dev, generator=gen), torch.randn(shape, device=dev, generator=gen))
            for _ in range(2)]
    for i in range(12):
        engine.step(*ring[i % 2])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    steps = 3
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA,
                                            torch.profiler.ProfilerActivity.CPU]) as prof:
        for i in range(steps):
            with torch.profiler.record_function(f"STEP{i}"):
                engine.step(*ring[i % 2])
        torch.cuda.synchronize()
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    trace = ROOT / "gpurun_out" / f"r02_chrome_trace_{arch}_n{world}.json"
    trace.parent.mkdir(exist_ok=True)
    prof.export_chrome_trace(str(trace))
    raw = json.loads(trace.read_text())
    trace.unlink()       # tens of MB: keep the compact list below instead
    kernels = sorted([[int(e.get("args", {}).get("stream", -1)), float(e["ts"]), float(e.get("dur", 0.0)), e["name"][:90]]
                      for e in raw["traceEvents"]
                      if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")],
                     key=lambda r: r[1])
    t0 = kernels[0][1]
    for k in kernels:
        k[1] -= t0
    out = ROOT / "gpurun_out" / f"r02_timeline_{arch}_n{world}.json"
    out.parent.mkdir(exist_ok=True)
    out.write_text(json.dumps({"arch": arch, "batch": B, "world": world, "steps": steps, "kernels": kernels}))
    # ---- summary -------------------------------------------------------------------------------------------------
    span = kernels[-1][1] + kernels[-1][2] - kernels[0][1]
    per_stream = defaultdict(float)
    nccl = [k for k in kernels if "nccl" in k[3].lower()]
    for s, st, du, nm in kernels:
        per_stream[s] += du
    print(f"[timeline] {arch} B={B} world={world}: {len(kernels)} device activities over {span / 1e3:.3f} ms "
          f"({span / 1e3 / steps:.3f} ms per step)")
    for s, du in sorted(per_stream.items(), key=lambda kv: -kv[1]):
        print(f"  stream {s}: busy {du / 1e3 / steps:.3f} ms per step")
    print(f"  NCCL kernels per step: {len(nccl) / steps:.1f}, {sum(k[2] for k in nccl) / 1e3 / steps:.3f} ms per step")
    agg = defaultdict(lambda: [0, 0.0])
    for s, st, du, nm in kernels:
        key = nm.split("(")[0][:60]
        agg[key][0] += 1
        agg[key][1] += du
    for nm, (cnt, du) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
        print(f"  {du / 1e3 / steps:8.3f} ms/step  x{cnt / steps:6.1f}  {nm}")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
