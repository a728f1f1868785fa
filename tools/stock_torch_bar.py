"""The honest GPU bar: the UNMODIFIED reference modules (baseline/_ref, loaded by
oracle/ref_loader.py) under stock torch / cuDNN on the same B200, same workload as bench.py's default line
(MoCoDiffLossTwoFc + R3D-18, batch 64, 2x16x112x112 clips, K=16384, SGD), in three numeric modes:

    fp32      cudnn.allow_tf32 = matmul.allow_tf32 = False          (the reference's published arithmetic)
    tf32      both True                                              (torch's default conv behaviour on Ampere+)
    bf16_cl3d torch.autocast(bfloat16) + channels_last_3d weights    (the best stock-library configuration)

usage (GPU box):  python tools/stock_torch_bar.py [arch] [batch] [steps] > gpurun_out/stock_bar.json
Prints one JSON line per mode.  Test / measurement infrastructure: nothing in rspnet_b200/ is imported.
"""
import json
import os
import statistics
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import ref_loader  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "resnet18"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
size = int(sys.argv[4]) if len(sys.argv) > 4 else 112
frames = int(sys.argv[5]) if len(sys.argv) > 5 else 32
HYPER = dict(dim=128, K=16384, m=0.999, T=0.07, margin=2.0, lr=0.1, momentum=0.9, weight_decay=1e-4)


def run(mode):
    torch.backends.cudnn.allow_tf32 = mode != "fp32"
    torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    model = ref_loader.build_reference_moco(arch, dim=HYPER["dim"], K=HYPER["K"], m=HYPER["m"], T=HYPER["T"]).cuda()
    if mode == "bf16_cl3d":
        model = model.to(memory_format=torch.channels_last_3d)
    ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[0], find_unused_parameters=True)
    crit = ref_loader.build_reference_loss(HYPER["margin"], 1.0, 1.0)
    opt = torch.optim.SGD(ddp.parameters(), lr=HYPER["lr"] * B / 64, momentum=HYPER["momentum"], dampening=0,
                          weight_decay=HYPER["weight_decay"], nesterov=False)
    gen = torch.Generator(device="cuda").manual_seed(1234)
    shape = (B, 3, frames, size, size)
    ring = [(torch.randn(shape, device="cuda", generator=gen), torch.randn(shape, device="cuda", generator=gen))
            for _ in range(2)]

    def step(i):
        q, k = ring[i % 2]
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16_cl3d")):
            output, target, rl, rt = ddp(q, k)
            loss, _, _ = crit(output, target, rl, rt)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    for i in range(4):
        step(i)
    torch.cuda.synchronize()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    marks[0].record()
    for i in range(steps):
        loss = step(i)
        marks[i + 1].record()
    torch.cuda.synchronize()
    per = [marks[i].elapsed_time(marks[i + 1]) for i in range(steps)]
    ms = statistics.median(per)
    print(json.dumps({"impl": "reference modules, stock torch %s / cuDNN %s on %s" % (
        torch.__version__, torch.backends.cudnn.version(), torch.cuda.get_device_name(0)), "mode": mode, "arch": arch,
        "batch": B, "clip": f"2x{frames // 2}x{size}x{size}", "K": HYPER["K"], "ms_per_step": ms,
        "clips_per_s": B / (ms / 1e3), "loss": float(loss), "steps": steps}), flush=True)
    del ddp, model, opt, ring
    torch.cuda.empty_cache()


def main():
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29577")
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    for mode in ("fp32", "tf32", "bf16_cl3d"):
        try:
            run(mode)
        except Exception as e:  # keep the other modes
            print(json.dumps({"mode": mode, "error": repr(e)[:300]}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
