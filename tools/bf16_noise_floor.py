"""Calibration of the stated bf16 tolerance: how far does the UNMODIFIED reference drift from its own
fp32 result when stock torch runs it in bf16 (torch.autocast + cuDNN bf16 kernels, BatchNorm kept in fp32 by autocast)?
Same weights, same inputs, same random draws; compares logits, loss and the gradient direction of one step at BASELINE
config 1 (R3D-18, batch 4) and at the bench shape (batch 64).  The product's own deviations from the fp32 fixture
(tests/test_parity_gpu.py::test_step_matches_reference_at_baseline_sizes) are gated against these figures.

usage (GPU box):  python tools/bf16_noise_floor.py > gpurun_out/r02_bf16_noise_floor.txt
Test / measurement infrastructure: nothing in rspnet_b200/ is imported.
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import initialize_seed, load_golden, make_inputs  # noqa: E402
from oracle import ref_loader  # noqa: E402


def cos(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return float(a @ b / (a.norm() * b.norm() + 1e-30))


def one_step(cfg, hyper, rec, mode):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    initialize_seed(cfg["seed"])
    model = ref_loader.build_reference_moco(cfg["arch"], dim=hyper["dim"], K=cfg["K"], m=hyper["m"], T=hyper["T"],
                                            diff_speed=hyper["diff_speed"]).cuda()
    crit = ref_loader.build_reference_loss(hyper["margin"], hyper["A"], hyper["M"])
    im_q, im_k = make_inputs(cfg, 0, 0)
    draws = iter([rec["perm"], rec["idx_shuffle_neg"], rec["idx_shuffle_pos"]])
    orig = torch.randperm
    torch.randperm = lambda n, *a, **k: (lambda r: r.to(k["device"]) if "device" in k else r.clone())(next(draws))
    try:
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
            output, target, rl, rt = model(im_q.cuda(), im_k.cuda())
            loss, ce, rk = crit(output, target, rl, rt)
    finally:
        torch.randperm = orig
    loss.backward()
    grads = {k: p.grad.detach().float().cpu() for k, p in model.named_parameters() if p.grad is not None}
    return output[0].detach().float().cpu(), torch.stack([loss, ce, rk]).detach().float().cpu(), grads


def main():
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29579")
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    for name in sys.argv[1:] or ["r3d18_cfg1", "r3d18_b64", "c3d_b64", "r3d18_w1"]:
        g = load_golden(name)
        cfg, hyper, rec = g["config"], g["hyper"], g["ranks"][0]["steps"][0]
        l32, loss32, g32 = one_step(cfg, hyper, rec, "fp32")
        l16, loss16, g16 = one_step(cfg, hyper, rec, "bf16")
        small_a, small_b, worst = [], [], (1.0, None)
        for k in g32:
            if g32[k].numel() > 4096 or g32[k].abs().max() < 1e-7 or k.endswith("conv1.bias"):
                continue
            c = cos(g16[k], g32[k])
            worst = min(worst, (c, k))
            small_a.append(g16[k].flatten())
            small_b.append(g32[k].flatten())
        big_a = torch.cat([g16[k].flatten() for k in g32 if g32[k].numel() > 4096])
        big_b = torch.cat([g32[k].flatten() for k in g32 if g32[k].numel() > 4096])
        print(f"[{name}] stock torch bf16 autocast vs its own fp32 (same weights / inputs / draws): "
              f"|dlogits| {(l16 - l32).abs().max():.4f} |dloss| {(loss16 - loss32).abs().max():.4f}; gradient cosine "
              f"small tensors {cos(torch.cat(small_a), torch.cat(small_b)):.4f} (worst {worst[0]:.4f} {worst[1]}), "
              f"large tensors {cos(big_a, big_b):.4f}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
