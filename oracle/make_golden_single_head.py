"""TEST INFRASTRUCTURE ONLY — golden vectors for the single-head builder ``MoCoDiffLoss`` (reference
moco/builder_diffspeed_diffloss.py:11-245), produced by running the UNMODIFIED reference on CPU under gloo (world 1):
R3D-18 with ``num_classes=dim`` as the encoder, one forward / loss / backward on 2 videos of 2x16x112x112 frames
(the reference's AvgPool3d((1,4,4)) fixes the clip size), K = 32.

Run in the build container only:  python oracle/make_golden_single_head.py  ->  tests/golden/r3d18_single_head.pt
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import ref_loader  # noqa: E402
from oracle.make_golden import HYPER, initialize_seed, make_inputs, pack, summarize  # noqa: E402

CFG = dict(arch="resnet18", world=1, batch=2, frames=32, size=112, K=32, steps=1, seed=0)


def main():
    assert ref_loader.available(), "/root/reference is required to (re)generate goldens"
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = "29671"
    torch.set_num_threads(8)
    dist.init_process_group("gloo", rank=0, world_size=1)
    ref_loader.install_shims()
    initialize_seed(CFG["seed"])
    builder = ref_loader.modules()["builder"]
    model = builder.MoCoDiffLoss(ref_loader.backbone_ctor("resnet18"), dim=HYPER["dim"], K=CFG["K"], m=HYPER["m"],
                                 T=HYPER["T"], diff_speed=list(HYPER["diff_speed"]))
    criterion = ref_loader.build_reference_loss(HYPER["margin"], HYPER["A"], HYPER["M"])
    init = {k: summarize(v.float()) for k, v in model.state_dict().items()}
    draws = []
    orig = torch.randperm

    def rec(*a, **k):
        r = orig(*a, **k)
        draws.append(r.clone())
        return r

    torch.randperm = rec
    im_q, im_k = make_inputs(CFG, 0, 0)
    output, target, ranking_logits, ranking_target = model(im_q, im_k)
    loss, loss_a, loss_m = criterion(output, target, ranking_logits, ranking_target)
    loss.backward()
    torch.randperm = orig
    sd = model.state_dict()
    ptr = int(sd["queue_ptr"])
    first = (ptr - CFG["batch"]) % CFG["K"]
    step = dict(perm=draws[0], idx_shuffle_neg=draws[1], idx_shuffle_pos=draws[2], n_randperm=len(draws),
                logits1=output[0].detach().clone(), logits2=output[1].detach().clone(),
                l_pos=ranking_logits[0].detach().clone(), l_neg_speed=ranking_logits[1].detach().clone(),
                target=target.clone(), ranking_target=ranking_target.clone(),
                loss=torch.stack([loss.detach(), loss_a.detach(), loss_m.detach()]), queue_ptr=ptr,
                queue_cols=sd["queue"][:, first:first + CFG["batch"]].clone(),
                grads={k: pack(p.grad) for k, p in model.named_parameters() if p.grad is not None})
    out = ROOT / "tests" / "golden" / "r3d18_single_head.pt"
    torch.save(dict(name="r3d18_single_head", config=CFG, hyper=HYPER, torch_version=torch.__version__, init=init,
                    step=step), out)
    print("loss", step["loss"].tolist(), "queue_ptr", ptr, f"({out.stat().st_size / 1024:.0f} KiB)")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
