"""2-GPU parity run (launch: torchrun --nproc-per-node 2 tools/ddp_parity.py) against the r3d18_w2 fixture recorded
from the unmodified reference under 2-rank gloo DDP: shuffle permutation exchange, key gather order / queue columns,
per-rank logits and loss, DDP-averaged gradients."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import build_product_moco, load_golden, make_inputs  # noqa: E402
from rspnet_b200.moco import FlatDDP, Loss  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = load_golden("r3d18_w2")
    cfg, hyper = g["config"], g["hyper"]
    assert world == cfg["world"]
    model = build_product_moco(cfg, hyper, rank=rank).cuda()
    ddp = FlatDDP(model)  # broadcasts rank 0's parameters / buffers like DDP's constructor
    crit = Loss(hyper["margin"], hyper["A"], hyper["M"])
    rec = g["ranks"][rank]["steps"][0]
    # every rank replays ITS OWN recorded draws: the two shuffle permutations differ per rank and rank 0's must win
    draws = [rec["perm"], rec["idx_shuffle_neg"], rec["idx_shuffle_pos"]]
    it = iter(draws)
    orig = torch.randperm
    torch.randperm = lambda n, *a, **k: (lambda r: r.to(k["device"]) if "device" in k else r.clone())(next(it))
    im_q, im_k = make_inputs(cfg, rank, 0)
    try:
        output, target, rl, rt = ddp(im_q.cuda(), im_k.cuda())
    finally:
        torch.randperm = orig
    loss, ce, rk = crit(output, target, rl, rt)
    loss.backward()
    torch.cuda.synchronize()
    ok = True

    def check(name, cond, info=""):
        nonlocal ok
        ok = ok and bool(cond)
        print(f"[rank {rank}] {'PASS' if cond else 'FAIL'} {name} {info}", flush=True)

    d = (output[0].detach().cpu() - rec["logits1"]).abs().max().item()
    check("logits1 vs reference fixture (bf16 tol 0.35)", d < 0.35, f"max diff {d:.4f}")
    d = (torch.stack([loss, ce, rk]).detach().cpu() - rec["loss"]).abs().max().item()
    check("loss triple (tol 0.15)", d < 0.15, f"max diff {d:.4f}")
    check("queue_ptr", int(model.queue_ptr) == rec["queue_ptr"], str(int(model.queue_ptr)))
    first = (rec["queue_ptr"] - cfg["batch"] * world) % cfg["K"]
    d = (model.queue[:, first:first + cfg["batch"] * world].cpu() - rec["queue_cols"]).abs().max().item()
    check("queue columns = gathered keys in rank order (tol 0.03)", d < 0.03, f"max diff {d:.4f}")
    q = model.queue.clone()
    dist.broadcast(q, src=0)
    check("queue bit-identical across ranks", torch.equal(q, model.queue))
    named = dict(model.named_parameters())
    worst = 1.0
    for k, ref in rec["grads"].items():
        if isinstance(ref, dict):
            continue
        got = named[k].grad
        if ref.abs().max() < 1e-6:
            continue
        c = float((got.cpu().flatten().double() @ ref.flatten().double()) /
                  (got.norm().double().cpu() * ref.norm().double() + 1e-30))
        worst = min(worst, c)
    check("DDP-averaged gradients vs fixture (cos >= 0.80; 2 clips per rank => BatchNorm over 2 samples)", worst > 0.80, f"worst cos {worst:.4f}")
    for k in rec["params_without_grad"]:
        if named[k].grad is not None:
            check(f"{k}.grad is None", False)
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if float(flag) != 1.0:
        sys.exit(1)
    if rank == 0:
        print("DDP PARITY OK")


if __name__ == "__main__":
    main()
