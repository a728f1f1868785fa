"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.pt by running the UNMODIFIED reference
(/root/reference, loaded by oracle/ref_loader.py) on CPU under gloo, exactly as pretrain.py drives it:
``MoCoDiffLossTwoFc`` (moco/__init__.py:19-46) wrapped in DistributedDataParallel(find_unused_parameters=True),
``Loss(margin=2.0, A, M)`` (pretrain.py:49-53), SGD(momentum .9, wd 1e-4) (pretrain.py:65-72).

Run in the build container only:  python oracle/make_golden.py
The fixtures are small (K is shrunk to 64/32 and clips to 32x32) so they can be committed; inputs and initial
weights are regenerated from seeds by the tests, and the fixture stores per-key checksums of the initial state
so that initialisation parity is checked too.
"""
import os
import random
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import ref_loader  # noqa: E402

GOLDEN_DIR = ROOT / "tests" / "golden"

CONFIGS = {
    # name: arch, world, per-rank batch, loaded frames T, H=W, K, steps
    "r3d18_w1": dict(arch="resnet18", world=1, batch=4, frames=8, size=64, K=64, steps=2, seed=0),
    # multi-rank fixtures are conditioned so that no BatchNorm normalises over a handful of values: 8 (4) clips per
    # rank, 4 output frames at 64x64 -> layer4 sees 8*1*2*2 = 32 (16) values per channel and rank
    "r3d18_w2": dict(arch="resnet18", world=2, batch=8, frames=8, size=64, K=64, steps=2, seed=0),
    "r3d18_w4": dict(arch="resnet18", world=4, batch=4, frames=8, size=64, K=64, steps=1, seed=0, slim=True),
    "r3d18_w8": dict(arch="resnet18", world=8, batch=4, frames=8, size=64, K=64, steps=1, seed=0, slim=True),
    "c3d_w1": dict(arch="c3d", world=1, batch=2, frames=16, size=64, K=32, steps=2, seed=0),
    "r2plus1d_w1": dict(arch="r2plus1d-vcop", world=1, batch=2, frames=16, size=64, K=32, steps=1, seed=0),
    "s3dg_w1": dict(arch="s3dg", world=1, batch=4, frames=16, size=128, K=32, steps=1, seed=0),
    # BASELINE.json configs at their real sizes (logits stored as slices + row statistics, see `big`)
    "r3d18_cfg1": dict(arch="resnet18", world=1, batch=4, frames=32, size=112, K=16384, steps=1, seed=0, big=True),
    "r3d18_b64": dict(arch="resnet18", world=1, batch=64, frames=32, size=112, K=16384, steps=1, seed=0, big=True),
    "c3d_b64": dict(arch="c3d", world=1, batch=64, frames=32, size=112, K=16384, steps=1, seed=0, big=True),
}
HYPER = dict(dim=128, m=0.999, T=0.07, diff_speed=[2], margin=2.0, A=1.0, M=1.0, lr=0.1, momentum=0.9,
             weight_decay=1e-4)

BIG = 4096  # tensors above this many elements are stored as checksums + a head slice


def summarize(t: torch.Tensor):
    t = t.detach().double().flatten()
    return dict(sum=float(t.sum()), abssum=float(t.abs().sum()), n=t.numel(), head=t[:32].float().clone())


def pack(t: torch.Tensor):
    return t.detach().clone() if t.numel() <= BIG else summarize(t)


def pack_logits(t, big):
    """Full [N, 1+K] logits for the small fixtures; for the BASELINE-sized ones the positive column, the first 255
    negatives, the row logsumexp / max and the checksums."""
    t = t.detach()
    if not big:
        return t.clone()
    return dict(head=t[:, :256].clone(), lse=torch.logsumexp(t.double(), 1).float(), rowmax=t.max(1).values.clone(),
                sum=float(t.double().sum()), abssum=float(t.double().abs().sum()), shape=tuple(t.shape))


def make_inputs(cfg, rank, step):
    """Synthetic clips as SURVEY.md §8d: randn seeded by 1234 + rank (a fresh draw per step)."""
    g = torch.Generator().manual_seed(1234 + rank + 1000 * step)
    shape = (cfg["batch"], 3, cfg["frames"], cfg["size"], cfg["size"])
    return torch.randn(shape, generator=g), torch.randn(shape, generator=g)


def initialize_seed(seed):
    """framework/utils/reproduction.py:29-33 (python, numpy, torch)."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def worker(rank, name, cfg, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(max(1, (os.cpu_count() or 4) // cfg["world"]) if cfg["world"] > 1 else (os.cpu_count() or 4))
    dist.init_process_group("gloo", rank=rank, world_size=cfg["world"])
    initialize_seed(cfg["seed"] + rank)  # pretrain.py:266-267
    model = ref_loader.build_reference_moco(cfg["arch"], dim=HYPER["dim"], K=cfg["K"], m=HYPER["m"], T=HYPER["T"],
                                            diff_speed=HYPER["diff_speed"])
    ddp = torch.nn.parallel.DistributedDataParallel(model, find_unused_parameters=True)
    criterion = ref_loader.build_reference_loss(HYPER["margin"], HYPER["A"], HYPER["M"])
    opt = torch.optim.SGD(ddp.parameters(), lr=HYPER["lr"], momentum=HYPER["momentum"], dampening=0,
                          weight_decay=HYPER["weight_decay"], nesterov=False)
    slim = cfg.get("slim", False) and rank > 0      # gradients / parameters are identical on all ranks after DDP
    big = cfg.get("big", False)
    rec = dict(rank=rank, init={} if (slim or big) else {k: summarize(v.float()) for k, v in model.state_dict().items()},
               steps=[])
    # what encoder_k really receives (the shuffled batch, builder:383-387) and returns before the unshuffle
    seen = {"in": [], "out": []}
    model.encoder_k.register_forward_pre_hook(lambda m, inp: seen["in"].append(inp[0][:, :, 0, 0, :4].detach().clone()))
    model.encoder_k.register_forward_hook(lambda m, inp, out: seen["out"].append(out[0].detach().clone()))

    draws = []
    orig_randperm = torch.randperm

    def recording_randperm(*a, **k):
        r = orig_randperm(*a, **k)
        draws.append(r.clone())
        return r

    torch.randperm = recording_randperm
    for step in range(cfg["steps"]):
        draws.clear()
        seen["in"].clear()
        seen["out"].clear()
        im_q, im_k = make_inputs(cfg, rank, step)
        output, target, ranking_logits, ranking_target = ddp(im_q, im_k)
        loss, loss_a, loss_m = criterion(output, target, ranking_logits, ranking_target)
        opt.zero_grad()
        loss.backward()
        grads = {} if slim else {k: pack(p.grad) for k, p in model.named_parameters() if p.grad is not None}
        no_grad = [k for k, p in model.named_parameters() if p.requires_grad and p.grad is None]
        opt.step()
        sd = model.state_dict()
        ptr = int(sd["queue_ptr"])
        n_all = cfg["batch"] * cfg["world"]
        first = (ptr - n_all) % cfg["K"]
        rec["steps"].append(dict(
            perm=draws[0].clone(), idx_shuffle_neg=draws[1].clone(), idx_shuffle_pos=draws[2].clone(),
            n_randperm=len(draws),
            logits1=pack_logits(output[0], big), logits2=pack_logits(output[1], big),
            shuffled_heads=[t.clone() for t in seen["in"]],      # [k_neg pass, k pass]: [B, 3, 4] input corners
            key_a_shuffled=[t.clone() for t in seen["out"]],     # encoder_k head-A outputs in shuffled order
            l_pos_m=ranking_logits[0].detach().clone(), l_neg_m=ranking_logits[1].detach().clone(),
            target=target.clone(), ranking_target=ranking_target.clone(),
            loss=torch.stack([loss.detach(), loss_a.detach(), loss_m.detach()]),
            queue_ptr=ptr, queue_cols=sd["queue"][:, first:first + n_all].clone(),
            grads=grads, params_without_grad=no_grad,
            params_after={} if (slim or big) else {k: summarize(v.float()) for k, v in sd.items() if k != "queue"},
        ))
    torch.randperm = orig_randperm
    torch.save(rec, out_dir / f"{name}.rank{rank}.pt")
    dist.barrier()
    dist.destroy_process_group()


def main():
    assert ref_loader.available(), "/root/reference is required to (re)generate goldens"
    GOLDEN_DIR.mkdir(parents=True, exist_ok=True)
    tmp = Path("/tmp/rsp_golden")
    tmp.mkdir(exist_ok=True)
    only = sys.argv[1:]
    for i, (name, cfg) in enumerate(CONFIGS.items()):
        if only and name not in only:
            continue
        port = 29650 + i
        if cfg["world"] == 1:
            worker(0, name, cfg, port, tmp)
        else:
            mp.spawn(worker, args=(name, cfg, port, tmp), nprocs=cfg["world"], join=True)
        ranks = [torch.load(tmp / f"{name}.rank{r}.pt", weights_only=False) for r in range(cfg["world"])]
        torch.save(dict(name=name, config=cfg, hyper=HYPER, torch_version=torch.__version__, ranks=ranks),
                   GOLDEN_DIR / f"{name}.pt")
        size = (GOLDEN_DIR / f"{name}.pt").stat().st_size
        print(f"{name}: losses rank0 = {[s['loss'].tolist() for s in ranks[0]['steps']]}  ({size / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
