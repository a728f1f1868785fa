"""W-GPU parity run (launch: torchrun --nproc-per-node W tools/ddp_parity.py, W in {2, 4, 8}) against the r3d18_w{W}
fixture recorded from the UNMODIFIED reference under W-rank gloo DDP (oracle/make_golden.py).

BIT-EXACT (north_star: "queue indices, shuffle permutations and gathered keys bit-exact"):
  * the batch each rank's key encoder receives after shuffle-BN, in the k_neg pass and in the k pass
      - against the rows the reference's encoder_k received (fixture, forward pre-hook), rounded to bf16 like the input
      - in full against the reference's own formula ``concat_all_gather(x)[idx_shuffle.view(W, -1)[rank]]``
        (builder:361-387) evaluated here with a plain NCCL all_gather
  * the un-shuffled keys each rank gets back and the all-rank key block, against ``concat_all_gather(k)[idx_unshuffle...]``
    (builder:389-406) evaluated with a plain NCCL all_gather
  * queue columns == the gathered k_neg head-A keys in rank order (builder:345-359); queue and queue_ptr identical on
    all ranks; labels
  * DDP-averaged gradients identical on all ranks
STATED bf16 TOLERANCE (conv path): logits / loss / queue columns / gradients against the fixture.
"""
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


def ref_all_gather(t):
    """concat_all_gather of the reference (builder:249-260): all_gather into a list, cat along dim 0."""
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t.contiguous())
    return torch.cat(parts, 0)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = load_golden(f"r3d18_w{world}")
    cfg, hyper = g["config"], g["hyper"]
    assert world == cfg["world"]
    B = cfg["batch"]
    model = build_product_moco(cfg, hyper, rank=rank).cuda()
    ddp = FlatDDP(model)  # broadcasts rank 0's parameters / buffers like DDP's constructor
    crit = Loss(hyper["margin"], hyper["A"], hyper["M"])
    rec = g["ranks"][rank]["steps"][0]
    rec0 = g["ranks"][0]["steps"][0]
    # every rank replays ITS OWN recorded draws: the two shuffle permutations differ per rank and rank 0's must win
    draws = [rec["perm"], rec["idx_shuffle_neg"], rec["idx_shuffle_pos"]]
    it = iter(draws)
    orig = torch.randperm
    torch.randperm = lambda n, *a, **k: (lambda r: r.to(k["device"]) if "device" in k else r.clone())(next(it))

    # ---- taps on the product's shuffle / unshuffle ----------------------------------------------------------
    taps = {"shuffle_in": [], "shuffle_out": [], "unshuffle_in": [], "unshuffle_out": []}
    orig_shuffle, orig_unshuffle = model._batch_shuffle_ddp, model._batch_unshuffle_ddp

    def tap_shuffle(x, *a, **k):
        out, idx_unshuffle = orig_shuffle(x, *a, **k)
        taps["shuffle_in"].append(x)
        taps["shuffle_out"].append((out, idx_unshuffle))
        return out, idx_unshuffle

    def tap_unshuffle(x, idx_unshuffle, return_all=False):
        res = orig_unshuffle(x, idx_unshuffle, True)
        taps["unshuffle_in"].append((x, idx_unshuffle))
        taps["unshuffle_out"].append(res)
        return res if return_all else res[0]

    model._batch_shuffle_ddp, model._batch_unshuffle_ddp = tap_shuffle, tap_unshuffle
    kneg_ptr = []
    orig_views = model._speed_views

    def tap_views(*a, **k):
        q, kk, kn = orig_views(*a, **k)
        kneg_ptr.append(kn.data_ptr())
        return q, kk, kn

    model._speed_views = tap_views
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
        print(f"[rank {rank}/{world}] {'PASS' if cond else 'FAIL'} {name} {info}", flush=True)

    # ---- bit-exact: shuffle-BN routing -----------------------------------------------------------------------
    idx_dev = [rec0["idx_shuffle_neg"].cuda(), rec0["idx_shuffle_pos"].cuda()]
    # forward() issues the k pull (on its own stream) before the k_neg pass: order the taps by pass
    order = sorted(range(2), key=lambda i: 0 if taps["shuffle_in"][i].data_ptr() == kneg_ptr[0] else 1)
    for p_i, pname in enumerate(("k_neg", "k")):
        x_local = taps["shuffle_in"][order[p_i]]                # bf16 NDHWC [B, T, H, W, 4]
        got, idx_unshuffle = taps["shuffle_out"][order[p_i]]
        want = ref_all_gather(x_local)[idx_dev[p_i].view(world, -1)[rank]]          # builder:383-387
        check(f"shuffled batch of the {pname} pass == concat_all_gather(x)[idx_shuffle.view(W,-1)[rank]] (bit-exact, "
              f"{got.numel() * 2 / 1e6:.1f} MB)", torch.equal(got, want))
        heads = got[:, 0, 0, :4, :3].permute(0, 2, 1).float().cpu()                # [B, 3 (C), 4 (W)] at t = h = 0
        check(f"shuffled batch of the {pname} pass == rows the reference's encoder_k received (fixture, bf16-rounded)",
              torch.equal(heads, rec["shuffled_heads"][p_i].bfloat16().float()))
        check(f"idx_unshuffle of the {pname} pass == argsort(idx_shuffle)",
              torch.equal(idx_unshuffle.cpu(), torch.argsort([rec0["idx_shuffle_neg"], rec0["idx_shuffle_pos"]][p_i])))
    # ---- bit-exact: key gather / unshuffle -------------------------------------------------------------------
    for p_i, pname in enumerate(("k_neg", "k")):
        feats, idx_unshuffle = taps["unshuffle_in"][p_i]        # [B, 256] = (head A | head M) in shuffled order
        mine, everyone = taps["unshuffle_out"][p_i]
        gathered = ref_all_gather(feats)
        iu = idx_unshuffle.to(feats.device)
        check(f"un-shuffled {pname} keys == concat_all_gather(k)[idx_unshuffle.view(W,-1)[rank]] (bit-exact)",
              torch.equal(mine, gathered[iu.view(world, -1)[rank]]))
        check(f"all-rank {pname} key block == concat_all_gather(k)[idx_unshuffle] (bit-exact)",
              torch.equal(everyone, gathered[iu]))
    keys_all = taps["unshuffle_out"][0][1][:, :hyper["dim"]]    # k_neg head A of every rank, global batch order
    first = (rec["queue_ptr"] - B * world) % cfg["K"]
    check("queue_ptr", int(model.queue_ptr) == rec["queue_ptr"], str(int(model.queue_ptr)))
    check("queue columns == gathered k_neg_A keys in rank order (bit-exact)",
          torch.equal(model.queue[:, first:first + B * world], keys_all.float().T))
    q = model.queue.clone()
    dist.broadcast(q, src=0)
    check("queue bit-identical across ranks", torch.equal(q, model.queue))
    check("labels", torch.equal(target.cpu(), rec["target"]) and torch.equal(rt.cpu(), rec["ranking_target"]))
    fg = ddp.flat_grad.clone()
    dist.broadcast(fg, src=0)
    check("DDP-averaged flat gradient bit-identical across ranks", torch.equal(fg, ddp.flat_grad))
    # ---- stated bf16 tolerance against the reference fixture ---------------------------------------------------
    d = (output[0].detach().cpu() - rec["logits1"]).abs().max().item()
    d2 = (output[1].detach().cpu() - rec["logits2"]).abs().max().item()
    check("logits1 / logits2 vs reference fixture (bf16 tol 0.22 = 2x observed)", max(d, d2) < 0.22,
          f"max diff {d:.4f} / {d2:.4f}")
    d = (torch.stack([loss, ce, rk]).detach().cpu() - rec["loss"]).abs().max().item()
    check("loss triple (tol 0.045)", d < 0.045, f"max diff {d:.4f}")
    d = (model.queue[:, first:first + B * world].cpu() - rec["queue_cols"]).abs().max().item()
    check("queue columns vs reference fixture (tol 0.025)", d < 0.025, f"max diff {d:.4f}")
    named = dict(model.named_parameters())
    worst, gots, refs = (1.0, None), [], []
    for k, ref in rec0["grads"].items():
        if isinstance(ref, dict) or ref.abs().max() < 1e-6 or k.endswith("conv1.bias"):
            continue
        got = named[k].grad.detach().float().cpu()
        c = float((got.flatten().double() @ ref.flatten().double()) / (got.norm().double() * ref.norm().double() + 1e-30))
        worst = min(worst, (c, k))
        gots.append(got.flatten())
        refs.append(ref.flatten())
    ga, ra = torch.cat(gots).double(), torch.cat(refs).double()
    overall = float(ga @ ra / (ga.norm() * ra.norm()))
    # Calibration (profiles/r02_bf16_noise_floor.txt): stock torch bf16 autocast of the UNMODIFIED reference keeps cosine
    # 0.911 (worst tensor 0.851) against its own fp32 gradients at this clip size — collapsed features at random init make
    # the useful gradient a small tangential component.  Gate = 2x the deviation observed for the product (0.916 / 0.882).
    check("DDP-averaged gradients vs reference fixture: all small tensors together cos >= 0.84, each >= 0.77",
          overall >= 0.84 and worst[0] >= 0.77, f"overall {overall:.4f}, worst {worst[0]:.4f} ({worst[1]})")
    for k in rec0["params_without_grad"]:
        if named[k].grad is not None:
            check(f"{k}.grad is None", False)
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if float(flag) != 1.0:
        sys.exit(1)
    if rank == 0:
        print(f"DDP PARITY OK (world {world})")


if __name__ == "__main__":
    main()
