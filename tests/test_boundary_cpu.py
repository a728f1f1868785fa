"""Host-side logic of the drop-in boundary, runnable without a GPU: config loading, registry, state_dict surface,
the permutation-exchange plan (incl. a 2-rank gloo run), flat parameter storage."""
import os
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT

CONFIG_DIR = Path("/root/reference/config")


def test_jsonnet_subset_semantics():
    """Language features the reference configs rely on, on hand-written fixtures (tests/golden/jsonnet)."""
    from rspnet_b200.config import get_config
    d = ROOT / "tests" / "golden" / "jsonnet"
    base = get_config(d / "base.jsonnet")
    assert base.get_string("child.name") == "base" and base.get_string("child.twice") == "basebase"
    assert base.get_float("opt.rate") == 0.5 and base.get_list("opt.flags") == [True, False, None]
    assert base.get_list("mean") == [1, 2, 3] and base.get_int("count") == 12
    assert base.get_int("window.size") == 32 and "_unit" not in base.get_config("window").keys()
    der = get_config(d / "derived.jsonnet", ["add.no_decay", "{speeds: []}"])
    assert der.get_string("child.name") == "derived"            # `$` is late-bound through inheritance
    assert der.get_float("opt.rate") == 0.25 and der.get_float("opt.decay") == 0   # `+:` merges, -x patches apply
    assert der.get_int("window.size") == 3 and "_unit" not in der.get_config("window").keys()  # stays hidden


@pytest.mark.skipif(not CONFIG_DIR.exists(), reason="/root/reference not present")
def test_jsonnet_pretrain_configs_resolve_like_the_reference():
    """The reference's own config/pretrain/*.jsonnet, loaded unchanged: resolved values of SURVEY.md appendix D."""
    from rspnet_b200.config import get_config, trim_moco_k
    c = get_config(CONFIG_DIR / "pretrain" / "resnet18.jsonnet")
    assert c.get_string("arch") == "resnet18" and c.get_config("model").get_string("arch") == "resnet18"
    assert c.get_int("batch_size") == 64 and c.get_int("num_workers") == 8 and c.get_int("num_epochs") == 200
    assert c.get_float("optimizer.lr") == 0.1 and c.get_float("optimizer.momentum") == 0.9
    assert c.get_float("optimizer.weight_decay") == 1e-4 and c.get_bool("optimizer.nesterov") is False
    assert c.get_int("moco.dim") == 128 and c.get_int("moco.k") == 16384 and c.get_float("moco.m") == 0.999
    assert c.get_float("moco.t") == 0.07 and c.get_list("moco.diff_speed") == [2]
    assert c.get_string("moco.fc_type") == "linear"
    assert c.get_int("temporal_transforms.size") == 32 and c.get_int("spatial_transforms.size") == 112
    assert "_size" not in c.get_config("temporal_transforms").keys()   # hidden field
    assert c.get("dataset.mean") == [0.485, 0.456, 0.406] and c.get_float("loss_lambda.A") == 1.0
    c3d = get_config(CONFIG_DIR / "pretrain" / "c3d.jsonnet")
    assert c3d.get_config("model").get_string("arch") == "c3d" and c3d.get_int("batch_size") == 32  # late-bound $.arch
    s3d = get_config(CONFIG_DIR / "pretrain" / "s3dg.jsonnet")
    assert s3d.get_float("optimizer.lr") == 0.05 and s3d.get_int("spatial_transforms.size") == 224
    r21 = get_config(CONFIG_DIR / "pretrain" / "r2plus1d.jsonnet")
    assert r21.get_string("model.arch") == "r2plus1d-vcop" and r21.get_int("temporal_transforms.size") == 32
    ext = get_config(CONFIG_DIR / "pretrain" / "resnet18.jsonnet", ["add.M0", "{batch_size: 4, moco+: {k: 70}}"])
    assert ext.get_float("loss_lambda.M") == 0 and ext.get_int("batch_size") == 4 and ext.get_float("moco.t") == 0.07
    assert trim_moco_k(ext.get_int("moco.k"), 4, 2) == 64
    ext.put("moco.k", 64)
    assert ext.get_int("moco.k") == 64


def test_registry_and_state_dict_surface():
    from rspnet_b200.models import get_model_class
    from rspnet_b200.moco import MoCoDiffLossTwoFc, MultiTaskWrapper
    with pytest.raises(ValueError):
        get_model_class(arch="nope")
    for arch, n_params in (("resnet18", 33_335_745), ("c3d", 27_793_281), ("r2plus1d-vcop", 14_497_144)):
        base = get_model_class(arch=arch)
        enc = MultiTaskWrapper(base, num_classes=128)
        assert sum(p.numel() for p in enc.parameters()) == n_params, arch
        assert hasattr(enc.encoder, "get_feature") and hasattr(enc.encoder, "feature_ndhwc")
    moco = MoCoDiffLossTwoFc(lambda num_classes=128: MultiTaskWrapper(get_model_class(arch="resnet18"), num_classes),
                             K=64, diff_speed=[2])
    sd = moco.state_dict()
    assert len(sd) == 254 and "queue" in sd and "queue_ptr" in sd
    assert sd["queue"].shape == (128, 64) and sd["queue_ptr"].dtype == torch.long
    assert "encoder_q.encoder.layer4.1.bn2.running_var" in sd and "encoder_k.fc2.2.bias" in sd
    assert all(not p.requires_grad for p in moco.encoder_k.parameters())
    torch.testing.assert_close(sd["queue"].norm(dim=0), torch.ones(64))


def test_head_variants_state_dict_matches_reference_fixture():
    """fc_type conv / convbn / finetune wrappers: same state_dict keys, order and initial values under the same seed as
    the reference (fixture recorded by oracle/make_golden_heads.py from moco/split_wrapper.py)."""
    from helpers import check_packed, initialize_seed, load_golden
    from rspnet_b200.models import get_model_class
    from rspnet_b200.moco import MultiTaskWrapper
    g = load_golden("r3d18_heads")
    for rec in g["cases"]:
        initialize_seed(g["seed"])
        model = MultiTaskWrapper(get_model_class(arch="resnet18"), num_classes=128, **rec["case"])
        sd = model.state_dict()
        assert list(sd.keys()) == rec["keys"], rec["case"]
        for k, v in sd.items():
            check_packed(v.float(), rec["init"][k], rtol=0, atol=0, what=f"{rec['case']} {k}")


def test_checkpoint_hand_off_to_reference_loaders(tmp_path):
    """PretrainEngine.checkpoint_state writes the dictionary of pretrain.py:249-259; the loaders of finetune.py:273-303
    and retrieval.py:84-101 (their key filters restated here, and the reference's own modules when /root/reference is
    present) accept it, torch.optim.SGD / CosineAnnealingLR load its optimizer / scheduler entries, and
    load_checkpoint round-trips it."""
    import math
    from helpers import build_product_moco
    from rspnet_b200.engine import PretrainEngine
    from rspnet_b200.models import get_model_class
    from rspnet_b200.moco import Loss, MultiTaskWrapper
    cfg = dict(arch="resnet18", seed=0, K=64)
    hyper = dict(dim=128, m=0.999, T=0.07, diff_speed=[2])
    model = build_product_moco(cfg, hyper)
    eng = PretrainEngine(model, Loss(2.0, 1.0, 1.0), lr=0.1, momentum=0.9, weight_decay=1e-4, num_epochs=200)
    # state as after some steps: momentum everywhere except the unused backbone fc (it never receives a gradient)
    names = [n for n, _ in model.encoder_q.named_parameters()]
    used = [i for i, n in enumerate(names) if not n.startswith("encoder.fc.")]
    torch.manual_seed(5)
    eng.momentum_buf.normal_()
    eng.ddp.used_parameter_ids = used
    eng._first = False
    path = tmp_path / "checkpoint.pth.tar"
    torch.save(eng.checkpoint_state(epoch=7, arch="resnet18", best_loss=3.5), path)
    cp = torch.load(path, weights_only=False)
    assert set(cp) == {"epoch", "arch", "model", "best_loss", "optimizer", "scheduler"} and cp["epoch"] == 7
    assert list(cp["model"].keys()) == list(model.state_dict().keys())

    def filtered(prefix, blacklist):
        keep = lambda k: k.startswith(prefix) and not any(k.startswith(f"{prefix}{fc}") for fc in blacklist)
        return {k[len(prefix):]: v for k, v in cp["model"].items() if keep(k)}

    # finetune.py:273-303
    ft_state = filtered("encoder_q.", ["fc.", "linear", "head", "new_fc", "fc8", "encoder_fuse"])
    targets = [MultiTaskWrapper(get_model_class(arch="resnet18"), num_classes=101, finetune=True)]
    # retrieval.py:84-101
    rt_state = filtered("encoder_q.encoder.", ["fc", "linear", "head", "new_fc"])
    backbones = [get_model_class(arch="resnet18")(num_classes=101)]
    if Path("/root/reference").exists():
        from oracle import ref_loader
        mods = ref_loader.modules()
        targets.append(mods["wrapper"].MultiTaskWrapper(ref_loader.backbone_ctor("resnet18"), num_classes=101,
                                                        finetune=True))
        backbones.append(ref_loader.backbone_ctor("resnet18")(num_classes=101))
    for tgt in targets:
        msg = tgt.load_state_dict(ft_state, strict=False)
        assert set(msg.missing_keys) == {"fc.weight", "fc.bias"}, msg
        assert all(k.startswith(("fc1.", "fc2.")) for k in msg.unexpected_keys), msg
        assert torch.equal(tgt.state_dict()["encoder.layer4.1.conv2.weight"],
                           cp["model"]["encoder_q.encoder.layer4.1.conv2.weight"])
    for tgt in backbones:
        msg = tgt.load_state_dict(rt_state, strict=False)
        assert set(msg.missing_keys) == {"fc.weight", "fc.bias"} and not msg.unexpected_keys, msg

    # optimizer / scheduler entries in torch's own format (pretrain.py:64-79,123-124)
    other = build_product_moco(cfg, hyper)
    opt = torch.optim.SGD(other.parameters(), lr=1.0, momentum=0.5)
    opt.load_state_dict(cp["optimizer"])
    lr7 = 1e-4 + (0.1 - 1e-4) * (1 + math.cos(math.pi * 7 / 200)) / 2
    assert abs(opt.param_groups[0]["lr"] - lr7) < 1e-12 and opt.param_groups[0]["momentum"] == 0.9
    params = list(other.parameters())
    views = eng._momentum_views()
    assert sum("momentum_buffer" in opt.state.get(p, {}) for p in params) == len(used)
    for i in (used[0], used[-1]):
        assert torch.equal(opt.state[params[i]]["momentum_buffer"], views[i])
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=10, eta_min=0.0)
    sched.load_state_dict(cp["scheduler"])
    assert sched.last_epoch == 7 and sched.T_max == 200 and abs(sched.get_last_lr()[0] - lr7) < 1e-12

    # round trip into a fresh engine
    eng2 = PretrainEngine(other, Loss(2.0, 1.0, 1.0), lr=0.05, momentum=0.0, weight_decay=0.0, num_epochs=5)
    with pytest.raises(ValueError):
        eng2.load_checkpoint(cp, arch="c3d")
    assert eng2.load_checkpoint(cp, arch="resnet18") == 7
    assert abs(eng2.lr - lr7) < 1e-12 and eng2.momentum == 0.9 and eng2.weight_decay == 1e-4 and not eng2._first
    for i in used:
        assert torch.equal(eng2._momentum_views()[i], views[i])
    for (k, a), b in zip(other.state_dict().items(), model.state_dict().values()):
        assert torch.equal(a, b), k


def test_flat_parameters_are_views_and_keep_values():
    from rspnet_b200.models import get_model_class
    from rspnet_b200.moco import MoCoDiffLossTwoFc, MultiTaskWrapper
    moco = MoCoDiffLossTwoFc(lambda num_classes=128: MultiTaskWrapper(get_model_class(arch="resnet18"), num_classes),
                             K=64, diff_speed=[2])
    before = {k: v.clone() for k, v in moco.state_dict().items()}
    fq, fk = moco.flat_parameters()
    assert fq.numel() == fk.numel() and fq.numel() % 4 == 0
    for k, v in moco.state_dict().items():
        assert torch.equal(v, before[k]), k
    p = moco.encoder_q.encoder.conv1.weight
    assert p.data_ptr() == fq.data_ptr()
    fq.zero_()
    assert p.abs().sum() == 0


def _exchange_reference(xs, idx, rank, world):
    return torch.cat(xs, 0)[idx.view(world, -1)[rank]]


def test_exchange_plan_is_the_reference_shuffle():
    from rspnet_b200.moco.exchange import plan_exchange
    g = torch.Generator().manual_seed(0)
    for world, batch in ((1, 5), (2, 3), (4, 8), (8, 64)):
        xs = [torch.randn(batch, 7, generator=g) for _ in range(world)]
        idx = torch.randperm(world * batch, generator=g)
        plans = [plan_exchange(idx, r, world) for r in range(world)]
        for r in range(world):
            recv = []
            for src in range(world):
                p = plans[src]
                start = sum(p.send_counts[:r])
                rows = p.send_index[start:start + p.send_counts[r]]
                assert plans[r].recv_counts[src] == rows.numel()
                recv.append(xs[src][rows])
            got = torch.cat(recv, 0)[plans[r].unpack_index]
            assert torch.equal(got, _exchange_reference(xs, idx, r, world))
        assert sum(sum(p.send_counts) for p in plans) == world * batch


def _gloo_worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rspnet_b200.moco import exchange
    batch = 6
    torch.manual_seed(100 + rank)             # per-rank RNG streams differ, as in pretrain.py:266-267
    x = torch.randn(batch, 4, 3) + 10 * rank
    idx = exchange.broadcast_permutation(torch.randperm(batch * world))
    mine = exchange.exchange_rows(x, idx, lambda s, i: s[i])
    gathered = exchange.all_gather_rows(x)
    # the builder-facing transport object (all_to_all mode under gloo): two steps, two key batches per step
    ex = exchange.ShuffleExchange(batch, (4, 3), torch.float32, torch.device("cpu"))
    steps = []
    for step in range(2):
        kneg_buf, k_buf = ex.begin_step()
        kneg_buf.copy_(x + step)
        k_buf.copy_(-x - step)
        idx2, host2 = ex.draw(batch * world, count=2)
        ex.publish(idx2)
        steps.append(dict(idx=idx2.clone(), kneg=ex.pull(ex.SLOT_KNEG, idx2[0], host2[0], lambda s, i: s[i]),
                          k=ex.pull(ex.SLOT_K, idx2[1], host2[1], lambda s, i: s[i])))
    torch.save(dict(x=x, idx=idx, mine=mine, gathered=gathered, steps=steps, mode=ex.mode), Path(tmp) / f"r{rank}.pt")
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_two_ranks_gloo(tmp_path):
    world = 2
    mp.spawn(_gloo_worker, args=(world, 29731, str(tmp_path)), nprocs=world, join=True)
    recs = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    assert torch.equal(recs[0]["idx"], recs[1]["idx"])          # rank 0's permutation everywhere
    xs = [r["x"] for r in recs]
    for r in range(world):
        assert torch.equal(recs[r]["gathered"], torch.cat(xs, 0))
        assert torch.equal(recs[r]["mine"], _exchange_reference(xs, recs[0]["idx"], r, world))
        assert recs[r]["mode"] == "a2a"
        for step, st in enumerate(recs[r]["steps"]):
            assert torch.equal(st["idx"], recs[0]["steps"][step]["idx"])       # rank 0's two permutations everywhere
            assert torch.equal(st["kneg"], _exchange_reference([x + step for x in xs], st["idx"][0], r, world))
            assert torch.equal(st["k"], _exchange_reference([-x - step for x in xs], st["idx"][1], r, world))
