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
    torch.save(dict(x=x, idx=idx, mine=mine, gathered=gathered), Path(tmp) / f"r{rank}.pt")
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
