"""Shared helpers for the parity tests (test infrastructure)."""
import random
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"


def load_golden(name):
    return torch.load(GOLDEN / f"{name}.pt", weights_only=False)


def initialize_seed(seed):
    """framework/utils/reproduction.py:29-33."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def make_inputs(cfg, rank, step):
    """Same synthetic clips as oracle/make_golden.py::make_inputs."""
    g = torch.Generator().manual_seed(1234 + rank + 1000 * step)
    shape = (cfg["batch"], 3, cfg["frames"], cfg["size"], cfg["size"])
    return torch.randn(shape, generator=g), torch.randn(shape, generator=g)


def build_product_moco(cfg, hyper, rank=0):
    """The product's MoCoDiffLossTwoFc built on CPU under the reference's seeding (construction has no CUDA dependency)."""
    from rspnet_b200.models import get_model_class
    from rspnet_b200.moco import MoCoDiffLossTwoFc, MultiTaskWrapper
    initialize_seed(cfg["seed"] + rank)
    base = get_model_class(arch=cfg["arch"])

    def model_class(num_classes=128):
        return MultiTaskWrapper(base, num_classes=num_classes, fc_type="linear", finetune=False, groups=1)

    return MoCoDiffLossTwoFc(model_class, dim=hyper["dim"], K=cfg["K"], m=hyper["m"], T=hyper["T"],
                             diff_speed=list(hyper["diff_speed"]))


def build_product_single_head(cfg, hyper, rank=0):
    """The product's MoCoDiffLoss (single head: the backbone's own fc is the projection) under the reference's seeding."""
    from rspnet_b200.models import get_model_class
    from rspnet_b200.moco import MoCoDiffLoss
    initialize_seed(cfg["seed"] + rank)
    return MoCoDiffLoss(get_model_class(arch=cfg["arch"]), dim=hyper["dim"], K=cfg["K"], m=hyper["m"], T=hyper["T"],
                        diff_speed=list(hyper["diff_speed"]))


def summarize(t):
    t = t.detach().double().flatten()
    return dict(sum=float(t.sum()), abssum=float(t.abs().sum()), n=t.numel(), head=t[:32].float().clone())


def check_packed(got: torch.Tensor, packed, rtol, atol, what="", norm_only=False):
    """Compare a tensor with a golden entry that is either a full tensor or a checksum dict.
    norm_only: compare in relative L2 norm (full tensors) / abs-sum (checksums) instead of element-wise."""
    if norm_only:
        if isinstance(packed, dict):
            a = summarize(got)["abssum"]
            assert abs(a - packed["abssum"]) <= rtol * packed["abssum"] + atol * packed["n"], what
        else:
            ref = packed.float()
            err = (got.detach().float() - ref).norm().item()
            assert err <= rtol * ref.norm().item() + atol * ref.numel() ** 0.5, f"{what}: L2 err {err}"
        return
    if isinstance(packed, dict):
        s = summarize(got)
        assert s["n"] == packed["n"], what
        torch.testing.assert_close(s["head"], packed["head"], rtol=rtol, atol=atol, msg=lambda m: f"{what}: {m}")
        scale = max(packed["abssum"], 1e-12)
        assert abs(s["abssum"] - packed["abssum"]) <= rtol * scale + atol * packed["n"], \
            f"{what}: abssum {s['abssum']} vs {packed['abssum']}"
        assert abs(s["sum"] - packed["sum"]) <= rtol * scale + atol * packed["n"], \
            f"{what}: sum {s['sum']} vs {packed['sum']}"
    else:
        torch.testing.assert_close(got.detach().float(), packed.float(), rtol=rtol, atol=atol,
                                   msg=lambda m: f"{what}: {m}")
