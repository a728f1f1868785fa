"""TEST / BASELINE INFRASTRUCTURE ONLY — loads the *unmodified* reference modules by file path, from /root/reference
in the build container or from the git-ignored byte-for-byte copy ``baseline/_ref`` (oracle/install_ref.py) on the
GPU box.

Used (a) to pin oracle/rspnet_oracle.py against the real reference, (b) to generate the golden vectors under
tests/golden/ (oracle/make_golden*.py), (c) by ``bench.py --impl reference`` / ``cpu_baseline`` and
tools/stock_torch_bar.py to time the reference itself.  The product package never imports this module.

Shims applied, none of which changes the reference's arithmetic:
  * modules are loaded with importlib from their file paths because ``import moco`` / ``import models`` need
    pyhocon, which is not installed;
  * ``torch.Tensor.cuda`` becomes a no-op on CPU (``_batch_shuffle_ddp`` hard-codes ``.cuda()``,
    moco/builder_diffspeed_diffloss.py:375);
  * ``F.margin_ranking_loss`` accepts the reference's (N,1),(N,1),(N,) shapes with torch-1.6 broadcasting
    semantics: ``clamp(-y*(x1-x2)+margin, min=0).mean()`` (builder:281).
"""
import importlib.util
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

_LIVE = Path("/root/reference")
_VENDORED = Path(__file__).resolve().parent.parent / "baseline" / "_ref"   # oracle/install_ref.py (git-ignored copy)
REFERENCE_ROOT = _LIVE if (_LIVE / "moco" / "builder_diffspeed_diffloss.py").exists() else _VENDORED


def available() -> bool:
    return (REFERENCE_ROOT / "moco" / "builder_diffspeed_diffloss.py").exists()


def _load(name: str, rel: str):
    spec = importlib.util.spec_from_file_location(name, str(REFERENCE_ROOT / rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_cache = {}


def modules():
    """Returns dict(resnet, c3d, s3dg, r2plus1d, wrapper, builder) of reference modules."""
    if not _cache:
        if not available():
            raise RuntimeError("neither /root/reference nor baseline/_ref is present")
        _cache["resnet"] = _load("_ref_resnet", "models/resnet.py")
        _cache["c3d"] = _load("_ref_c3d", "models/c3d.py")
        _cache["s3dg"] = _load("_ref_s3dg", "models/s3dg.py")
        _cache["r2plus1d"] = _load("_ref_r2plus1d", "models/r2plus1d_vcop.py")
        _cache["wrapper"] = _load("_ref_split_wrapper", "moco/split_wrapper.py")
        _cache["builder"] = _load("_ref_builder", "moco/builder_diffspeed_diffloss.py")
    return _cache


_orig_mrl = F.margin_ranking_loss
_shimmed = False
FORCE_CPU = False   # set before install_shims() to run the reference on the host cores of a box that has a GPU


def install_shims():
    global _shimmed
    if _shimmed:
        return
    _shimmed = True

    def margin_ranking_loss(input1, input2, target, margin=0.0, size_average=None, reduce=None, reduction="mean"):
        if input1.dim() != target.dim():
            assert reduction == "mean"
            return torch.clamp(-target * (input1 - input2) + margin, min=0).mean()
        return _orig_mrl(input1, input2, target, margin=margin, reduction=reduction)

    F.margin_ranking_loss = margin_ranking_loss
    torch.nn.functional.margin_ranking_loss = margin_ranking_loss
    if not torch.cuda.is_available() or FORCE_CPU:
        torch.Tensor.cuda = lambda self, *a, **k: self


def backbone_ctor(arch: str):
    m = modules()
    if arch == "resnet18":
        return m["resnet"].resnet18
    if arch == "c3d":
        return m["c3d"].C3D
    if arch == "s3dg":
        return m["s3dg"].S3D_G
    if arch == "r2plus1d-vcop":
        return lambda num_classes=128: m["r2plus1d"].R2Plus1DNet((1, 1, 1, 1), with_classifier=True,
                                                                 num_classes=num_classes)
    raise ValueError(arch)


def build_reference_moco(arch: str, dim=128, K=16384, m=0.999, T=0.07, diff_speed=(2,)):
    """MoCoDiffLossTwoFc exactly as moco/__init__.py:19-46 builds it (minus .cuda()/DDP)."""
    install_shims()
    mods = modules()
    base = backbone_ctor(arch)

    def model_class(num_classes=128):
        return mods["wrapper"].MultiTaskWrapper(base, num_classes=num_classes, fc_type="linear", finetune=False,
                                                groups=1)

    return mods["builder"].MoCoDiffLossTwoFc(model_class, dim=dim, K=K, m=m, T=T, diff_speed=list(diff_speed))


def build_reference_loss(margin=2.0, A=1.0, M=1.0):
    install_shims()
    return modules()["builder"].Loss(margin=margin, A=A, M=M)
