"""Models registry with the reference's interface (``models/__init__.py:16-75``):
``get_model_class(**model_cfg)`` maps an ``arch`` string to a constructor called as ``cls(num_classes=int)``.
Instances provide ``forward(x)``, ``get_feature(x)`` ([B,3,T,H,W] fp32 -> [B,C,t,h,w] fp32) and the
channels-last fast path ``feature_ndhwc(x)`` used by the MoCo wrapper.
"""
import logging
from typing import Callable

from torch import nn

logger = logging.getLogger(__name__)

PRETRAIN_ARCHS = ("resnet18", "resnet34", "resnet50", "c3d", "s3dg", "r2plus1d-vcop")


def get_model_class(**kwargs) -> Callable[[int], nn.Module]:
    logger.info(f'Using global get_model_class({kwargs})')
    arch = str(kwargs['arch'])
    if arch in ('resnet18', 'resnet34', 'resnet50'):
        from . import resnet
        return getattr(resnet, arch)
    if arch == 'c3d':
        from .c3d import C3D
        return C3D
    if arch == 's3dg':
        from .s3dg import S3D_G
        return S3D_G
    if arch == 'r2plus1d-vcop':
        from .r2plus1d_vcop import R2Plus1DNet
        return lambda num_classes=128: R2Plus1DNet((1, 1, 1, 1), with_classifier=True, num_classes=num_classes)
    if arch in ('torchvision-resnet18', 'mfnet', 'tsm') or arch.startswith('SLOWFAST'):
        raise NotImplementedError(
            f'arch "{arch}" has no get_feature() in the reference and cannot be used for RSPNet pretraining; '
            'it is outside the scope of rspnet_b200')
    raise ValueError(f'Unknown model architecture "{arch}"')
