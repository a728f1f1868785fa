"""``MultiTaskWrapper`` — backbone plus two projection heads (A-VID, RSP), interface and state_dict names of the
reference's ``moco/split_wrapper.py`` (:66-190).  With ``fc_type='linear'`` (the shipped configs) the pooled
features, both Linear heads and both L2 normalisations run as one fused kernel (rspnet_b200.nn.HeadsFn); the
'conv' / 'convbn' heads (:17-64) run their 3x3x3 convolutions (and BatchNorm) on the conv kernels of the backbone.
"""
import logging
from typing import Callable, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from .. import nn as rnn

logger = logging.getLogger(__name__)


class Flatten(nn.Module):
    def forward(self, x: Tensor):
        return x.flatten(1)


class ConvFc(nn.Module):
    """conv -> relu -> conv -> global average -> linear (split_wrapper.py:17-40); input: bf16 NDHWC feature map."""

    def __init__(self, feat_dim: int, moco_dim: int, kernel_size: Tuple[int, int, int], padding: Tuple[int, int, int]):
        super().__init__()
        self.conv1 = nn.Conv3d(feat_dim, feat_dim, kernel_size, padding=padding)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv3d(feat_dim, feat_dim, kernel_size, padding=padding)
        self.avg_pool = nn.AdaptiveAvgPool3d((1, 1, 1))
        self.linear = nn.Linear(feat_dim, moco_dim)

    def forward(self, x: Tensor):
        out = rnn.conv_bias_act(x, self.conv1, relu=True)
        out = rnn.conv_bias_act(out, self.conv2, relu=False)
        out = rnn.ToNCDHW.apply(out, self.conv2.out_channels)
        return self.linear(self.avg_pool(out).flatten(1))


class ConvBnFc(nn.Module):
    """conv -> bn -> relu -> global average -> linear (split_wrapper.py:43-64); input: bf16 NDHWC feature map."""

    def __init__(self, feat_dim: int, moco_dim: int, kernel_size: Tuple[int, int, int], padding: Tuple[int, int, int]):
        super().__init__()
        self.conv1 = nn.Conv3d(feat_dim, feat_dim, kernel_size, padding=padding)
        self.bn = nn.BatchNorm3d(feat_dim)
        self.relu = nn.ReLU(inplace=True)
        self.avg_pool = nn.AdaptiveAvgPool3d((1, 1, 1))
        self.linear = nn.Linear(feat_dim, moco_dim)

    def forward(self, x: Tensor):
        out = rnn.conv_bn_act(x, self.conv1, self.bn, relu=True)
        out = rnn.ToNCDHW.apply(out, self.conv1.out_channels)
        return self.linear(self.avg_pool(out).flatten(1))


class MultiTaskWrapper(nn.Module):
    def __init__(self, base_encoder: Callable[[int], nn.Module], num_classes: int = 128, finetune: bool = False,
                 fc_type: str = 'linear', groups: int = 1):
        super().__init__()
        logger.info('Using MultiTask Wrapper')
        self.finetune = finetune
        self.moco_dim = num_classes
        self.num_classes = num_classes
        self.groups = groups
        self.fc_type = fc_type
        self.feat = None
        self.encoder = base_encoder(num_classes=1)
        feat_dim = self._get_feat_dim(self.encoder) // groups
        if self.finetune:
            self.avg_pool = nn.AdaptiveAvgPool3d((1, 1, 1))
            self.fc = nn.Linear(feat_dim, num_classes)
        elif fc_type == 'linear':
            self.fc1 = self._get_linear_fc(feat_dim, self.moco_dim)
            self.fc2 = self._get_linear_fc(feat_dim, self.moco_dim)
        elif fc_type == 'mlp':
            self.fc1 = self._get_mlp_fc(feat_dim, self.moco_dim)
            self.fc2 = self._get_mlp_fc(feat_dim, self.moco_dim)
        elif fc_type == 'conv':
            self.fc1 = ConvFc(feat_dim, self.moco_dim, (3, 3, 3), (1, 1, 1))
            self.fc2 = ConvFc(feat_dim, self.moco_dim, (3, 3, 3), (1, 1, 1))
        elif fc_type == 'convbn':
            self.fc1 = ConvBnFc(feat_dim, self.moco_dim, (3, 3, 3), (1, 1, 1))
            self.fc2 = ConvBnFc(feat_dim, self.moco_dim, (3, 3, 3), (1, 1, 1))
        elif fc_type == 'speednet':
            self.fc1 = self._get_linear_fc(feat_dim, self.moco_dim)
            self.fc2 = self._get_linear_fc(feat_dim, 1)
        # (any other fc_type leaves the wrapper without heads, as in the reference)

    def forward(self, x: Tensor):
        feat = self.encoder.feature_ndhwc(x)  # bf16 NDHWC
        self.feat = feat
        counters = getattr(self, "_rsp_bn_counters", None)
        if counters is not None:   # all BatchNorm num_batches_tracked of this encoder, one launch (see nn.batch_bn_counters)
            counters += 1
        fused = (not self.finetune and self.fc_type == 'linear' and self.groups == 1)
        if fused:
            l1, l2 = self.fc1[2], self.fc2[2]
            return rnn.HeadsFn.apply(feat, l1.weight, l1.bias, l2.weight, l2.bias, l1.in_features)
        if not self.finetune and self.fc_type in ('conv', 'convbn'):
            # conv heads consume the bf16 NDHWC map directly (the chunk of groups == 2 is a channel slice)
            if self.groups == 1:
                x1, x2 = self.fc1(feat), self.fc2(feat)
            elif self.groups == 2:
                half = self.encoder.feature_channels // 2
                if half % 64 != 0:
                    raise NotImplementedError("rspnet_b200: groups=2 conv heads need feature_channels / 2 % 64 == 0")
                x1, x2 = self.fc1(feat[..., :half].contiguous()), self.fc2(feat[..., half:2 * half].contiguous())
            else:
                raise Exception
            return F.normalize(x1, dim=1), F.normalize(x2, dim=1)
        # generic (non-hot) variants run on the reference layout
        f = rnn.ToNCDHW.apply(feat, self.encoder.feature_channels)
        if self.finetune:
            return self.fc(self.avg_pool(f).flatten(1))
        if self.groups == 1:
            x1, x2 = self.fc1(f), self.fc2(f)
        elif self.groups == 2:
            f1, f2 = f.chunk(2, 1)
            x1, x2 = self.fc1(f1), self.fc2(f2)
        else:
            raise Exception
        x1 = F.normalize(x1, dim=1)
        x2 = torch.sigmoid(x2) if self.fc_type == 'speednet' else F.normalize(x2, dim=1)
        return x1, x2

    def _get_last_feature(self):
        """Feature map of the last forward in the reference layout ([B,C,t,h,w] fp32)."""
        return rnn.ToNCDHW.apply(self.feat, self.encoder.feature_channels)

    def _get_fc_weight(self) -> Tuple[Tensor, Tensor]:
        with torch.no_grad():
            return self.fc1[2].weight.data, self.fc2[2].weight.data

    @staticmethod
    def _get_linear_fc(feat_dim: int, moco_dim: int):
        return nn.Sequential(nn.AdaptiveAvgPool3d((1, 1, 1)), Flatten(), nn.Linear(feat_dim, moco_dim))

    @staticmethod
    def _get_mlp_fc(feat_dim: int, moco_dim: int):
        return nn.Sequential(nn.AdaptiveAvgPool3d((1, 1, 1)), Flatten(), nn.Linear(feat_dim, feat_dim),
                             nn.ReLU(inplace=True), nn.Linear(feat_dim, moco_dim))

    @staticmethod
    def _get_feat_dim(encoder):
        feat_dim = 512
        for fc_name in ('fc', 'new_fc', 'classifier'):
            if hasattr(encoder, fc_name):
                feat_dim = getattr(encoder, fc_name).in_features
                logger.info(f'Found fc: {fc_name} with in_features: {feat_dim}')
                break
        return feat_dim
