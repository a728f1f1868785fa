"""C3D with BatchNorm, parameter names as in the reference's ``models/c3d.py`` (:13-150).

Eight Conv3d(3x3x3, pad 1, bias=True)+BN+ReLU blocks and four max-pools; ``get_feature`` stops before pool5.
Executed as fused tcgen05 conv + BN/ReLU kernels on bf16 NDHWC activations (rspnet_b200.nn).
"""
import torch
from torch import nn

from .. import nn as rnn


class C3D(nn.Module):
    """C3D with BN and pool5 = AdaptiveAvgPool3d(1)."""

    _PLAN = [("1", 3, 64), ("2", 64, 128), ("3a", 128, 256), ("3b", 256, 256), ("4a", 256, 512), ("4b", 512, 512),
             ("5a", 512, 512), ("5b", 512, 512)]
    _POOL_AFTER = {"1": "pool1", "2": "pool2", "3b": "pool3", "4b": "pool4"}

    def __init__(self, with_classifier=True, return_conv=False, num_classes=101):
        super().__init__()
        self.with_classifier = with_classifier
        self.num_classes = num_classes
        self.return_conv = return_conv
        pools = {"pool1": ((1, 2, 2), (1, 2, 2)), "pool2": ((2, 2, 2), (2, 2, 2)), "pool3": ((2, 2, 2), (2, 2, 2)),
                 "pool4": ((2, 2, 2), (2, 2, 2))}
        # registration order follows the reference: conv, bn, relu, (pool) per stage
        for tag, cin, cout in self._PLAN:
            setattr(self, "conv" + tag, nn.Conv3d(cin, cout, kernel_size=(3, 3, 3), padding=(1, 1, 1)))
            setattr(self, "bn" + tag, nn.BatchNorm3d(cout))
            setattr(self, "relu" + tag, nn.ReLU())
            pool = self._POOL_AFTER.get(tag)
            if pool:
                k, s = pools[pool]
                setattr(self, pool, nn.MaxPool3d(kernel_size=k, stride=s))
        if self.return_conv:
            self.feature_pool = nn.MaxPool3d(kernel_size=(1, 2, 2), stride=(1, 2, 2))
        self.pool5 = nn.AdaptiveAvgPool3d(1)
        if self.with_classifier:
            self.linear = nn.Linear(512, self.num_classes)

    feature_channels = 512

    def feature_ndhwc(self, x):
        x = rnn.as_ndhwc(x)
        for tag, _, _ in self._PLAN:
            pool = self._POOL_AFTER.get(tag)
            if pool:
                x = rnn.conv_bn_relu_pool(x, getattr(self, "conv" + tag), getattr(self, "bn" + tag), getattr(self, pool))
            else:
                x = rnn.conv_bn_act(x, getattr(self, "conv" + tag), getattr(self, "bn" + tag), relu=True)
        if self.return_conv:
            x = rnn.max_pool3d(x, self.feature_pool)
        return x

    def get_feature(self, x):
        x = rnn.ToNCDHW.apply(self.feature_ndhwc(x), 512)
        if self.return_conv:
            return x.view(x.shape[0], -1)
        return x

    def forward(self, x):
        x = self.get_feature(x)
        if self.return_conv:
            return x
        x = self.pool5(x).view(-1, 512)
        if self.with_classifier:
            x = self.linear(x)
        return x
