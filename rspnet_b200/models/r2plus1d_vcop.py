"""R(2+1)D (VCOP variant) with the parameter names of the reference's ``models/r2plus1d_vcop.py``
(``SpatioTemporalConv`` :13-72, ``SpatioTemporalResBlock`` :75-123, ``R2Plus1DNet`` :160-224).

Every (1,k,k) / (k,1,1) convolution is followed by a BatchNorm, so the whole network is a chain of fused
Conv3d+BN(+residual)(+ReLU) kernels.  The odd intermediate widths (83, 144, 230, ...) are zero-padded to a
multiple of 64 channels in the packed operands; padded lanes stay exactly zero through BN.
"""
import math

from torch import nn
from torch.nn.modules.utils import _triple

from .. import nn as rnn


class SpatioTemporalConv(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=False, first_conv=False):
        super().__init__()
        k, s, p = _triple(kernel_size), _triple(stride), _triple(padding)
        mid = int(math.floor((k[0] * k[1] * k[2] * in_channels * out_channels) /
                             (k[1] * k[2] * in_channels + k[0] * out_channels)))
        self.spatial_conv = nn.Conv3d(in_channels, mid, (1, k[1], k[2]), stride=(1, s[1], s[2]),
                                      padding=(0, p[1], p[2]), bias=bias)
        self.bn = nn.BatchNorm3d(mid)
        self.relu = nn.ReLU()
        self.temporal_conv = nn.Conv3d(mid, out_channels, (k[0], 1, 1), stride=(s[0], 1, 1), padding=(p[0], 0, 0),
                                       bias=bias)

    def fused(self, x, bn, relu, residual=None):
        """spatial conv -> bn -> relu -> temporal conv -> ``bn`` (the caller's) (+residual) (-> relu)."""
        x = rnn.conv_bn_act(x, self.spatial_conv, self.bn, relu=True)
        return rnn.conv_bn_act(x, self.temporal_conv, bn, relu=relu, residual=residual)


class SpatioTemporalResBlock(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, downsample=False):
        super().__init__()
        self.downsample = downsample
        padding = kernel_size // 2
        if self.downsample:
            self.downsampleconv = SpatioTemporalConv(in_channels, out_channels, 1, stride=2)
            self.downsamplebn = nn.BatchNorm3d(out_channels)
            self.conv1 = SpatioTemporalConv(in_channels, out_channels, kernel_size, padding=padding, stride=2)
        else:
            self.conv1 = SpatioTemporalConv(in_channels, out_channels, kernel_size, padding=padding)
        self.bn1 = nn.BatchNorm3d(out_channels)
        self.relu1 = nn.ReLU()
        self.conv2 = SpatioTemporalConv(out_channels, out_channels, kernel_size, padding=padding)
        self.bn2 = nn.BatchNorm3d(out_channels)
        self.outrelu = nn.ReLU()

    def forward(self, x):
        x = rnn.as_ndhwc(x)
        shortcut = x
        if self.downsample:
            shortcut = self.downsampleconv.fused(x, self.downsamplebn, relu=False)
        res = self.conv1.fused(x, self.bn1, relu=True)
        return self.conv2.fused(res, self.bn2, relu=True, residual=shortcut)


class SpatioTemporalResLayer(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, layer_size, block_type=SpatioTemporalResBlock,
                 downsample=False):
        super().__init__()
        self.block1 = block_type(in_channels, out_channels, kernel_size, downsample)
        self.blocks = nn.ModuleList([])
        for _ in range(layer_size - 1):
            self.blocks += [block_type(out_channels, out_channels, kernel_size)]

    def forward(self, x):
        x = self.block1(x)
        for block in self.blocks:
            x = block(x)
        return x


class R2Plus1DNet(nn.Module):
    def __init__(self, layer_sizes, block_type=SpatioTemporalResBlock, with_classifier=False, return_conv=False,
                 num_classes=101):
        super().__init__()
        self.with_classifier = with_classifier
        self.return_conv = return_conv
        self.num_classes = num_classes
        self.conv1 = SpatioTemporalConv(3, 64, (3, 7, 7), stride=(1, 2, 2), padding=(1, 3, 3))
        self.bn1 = nn.BatchNorm3d(64)
        self.relu1 = nn.ReLU()
        self.conv2 = SpatioTemporalResLayer(64, 64, 3, layer_sizes[0], block_type=block_type)
        self.conv3 = SpatioTemporalResLayer(64, 128, 3, layer_sizes[1], block_type=block_type, downsample=True)
        self.conv4 = SpatioTemporalResLayer(128, 256, 3, layer_sizes[2], block_type=block_type, downsample=True)
        self.conv5 = SpatioTemporalResLayer(256, 512, 3, layer_sizes[3], block_type=block_type, downsample=True)
        if self.return_conv:
            self.feature_pool = nn.MaxPool3d(kernel_size=(1, 2, 2), stride=(1, 2, 2))
        self.pool = nn.AdaptiveAvgPool3d(1)
        if self.with_classifier:
            self.linear = nn.Linear(512, self.num_classes)

    feature_channels = 512

    def feature_ndhwc(self, x):
        x = rnn.as_ndhwc(x)
        x = self.conv1.fused(x, self.bn1, relu=True)
        x = self.conv2(x)
        x = self.conv3(x)
        x = self.conv4(x)
        return self.conv5(x)

    def get_feature(self, x):
        return rnn.ToNCDHW.apply(self.feature_ndhwc(x), 512)

    def forward(self, x):
        x = self.feature_ndhwc(x)
        if self.return_conv:
            x = rnn.ToNCDHW.apply(rnn.max_pool3d(x, self.feature_pool), 512)
            return x.view(x.shape[0], -1)
        x = self.pool(rnn.ToNCDHW.apply(x, 512)).view(-1, 512)
        if self.with_classifier:
            x = self.linear(x)
        return x
