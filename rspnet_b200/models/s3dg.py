"""S3D-G (separable 3-D inception with self-gating), parameter names of the reference's ``models/s3dg.py``
(``BasicConv3d`` :6-33, ``sep_conv`` :36-72, ``sep_inc`` :74-99, ``S3D_G`` :102-153).

Every BasicConv3d is a fused Conv3d+BN(eps 1e-3, momentum 0.001)+ReLU tcgen05 block; the gate (squeeze ->
1x1x1 excitation conv -> sigmoid -> scale) and the inception concat run as dedicated NDHWC kernels.  Branch widths
that are not multiples of 64 are zero-padded in the stored layout only.
"""
from collections import OrderedDict

import torch
from torch import nn

from .. import nn as rnn


class BasicConv3d(nn.Module):
    def __init__(self, in_channel, out_channel, kernel_size=1, stride=1, padding=0, use_bias=False, use_bn=True,
                 activation='rule'):
        super().__init__()
        if not use_bn or activation != 'rule':
            raise NotImplementedError("S3D-G only instantiates conv+BN+ReLU blocks")
        self.use_bn = use_bn
        self.conv3d = nn.Conv3d(in_channel, out_channel, kernel_size=kernel_size, stride=stride, padding=padding,
                                bias=use_bias)
        self.bn = nn.BatchNorm3d(out_channel, eps=1e-3, momentum=0.001, affine=True)
        self.activation = nn.ReLU()
        self.out_channels = out_channel

    def forward(self, x):
        return rnn.conv_bn_act(rnn.as_ndhwc(x), self.conv3d, self.bn, relu=True)


class sep_conv(nn.Module):
    def __init__(self, in_channel, out_channel, kernel_size, stride=1, padding=0, use_bias=True, use_bn=True,
                 activation='rule', gate=True):
        super().__init__()
        down = BasicConv3d(in_channel, out_channel, (1, kernel_size, kernel_size), stride=stride,
                           padding=(0, padding, padding), use_bias=False, use_bn=True)
        up = BasicConv3d(out_channel, out_channel, (kernel_size, 1, 1), stride=1, padding=(padding, 0, 0),
                         use_bias=False, use_bn=True)
        self.sep_conv = nn.Sequential(down, up)
        self.out_channels = out_channel
        if gate:
            self.gate = gate
            self.squeeze = nn.AdaptiveAvgPool3d(1)
            self.excitation = nn.Conv3d(out_channel, out_channel, 1)
            self.sigmoid = nn.Sigmoid()
        else:
            self.gate = False

    def forward(self, x):
        x = self.sep_conv(x)
        if self.gate:
            x = rnn.GateFn.apply(x, self.excitation.weight, self.excitation.bias, self.out_channels)
        return x


class sep_inc(nn.Module):
    def __init__(self, in_channel, out_channel, gate=True):
        super().__init__()
        self.branch0 = BasicConv3d(in_channel, out_channel[0], kernel_size=(1, 1, 1), stride=1, padding=0)
        branch1_conv1 = BasicConv3d(in_channel, out_channel[1], kernel_size=(1, 1, 1), stride=1, padding=0)
        branch1_sep_conv = sep_conv(out_channel[1], out_channel[2], kernel_size=3, stride=1, padding=1, gate=gate)
        self.branch1 = nn.Sequential(branch1_conv1, branch1_sep_conv)
        branch2_conv1 = BasicConv3d(in_channel, out_channel[3], kernel_size=(1, 1, 1), stride=1, padding=0)
        branch2_sep_conv = sep_conv(out_channel[3], out_channel[4], kernel_size=3, stride=1, padding=1, gate=gate)
        self.branch2 = nn.Sequential(branch2_conv1, branch2_sep_conv)
        branch3_pool = nn.MaxPool3d(kernel_size=3, stride=1, padding=1)
        branch3_conv = BasicConv3d(in_channel, out_channel[5], kernel_size=(1, 1, 1))
        self.branch3 = nn.Sequential(branch3_pool, branch3_conv)
        self.widths = (out_channel[0], out_channel[2], out_channel[4], out_channel[5])

    def forward(self, x):
        out_0 = self.branch0(x)
        out_1 = self.branch1(x)
        out_2 = self.branch2(x)
        out_3 = self.branch3[1](rnn.max_pool3d(x, self.branch3[0]))
        return rnn.concat_channels((out_0, out_1, out_2, out_3), self.widths)


class S3D_G(nn.Module):
    def __init__(self, num_classes=400, drop_prob=0.5, in_channel=3, gate=True):
        super().__init__()
        self.feature = nn.Sequential(OrderedDict([
            ('sepConv1', sep_conv(in_channel, 64, kernel_size=7, stride=2, padding=3, gate=gate)),
            ('maxPool1', nn.MaxPool3d(kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1))),
            ('basicConv3d', BasicConv3d(64, 64, kernel_size=1, stride=1)),
            ('sep_conv2', sep_conv(64, 192, kernel_size=3, stride=1, padding=1, gate=gate)),
            ('maxPool2', nn.MaxPool3d(kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1))),
            ('sepInc_3b', sep_inc(192, [64, 96, 128, 16, 32, 32], gate=gate)),
            ('sepInc_3c', sep_inc(256, [128, 128, 192, 32, 96, 64], gate=gate)),
            ('maxPool3', nn.MaxPool3d(kernel_size=(3, 3, 3), stride=(2, 2, 2), padding=(1, 1, 1))),
            ('sepInc_4b', sep_inc(480, [192, 96, 208, 16, 48, 64], gate=gate)),
            ('sepInc_4c', sep_inc(512, [160, 112, 224, 24, 64, 64], gate=gate)),
            ('sepInc_4d', sep_inc(512, [128, 128, 256, 24, 64, 64], gate=gate)),
            ('sepInc_4e', sep_inc(512, [112, 144, 288, 32, 64, 64], gate=gate)),
            ('sepInc_4f', sep_inc(528, [256, 160, 320, 32, 128, 128], gate=gate)),
            ('maxpool4', nn.MaxPool3d(kernel_size=(2, 2, 2), stride=(2, 2, 2), padding=(0, 0, 0))),
            ('sepInc_5b', sep_inc(832, [256, 160, 320, 32, 128, 128], gate=gate)),
            ('sepInc_5c', sep_inc(832, [384, 192, 384, 48, 128, 128], gate=gate)),
        ]))
        self.avg_pool = nn.AdaptiveAvgPool3d((1, 1, 1))
        self.drop = nn.Dropout(drop_prob)
        self.fc = nn.Linear(1024, num_classes)

    feature_channels = 1024

    def feature_ndhwc(self, x):
        x = rnn.as_ndhwc(x)
        for layer in self.feature:
            x = rnn.max_pool3d(x, layer) if isinstance(layer, nn.MaxPool3d) else layer(x)
        return x

    def get_feature(self, x):
        return rnn.ToNCDHW.apply(self.feature_ndhwc(x), 1024)

    def forward(self, x):
        out = self.avg_pool(self.get_feature(x)).flatten(1)
        return self.fc(self.drop(out))
