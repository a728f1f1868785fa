"""3-D ResNets (Hara et al.) with the parameter / buffer names of the reference's ``models/resnet.py``
(``ResNet`` :119-223, ``BasicBlock`` :48-77, ``Bottleneck`` :80-116, constructors :247-301) so that
``state_dict()`` round-trips with reference checkpoints, while every block executes as fused
Conv3d+BN(+residual)+ReLU tcgen05 kernels on bf16 NDHWC activations (rspnet_b200.nn).
"""
import math

import torch
from torch import nn

from .. import nn as rnn

__all__ = ['ResNet', 'resnet10', 'resnet18', 'resnet34', 'resnet50', 'resnet101', 'resnet152', 'resnet200']


def _conv3(inp, out, stride=1):
    return nn.Conv3d(inp, out, kernel_size=3, stride=stride, padding=1, bias=False)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = _conv3(inplanes, planes, stride)
        self.bn1 = nn.BatchNorm3d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = _conv3(planes, planes)
        self.bn2 = nn.BatchNorm3d(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        x = rnn.as_ndhwc(x)
        shortcut = x
        if self.downsample is not None:
            shortcut = rnn.conv_bn_act(x, self.downsample[0], self.downsample[1], relu=False)
        out = rnn.conv_bn_act(x, self.conv1, self.bn1, relu=True)
        return rnn.conv_bn_act(out, self.conv2, self.bn2, relu=True, residual=shortcut)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv3d(inplanes, planes, kernel_size=1, bias=False)
        self.bn1 = nn.BatchNorm3d(planes)
        self.conv2 = nn.Conv3d(planes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm3d(planes)
        self.conv3 = nn.Conv3d(planes, planes * 4, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm3d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        x = rnn.as_ndhwc(x)
        shortcut = x
        if self.downsample is not None:
            shortcut = rnn.conv_bn_act(x, self.downsample[0], self.downsample[1], relu=False)
        out = rnn.conv_bn_act(x, self.conv1, self.bn1, relu=True)
        out = rnn.conv_bn_act(out, self.conv2, self.bn2, relu=True)
        return rnn.conv_bn_act(out, self.conv3, self.bn3, relu=True, residual=shortcut)


class ResNet(nn.Module):

    def __init__(self, block, layers, sample_size=112, sample_duration=16, shortcut_type='B', num_classes=400):
        super().__init__()
        if shortcut_type != 'B':
            raise NotImplementedError("only shortcut type B (1x1x1 conv + BN) is used by the reference constructors")
        self.inplanes = 64
        self.conv1 = nn.Conv3d(3, 64, kernel_size=7, stride=(1, 2, 2), padding=(3, 3, 3), bias=False)
        self.bn1 = nn.BatchNorm3d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool3d(kernel_size=(3, 3, 3), stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=2)
        last_duration = int(math.ceil(sample_duration / 16))
        last_size = int(math.ceil(sample_size / 32))
        self.avgpool = nn.AvgPool3d((last_duration, last_size, last_size), stride=1)
        self.fc = nn.Linear(512 * block.expansion, num_classes)
        # same draw order as the reference (:153-158): kaiming-normal(fan_out) per conv in module order
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                m.weight = nn.init.kaiming_normal_(m.weight, mode='fan_out')
            elif isinstance(m, nn.BatchNorm3d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _make_layer(self, block, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                nn.Conv3d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False),
                nn.BatchNorm3d(planes * block.expansion))
        stack = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * block.expansion
        stack += [block(self.inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*stack)

    # ---- B200 path -------------------------------------------------------------------------------------------
    def feature_ndhwc(self, x):
        """get_feature on channels-last bf16 activations; returns bf16 [N, t, h, w, 512*expansion]."""
        x = rnn.as_ndhwc(x)
        x = rnn.conv_bn_relu_pool(x, self.conv1, self.bn1, self.maxpool)
        x = self.layer1(x)
        x = self.layer2(x)
        x = self.layer3(x)
        return self.layer4(x)

    # ``fc`` may have been wrapped into Sequential(Linear, ReLU, fc) by the builder's mlp option (builder:45-48)
    feature_channels = property(lambda self: (self.fc[0] if isinstance(self.fc, nn.Sequential) else self.fc).in_features)

    # ---- reference-facing API (NCDHW fp32 in / out) ----------------------------------------------------------
    def get_feature(self, x):
        return rnn.ToNCDHW.apply(self.feature_ndhwc(x), self.feature_channels)

    def get_output_and_feature(self, x):
        feat = self.get_feature(x)
        pooled = self.avgpool(feat)
        return self.fc(pooled.view(pooled.size(0), -1)), feat

    def forward(self, x):
        return self.get_output_and_feature(x)[0]


def resnet10(**kw):
    return ResNet(BasicBlock, [1, 1, 1, 1], **kw)


def resnet18(**kw):
    return ResNet(BasicBlock, [2, 2, 2, 2], **kw)


def resnet34(**kw):
    return ResNet(BasicBlock, [3, 4, 6, 3], **kw)


def resnet50(**kw):
    return ResNet(Bottleneck, [3, 4, 6, 3], **kw)


def resnet101(**kw):
    return ResNet(Bottleneck, [3, 4, 23, 3], **kw)


def resnet152(**kw):
    return ResNet(Bottleneck, [3, 8, 36, 3], **kw)


def resnet200(**kw):
    return ResNet(Bottleneck, [3, 24, 36, 3], **kw)
