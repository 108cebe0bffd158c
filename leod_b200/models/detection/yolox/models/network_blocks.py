"""Conv-BN-SiLU building blocks of the YOLOX neck/head.

Mirror of models/detection/yolox/models/network_blocks.py:29-54 (BaseConv), :79-101 (Bottleneck),
:104-142 (CSPLayer) with identical sub-module names, so state_dict keys match the reference.
Depthwise variants are not built (depthwise: False in every shipped config).
"""
import torch
import torch.nn as nn


class BaseConv(nn.Module):
    """conv (bias-free, 'same' padding) -> BatchNorm2d -> SiLU."""

    def __init__(self, in_channels, out_channels, ksize, stride, act='silu'):
        super().__init__()
        if act != 'silu':
            raise NotImplementedError(act)
        self.conv = nn.Conv2d(in_channels, out_channels, ksize, stride, (ksize - 1) // 2, bias=False)
        self.bn = nn.BatchNorm2d(out_channels)
        self.act = nn.SiLU(inplace=True)

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))


class Bottleneck(nn.Module):
    def __init__(self, in_channels, out_channels, shortcut=True, expansion=0.5, act='silu'):
        super().__init__()
        hidden = int(out_channels * expansion)
        self.conv1 = BaseConv(in_channels, hidden, 1, 1, act)
        self.conv2 = BaseConv(hidden, out_channels, 3, 1, act)
        self.use_add = shortcut and in_channels == out_channels

    def forward(self, x):
        y = self.conv2(self.conv1(x))
        return y + x if self.use_add else y


class CSPLayer(nn.Module):
    def __init__(self, in_channels, out_channels, n=1, shortcut=True, expansion=0.5, act='silu'):
        super().__init__()
        hidden = int(out_channels * expansion)
        self.conv1 = BaseConv(in_channels, hidden, 1, 1, act)
        self.conv2 = BaseConv(in_channels, hidden, 1, 1, act)
        self.conv3 = BaseConv(2 * hidden, out_channels, 1, 1, act)
        self.m = nn.Sequential(*[Bottleneck(hidden, hidden, shortcut, 1.0, act) for _ in range(n)])

    def forward(self, x):
        return self.conv3(torch.cat((self.m(self.conv1(x)), self.conv2(x)), dim=1))
