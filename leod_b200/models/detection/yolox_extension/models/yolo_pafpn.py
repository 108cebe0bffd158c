"""YOLOX PAFPN neck.  Mirror of models/detection/yolox_extension/models/yolo_pafpn.py:18-140
(same constructor keywords, same sub-module names, same forward data flow)."""
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from ...yolox.models.network_blocks import BaseConv, CSPLayer


class YOLOPAFPN(nn.Module):
    def __init__(self, depth: float = 1.0, in_stages: Tuple[int, ...] = (2, 3, 4),
                 in_channels: Tuple[int, ...] = (256, 512, 1024), depthwise: bool = False, act: str = 'silu',
                 compile_cfg: Optional[Dict] = None):
        super().__init__()
        assert len(in_stages) == len(in_channels) == 3
        if depthwise:
            raise NotImplementedError('depthwise PAFPN is not used by any shipped config')
        if compile_cfg is not None and compile_cfg.get('enable', False):
            raise NotImplementedError('torch.compile is not part of the B200-native path')
        self.in_features = tuple(in_stages)
        self.in_channels = tuple(in_channels)
        c0, c1, c2 = in_channels
        n = round(3 * depth)
        self.lateral_conv0 = BaseConv(c2, c1, 1, 1, act)
        self.C3_p4 = CSPLayer(2 * c1, c1, n, False, act=act)
        self.reduce_conv1 = BaseConv(c1, c0, 1, 1, act)
        self.C3_p3 = CSPLayer(2 * c0, c0, n, False, act=act)
        self.bu_conv2 = BaseConv(c0, c0, 3, 2, act)
        self.C3_n3 = CSPLayer(2 * c0, c1, n, False, act=act)
        self.bu_conv1 = BaseConv(c1, c1, 3, 2, act)
        self.C3_n4 = CSPLayer(2 * c1, c2, n, False, act=act)

    @staticmethod
    def upsample(x):
        return F.interpolate(x, scale_factor=2, mode='nearest-exact')

    def forward(self, input: Dict[int, torch.Tensor]):
        x2, x1, x0 = (input[f] for f in self.in_features)
        fpn_out0 = self.lateral_conv0(x0)
        f_out0 = self.C3_p4(torch.cat((self.upsample(fpn_out0), x1), 1))
        fpn_out1 = self.reduce_conv1(f_out0)
        pan_out2 = self.C3_p3(torch.cat((self.upsample(fpn_out1), x2), 1))
        pan_out1 = self.C3_n3(torch.cat((self.bu_conv2(pan_out2), fpn_out1), 1))
        pan_out0 = self.C3_n4(torch.cat((self.bu_conv1(pan_out1), fpn_out0), 1))
        return pan_out2, pan_out1, pan_out0
