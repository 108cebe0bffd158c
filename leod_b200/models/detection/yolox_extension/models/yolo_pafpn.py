"""YOLOX PAFPN neck.  Same constructor keywords and sub-module / parameter names as
models/detection/yolox_extension/models/yolo_pafpn.py:18-140; the forward data flow of :109-140 (top-down + bottom-up
PAN) is executed by the CUDA library together with the head (`YoloXDetector.forward_detect` -> DetectEngine ->
leod_fpn_head_fwd/bwd), so this module holds configuration and parameters only."""
from typing import Dict, Optional, Tuple

import torch.nn as nn


class YOLOPAFPN(nn.Module):
    def __init__(self, depth: float = 1.0, in_stages: Tuple[int, ...] = (2, 3, 4),
                 in_channels: Tuple[int, ...] = (256, 512, 1024), depthwise: bool = False, act: str = 'silu',
                 compile_cfg: Optional[Dict] = None):
        super().__init__()
        assert len(in_stages) == len(in_channels) == 3, 'Current implementation only for 3 feature maps'
        if depthwise:
            raise NotImplementedError('depthwise PAFPN is not used by any shipped config')
        if act != 'silu':
            raise NotImplementedError(act)
        if compile_cfg is not None and compile_cfg.get('enable', False):
            raise NotImplementedError('torch.compile is not part of the B200-native path')
        self.in_features = tuple(in_stages)
        self.in_channels = tuple(in_channels)
        self.depth = depth
        self.n_bottleneck = round(3 * depth)     # yolo_pafpn.py:55

    def forward(self, input):
        raise RuntimeError('the PAFPN runs fused with the head inside the CUDA library: call YoloXDetector.forward_detect')
