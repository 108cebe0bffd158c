"""YOLOX decoupled head with SimOTA loss.  Same constructor keywords, sub-module / parameter names and bias initialisation
as models/detection/yolox/models/yolo_head.py:21-193.  forward (:195-287), decode (:310-332) and the losses (:403-597,
:776-1148, plain and ignore-label variants) are executed by the CUDA library (kernels_simota.cu) behind
`YoloXDetector.forward_detect`; this module holds configuration and parameters only.

Options left off by every shipped config (obj_focal_loss, bbox_loss_weighting, ignore_bg_k, use_l1, depthwise) raise
NotImplementedError.
"""
from typing import Dict, Optional

import torch.nn as nn


class YOLOXHead(nn.Module):
    def __init__(self, num_classes=80, strides=(8, 16, 32), in_channels=(256, 512, 1024), act='silu', depthwise=False,
                 compile_cfg: Optional[Dict] = None, obj_focal_loss=False, bbox_loss_weighting='', ignore_bg_k=-1,
                 reg_weight=5.0, obj_weight=1.0, cls_weight=1.0, ignore_bbox_thresh=None, ignore_label=1024):
        super().__init__()
        if depthwise or obj_focal_loss or bbox_loss_weighting or (ignore_bg_k is not None and ignore_bg_k > 0):
            raise NotImplementedError('depthwise / focal / bbox_loss_weighting / ignore_bg_k are off in every shipped config')
        if act != 'silu':
            raise NotImplementedError(act)
        if compile_cfg is not None and compile_cfg.get('enable', False):
            raise NotImplementedError('torch.compile is not part of the B200-native path')
        self.num_classes = num_classes
        self.decode_in_inference = True
        self.strides = tuple(strides)
        self.in_channels = tuple(in_channels)
        self.hidden_dim = int(256 * in_channels[-1] / 1024)     # yolo_head.py:61-66
        self.reg_weight, self.obj_weight, self.cls_weight = reg_weight, obj_weight, cls_weight
        self.ignore_bbox_thresh = list(ignore_bbox_thresh) if ignore_bbox_thresh else None
        self.ignore_label = ignore_label
        self.use_l1 = False

    def forward(self, xin, labels=None, pred_probs=None):
        raise RuntimeError('the head runs fused with the PAFPN inside the CUDA library: call YoloXDetector.forward_detect')
