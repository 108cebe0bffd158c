"""Model facade — the drop-in boundary.  Mirror of
models/detection/yolox_extension/models/detector.py:18-91 (YoloXDetector): same constructor, same
`forward_backbone` / `forward_detect` / `forward` signatures and return values, same state_dict keys
(`backbone.*`, `fpn.*`, `yolox_head.*`)."""
from typing import Dict, Optional, Tuple, Union

import torch as th

from ...recurrent_backbone import build_recurrent_backbone
from .build import build_yolox_fpn, build_yolox_head


class YoloXDetector(th.nn.Module):
    """RNN-based MaxViT backbone (CUDA library) + YOLOX PAFPN/head."""

    def __init__(self, model_cfg, ssod: bool = False):
        super().__init__()
        backbone_cfg, fpn_cfg, head_cfg = model_cfg.backbone, model_cfg.fpn, model_cfg.head
        self.backbone = build_recurrent_backbone(backbone_cfg)
        in_channels = self.backbone.get_stage_dims(tuple(fpn_cfg.in_stages))
        self.fpn = build_yolox_fpn(fpn_cfg, in_channels=in_channels)
        strides = self.backbone.get_strides(tuple(fpn_cfg.in_stages))
        self.yolox_head = build_yolox_head(head_cfg, in_channels=in_channels, strides=strides, ssod=ssod)

    def forward_backbone(self, x: th.Tensor, previous_states=None, token_mask: Optional[th.Tensor] = None):
        """-> ({stage: [B,C,h,w]}, [(h,c)]*4)   (detector.py:35-53)"""
        return self.backbone(x, previous_states, token_mask)

    def forward_detect(self, backbone_features: Dict[int, th.Tensor], targets: Optional[th.Tensor] = None,
                       soft_targets: Optional[th.Tensor] = None) -> Tuple[th.Tensor, Union[Dict[str, th.Tensor], None]]:
        """-> (predictions [B,A,4+1+num_cls] decoded, losses dict | None)   (detector.py:55-77)"""
        feats = {k: v.float() for k, v in backbone_features.items()}  # neck/head run in fp32 (TF32 convs)
        fpn_features = self.fpn(feats)
        if self.training:
            assert targets is not None
            return self.yolox_head(fpn_features, targets, soft_targets)
        outputs, losses = self.yolox_head(fpn_features)
        assert losses is None
        return outputs, losses

    def forward(self, x: th.Tensor, previous_states=None, retrieve_detections: bool = True,
                targets: Optional[th.Tensor] = None):
        backbone_features, states = self.forward_backbone(x, previous_states)
        if not retrieve_detections:
            assert targets is None
            return None, None, states
        outputs, losses = self.forward_detect(backbone_features=backbone_features, targets=targets)
        return outputs, losses, states
