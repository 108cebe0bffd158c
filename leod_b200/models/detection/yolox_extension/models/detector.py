"""Model facade — the drop-in boundary.  Mirror of
models/detection/yolox_extension/models/detector.py:18-91 (YoloXDetector): same constructor, same
`forward_backbone` / `forward_detect` / `forward` signatures and return values, same state_dict keys
(`backbone.*`, `fpn.*`, `yolox_head.*`).  Both halves run in the CUDA library: the recurrent backbone through
`leod_backbone_*`, the PAFPN neck + YOLOX head + SimOTA loss through `leod_fpn_head_*` / `leod_simota_loss_*`."""
from typing import Dict, Optional, Tuple, Union

import torch as th

from ...recurrent_backbone import build_recurrent_backbone
from .build import build_yolox_fpn, build_yolox_head
from .detect_engine import DetectEngine


class YoloXDetector(th.nn.Module):
    """RNN-based MaxViT backbone + YOLOX PAFPN/head, all hand-written CUDA behind the reference interface."""

    def __init__(self, model_cfg, ssod: bool = False):
        super().__init__()
        backbone_cfg, fpn_cfg, head_cfg = model_cfg.backbone, model_cfg.fpn, model_cfg.head
        self.backbone = build_recurrent_backbone(backbone_cfg)
        in_stages = tuple(fpn_cfg.in_stages)
        in_channels = self.backbone.get_stage_dims(in_stages)
        self.fpn = build_yolox_fpn(fpn_cfg, in_channels=in_channels)
        strides = self.backbone.get_strides(in_stages)
        self.yolox_head = build_yolox_head(head_cfg, in_channels=in_channels, strides=strides, ssod=ssod)
        # not a sub-module: its parameters / buffers are registered INSIDE self.fpn / self.yolox_head under the reference's names
        object.__setattr__(self, 'detect_engine', DetectEngine(self.fpn, self.yolox_head, in_channels, strides, self.backbone.in_res_hw,
                                                               self.backbone.compute_dtype))

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        self.detect_engine.reflatten()
        return self

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        out = super().load_state_dict(state_dict, strict=strict, assign=assign)
        if assign:
            self.detect_engine.reflatten()
        self.detect_engine.mark_params_updated()
        return out

    def forward_backbone(self, x: th.Tensor, previous_states=None, token_mask: Optional[th.Tensor] = None):
        """-> ({stage: [B,C,h,w]}, [(h,c)]*4)   (detector.py:35-53)"""
        return self.backbone(x, previous_states, token_mask)

    def forward_detect(self, backbone_features: Dict[int, th.Tensor], targets: Optional[th.Tensor] = None,
                       soft_targets: Optional[th.Tensor] = None) -> Tuple[th.Tensor, Union[Dict[str, th.Tensor], None]]:
        """-> (predictions [B,A,4+1+num_cls] decoded, losses dict | None)   (detector.py:55-77)"""
        if soft_targets is not None:
            raise NotImplementedError('soft_targets are unused by the reference (yolo_head.py:196 asserts pred_probs is None)')
        feats = [backbone_features[k] for k in self.fpn.in_features]
        if self.training:
            assert targets is not None
            return self.detect_engine.detect(feats, targets, training=True)
        return self.detect_engine.detect(feats, None, training=False)

    def forward(self, x: th.Tensor, previous_states=None, retrieve_detections: bool = True,
                targets: Optional[th.Tensor] = None):
        backbone_features, states = self.forward_backbone(x, previous_states)
        if not retrieve_detections:
            assert targets is None
            return None, None, states
        outputs, losses = self.forward_detect(backbone_features=backbone_features, targets=targets)
        return outputs, losses, states
