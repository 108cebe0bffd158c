"""Mirror of models/detection/yolox_extension/models/build.py:9-30 (same factory signatures)."""
from typing import Tuple

from .yolo_pafpn import YOLOPAFPN
from ...yolox.models.yolo_head import YOLOXHead


def _to_dict(cfg):
    try:
        from omegaconf import OmegaConf
        if OmegaConf.is_config(cfg):
            return OmegaConf.to_container(cfg, resolve=True, throw_on_missing=True)
    except ImportError:
        pass
    return {k: (dict(v) if isinstance(v, dict) else v) for k, v in dict(cfg).items()}


def build_yolox_head(head_cfg, in_channels: Tuple[int, ...], strides: Tuple[int, ...], ssod: bool = False):
    assert not ssod  # same contract as the reference (build.py:10)
    d = _to_dict(head_cfg)
    d.pop('name')
    d.pop('version', None)
    d.update(in_channels=in_channels, strides=strides, compile_cfg=d.pop('compile', None))
    return YOLOXHead(**d)


def build_yolox_fpn(fpn_cfg, in_channels: Tuple[int, ...]):
    d = _to_dict(fpn_cfg)
    name = d.pop('name')
    if name not in {'PAFPN', 'pafpn'}:
        raise NotImplementedError(name)
    d.update(in_channels=in_channels, compile_cfg=d.pop('compile', None))
    return YOLOPAFPN(**d)
