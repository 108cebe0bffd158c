"""Mirror of models/detection/recurrent_backbone/__init__.py:6-11."""
from .maxvit_rnn import RNNDetector as MaxViTRNNDetector


def build_recurrent_backbone(backbone_cfg):
    name = backbone_cfg.name
    if name == 'MaxViTRNN':
        return MaxViTRNNDetector(backbone_cfg)
    raise NotImplementedError(name)
