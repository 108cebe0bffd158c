"""Shape/config record for the oracle (values: config/model/maxvit_yolox/default.yaml:1-64,
config/experiment/gen{1,4}/{tiny,small,base}.yaml, config/modifier.py:49-64)."""
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple


@dataclass
class ModelCfg:
    input_channels: int = 20
    embed_dim: int = 48                       # 32 tiny / 48 small / 64 base
    dim_multiplier: Tuple[int, ...] = (1, 2, 4, 8)
    dim_head: int = 24                        # 32 tiny,base / 24 small
    partition_size: Tuple[int, int] = (8, 10)  # (8,10) gen1 / (6,10) gen4
    mlp_ratio: int = 4
    norm_eps: float = 1e-5
    num_classes: int = 2
    fpn_depth: float = 0.33
    in_stages: Tuple[int, ...] = (2, 3, 4)
    ignore_label: int = 1024
    ignore_bbox_thresh: Optional[Sequence[float]] = None
    reg_weight: float = 5.0
    obj_weight: float = 1.0
    cls_weight: float = 1.0

    @property
    def stage_dims(self):
        return tuple(self.embed_dim * m for m in self.dim_multiplier)

    @property
    def strides(self):
        return (4, 8, 16, 32)

    @staticmethod
    def named(size: str, dataset: str = 'gen1', **kw):
        embed, dh, depth = {'tiny': (32, 32, 0.33), 'small': (48, 24, 0.33), 'base': (64, 32, 0.67)}[size]
        part, ncls = {'gen1': ((8, 10), 2), 'gen4': ((6, 10), 3)}[dataset]
        base = dict(embed_dim=embed, dim_head=dh, fpn_depth=depth, partition_size=part, num_classes=ncls)
        base.update(kw)
        return ModelCfg(**base)
