"""Profiling driver: one forward_sequence (+ optional backward) of RVT-small Gen1 at the bench shape (B=8, L=21),
no warm-up, so kernel launch indices are predictable under `ncu -k regex:... -c N`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from leod_b200.config import make_model_cfg
from leod_b200.models.detection.yolox_extension.models.detector import YoloXDetector

bwd = len(sys.argv) > 1 and sys.argv[1] == 'bwd'
torch.manual_seed(0)
m = YoloXDetector(make_model_cfg(size='small', dataset='gen1')).cuda().train()
x = (torch.rand(21, 8, 20, 240, 304, device='cuda') < 0.1).to(torch.uint8) * 2
with torch.set_grad_enabled(bwd):
    feats, states = m.backbone.forward_sequence(x, None)
    if bwd:
        sum(f.float().mean() for f in feats.values()).backward()
torch.cuda.synchronize()
print('done')
