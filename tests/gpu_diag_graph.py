"""Diagnostic (not a test): find the op that breaks CUDA-graph capture of the neck+head+loss backward."""
import logging
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from leod_b200.config import make_model_cfg
from leod_b200.models.detection.yolox_extension.models.detector import YoloXDetector

log = logging.getLogger('torch.autograd.graph')
log.setLevel(logging.DEBUG)
h = logging.StreamHandler(sys.stdout)
log.addHandler(h)

torch.manual_seed(0)
m = YoloXDetector(make_model_cfg(size='small', dataset='gen1')).cuda().train()
B = 4
shapes = m.backbone._state_shapes(B)
feats = [torch.randn(shapes[s], device='cuda').requires_grad_(True) for s in (1, 2, 3)]
labels = torch.zeros(B, 16, 7, device='cuda')
labels[:, 0] = torch.tensor([0, 100., 100., 40., 30., 1., 1.], device='cuda')
labels[:, 1] = torch.tensor([1, 200., 150., 60., 50., 1., 1.], device='cuda')
params = [p for mm in (m.fpn, m.yolox_head) for p in mm.parameters()]
log.setLevel(logging.WARNING)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        _, losses = m._detect_eager(feats, labels)
        torch.autograd.grad(losses['loss'], feats + params, allow_unused=True)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
print('warm-up ok, loss', float(losses['loss']))
for what in ('fwd', 'fwd+bwd'):
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            _, losses = m._detect_eager(feats, labels)
            if what == 'fwd+bwd':
                log.setLevel(logging.DEBUG)
                grads = torch.autograd.grad(losses['loss'], feats + params, allow_unused=True)
                log.setLevel(logging.WARNING)
        g.replay()
        torch.cuda.synchronize()
        print(what, 'capture OK, loss', float(losses['loss']))
    except Exception as e:  # noqa: BLE001
        log.setLevel(logging.WARNING)
        print(what, 'capture FAILED:', str(e).split('\n')[0])
        break
