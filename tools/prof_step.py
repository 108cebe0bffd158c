"""Profiling driver: ONE full training step of the bench workload (RVT-small Gen1, B=8, L=21, labels on 2 frames: backbone window
forward, neck + head + SimOTA loss, backward, AdamW) with no warm-up, so that `ncu -k regex:... ` sees exactly one step."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from leod_b200.config import Node, make_model_cfg  # noqa: E402
from leod_b200.modules.detection import FlatOptimizer, Module  # noqa: E402

wl = bench.WORKLOADS['train'] if hasattr(bench, 'WORKLOADS') else None
assert wl is not None
torch.manual_seed(0)
dev = torch.device('cuda')
cfg = Node(model=make_model_cfg(size=wl['size'], dataset=wl['dataset'], compute_dtype='bf16'), dataset=dict(sequence_length=wl['L'], name=wl['dataset']))
module = Module(cfg).to(dev).train()
opt = FlatOptimizer(module.mdl, lr=2e-4, weight_decay=0.0, clip_value=1.0)
ev, boxes, first = bench.synth_batch(wl, 0)
batch = bench.make_batch(wl, ev.to(dev), boxes, first.to(dev))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for _ in range(steps):
    opt.zero_grad()
    out = module.training_step(batch)
    out['loss'].backward()
    opt.step()
torch.cuda.synchronize()
print('done, loss', float(out['loss']))
