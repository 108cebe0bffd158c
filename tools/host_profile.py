"""Where the host time of a training step goes: cProfile over 10 steps of the bench workload (no sync inside the loop)."""
import cProfile, os, pstats, sys, io
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from leod_b200.config import Node, make_model_cfg
from leod_b200.modules.detection import FlatOptimizer, Module
wl = bench.WORKLOADS['train']
dev = torch.device('cuda')
cfg = Node(model=make_model_cfg(size=wl['size'], dataset=wl['dataset'], compute_dtype='bf16'), dataset=dict(sequence_length=wl['L'], name=wl['dataset']))
module = Module(cfg).to(dev).train()
opt = FlatOptimizer(module.mdl, lr=2e-4, weight_decay=0.0, clip_value=1.0)
ev, boxes, first = bench.synth_batch(wl, 0)
batch = bench.make_batch(wl, ev.to(dev), boxes, first)
def step():
    opt.zero_grad()
    out = module.training_step(batch)
    out['loss'].backward()
    opt.step()
for _ in range(4): step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step()
pr.disable()
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f'host time per step (under cProfile): {(t1 - t0) * 100:.2f} ms')
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45); print(s.getvalue()[:9000])
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(25); print(s.getvalue()[:6000])
