"""Time the neck + head + loss forward/backward alone (RVT-S Gen1 shapes) with CUDA events; under
`ncu --profile-from-start off` the last iteration is the profiled one."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import torch
from leod_b200.config import make_model_cfg
from leod_b200.models.detection.yolox_extension.models.detector import YoloXDetector

size, ds = (sys.argv[2], sys.argv[3]) if len(sys.argv) > 3 else ('small', 'gen1')
Bs = [int(v) for v in sys.argv[1].split(',')] if len(sys.argv) > 1 else [16]
m = YoloXDetector(make_model_cfg(size=size, dataset=ds, compute_dtype='bf16')).cuda().train()
H, W = m.backbone.in_res_hw
dims = m.backbone.get_stage_dims((2, 3, 4))
for B in Bs:
    g = torch.Generator(device='cuda').manual_seed(0)
    feats = {s: torch.randn(B, c, H // st, W // st, device='cuda', generator=g).to(torch.bfloat16).permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2).requires_grad_(True)
             for s, c, st in zip((2, 3, 4), dims, (8, 16, 32))}
    lab = torch.zeros(B, 16, 7, device='cuda')
    for b in range(B):
        for i in range(1 + b % 8):
            lab[b, i] = torch.tensor([i % 2, 40. + 30 * i, 50. + 20 * i, 30. + 8 * i, 24. + 6 * i, 1., 1.])
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    N = 20
    for it in range(N + 5):
        prof = it == N + 4
        if prof:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        ev[0].record()
        preds, losses = m.forward_detect(feats, targets=lab)
        ev[1].record()
        losses['loss'].backward()
        ev[2].record()
        torch.cuda.synchronize()
        if prof:
            torch.cuda.profiler.stop()
        if 5 <= it < N + 4:
            tf += ev[0].elapsed_time(ev[1]) / (N - 1)
            tb += ev[1].elapsed_time(ev[2]) / (N - 1)
    print(f'{size}/{ds} B\'={B}: neck+head+loss fwd {tf:.3f} ms, bwd {tb:.3f} ms, loss {float(losses["loss"]):.4f}')
    # same work replayed from a CUDA graph: device time without host launch overhead
    for p in m.parameters():
        p.grad = None
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    feats = {k: v.detach().clone().requires_grad_(True) for k, v in feats.items()}    # fresh leaves: no AccumulateGrad node on the default stream
    fl = [feats[s] for s in (2, 3, 4)]
    with torch.cuda.stream(side):
        for _ in range(3):
            _, ls = m.forward_detect(feats, targets=lab)
            torch.autograd.grad(ls['loss'], fl)
    torch.cuda.current_stream().wait_stream(side)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        _, ls = m.forward_detect(feats, targets=lab)
        gg = torch.autograd.grad(ls['loss'], fl)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gr.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f'   graph replay fwd+bwd: {e0.elapsed_time(e1) / 20:.3f} ms, loss {float(ls["loss"]):.4f}')
