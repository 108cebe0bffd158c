"""Diagnostic (not a pytest file): per-parameter gradient error of the CUDA backbone against the CPU
oracle for T = 1, 2, 3 timesteps, fp32 and bf16."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from helpers import load_net_fixture, rel_err
from test_host_cpu import product_cfg
from oracle import rvt, yolox
from leod_b200.models.detection.yolox_extension.models.detector import YoloXDetector

z, cfg, sd, d = load_net_fixture()
x = torch.from_numpy(z['x']).float()
labels = torch.from_numpy(z['train_plain/labels'])
for T in (1, 2, 3):
    psd = {k: v.clone().requires_grad_(v.is_floating_point() and 'running' not in k) for k, v in sd.items()}
    states = None
    for t in range(T):
        feats, states = rvt.backbone_forward(x[t], states, psd, cfg)
    _, ol = yolox.detect_forward(feats, psd, cfg, targets=labels, training=True)
    ol['loss'].backward()
    for dtype in ('fp32', 'bf16'):
        m = YoloXDetector(product_cfg(cfg, (d['H'], d['W']), compute_dtype=dtype))
        m.load_state_dict(sd)
        m.cuda().train()
        st = None
        for t in range(T):
            f, st = m.forward_backbone(x[t].cuda(), st)
        _, losses = m.forward_detect(f, targets=labels.cuda())
        losses['loss'].backward()
        torch.cuda.synchronize()
        print(f'=== T={T} {dtype}: loss {float(losses["loss"]):.5f} vs {float(ol["loss"]):.5f}')
        worst = []
        for k, p in m.named_parameters():
            if not k.startswith('backbone'):
                continue
            e = rel_err(p.grad.float().cpu(), psd[k].grad)
            worst.append((e, k))
        worst.sort(reverse=True)
        bad = [w for w in worst if w[0] > (1e-3 if dtype == 'fp32' else 5e-2)]
        print(f'   {len(bad)} / {len(worst)} parameters above tolerance')
        for e, k in worst[:12]:
            print(f'   {e:.3e}  {k}')

print('######## smooth loss, T=3')
g = torch.Generator().manual_seed(0)
psd = {k: v.clone().requires_grad_(k.startswith('backbone')) for k, v in sd.items()}
states = None
for t in range(3):
    feats, states = rvt.backbone_forward(x[t], states, psd, cfg)
R = [(torch.randn(h.shape, generator=g), torch.randn(c.shape, generator=g)) for h, c in states]
sum((h * rh).sum() + (c * rc).sum() for (h, c), (rh, rc) in zip(states, R)).backward()
for dtype in ('fp32', 'bf16'):
    m = YoloXDetector(product_cfg(cfg, (d['H'], d['W']), compute_dtype=dtype))
    m.load_state_dict(sd)
    m.cuda().train()
    st = None
    for t in range(3):
        f, st = m.forward_backbone(x[t].cuda(), st)
    sum((h.float() * rh.cuda()).sum() + (c.float() * rc.cuda()).sum() for (h, c), (rh, rc) in zip(st, R)).backward()
    torch.cuda.synchronize()
    worst = sorted(((rel_err(p.grad.float().cpu(), psd[k].grad), k) for k, p in m.named_parameters() if k.startswith('backbone')), reverse=True)
    print(f'=== smooth {dtype}: feature err', [round(rel_err(st[s][0].float().cpu(), states[s][0].detach()), 5) for s in range(4)])
    for e, k in worst[:8]:
        print(f'   {e:.3e}  {k}')
    print('   median', worst[len(worst) // 2][0])
