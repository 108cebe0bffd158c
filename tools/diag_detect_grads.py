"""Diagnostics: bf16 vs fp32 CUDA-path parameter gradients of the neck/head, per tensor (same weights, same inputs)."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import torch
from test_gpu_detect import CASES, build_detector, rand_feats, make_labels

case, fill = sys.argv[1], sys.argv[2] == 'o1'
cfg, hw, B = CASES[case]
res = {}
for dt in ('fp32', 'bf16'):
    torch.manual_seed(11)
    m = build_detector(cfg, hw, dt, seed_fill=fill).train()
    feats = {k: v.bfloat16().float().cuda().requires_grad_(True) for k, v in rand_feats(cfg, hw, B, 2).items()}
    labels = make_labels(B, 6, hw[0], hw[1], cfg.num_classes, seed=3).cuda()
    out, losses = m.forward_detect(feats, targets=labels)
    a, _ = m.detect_engine.last_assignment(B)
    losses['loss'].backward()
    torch.cuda.synchronize()
    res[dt] = ({n: p.grad.clone() for n, p in m.named_parameters() if not n.startswith('backbone')}, a.clone(), float(losses['loss']),
               {k: v.grad.float().clone() for k, v in feats.items()})
print('loss', res['fp32'][2], res['bf16'][2], 'assignment flips', int((res['fp32'][1] != res['bf16'][1]).sum()))
rows = []
for n, g in res['fp32'][0].items():
    b = res['bf16'][0][n]
    e = float((g - b).abs().max() / (g.abs().max() + 1e-20))
    rows.append((e, n, float(g.abs().max()), tuple(g.shape)))
for e, n, mx, shp in sorted(rows, reverse=True)[:25]:
    print(f'{e:9.3e}  {n:45s} max|g| {mx:.3e} {shp}')
for k in res['fp32'][3]:
    g, b = res['fp32'][3][k], res['bf16'][3][k]
    print('feat', k, float((g - b).abs().max() / g.abs().max()))
n = 'fpn.lateral_conv0.conv.weight'
g, b = res['fp32'][0][n][:, :, 0, 0], res['bf16'][0][n][:, :, 0, 0]
d = (g - b).abs()
print('lateral_conv0 err by row block (64):', [float(d[i:i + 64].max()) for i in range(0, g.shape[0], 64)])
print('lateral_conv0 err by col block (64):', [float(d[:, i:i + 64].max()) for i in range(0, g.shape[1], 64)])
for n in ['yolox_head.reg_preds.2.bias', 'yolox_head.obj_preds.2.bias', 'yolox_head.cls_preds.2.bias', 'yolox_head.reg_preds.0.bias', 'yolox_head.obj_preds.0.bias',
          'yolox_head.cls_preds.0.bias']:
    print(n, res['fp32'][0][n].tolist(), res['bf16'][0][n].tolist())
n = 'yolox_head.reg_preds.2.weight'
print(n, res['fp32'][0][n].flatten()[:8].tolist(), res['bf16'][0][n].flatten()[:8].tolist())
n = 'yolox_head.obj_preds.2.weight'
print(n, res['fp32'][0][n].flatten()[:8].tolist(), res['bf16'][0][n].flatten()[:8].tolist())
