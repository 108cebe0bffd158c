"""Diagnostics: where does the bf16 neck/head error come from (adversarial O(1) weights vs reference-like init)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
from test_gpu_detect import CASES, build_detector, rand_feats, make_labels
from oracle import yolox as oy

for case in ('fixture', 'small_gen1'):
    for fill in (True, False):
        cfg, hw, B = CASES[case]
        torch.manual_seed(0)
        m = build_detector(cfg, hw, 'bf16', seed_fill=fill).train()
        sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items() if not k.startswith('backbone')}
        feats = {k: v.bfloat16().float() for k, v in rand_feats(cfg, hw, B, 2).items()}
        labels = make_labels(B, 6, hw[0], hw[1], cfg.num_classes, seed=3)
        out, losses = m.forward_detect({k: v.cuda() for k, v in feats.items()}, targets=labels.cuda())
        with torch.no_grad():
            ref, rl = oy.detect_forward(feats, sd, cfg, targets=labels, training=True)
        out = out.cpu()
        d = (out - ref).abs()
        print(f'--- {case} fill={fill}: loss {float(losses["loss"]):.4f} ref {float(rl["loss"]):.4f}')
        for c, n in enumerate(['cx', 'cy', 'w', 'h', 'obj', 'cls0', 'cls1']):
            print(f'  {n}: max|ref| {float(ref[..., c].abs().max()):.3e}  max err {float(d[..., c].max()):.3e}  mean err {float(d[..., c].mean()):.3e}  '
                  f'rel {float(d[..., c].max() / ref[..., c].abs().max()):.2e}')
        lw = (out[..., 2] / ref[..., 2]).log().abs()
        print(f'  log(w/w_ref): max {float(lw.max()):.3e} mean {float(lw.mean()):.3e}')
