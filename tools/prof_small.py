"""Profiling driver for the kernels outside the training step (ncu evidence): event binning, NMS post-processing,
pred2label, TTA merge, optimizer, augmentation, small uploads, tracker, evaluation — each at a production-like size, twice."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from leod_b200 import _lib  # noqa: E402
from leod_b200.data.utils.augmentor import AugmentationState, BatchSpatialAugmentor, ZoomState  # noqa: E402
from leod_b200.data.utils.representations import StackedHistogram  # noqa: E402
from leod_b200.models.detection.yolox.utils.boxes import postprocess_packed  # noqa: E402
from leod_b200.modules.utils.ssod import fused_adamw_ema, pred2label_packed, tta_merge_packed  # noqa: E402

dev = torch.device('cuda')
g = torch.Generator().manual_seed(0)
for rep in range(2):
    # event binning: 2 M events of a 50 ms Gen1 window
    n = 2_000_000
    sh = StackedHistogram(10, 240, 304)
    x = torch.randint(0, 304, (n,), generator=g).to(dev)
    y = torch.randint(0, 240, (n,), generator=g).to(dev)
    p = torch.randint(0, 2, (n,), generator=g).to(dev)
    t = torch.sort(torch.randint(0, 50000, (n,), generator=g))[0].to(dev)
    sh.construct(x, y, p, t)
    # post-processing of one sweep chunk: 672 view-frames x 1680 anchors, ~25 % positive logits
    B, A, K = 672, 1680, 2
    pred = torch.rand(B, A, 5 + K, generator=g)
    pred[..., :2] *= torch.tensor([304.0, 240.0])
    pred[..., 2:4] = pred[..., 2:4] * 80 + 8
    pred[..., 4:] = torch.sigmoid(torch.randn(B, A, 1 + K, generator=g) * 2 - 1.5)
    pred = pred.to(dev)
    dets, cnt = postprocess_packed(pred, K, conf_thre=0.01, nms_thre=0.45)
    labels, nl = pred2label_packed(dets, cnt, [0.6, 0.3], [0.6, 0.3], (240, 304))
    both = torch.cat((labels[:336], labels[336:]), 1).contiguous()
    tta_merge_packed(both, (nl[:336] + nl[336:]).to(torch.int32), 0.01, 0.45)
    # optimizer + EMA over the RVT-small flat buffer (9.9 M parameters)
    P = 9_900_000
    bufs = [torch.randn(P, device=dev) for _ in range(5)]
    fused_adamw_ema(bufs[0], bufs[1], bufs[2], bufs[3].abs_(), step=3, lr=2e-4, clip_value=1.0, ema=bufs[4], ema_alpha=0.999)
    # augmentation of a training batch
    ev = (torch.rand(21, 8, 20, 240, 304, device=dev) < 0.1).to(torch.uint8)
    aug = BatchSpatialAugmentor((240, 304), dict(prob_hflip=0.5, prob_tflip=0, rotate=dict(prob=0), zoom=dict(
        prob=0.8, zoom_in=dict(weight=8, factor=dict(min=1, max=1.5)), zoom_out=dict(weight=2, factor=dict(min=1, max=1.2)))), 8)
    states = [AugmentationState(apply_h_flip=b % 2 == 0,
                                zoom_in=ZoomState(active=b % 3 == 0, x0=10, y0=12, factor=1.3),
                                zoom_out=ZoomState(active=b % 3 == 1, x0=5, y0=7, factor=1.15)) for b in range(8)]
    aug(ev, None, states=states)
    _lib.upload_small(torch.randn(16, 8, 7), dev)
    # evaluation of a 3 000-frame validation buffer (about 4 gt boxes and 15 detections per frame)
    from leod_b200.utils.evaluation.prophesee.evaluator import FrameBoxes, coco_eval_device  # noqa: E402
    rng = np.random.default_rng(100 + rep)
    rec_dt = np.dtype([('t', '<i8'), ('x', '<f4'), ('y', '<f4'), ('w', '<f4'), ('h', '<f4'), ('class_id', '<u4'), ('class_confidence', '<f4')])

    def frame_records(n, t):
        r = np.zeros(n, dtype=rec_dt)
        r['t'] = t
        r['w'], r['h'] = rng.uniform(8, 120, n), rng.uniform(8, 100, n)
        r['x'], r['y'] = rng.uniform(0, 304 - r['w']), rng.uniform(0, 240 - r['h'])
        r['class_id'] = rng.integers(0, 2, n)
        r['class_confidence'] = rng.uniform(0.05, 1, n)
        return r
    gts = [frame_records(int(rng.integers(1, 8)), 600000 + 50000 * f) for f in range(3000)]
    dts = [np.concatenate((gt, frame_records(int(rng.poisson(11)), gt['t'][0]))) for gt in gts]
    coco_eval_device(FrameBoxes(gts, dev, False), FrameBoxes(dts, dev, True), 3000, 2, 'gen1', False)
    # tracker / short-track filter on 64 pseudo-labelled sequences
    from leod_b200.modules.tracking import track_filter_sequences  # noqa: E402
    rng = np.random.default_rng(rep)
    seqs = []
    for s in range(64):
        frames, rows = list(range(0, 300, 1)), []
        for f in frames:
            k = int(rng.integers(1, 6))
            xy = rng.uniform(0, 200, (k, 2))
            wh = rng.uniform(15, 60, (k, 2))
            r = np.concatenate((np.zeros((k, 1)), xy, wh, rng.integers(0, 2, (k, 1)), rng.uniform(0.3, 1, (k, 2))), 1).astype(np.float32)
            rows.append(torch.from_numpy(r))
        seqs.append((frames, rows))
    track_filter_sequences(seqs, [(240, 304)] * 64, min_track_len=6, track_method='forward or backward', inpaint=True, ignore_label=1024)
    torch.cuda.synchronize()
print('done')
