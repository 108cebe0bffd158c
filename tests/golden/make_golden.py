"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE (/root/reference).

Run in the build container only (the reference tree is absent on the GPU box):
    python tests/golden/make_golden.py
Everything written here is small (< 2 MB total) and committed; tests/test_oracle_golden.py pins
oracle/ against it, and the -m gpu tests pin the CUDA path against the same files.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '_refshim'))
import refimport  # noqa: E402

refimport.setup()

from omegaconf import DictConfig  # noqa: E402  (the stand-in)
from models.detection.yolox_extension.models.detector import YoloXDetector  # noqa: E402
from models.detection.yolox.utils.boxes import postprocess  # noqa: E402
import torchvision  # noqa: E402


def model_cfg(embed, dh, part, ncls, depth, inch, ignore_thresh=None):
    return DictConfig(dict(
        backbone=dict(name='MaxViTRNN', compile=dict(enable=False, args=dict(mode='reduce-overhead')),
                      input_channels=inch, enable_masking=False, partition_split_32=1, embed_dim=embed,
                      dim_multiplier=[1, 2, 4, 8], num_blocks=[1, 1, 1, 1], T_max_chrono_init=[4, 8, 16, 32],
                      stem=dict(patch_size=4),
                      stage=dict(downsample=dict(type='patch', overlap=True, norm_affine=True),
                                 attention=dict(use_torch_mha=False, partition_size=part, dim_head=dh,
                                                attention_bias=True, mlp_activation='gelu', mlp_gated=False,
                                                mlp_bias=True, mlp_ratio=4, drop_mlp=0, drop_path=0,
                                                ls_init_value=1e-5),
                                 lstm=dict(dws_conv=False, dws_conv_only_hidden=True, dws_conv_kernel_size=3,
                                           drop_cell_update=0))),
        fpn=dict(name='PAFPN', compile=dict(enable=False, args={}), depth=depth, in_stages=[2, 3, 4],
                 depthwise=False, act='silu'),
        head=dict(name='YoloX', compile=dict(enable=False, args={}), depthwise=False, act='silu',
                  num_classes=ncls, obj_focal_loss=False, bbox_loss_weighting='',
                  ignore_bbox_thresh=ignore_thresh, ignore_label=1024, ignore_bg_k=0),
        postprocess=dict(confidence_threshold=0.1, nms_threshold=0.45)))


def randomize(model, g):
    """Reference init leaves LayerScale at 1e-5 and BN at identity, which would hide errors in
    those paths: overwrite every parameter/buffer with O(1) seeded values."""
    with torch.no_grad():
        for k, v in model.state_dict().items():
            if k.endswith('num_batches_tracked'):
                continue
            if k.endswith('running_var'):
                v.copy_(torch.rand(v.shape, generator=g) + 0.5)
            elif k.endswith('gamma'):
                v.copy_(torch.rand(v.shape, generator=g) * 0.5 + 0.25)
            elif k.endswith('.norm.weight') or k.endswith('norm1.weight') or k.endswith('norm2.weight') \
                    or k.endswith('bn.weight'):
                v.copy_(torch.rand(v.shape, generator=g) + 0.5)
            elif v.dim() >= 2:
                fan_in = v[0].numel()
                v.copy_(torch.randn(v.shape, generator=g) * (1.0 / fan_in ** 0.5))
            else:
                v.copy_(torch.randn(v.shape, generator=g) * 0.2)


def make_labels(g, B, N, ncls, hw, with_ignore):
    """[B,N,7] = (cls, cx, cy, w, h, obj_conf, cls_conf), zero padded at the end."""
    H, W = hw
    lab = torch.zeros(B, N, 7)
    for b in range(B):
        n = int(torch.randint(1, N + 1, (1,), generator=g))
        if with_ignore and b == 1:
            n = 0                              # an empty image
        for i in range(n):
            w = float(torch.rand(1, generator=g)) * 0.4 * W + 6
            h = float(torch.rand(1, generator=g)) * 0.4 * H + 6
            cx = float(torch.rand(1, generator=g)) * (W - w) + w / 2
            cy = float(torch.rand(1, generator=g)) * (H - h) + h / 2
            cls = int(torch.randint(0, ncls, (1,), generator=g))
            lab[b, i] = torch.tensor([cls, cx, cy, w, h, float(torch.rand(1, generator=g)) * 0.7 + 0.3,
                                      float(torch.rand(1, generator=g)) * 0.7 + 0.3])
        if with_ignore and n > 0:
            if b == 2:
                lab[b, :n, 0] = 1024           # an all-ignore image
            elif n > 1:
                lab[b, 0, 0] = 1024            # mixed
    return lab


def gen_net():
    g = torch.Generator().manual_seed(1234)
    embed, dh, part, ncls, depth, inch = 8, 4, (2, 3), 2, 0.33, 6
    H, W, B, T = 64, 96, 4, 3
    out = dict(meta=np.array([embed, dh, part[0], part[1], ncls, inch, H, W, B, T], np.int64),
               fpn_depth=np.float32(depth))
    mcfg = model_cfg(embed, dh, part, ncls, depth, inch)
    ref = YoloXDetector(mcfg)
    randomize(ref, g)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    for k, v in sd.items():
        out['sd/' + k] = v.numpy()
    x = (torch.rand(T, B, inch, H, W, generator=g) < 0.15).float() * torch.randint(1, 12, (T, B, inch, H, W), generator=g)
    out['x'] = x.numpy().astype(np.uint8)

    # --- eval: backbone over T steps, head on the last, postprocess
    ref.eval()
    states = None
    with torch.no_grad():
        for t in range(T):
            feats, states = ref.forward_backbone(x[t], states)
            for s in (1, 2, 3, 4):
                out[f'eval/feat{s}_t{t}'] = feats[s].numpy()
        for s in range(4):
            out[f'eval/c{s}'] = states[s][1].numpy()
        preds, losses = ref.forward_detect(feats)
        assert losses is None
        out['eval/preds'] = preds.numpy()
        dets = postprocess(preds.clone(), ncls, 0.001, 0.45)
        for b, d in enumerate(dets):
            out[f'eval/det{b}'] = (d if d is not None else torch.zeros(0, 7)).numpy()

    # --- train: fresh instances (decode-grid cache), plain labels and ignore-aware labels
    for tag, with_ignore, thr in (('plain', False, None), ('ignore', True, None), ('thresh', False, [0.7, 0.35])):
        m = YoloXDetector(model_cfg(embed, dh, part, ncls, depth, inch, ignore_thresh=thr))
        m.load_state_dict(sd)
        m.train()
        labels = make_labels(g, B, 5, ncls, (H, W), with_ignore)
        out[f'train_{tag}/labels'] = labels.numpy()
        states = None
        for t in range(T):
            feats, states = m.forward_backbone(x[t], states)
        preds, losses = m.forward_detect(feats, targets=labels.clone())
        for k in ('loss', 'iou_loss', 'conf_loss', 'cls_loss', 'num_fg'):
            out[f'train_{tag}/{k}'] = np.float32(float(losses[k]))
        out[f'train_{tag}/preds'] = preds.detach().numpy()
        losses['loss'].backward()
        grads = dict(m.named_parameters())
        for k in ('backbone.stages.0.downsample_cf2cl.conv.weight',
                  'backbone.stages.0.att_blocks.0.att_window.self_attn.qkv.weight',
                  'backbone.stages.0.att_blocks.0.att_grid.ls1.gamma',
                  'backbone.stages.1.att_blocks.0.att_grid.norm1.weight',
                  'backbone.stages.2.att_blocks.0.att_window.mlp.net.0.0.bias',
                  'backbone.stages.3.lstm.conv1x1.weight',
                  'backbone.stages.0.lstm.conv1x1.bias',
                  'fpn.C3_p4.m.0.conv2.conv.weight', 'fpn.bu_conv1.bn.weight',
                  'yolox_head.stems.0.conv.weight', 'yolox_head.obj_preds.1.bias',
                  'yolox_head.cls_preds.2.weight'):
            out[f'train_{tag}/grad/{k}'] = grads[k].grad.numpy()
        if tag == 'plain':
            for k, v in m.state_dict().items():
                if k.endswith('running_mean') or k.endswith('running_var'):
                    if k.startswith('fpn.lateral_conv0') or k.startswith('yolox_head.stems.2'):
                        out[f'train_plain/bn/{k}'] = v.numpy()
    np.savez_compressed(os.path.join(HERE, 'net_small.npz'), **out)
    print('net_small.npz', sum(v.nbytes for v in out.values()) / 1e6, 'MB raw')


def gen_nms():
    g = torch.Generator().manual_seed(7)
    out = {}
    cases = []
    for n, ncls, tie in ((0, 2, False), (1, 2, False), (50, 2, False), (300, 3, False), (64, 1, True),
                         (200, 2, True), (129, 3, True), (1000, 2, False)):
        xy = torch.rand(n, 2, generator=g) * 250
        wh = torch.rand(n, 2, generator=g) * 80 + 2
        boxes = torch.cat((xy, xy + wh), 1)
        scores = torch.rand(n, generator=g)
        cls = torch.randint(0, ncls, (n,), generator=g).float()
        if tie and n > 0:
            scores = (scores * 8).round() / 8          # many equal scores
            boxes[n // 2:] = boxes[:n - n // 2].clone()         # duplicated boxes
            boxes = (boxes * 2).round() / 2
        cases.append((boxes, scores, cls))
    # a constructed threshold-boundary case: IoU exactly 0.5 with thr 0.5 must NOT suppress
    cases.append((torch.tensor([[0., 0., 10., 10.], [0., 0., 10., 5.], [0., 5., 10., 10.]]),
                  torch.tensor([0.9, 0.8, 0.7]), torch.zeros(3)))
    for i, (boxes, scores, cls) in enumerate(cases):
        thr = 0.5 if i == len(cases) - 1 else 0.45
        if boxes.shape[0]:
            # torchvision's sort is not guaranteed stable: feed it strictly distinct keys that
            # preserve the stable order, so the fixture encodes "stable descending score order".
            order = torch.sort(scores, descending=True, stable=True).indices
            rank = torch.empty_like(order)
            rank[order] = torch.arange(len(order))
            keys = -rank.float()
            keep = torchvision.ops.batched_nms(boxes, keys, cls, thr)
        else:
            keep = torch.zeros(0, dtype=torch.long)
        out[f'{i}/boxes'], out[f'{i}/scores'], out[f'{i}/cls'] = boxes.numpy(), scores.numpy(), cls.numpy()
        out[f'{i}/thr'] = np.float32(thr)
        out[f'{i}/keep'] = keep.numpy()
    out['n'] = np.int64(len(cases))
    np.savez_compressed(os.path.join(HERE, 'nms_cases.npz'), **out)
    print('nms_cases.npz', len(cases), 'cases')


def gen_pred2label():
    from modules.utils.ssod import pred2label, filter_pred_boxes
    from modules.pseudo_labeler import tta_postprocess
    from data.genx_utils.labels import ObjectLabels
    from functools import partial
    g = torch.Generator().manual_seed(11)
    out = {}
    for ci, (dataset, ds2, hw, ncls, othr, cthr) in enumerate((('gen1', False, (240, 304), 2, [0.6, 0.3], [0.6, 0.3]),
                                                              ('gen4', True, (360, 640), 3, [0.5, 0.5, 0.4], 0.45))):
        dets = []
        for b in range(4):
            n = [0, 7, 40, 25][b]
            xy = torch.rand(n, 2, generator=g) * torch.tensor([hw[1] * 1.2, hw[0] * 1.2]) - 20
            wh = torch.rand(n, 2, generator=g) * torch.tensor([hw[1] * 1.0, hw[0] * 0.5]) + 1
            d = torch.cat((xy, xy + wh, torch.rand(n, 2, generator=g),
                           torch.randint(0, ncls, (n, 1), generator=g).float()), 1)
            dets.append(d)
            out[f'{ci}/det{b}'] = d.numpy()
        fn = partial(filter_pred_boxes, dataset_name=dataset, downsampled_by_2=ds2)
        labs = pred2label([d.clone() for d in dets], obj_thresh=othr, cls_thresh=cthr, filter_bbox_fn=fn, hw=hw)
        for b, l in enumerate(labs):
            out[f'{ci}/label{b}'] = l.object_labels.numpy()
        out[f'{ci}/hw'] = np.array(hw, np.int64)
        out[f'{ci}/obj_thresh'] = np.array(othr, np.float32)
        out[f'{ci}/cls_thresh'] = np.array(cthr if isinstance(cthr, list) else [cthr] * ncls, np.float32)
        # TTA merge: concatenate the labels of two "views" of one frame and NMS again
        merged_in = torch.cat([labs[2].object_labels, labs[3].object_labels, labs[2].object_labels * 1.0], 0)
        res = tta_postprocess([ObjectLabels(merged_in.clone(), hw)], conf_thre=0.1, nms_thre=0.45)
        out[f'{ci}/tta_in'] = merged_in.numpy()
        out[f'{ci}/tta_out'] = res[0].object_labels.numpy()
    out['n'] = np.int64(2)
    np.savez_compressed(os.path.join(HERE, 'pred2label_cases.npz'), **out)
    print('pred2label_cases.npz')


def gen_binning():
    from data.utils.representations import StackedHistogram
    g = torch.Generator().manual_seed(5)
    out = {}
    cases = ((0, 10, 24, 32, None, True), (5000, 10, 24, 32, None, True), (20000, 5, 16, 20, 10, True),
             (30000, 10, 12, 16, None, False), (1, 3, 8, 8, None, True), (70000, 2, 4, 4, None, True))
    for i, (n, bins, H, W, cutoff, fast) in enumerate(cases):
        x = torch.randint(0, W, (n,), generator=g)
        y = torch.randint(0, H, (n,), generator=g)
        p = torch.randint(0, 2, (n,), generator=g)
        t = torch.sort(torch.randint(1_000_000, 1_050_000, (n,), generator=g)).values
        if n > 1000:   # hot pixel cluster to hit the wrap / clamp
            x[: n // 3] = 3
            y[: n // 3] = 2
        rep = StackedHistogram(bins=bins, height=H, width=W, count_cutoff=cutoff, fastmode=fast).construct(x, y, p, t)
        out[f'{i}/x'], out[f'{i}/y'], out[f'{i}/p'], out[f'{i}/t'] = (a.numpy().astype(np.int32) for a in (x, y, p, t))
        out[f'{i}/cfg'] = np.array([bins, H, W, -1 if cutoff is None else cutoff, int(fast)], np.int64)
        out[f'{i}/rep'] = rep.numpy()
    out['n'] = np.int64(len(cases))
    np.savez_compressed(os.path.join(HERE, 'binning_cases.npz'), **out)
    print('binning_cases.npz')


def gen_optim():
    from modules.utils.ssod import ema_model_update
    g = torch.Generator().manual_seed(3)
    out = {}
    student = torch.nn.Sequential(torch.nn.Linear(13, 7), torch.nn.BatchNorm1d(7), torch.nn.Linear(7, 3))
    teacher = torch.nn.Sequential(torch.nn.Linear(13, 7), torch.nn.BatchNorm1d(7), torch.nn.Linear(7, 3))
    with torch.no_grad():
        for p in list(student.parameters()) + list(teacher.parameters()):
            p.copy_(torch.randn(p.shape, generator=g))
    for i, p in enumerate(student.parameters()):
        out[f'ema/student{i}'] = p.detach().numpy().copy()
    for i, p in enumerate(teacher.parameters()):
        out[f'ema/teacher{i}'] = p.detach().numpy().copy()
    for step in (0, 5, 5000):
        t2 = [p.detach().clone() for p in teacher.parameters()]

        class _M:  # minimal .parameters() holder
            def __init__(self, ps): self.ps = ps
            def parameters(self): return self.ps
        ema_model_update(student, _M(t2), global_step=step, alpha=0.999)
        for i, p in enumerate(t2):
            out[f'ema/step{step}/teacher{i}'] = p.numpy()
    # AdamW with clip-by-value 1.0 (train.py:236-237), 3 steps
    p = torch.nn.Parameter(torch.randn(257, generator=g))
    out['adamw/p0'] = p.detach().numpy().copy()
    opt = torch.optim.AdamW([p], lr=2e-4, weight_decay=0)
    for s in range(3):
        grad = torch.randn(257, generator=g) * 2
        out[f'adamw/g{s}'] = grad.numpy().copy()
        p.grad = grad.clone()
        torch.nn.utils.clip_grad_value_([p], 1.0)
        opt.step()
        out[f'adamw/p{s + 1}'] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, 'optim_cases.npz'), **out)
    print('optim_cases.npz')


def synth_track_sequence(seed, n_frames, hw, n_obj, ncls, miss_p=0.2, fp_rate=0.6, gt_frames=(), drop_frames=()):
    """Synthetic per-frame detections of a sequence, as ObjectLabels rows (t, x, y, w, h, cls, cls_conf, obj_conf),
    corner format: objects on straight tracks (some leaving the image), random misses, short-lived false positives."""
    rng = np.random.default_rng(seed)
    H, W = hw
    objs = []
    for _ in range(n_obj):
        w, h = rng.uniform(15, 70), rng.uniform(12, 50)
        objs.append(dict(x=rng.uniform(-20, W), y=rng.uniform(0, H - h), w=w, h=h, vx=rng.uniform(-6, 6), vy=rng.uniform(-2, 2),
                         cls=int(rng.integers(0, ncls)), t0=int(rng.integers(0, n_frames // 2)), t1=int(rng.integers(n_frames // 2, n_frames + 1))))
    frames, idx = [], []
    for f in range(n_frames):
        if f in drop_frames:
            continue
        rows = []
        gt = f in gt_frames
        for o in objs:
            if not (o['t0'] <= f < o['t1']) or (rng.random() < miss_p and not gt):
                continue
            x = o['x'] + o['vx'] * (f - o['t0']) + rng.normal(0, 0.7)
            y = o['y'] + o['vy'] * (f - o['t0']) + rng.normal(0, 0.7)
            x1, y1 = np.clip(x, 0, W - 1), np.clip(y, 0, H - 1)
            x2, y2 = np.clip(x + o['w'], 0, W - 1), np.clip(y + o['h'], 0, H - 1)
            if x2 - x1 < 5 or y2 - y1 < 5:
                continue
            rows.append([1.0 if gt else 0.0, x1, y1, x2 - x1, y2 - y1, o['cls'], 1.0 if gt else rng.uniform(0.4, 1), 1.0 if gt else rng.uniform(0.4, 1)])
        if not gt:
            for _ in range(rng.poisson(fp_rate)):
                w, h = rng.uniform(8, 40), rng.uniform(8, 40)
                rows.append([0.0, rng.uniform(0, W - w - 1), rng.uniform(0, H - h - 1), w, h, int(rng.integers(0, ncls)), rng.uniform(0.3, 0.7), rng.uniform(0.3, 0.7)])
        if rows:
            frames.append(np.asarray(rows, np.float32))
            idx.append(f)
    return frames, idx


def gen_tracking():
    """modules/pseudo_labeler.py:201-333 (EventSeqData._track / _track_filter) with modules/tracking/linear.py run on
    synthetic sequences: which boxes become `ignore` (class 1024), which boxes are in-painted, the final label lists."""
    from modules.pseudo_labeler import EventSeqData
    from data.genx_utils.labels import ObjectLabels
    cases = [dict(seed=1, n_frames=40, hw=(240, 304), n_obj=4, ncls=2, gt_frames=(), drop_frames=()),
             dict(seed=2, n_frames=60, hw=(240, 304), n_obj=7, ncls=2, gt_frames=(20, 40), drop_frames=(5, 6, 7, 30)),
             dict(seed=3, n_frames=50, hw=(360, 640), n_obj=10, ncls=3, miss_p=0.35, fp_rate=1.5, gt_frames=(25,), drop_frames=()),
             dict(seed=4, n_frames=12, hw=(240, 304), n_obj=2, ncls=2, miss_p=0.0, fp_rate=0.0),
             dict(seed=5, n_frames=30, hw=(240, 304), n_obj=3, ncls=2, miss_p=0.5, fp_rate=0.2)]
    out = {'n': np.int64(len(cases))}
    for ci, c in enumerate(cases):
        hw = c['hw']
        frames, idx = synth_track_sequence(**c)
        out[f'{ci}/hw'] = np.array(hw, np.int64)
        out[f'{ci}/frame_idx'] = np.array(idx, np.int64)
        out[f'{ci}/counts'] = np.array([len(f) for f in frames], np.int64)
        out[f'{ci}/rows'] = np.concatenate(frames, 0)
        for inpaint in (False, True):
            labels = [ObjectLabels(torch.from_numpy(f.copy()), hw) for f in frames]
            remove_idx, inpainted = EventSeqData._track(labels, list(idx), min_track_len=6, inpaint=inpaint)
            out[f'{ci}/remove_idx_inpaint{int(inpaint)}'] = np.array(sorted(remove_idx), np.int64)
            if inpaint:
                keys = sorted(inpainted.keys())
                out[f'{ci}/inpaint_frames'] = np.array(keys, np.int64)
                out[f'{ci}/inpaint_counts'] = np.array([len(inpainted[k]) for k in keys], np.int64)
                out[f'{ci}/inpaint_rows'] = np.concatenate([inpainted[k] for k in keys], 0) if keys else np.zeros((0, 8), np.float32)
        # the whole post-processing: forward or backward tracking, ignore labels, in-painting
        for method in ('forward', 'forward or backward'):
            seq = EventSeqData.__new__(EventSeqData)
            seq.filter_config = DictConfig(dict(min_track_len=6, track_method=method, inpaint=True, ignore_label=1024))
            seq.labels = [ObjectLabels(torch.from_numpy(f.copy()), hw) for f in frames]
            seq.frame_idx = list(idx)
            seq._track_filter()
            tag = 'fb' if 'backward' in method else 'f'
            out[f'{ci}/final_{tag}_frame_idx'] = np.array(seq.frame_idx, np.int64)
            out[f'{ci}/final_{tag}_counts'] = np.array([len(l) for l in seq.labels], np.int64)
            rows = []
            for l in seq.labels:
                l.numpy_()
                rows.append(np.asarray(l.object_labels, np.float32))
            out[f'{ci}/final_{tag}_rows'] = np.concatenate(rows, 0)
            if tag == 'fb':    # the on-disk records (pseudo_labeler.py:179-199 _summarize, labels.py:12-16 BBOX_DTYPE)
                packed, lbl_idx, repr_idx = seq._summarize()
                out[f'{ci}/packed_bytes'] = np.frombuffer(packed.tobytes(), np.uint8)
                out[f'{ci}/objframe_idx_2_label_idx'] = lbl_idx
                out[f'{ci}/objframe_idx_2_repr_idx'] = repr_idx
        print('tracking case', ci, 'boxes', int(out[f'{ci}/counts'].sum()), 'removed', len(out[f'{ci}/remove_idx_inpaint1']),
              'inpainted', int(out[f'{ci}/inpaint_counts'].sum()))
    np.savez_compressed(os.path.join(HERE, 'tracking_cases.npz'), **out)


def gen_fullsize(only=None):
    """Full-size runs of the REFERENCE itself (BASELINE configs[0]: RVT-tiny, 10 input channels, 1 x 240 x 304, one
    frame; and the configs[1] model RVT-small at batch 1, two frames) with name-seeded weights (tests/helpers.py:
    det_state_value), so only the small outputs are stored: stage-4 feature, final cell state of stage 4, decoded
    predictions and the post-processed detections."""
    sys.path.insert(0, os.path.dirname(HERE))
    from helpers import FULLSIZE_CASES, det_events, det_fill_
    sizes = {'tiny': (32, 32, 0.33), 'small': (48, 24, 0.33), 'base': (64, 32, 0.67)}
    dsets = {'gen1': ((8, 10), 2, (256, 320), (240, 304)), 'gen4': ((6, 10), 3, (384, 640), (360, 640))}
    out = {}
    path = os.path.join(HERE, 'fullsize_cases.npz')
    if only is not None and os.path.exists(path):      # add cases without touching the committed ones
        out = dict(np.load(path))
    for tag, (size, dataset, inch, B, L) in FULLSIZE_CASES.items():
        if only is not None and tag not in only:
            continue
        embed, dh, depth = sizes[size]
        part, ncls, in_res, frame = dsets[dataset]
        ref = YoloXDetector(model_cfg(embed, dh, part, ncls, depth, inch))
        det_fill_(ref)
        ref.eval()
        x = det_events(7, (L, B, inch, frame[0], frame[1]))
        xp = torch.zeros(L, B, inch, in_res[0], in_res[1])
        xp[..., :frame[0], :frame[1]] = x.float()          # utils/padding.py:33-58
        states = None
        with torch.no_grad():
            for t in range(L):
                feats, states = ref.forward_backbone(xp[t], states)
            preds, _ = ref.forward_detect(feats)
            dets = postprocess(preds.clone(), ncls, 0.001, 0.45)
        out[f'{tag}/feat4'] = feats[4].numpy()
        out[f'{tag}/c4'] = states[3][1].numpy()
        out[f'{tag}/preds'] = preds.numpy()
        for b in range(B):
            d = dets[b]
            out[f'{tag}/det{b}'] = (d if d is not None else torch.zeros(0, 7)).numpy()
        print(tag, 'preds', tuple(preds.shape), 'dets', [0 if d is None else len(d) for d in dets])
    np.savez_compressed(os.path.join(HERE, 'fullsize_cases.npz'), **out)


def gen_augment():
    """data/utils/augmentor.py run as a dataloader worker runs it (one RandomSpatialAugmentorGenX per sample, seeded torch RNG):
    per case the sampled state, the augmented event frames and the augmented label rows.  Also the time flip of
    sequence_base.py:208-225.  Inputs are regenerated by the tests from the stored seeds (tests/helpers.py: augment_inputs)."""
    import zlib
    from data.utils.augmentor import RandomSpatialAugmentorGenX
    from data.utils.types import DataType
    from data.genx_utils.labels import ObjectLabels, SparselyBatchedObjectLabels
    import types
    for name in ('torchdata.datapipes', 'torchdata.datapipes.map'):   # torchdata 0.11 dropped datapipes; only the base class name is needed
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['torchdata.datapipes.map'].MapDataPipe = object
    from data.genx_utils.sequence_base import SequenceBase
    sys.path[:0] = [os.path.join(HERE, '..'), os.path.join(HERE, '..', '..')]
    from helpers import AUGMENT_CASES, AUGM_CFG, augment_inputs
    out = {}
    for ci, (H, W, L, C, seed, tflip, store) in enumerate(AUGMENT_CASES):
        ev, label_rows = augment_inputs(H, W, L, C, seed)
        aug = RandomSpatialAugmentorGenX(dataset_hw=(H, W), automatic_randomization=False, augm_config=DictConfig(AUGM_CFG))
        torch.manual_seed(seed)
        aug.randomize_augmentation()
        aug.augm_state.apply_t_flip = False       # consumed by the dataset (dataset_rnd.py:102-113); the case's `tflip` plays that role
        data = {DataType.EV_REPR: [torch.from_numpy(e.copy()) for e in ev],
                DataType.OBJLABELS_SEQ: SparselyBatchedObjectLabels(
                    [None if r is None else ObjectLabels(torch.from_numpy(r.copy()), (H, W)) for r in label_rows]),
                DataType.IS_FIRST_SAMPLE: True, DataType.IS_PADDED_MASK: [False] * L, DataType.EV_IDX: list(range(L)),
                DataType.IS_REVERSED: bool(tflip)}
        if tflip:
            data = SequenceBase.time_flip_data(data)
        data = aug(data)
        st = data[DataType.AUGM_STATE].to_dict()
        out[f'{ci}/state'] = np.array([st['h_flip']['active'], st['zoom_in']['active'], st['zoom_in']['x0'], st['zoom_in']['y0'],
                                       st['zoom_out']['active'], st['zoom_out']['x0'], st['zoom_out']['y0']], np.int64)
        out[f'{ci}/factors'] = np.array([st['zoom_in']['factor'], st['zoom_out']['factor']], np.float64)
        res = torch.stack(data[DataType.EV_REPR]).numpy()
        out[f'{ci}/crc'] = np.int64(zlib.crc32(res.tobytes()))
        out[f'{ci}/nnz'] = np.int64((res != 0).sum())
        if store:
            out[f'{ci}/ev'] = res
        for t, lab in enumerate(data[DataType.OBJLABELS_SEQ]):
            if lab is not None:
                out[f'{ci}/label{t}'] = lab.object_labels.numpy()
        print(ci, (H, W), 'tflip', tflip, st)
    out['n'] = np.int64(len(AUGMENT_CASES))
    np.savez_compressed(os.path.join(HERE, 'augment_cases.npz'), **out)
    print('augment_cases.npz')


def gen_eval():
    """The reference's own half of the evaluation (everything before pycocotools): filter_boxes (io/box_filtering.py:18-36) with the
    thresholds of evaluation.py:24-33, _match_times and _to_coco_format (metrics/coco_eval.py:65-94, 140-194) on seeded per-frame
    buffers built like to_prophesee builds them (io/box_loading.py:58-107).  pycocotools is absent here: COCOeval itself is not run."""
    import importlib.util
    import types
    from utils.evaluation.prophesee.io.box_filtering import filter_boxes
    from data.genx_utils.labels import BBOX_DTYPE
    for name in ('pycocotools', 'pycocotools.coco', 'pycocotools.cocoeval'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['pycocotools.coco'].COCO = None
    sys.modules['pycocotools.cocoeval'].COCOeval = None
    torch.cuda.get_device_name = lambda *a, **k: 'cpu'          # metrics/coco_eval.py:17 queries it at import time
    spec = importlib.util.spec_from_file_location('ref_coco_eval', os.path.join(refimport.REF, 'utils/evaluation/prophesee/metrics/coco_eval.py'))
    ce = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ce)
    sys.path[:0] = [os.path.join(HERE, '..'), os.path.join(HERE, '..', '..')]
    from helpers import EVAL_CASES, eval_inputs
    out = {}
    for ci, (camera, ds2, F, seed) in enumerate(EVAL_CASES):
        gts, dts = eval_inputs(camera, ds2, F, seed)

        def rec(d, with_score):
            r = np.zeros(len(d['cls']), dtype=BBOX_DTYPE)
            r['t'], r['x'], r['y'], r['w'], r['h'] = d['t'], d['xywh'][:, 0], d['xywh'][:, 1], d['xywh'][:, 2], d['xywh'][:, 3]
            r['class_id'] = d['cls']
            r['class_confidence'] = d['score'] if with_score else 1.0
            return r
        diag, side = (60, 20) if camera == 'gen4' else (30, 10)          # evaluation.py:24-33
        if ds2:
            diag, side = diag // 2, side // 2
        fl = lambda x: filter_boxes(x, int(5e5), diag, side)  # noqa: E731
        flat_gt, flat_dt, frames = [], [], []
        for f, (g, d) in enumerate(zip(gts, dts)):                          # coco_eval.py:47-60
            g, d = fl(rec(g, False)), fl(rec(d, True))
            all_ts = np.unique(g['t'])
            gw, dw = ce._match_times(all_ts, g, d, 50000)
            flat_gt += gw
            flat_dt += dw
            frames += [f] * len(gw)
        cats = [{'id': i + 1, 'name': str(i), 'supercategory': 'none'} for i in range(3 if camera == 'gen4' else 2)]
        dataset, results = ce._to_coco_format(flat_gt, flat_dt, cats)
        out[f'{ci}/frames'] = np.array(frames, np.int64)
        out[f'{ci}/ann'] = np.array([[a['image_id'], a['category_id'], *a['bbox'], a['area']] for a in dataset['annotations']],
                                    np.float64).reshape(-1, 7)
        out[f'{ci}/res'] = np.array([[r['image_id'], r['category_id'], *r['bbox'], r['score']] for r in results], np.float64).reshape(-1, 7)
        print(ci, camera, ds2, 'images', len(frames), 'of', F, 'ann', len(dataset['annotations']), 'res', len(results))
    out['n'] = np.int64(len(EVAL_CASES))
    np.savez_compressed(os.path.join(HERE, 'eval_cases.npz'), **out)
    print('eval_cases.npz')


if __name__ == '__main__':
    torch.set_num_threads(4)
    gen_net()
    gen_nms()
    gen_pred2label()
    gen_binning()
    gen_optim()
    gen_tracking()
    gen_fullsize()
    gen_augment()
    gen_eval()
