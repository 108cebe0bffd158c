"""Shared helpers for the test-suite (fixtures loading, synthetic inputs)."""
import os

import numpy as np
import torch

from oracle.config import ModelCfg

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_net_fixture():
    z = np.load(os.path.join(GOLDEN, 'net_small.npz'))
    embed, dh, p0, p1, ncls, inch, H, W, B, T = (int(v) for v in z['meta'])
    cfg = ModelCfg(input_channels=inch, embed_dim=embed, dim_head=dh, partition_size=(p0, p1), num_classes=ncls,
                   fpn_depth=float(z['fpn_depth']))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('sd/')}
    return z, cfg, sd, dict(H=H, W=W, B=B, T=T)


def rel_err(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def det_state_value(name: str, shape, dtype=torch.float32) -> torch.Tensor:
    """Deterministic O(1) value for a state_dict entry, seeded by its NAME: the reference model (make_golden.py) and the
    models under test are filled with identical weights without storing multi-megabyte state dicts."""
    import zlib
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7fffffff)
    shape = tuple(shape)
    if name.endswith('num_batches_tracked'):
        return torch.zeros(shape, dtype=torch.long)
    if name.endswith('running_var'):
        return torch.rand(shape, generator=g) + 0.5
    if name.endswith('gamma'):
        return torch.rand(shape, generator=g) * 0.5 + 0.25
    if name.endswith('.norm.weight') or name.endswith('norm1.weight') or name.endswith('norm2.weight') or name.endswith('bn.weight'):
        return torch.rand(shape, generator=g) + 0.5
    if len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return torch.randn(shape, generator=g) * (1.0 / fan_in ** 0.5)
    return torch.randn(shape, generator=g) * 0.2


def det_fill_(model) -> None:
    with torch.no_grad():
        for k, v in model.state_dict().items():
            v.copy_(det_state_value(k, v.shape).to(v.dtype))


def det_events(seed: int, shape, density=0.1, vmax=6) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return ((torch.rand(shape, generator=g) < density).float() * torch.randint(1, vmax, shape, generator=g)).to(torch.uint8)


FULLSIZE_CASES = {   # tag -> (size, dataset, input_channels, B, L): BASELINE configs[0] (plumbing case) and the configs[1] model
    'tiny_gen1_c10': ('tiny', 'gen1', 10, 1, 1),
    'small_gen1': ('small', 'gen1', 20, 1, 2),
    'base_gen4': ('base', 'gen4', 20, 1, 1),          # BASELINE configs[2] model: partition (6, 10), dim_head 32, 3 classes, 384x640
    'small_gen1_l8': ('small', 'gen1', 20, 1, 8),      # eight recurrent steps
}


def canon_rows(d):
    """Detections [N,7] in a canonical order: descending score (obj * cls_conf), ties broken by the remaining columns."""
    d = np.asarray(d, np.float32)
    if d.shape[0] == 0:
        return d
    score = d[:, 4] * d[:, 5]
    order = np.lexsort((d[:, 3], d[:, 2], d[:, 1], d[:, 0], d[:, 6], -score))
    return d[order]


# ---- augmentation cases (tests/golden/augment_cases.npz): (H, W, L, C, seed, time flip, store the full output in the fixture)
AUGM_CFG = dict(prob_hflip=0.5, prob_tflip=0.0, rotate=dict(prob=0, min_angle_deg=2, max_angle_deg=6),
                zoom=dict(prob=0.8, zoom_in=dict(weight=8, factor=dict(min=1, max=1.5)),
                          zoom_out=dict(weight=2, factor=dict(min=1, max=1.2))))     # config/dataset/base.yaml:19-39 (random sampling)
AUGMENT_CASES = [(48, 64, 3, 4, s, s % 3 == 0, True) for s in range(1, 11)] + \
                [(45, 50, 2, 3, s, s % 2 == 0, True) for s in range(11, 17)] + \
                [(240, 304, 2, 20, s, s == 22, False) for s in range(20, 26)] + \
                [(360, 640, 2, 20, s, False, False) for s in range(30, 34)]


def augment_inputs(H, W, L, C, seed):
    """-> (ev uint8 [L, C, H, W] numpy, L-list of label rows fp32 [n, 8] or None); the last frame always has boxes."""
    rng = np.random.default_rng(1000 + seed)
    ev = ((rng.random((L, C, H, W)) < 0.15) * rng.integers(1, 256, (L, C, H, W))).astype(np.uint8)
    labels = []
    for t in range(L):
        if t < L - 1 and rng.random() < 0.4:
            labels.append(None)
            continue
        n = int(rng.integers(1, 6))
        w = rng.uniform(3, W * 0.5, n)
        h = rng.uniform(3, H * 0.5, n)
        x = rng.uniform(0, W - 1 - w)
        y = rng.uniform(0, H - 1 - h)
        rows = np.stack((np.full(n, 1000.0 + t), x, y, w, h, rng.integers(0, 2, n).astype(np.float64), np.ones(n), np.ones(n)), 1)
        labels.append(rows.astype(np.float32))
    return ev, labels


# ---- evaluation cases (tests/golden/eval_cases.npz): (camera, downsampled_by_2, frames, seed)
EVAL_CASES = [('gen1', False, 24, 1), ('gen4', True, 30, 2), ('gen4', False, 12, 3)]


def eval_inputs(camera, ds2, n_frames, seed, max_gt=8, det_noise=6.0, fp_rate=3.0, miss_p=0.15):
    """Per-frame ground truth and detections as the evaluator buffers hold them: dict(t int64, xywh fp32 [n,4], cls int64, score fp32).
    Detections = jittered copies of most gt boxes + false positives; some boxes violate the Prophesee size / time filters, some frames
    lose all their gt to the filter, some have no detection."""
    rng = np.random.default_rng(7000 + seed)
    H, W = ((360, 640) if ds2 else (720, 1280)) if camera == 'gen4' else (240, 304)
    K = 3 if camera == 'gen4' else 2
    gts, dts = [], []
    for f in range(n_frames):
        t = np.int64(100000 * (f + 1) + (f % 3) * 150000)          # the first frames lie before 0.5 s
        n = int(rng.integers(0 if f % 7 == 3 else 1, max_gt + 1))
        w = rng.uniform(4, W * 0.4, n).astype(np.float32)
        h = rng.uniform(4, H * 0.4, n).astype(np.float32)
        if f % 5 == 4:
            w[:], h[:] = 6, 6                                       # every gt box too small: the frame is no image
        x = rng.uniform(0, W - w).astype(np.float32)
        y = rng.uniform(0, H - h).astype(np.float32)
        cls = rng.integers(0, K, n)
        gts.append(dict(t=np.full(n, t), xywh=np.stack((x, y, w, h), 1).reshape(-1, 4).astype(np.float32), cls=cls.astype(np.int64)))
        keep = rng.random(n) > miss_p
        jit = rng.normal(0, det_noise, (int(keep.sum()), 4)).astype(np.float32)
        dx = gts[-1]['xywh'][keep] + jit
        dx[:, 2:] = np.maximum(dx[:, 2:], 2)
        dc = cls[keep].copy()
        flip = rng.random(len(dc)) < 0.1
        dc[flip] = (dc[flip] + 1) % K
        nf = int(rng.poisson(fp_rate)) if f % 6 != 5 else 0
        fw = rng.uniform(4, W * 0.3, nf).astype(np.float32)
        fh = rng.uniform(4, H * 0.3, nf).astype(np.float32)
        fx = np.stack((rng.uniform(0, W - fw), rng.uniform(0, H - fh), fw, fh), 1).astype(np.float32).reshape(-1, 4)
        xywh = np.concatenate((dx.reshape(-1, 4), fx), 0).astype(np.float32)
        c = np.concatenate((dc, rng.integers(0, K, nf))).astype(np.int64)
        score = np.round(rng.uniform(0.05, 1.0, len(c)), 2).astype(np.float32)      # rounded: ties between detections occur
        if f % 6 == 5:
            xywh, c, score = xywh[:0], c[:0], score[:0]
        p = rng.permutation(len(c))
        dts.append(dict(t=np.full(len(c), t), xywh=xywh[p], cls=c[p], score=score[p]))
    return gts, dts
