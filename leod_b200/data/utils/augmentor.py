"""Batch augmentation on the device: the reference augments every sample on a dataloader worker (CPU tensors,
data/utils/augmentor.py:125-478, one RandomSpatialAugmentorGenX per worker/sequence); here the host only SAMPLES the
per-sequence augmentation state (same torch RNG calls in the same order as randomize_augmentation / _zoom_in_and_rescale, so a
seeded run draws the same flips, factors and windows) and the pixels / boxes of the whole [L, B] batch are transformed by two
kernels (leod_augment_ev_repr, leod_augment_labels; bit-identical results, see tests/test_gpu_augment.py).

Rotation has probability 0 in every shipped config (config/dataset/base.yaml:24-27) and raises NotImplementedError.
There is no CPU fallback: the transforms need the CUDA library.
"""
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch as th

from leod_b200 import _lib
from leod_b200.data.labels import ObjectLabels, SparselyBatchedObjectLabels


def _get(cfg, name, default=None):
    if isinstance(cfg, dict):
        return cfg.get(name, default)
    return getattr(cfg, name, default)


def _uniform(lo: float, hi: float) -> float:
    """utils/helpers.py:14-18 (one th.rand draw unless the interval is empty)."""
    assert hi >= lo
    return lo if hi == lo else lo + (hi - lo) * th.rand(1).item()


@dataclass
class ZoomState:
    active: bool = False
    x0: int = 0
    y0: int = 0
    factor: float = 1.0


@dataclass
class AugmentationState:
    """augmentor.py:71-87 (same meaning; `apply_t_flip` is consumed by the dataset, `is_reversed` is what the batch carries)."""
    apply_h_flip: bool = False
    apply_t_flip: bool = False
    zoom_in: ZoomState = field(default_factory=ZoomState)
    zoom_out: ZoomState = field(default_factory=ZoomState)

    def to_dict(self):
        z = lambda s: {'active': s.active, 'x0': s.x0, 'y0': s.y0, 'factor': s.factor}  # noqa: E731
        return {'h_flip': {'active': self.apply_h_flip}, 'rotation': {'active': False, 'angle_deg': 0.0},
                'zoom_in': z(self.zoom_in), 'zoom_out': z(self.zoom_out)}


def zoom_window_for_label(xywh: Sequence[float], hw: Tuple[int, int], win_hw: Tuple[int, int]) -> Tuple[int, int]:
    """augmentor.py:521-562: a top-left corner such that the zoom window contains the whole label rectangle."""
    H, W = hw
    wh, ww = win_hw
    x0, y0, w, h = (float(v) for v in xywh)
    x1, y1 = x0 + w, y0 + h
    assert x0 >= 0 and y0 >= 0 and w > 0 and h > 0 and x1 <= W + 1e-2 - 1 and y1 <= H + 1e-2 - 1
    xa = max(x1 - max(ww, w), 0)
    ya = max(y1 - max(wh, h), 0)
    xb = max(min(x0 + max(ww, w), W - 1) - ww, xa)
    yb = max(min(y0 + max(wh, h), H - 1) - wh, ya)
    xs = int(_uniform(xa, xb))
    ys = int(_uniform(ya, yb))
    assert 0 <= xs < W and 0 <= ys < H
    return xs, ys


class RandomSpatialAugmentorGenX:
    """State sampler of ONE sequence (augmentor.py:125-208): same constructor arguments, same RNG consumption."""

    def __init__(self, dataset_hw: Tuple[int, int], automatic_randomization: bool, augm_config):
        assert len(dataset_hw) == 2 and all(x > 0 for x in dataset_hw)
        self.hw_tuple = tuple(dataset_hw)
        self.automatic_randomization = automatic_randomization
        self.h_flip_prob = _get(augm_config, 'prob_hflip')
        self.t_flip_prob = _get(augm_config, 'prob_tflip')
        rot = _get(augm_config, 'rotate')
        self.rot_prob = _get(rot, 'prob', 0) if rot is not None else 0
        if self.rot_prob > 0:
            raise NotImplementedError('rotation augmentation (prob 0 in every shipped config) is not built')
        zoom = _get(augm_config, 'zoom')
        self.zoom_prob = _get(zoom, 'prob')
        zo = _get(zoom, 'zoom_out')
        zi = _get(zoom, 'zoom_in')
        zo_w = _get(zo, 'weight', 1)
        self.zo_min, self.zo_max = _get(_get(zo, 'factor'), 'min'), _get(_get(zo, 'factor'), 'max')
        zi_w = _get(zi, 'weight') if zi is not None else 0
        self.zi_min = _get(_get(zi, 'factor'), 'min') if zi is not None else 1
        self.zi_max = _get(_get(zi, 'factor'), 'max') if zi is not None else 1
        assert 0 <= self.h_flip_prob <= 1 and 0 <= self.t_flip_prob <= 1 and 0 <= self.zoom_prob <= 1
        assert self.zi_max >= self.zi_min >= 1 and self.zo_max >= self.zo_min >= 1 and zi_w >= 0 and zo_w >= 0
        self.zoom_in_or_out_distribution = th.distributions.categorical.Categorical(probs=th.tensor([zi_w, zo_w], dtype=th.float32))
        self.augm_state = AugmentationState()

    def randomize_augmentation(self):
        """augmentor.py:176-207, draw for draw (the rotation draw is kept so that the stream of random numbers is the same)."""
        s = self.augm_state
        s.apply_h_flip = self.h_flip_prob > th.rand(1).item()
        s.apply_t_flip = self.t_flip_prob > th.rand(1).item()
        th.rand(1)                                         # rotation.active (probability 0)
        do_zoom = self.zoom_prob > th.rand(1).item()
        zoom_in = self.zoom_in_or_out_distribution.sample().item() == 0
        s.zoom_in = ZoomState(active=bool(zoom_in and do_zoom))
        s.zoom_out = ZoomState(active=bool((not zoom_in) and do_zoom))
        if s.zoom_out.active:
            f = _uniform(self.zo_min, self.zo_max)
            H, W = self.hw_tuple
            wh, ww = int(H / f), int(W / f)
            s.zoom_out.x0 = int(_uniform(0, W - ww))
            s.zoom_out.y0 = int(_uniform(0, H - wh))
            s.zoom_out.factor = f

    def choose_zoom_in_window(self, labels_of_sequence: Sequence[Optional[ObjectLabels]]):
        """augmentor.py:290-317: factor, then a window around a box of the most recent non-empty label frame (labels as they
        are BEFORE this augmentation; a time-flipped sequence passes its reversed label list)."""
        s = self.augm_state
        f = _uniform(self.zi_min, self.zi_max)
        latest = next((l for l in reversed(list(labels_of_sequence)) if l is not None and len(l) > 0), None)
        if f == 1 or latest is None:
            s.zoom_in = ZoomState()
            return
        H, W = self.hw_tuple
        win = (int(H / f), int(W / f))
        rows = latest.object_labels.to(th.float32)
        if s.apply_h_flip:       # the reference flips the sample (labels included, in fp32) before it zooms (augmentor.py:466-471)
            rows = rows.clone()
            rows[:, 1] = W - 1 - rows[:, 1] - rows[:, 3]
        cand = [zoom_window_for_label(rows[i, 1:5].tolist(), (H, W), win) for i in range(rows.shape[0])]
        k = 0 if len(cand) == 1 else th.randint(low=0, high=len(cand) - 1, size=(1,)).item()
        s.zoom_in = ZoomState(active=True, x0=cand[k][0], y0=cand[k][1], factor=f)


def _f32(v: float) -> float:
    return float(th.tensor(v, dtype=th.float64).to(th.float32))


def pack_state(state: AugmentationState, hw: Tuple[int, int], is_reversed: bool = False) -> _lib.AugmState:
    """AugmentationState -> leod_augm_state: window sizes as the tensor path computes them (augmentor.py:241, 322), label constants
    as the Python doubles of labels.py:371-411, 437-459, 482-497 (rounded to fp32 where torch does)."""
    H, W = hw
    r = _lib.AugmState()
    r.h_flip, r.t_flip = int(state.apply_h_flip), int(is_reversed)
    r.flip_c = float(W - 1)
    assert not (state.zoom_in.active and state.zoom_out.active)
    if state.zoom_in.active and state.zoom_in.factor != 1:
        f, z = state.zoom_in.factor, state.zoom_in
        r.zoom_mode, r.x0, r.y0, r.win_h, r.win_w = 1, z.x0, z.y0, int(H / f), int(W / f)
        zh, zw = H / f, W / f
        r.lo_x, r.hi_x = float(z.x0), min(z.x0 + zw, W - 1) - 1
        r.lo_y, r.hi_y = float(z.y0), min(z.y0 + zh, H - 1) - 1
        r.mul, r.cap_x, r.cap_y = f, f * zw - 1, f * zh - 1
    elif state.zoom_out.active and state.zoom_out.factor != 1:
        f, z = state.zoom_out.factor, state.zoom_out
        r.zoom_mode, r.x0, r.y0, r.win_h, r.win_w = 2, z.x0, z.y0, int(H / f), int(W / f)
        m = 1 / f
        r.mul, r.cap_x, r.cap_y = m, m * W - 1, m * H - 1
    return r


class BatchSpatialAugmentor:
    """Augment a whole device batch: `ev` uint8 [L, B, C, H, W] (cuda), `labels` an L-list of SparselyBatchedObjectLabels."""

    def __init__(self, dataset_hw: Tuple[int, int], augm_config, batch_size: int):
        self.hw = tuple(dataset_hw)
        self.samplers = [RandomSpatialAugmentorGenX(self.hw, False, augm_config) for _ in range(batch_size)]

    def randomize(self, labels: Optional[List[SparselyBatchedObjectLabels]] = None) -> List[AugmentationState]:
        """Draw a state for every sequence (sequence after sequence, as independent workers would)."""
        for b, s in enumerate(self.samplers):
            s.randomize_augmentation()
            if s.augm_state.zoom_in.active:
                assert labels is not None, 'zoom-in samples its window from the labels (augmentor.py:299-307)'
                s.choose_zoom_in_window([labels[t][b] for t in range(len(labels))])
        return [s.augm_state for s in self.samplers]

    def __call__(self, ev: th.Tensor, labels: Optional[List[SparselyBatchedObjectLabels]] = None,
                 states: Optional[List[AugmentationState]] = None, is_reversed: Optional[Sequence[bool]] = None):
        """-> (augmented ev, augmented labels (new objects; None where the input had None), states)."""
        assert ev.is_cuda and ev.dtype == th.uint8 and ev.dim() == 5, 'the augmentation kernels take a cuda uint8 [L,B,C,H,W] batch'
        L, B, C, H, W = ev.shape
        assert (H, W) == self.hw and B == len(self.samplers)
        if states is None:
            states = self.randomize(labels)
        rev = list(is_reversed) if is_reversed is not None else [False] * B
        recs = (_lib.AugmState * B)(*[pack_state(states[b], self.hw, bool(rev[b])) for b in range(B)])
        ev = ev.contiguous()
        out = th.empty_like(ev)
        with th.cuda.device(ev.device):
            _lib.check(_lib.lib().leod_augment_ev_repr(_lib.ptr(ev), _lib.ptr(out), L, B, C, H, W, recs, _lib.stream_ptr(ev.device)),
                       'augment_ev_repr')
        new_labels = self.augment_labels(labels, recs, rev, ev.device) if labels is not None else None
        return out, new_labels, states

    def augment_labels(self, labels, recs, rev, device):
        L, B = len(labels), len(self.samplers)
        src = []   # (t_out, b, ObjectLabels)
        for t in range(L):
            for b in range(B):
                lab = labels[L - 1 - t][b] if rev[b] else labels[t][b]       # time flip: the label list is reversed (labels.py:707-708)
                if lab is not None:
                    src.append((t, b, lab))
        out = [[None] * B for _ in range(L)]
        n = sum(len(l) for _, _, l in src)
        if n > 0:
            rows = th.cat([l.object_labels.to(th.float32) for _, _, l in src if len(l) > 0], 0)
            assert rows.shape[1] == 8, 'ObjectLabels rows are (t, x, y, w, h, class_id, class_confidence, objectness)'
            seq = th.cat([th.full((len(l),), b, dtype=th.int32) for _, b, l in src if len(l) > 0])
            d_rows, d_seq = _lib.upload_small(rows, device), _lib.upload_small(seq, device)
            keep = th.empty(n, dtype=th.uint8, device=device)
            with th.cuda.device(device):
                _lib.check(_lib.lib().leod_augment_labels(_lib.ptr(d_rows), _lib.ptr(d_seq), n, B, recs, _lib.ptr(keep),
                                                          _lib.stream_ptr(device)), 'augment_labels')
            rows, keep = d_rows.cpu(), keep.cpu().bool()
        k = 0
        for t, b, lab in src:
            m = len(lab)
            if m == 0:
                out[t][b] = ObjectLabels(lab.object_labels.clone(), lab.input_size_hw)
                continue
            out[t][b] = ObjectLabels(rows[k:k + m][keep[k:k + m]], lab.input_size_hw)
            k += m
        return [SparselyBatchedObjectLabels(r) for r in out]
