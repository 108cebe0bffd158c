"""Sequence post-processing of the pseudo-label sweep on the device.

Host mirror of `EventSeqData._track_filter` (modules/pseudo_labeler.py:268-333: linear tracking forward / backward, "ignore" class
for boxes on short tracks, in-painting of missed detections) and `EventSeqData._summarize` (:179-199: the `labels.npz` arrays), batched
over sequences: ONE library call post-processes all sequences handed in (leod_track_filter: one CTA per sequence and direction;
leod_pack_bbox for the records).  Bit-exact with the reference (tests/test_gpu_tracking.py against tests/golden/tracking_cases.npz).
"""
import ctypes
from typing import List, Sequence, Tuple

import numpy as np
import torch

from leod_b200 import _lib

BBOX_DTYPE = np.dtype({'names': ['t', 'x', 'y', 'w', 'h', 'class_id', 'class_confidence', 'objectness'],
                       'formats': ['<i8', '<f4', '<f4', '<f4', '<f4', '<u4', '<f4', '<f4'],
                       'offsets': [0, 8, 12, 16, 20, 24, 28, 32], 'itemsize': 40})        # data/genx_utils/labels.py:12-16
BBOX_DTYPE_PACKED = np.dtype([(n, BBOX_DTYPE[n]) for n in BBOX_DTYPE.names])             # 36 bytes: what np.concatenate makes of it


def track_filter_sequences(seqs: Sequence[Tuple[Sequence[int], Sequence[torch.Tensor]]], hw, min_track_len: int = 6,
                           track_method: str = 'forward or backward', inpaint: bool = True, ignore_label: int = 1024, q: float = 0.9,
                           min_conf: float = 0.55, iou_thr: float = 0.45, device='cuda'):
    """seqs: per sequence (frame_idx ascending, per-frame [n, 8] ObjectLabels rows).  hw: (H, W) or one per sequence.
    -> per sequence (frame_idx list, per-frame [n, 8] CUDA tensors), as `EventSeqData._track_filter` leaves `frame_idx` / `labels`."""
    S = len(seqs)
    if S == 0 or min_track_len <= 0:
        return [(list(fi), [torch.as_tensor(r) for r in rows]) for fi, rows in seqs]
    dev = torch.device(device)
    if dev.type != 'cuda':
        raise RuntimeError('leod_b200 tracking runs on CUDA only (no CPU fallback)')
    hws = [hw] * S if isinstance(hw[0], (int, np.integer)) else list(hw)
    frame_ptr, frame_idx, seq_ptr, rows_all = [0], [], [0], []
    for fi, rows in seqs:
        assert len(fi) == len(rows) and list(fi) == sorted(fi), 'frames must come in ascending order'
        for f, r in zip(fi, rows):
            r = torch.as_tensor(r, dtype=torch.float32).reshape(-1, 8)
            rows_all.append(r.cpu())
            frame_idx.append(int(f))
            frame_ptr.append(frame_ptr[-1] + r.shape[0])
        seq_ptr.append(len(frame_idx))
    total_rows, total_frames = frame_ptr[-1], len(frame_idx)
    max_len = max([fi[-1] + 1 for fi, _ in seqs if len(fi)] + [1])
    rows_host = torch.cat(rows_all) if rows_all else torch.zeros(0, 8)
    rows_dev = rows_host.to(dev).contiguous() if total_rows else torch.zeros(1, 8, device=dev)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)                       # noqa: E731
    frame_ptr_d, frame_idx_d, seq_ptr_d = i32(frame_ptr), i32(frame_idx if frame_idx else [0]), i32(seq_ptr)
    hw_d = i32([[int(h), int(w)] for h, w in hws])
    hole_cap = 64 + 8 * max(1, total_rows // max(S, 1))
    n_rows = [frame_ptr[seq_ptr[s + 1]] - frame_ptr[seq_ptr[s]] for s in range(S)]
    n_frames = [seq_ptr[s + 1] - seq_ptr[s] for s in range(S)]
    row_base = np.concatenate(([0], np.cumsum([n + hole_cap for n in n_rows]))).astype(np.int64)
    frame_base = np.concatenate(([0], np.cumsum([n + hole_cap for n in n_frames]))).astype(np.int64)
    out_rows = torch.zeros(int(row_base[-1]), 8, dtype=torch.float32, device=dev)
    out_fi = torch.zeros(int(frame_base[-1]), dtype=torch.int32, device=dev)
    out_fs = torch.zeros_like(out_fi)
    out_counts = torch.zeros(S, 2, dtype=torch.int32, device=dev)
    status = torch.zeros(S, dtype=torch.int32, device=dev)
    npow = max_len + 2
    qpow = (ctypes.c_double * npow)(*[q ** a for a in range(npow)])                      # the host's pow(), as CPython's float.__pow__
    row_base_d, frame_base_d = i32(row_base[:-1].tolist()), i32(frame_base[:-1].tolist())     # keep alive across the call
    l = _lib.lib()
    ws = torch.empty(int(l.leod_track_workspace_bytes(total_rows, S, hole_cap, npow)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(l.leod_track_filter(_lib.ptr(rows_dev), _lib.ptr(frame_ptr_d), _lib.ptr(frame_idx_d), _lib.ptr(seq_ptr_d), _lib.ptr(hw_d),
                                       total_rows, total_frames, S, qpow, npow, q, min_conf, iou_thr, int(min_track_len), int(bool(inpaint)),
                                       int('backward' in track_method), float(ignore_label), hole_cap, _lib.ptr(ws), _lib.ptr(row_base_d),
                                       _lib.ptr(frame_base_d), _lib.ptr(out_rows), _lib.ptr(out_fi), _lib.ptr(out_fs),
                                       _lib.ptr(out_counts), _lib.ptr(status), _lib.stream_ptr(dev)), 'track_filter')
    st = status.cpu().tolist()
    if any(st):
        bad = [(s, c) for s, c in enumerate(st) if c]
        raise RuntimeError(f'leod_track_filter: capacity exceeded (sequence, code) {bad[:4]}: 1 = more than 64 live tracks / boxes per frame, '
                           f'2 = in-painting capacity, 3 = q^age table')
    counts = out_counts.cpu().tolist()
    fi_h, fs_h = out_fi.cpu().numpy(), out_fs.cpu().numpy()
    res = []
    for s in range(S):
        nr, nf = counts[s]
        fi = fi_h[frame_base[s]:frame_base[s] + nf].tolist()
        starts = fs_h[frame_base[s]:frame_base[s] + nf].tolist() + [nr]
        rows = out_rows[row_base[s]:row_base[s] + nr]
        res.append((fi, [rows[starts[k]:starts[k + 1]] for k in range(nf)]))
    return res


def pack_labels(rows: torch.Tensor, packed: bool = True) -> np.ndarray:
    """[n, 8] CUDA ObjectLabels rows -> n label-file records (structured numpy array, BBOX_DTYPE fields)."""
    rows = rows.to(torch.float32).contiguous()
    n = rows.shape[0]
    stride = 36 if packed else 40
    out = torch.empty(max(n, 1) * stride, dtype=torch.uint8, device=rows.device)
    with torch.cuda.device(rows.device):
        _lib.check(_lib.lib().leod_pack_bbox(_lib.ptr(rows), n, _lib.ptr(out), stride, _lib.stream_ptr(rows.device)), 'pack_bbox')
    return np.frombuffer(out[:n * stride].cpu().numpy().tobytes(), dtype=BBOX_DTYPE_PACKED if packed else BBOX_DTYPE)


def summarize(frame_idx: Sequence[int], frames_rows: List[torch.Tensor]):
    """EventSeqData._summarize (pseudo_labeler.py:179-199) -> (`labels` records, `objframe_idx_2_label_idx`, `objframe_idx_2_repr_idx`):
    the arrays of a sequence's `labels_v2/labels.npz` and `objframe_idx_2_repr_idx.npy`."""
    starts, n = [], 0
    for r in frames_rows:
        starts.append(n)
        n += r.shape[0]
    rows = torch.cat(list(frames_rows)) if frames_rows else torch.zeros(0, 8, device='cuda')
    return pack_labels(rows, packed=True), np.asarray(starts, np.int64), np.asarray(list(frame_idx), np.int64)


class EventSeqData:
    """Labels of one event sequence collected during the sweep — mirror of modules/pseudo_labeler.py:94-400 (`update` with the
    hflip / tflip bookkeeping :107-155, `_aggregate_results` :157-177, `_track_filter` :268-333, `_summarize` :179-199, `save`
    :335-397).  Rows stay on the device; the TTA merge, the tracker and the record packer are library calls.  `finalize_sequences`
    post-processes many finished sequences with ONE tracker launch (the reference tracks them one by one in Python)."""

    def __init__(self, path: str, scale_ratio: float, filter_config, postproc_cfg, hw):
        self.path, self.scale_ratio, self.filter_config, self.postproc_cfg = path, scale_ratio, filter_config, postproc_cfg
        self.hw = tuple(hw)
        self._eoe, self._aug = False, False
        self.frame_idx_2_labels = {}
        self.frame_idx, self.labels = [], []

    @property
    def eoe(self) -> bool:
        return self._eoe

    def update(self, labels, ev_idx, is_last_sample: bool, is_padded_mask, is_hflip: bool, is_tflip: bool, tflip_offset: int = 0) -> None:
        """pseudo_labeler.py:107-155.  labels: per-timestep ObjectLabels | None of one chunk; ev_idx: their frame indices (-1 = padding)."""
        self._eoe = bool(is_last_sample)
        if is_hflip:
            for l in labels:
                if l is not None:
                    l.flip_lr_()
            self._aug = True
        if is_tflip:
            ev_idx = [i + tflip_offset for i in ev_idx]
            self._aug = True
        for tidx, (label, f) in enumerate(zip(labels, ev_idx)):
            if f < 0 or label is None or len(label) == 0:
                continue
            assert not is_padded_mask[tidx]
            rows = label.object_labels
            if self.scale_ratio != 1:                       # labels are stored at sensor resolution (ObjectLabels.scale_)
                rows = rows.clone()
                rows[:, 1:5] = rows[:, 1:5] * self.scale_ratio
            if f in self.frame_idx_2_labels:
                if bool((rows[:, 0] > 0).any()):            # ground truth is added once (:139-149)
                    continue
                self.frame_idx_2_labels[f] = torch.cat((self.frame_idx_2_labels[f], rows), 0)
            else:
                self.frame_idx_2_labels[f] = rows

    def aggregate_results(self, num_frames: int) -> None:
        """pseudo_labeler.py:157-177: sort by frame, merge the TTA views of every frame with a second NMS."""
        from leod_b200.data.labels import ObjectLabels
        from leod_b200.modules.pseudo_labeler import tta_postprocess
        assert self._eoe, 'Cannot aggregate results before the sequence ends.'
        self.frame_idx = sorted(i for i in self.frame_idx_2_labels if 0 <= i < num_frames)
        self.labels = [self.frame_idx_2_labels[i] for i in self.frame_idx]
        if self._aug and self.labels:
            hw = tuple(int(v * self.scale_ratio) for v in self.hw)
            merged = tta_postprocess([ObjectLabels(r, hw) for r in self.labels], conf_thre=self.postproc_cfg.confidence_threshold,
                                     nms_thre=self.postproc_cfg.nms_threshold)
            self.labels = [m.object_labels for m in merged]

    def save(self, save_dir: str, num_frames: int, labels_fn: str = 'labels_v2/labels.npz', repr_dir: str = 'event_representations_v2'):
        """The files of pseudo_labeler.py:335-381 for this sequence (the h5 soft link of :362 is the caller's: h5py is not a dependency):
        <save_dir>/<seq>/labels_v2/labels.npz {labels, objframe_idx_2_label_idx}, <save_dir>/<seq>/<repr_dir>/objframe_idx_2_repr_idx.npy."""
        import os
        if not self.labels and self.frame_idx_2_labels:
            finalize_sequences([self], [num_frames])
        labels, lbl_idx, repr_idx = summarize(self.frame_idx, self.labels)
        seq_dir = os.path.join(save_dir, os.path.basename(self.path))
        os.makedirs(os.path.join(seq_dir, os.path.dirname(labels_fn)), exist_ok=True)
        os.makedirs(os.path.join(seq_dir, repr_dir), exist_ok=True)
        np.save(os.path.join(seq_dir, repr_dir, 'objframe_idx_2_repr_idx.npy'), repr_idx)
        np.savez(os.path.join(seq_dir, labels_fn), labels=labels, objframe_idx_2_label_idx=lbl_idx)
        return seq_dir


def finalize_sequences(seqs: Sequence[EventSeqData], num_frames: Sequence[int]) -> None:
    """`_aggregate_results` per sequence, then `_track_filter` for ALL of them in one tracker launch (pseudo_labeler.py:366-367)."""
    for s, n in zip(seqs, num_frames):
        s.aggregate_results(n)
    if not seqs:
        return
    fc = seqs[0].filter_config
    mtl = int(fc.get('min_track_len', 6))
    if mtl <= 0:
        return
    res = track_filter_sequences([(s.frame_idx, s.labels) for s in seqs], [tuple(int(v * s.scale_ratio) for v in s.hw) for s in seqs],
                                 min_track_len=mtl, track_method=fc.get('track_method', 'forward or backward'), inpaint=bool(fc.get('inpaint', True)),
                                 ignore_label=int(fc.get('ignore_label', 1024)))
    for s, (fi, rows) in zip(seqs, res):
        s.frame_idx, s.labels = fi, rows
