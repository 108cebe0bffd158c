"""Detection evaluation on the device.  Same class name, buffer methods and result keys as the reference's
utils/evaluation/prophesee/evaluator.py:25-110 (PropheseeEvaluator: add_labels / add_predictions per labelled frame,
evaluate_buffer -> {'AP', 'AP_50', 'AP_75', 'AP_S', 'AP_M', 'AP_L'} overall and with a `_<class>` suffix per class), but the box
filter, the COCO matching and the precision/recall accumulation run in leod_coco_eval (csrc/kernels_eval.cu) instead of
numpy + pycocotools.  The final means over the [10,101,K,4,3] precision array are numpy's (COCOeval.summarize).
No CPU fallback: evaluate_buffer needs the CUDA library."""
from typing import Any, Dict, List, Optional
from warnings import warn

import numpy as np
import torch

from leod_b200 import _lib

LABELMAP = {'gen1': ('car', 'ped'), 'gen4': ('ped', 'cyc', 'car')}
IOU_THRS = np.linspace(.5, 0.95, int(np.round((0.95 - .5) / .05)) + 1, endpoint=True)       # pycocotools Params.setDetParams
REC_THRS = np.linspace(.0, 1.00, int(np.round((1.00 - .0) / .01)) + 1, endpoint=True)
OUT_KEYS = ('AP', 'AP_50', 'AP_75', 'AP_S', 'AP_M', 'AP_L')


def get_labelmap(dst_name: str = None, num_cls: int = None):
    assert dst_name is None or num_cls is None
    if dst_name is not None:
        return LABELMAP[dst_name.lower()]
    assert num_cls in (2, 3), f'Invalid number of classes: {num_cls}'
    return LABELMAP['gen1'] if num_cls == 2 else LABELMAP['gen4']


def filter_thresholds(camera: str, downsampled_by_2: bool):
    """evaluation.py:24-33 -> (skip_ts, min_box_diag, min_box_side)."""
    diag, side = (60, 20) if camera == 'gen4' else (30, 10)
    if downsampled_by_2:
        diag, side = diag // 2, side // 2
    return int(5e5), diag, side


def summarize(precision: np.ndarray, recall: np.ndarray) -> np.ndarray:
    """COCOeval.summarize for iouType 'bbox': 12 numbers (the reference reports the first six, coco_eval.py:103-120)."""
    def one(ap, iou=None, area=0, m=2):
        s = precision if ap else recall
        if iou is not None:
            s = s[np.where(iou == IOU_THRS)[0]]
        s = s[:, :, :, area, m] if ap else s[:, :, area, m]
        return -1 if len(s[s > -1]) == 0 else np.mean(s[s > -1])
    return np.array([one(1), one(1, .5), one(1, .75), one(1, area=1), one(1, area=2), one(1, area=3),
                     one(0, m=0), one(0, m=1), one(0, m=2), one(0, area=1), one(0, area=2), one(0, area=3)])


class FrameBoxes:
    """Flat device copy of per-frame box lists: t int64 [N], xywh fp32 [N,4], cls int32 [N], score fp32 [N], ptr int32 [F+1]."""

    def __init__(self, frames: List[Any], device, with_score: bool):
        n = [len(f) for f in frames]
        self.ptr = torch.tensor(np.concatenate(([0], np.cumsum(n))), dtype=torch.int32)
        cat = (lambda k, dt: np.concatenate([np.asarray(f[k], dt) for f in frames]) if sum(n) else np.zeros(0, dt))
        self.t = torch.from_numpy(cat('t', np.int64))
        self.xywh = torch.from_numpy(np.stack([cat(k, np.float32) for k in ('x', 'y', 'w', 'h')], 1)) if sum(n) else torch.zeros(0, 4)
        self.cls = torch.from_numpy(cat('class_id', np.int64).astype(np.int32))
        self.score = torch.from_numpy(cat('class_confidence', np.float32)) if with_score else torch.zeros(0)
        for k in ('ptr', 't', 'xywh', 'cls', 'score'):
            v = getattr(self, k).contiguous()
            setattr(self, k, v.to(device) if v.numel() else torch.zeros((1,) + tuple(v.shape[1:]), dtype=v.dtype, device=device))


def coco_eval_device(gt: FrameBoxes, dt: FrameBoxes, num_frames: int, num_classes: int, camera: str, downsampled_by_2: bool,
                     only_class: int = -1):
    """-> (precision [10,101,K,4,3], recall [10,K,4,3], counts [images, detections, status]) as numpy arrays."""
    dev = gt.ptr.device
    skip_ts, diag, side = filter_thresholds(camera, downsampled_by_2)
    L = _lib.lib()
    ws = torch.empty(int(L.leod_coco_eval_workspace_bytes(num_frames, num_classes)), dtype=torch.uint8, device=dev)
    precision = torch.empty((10, 101, num_classes, 4, 3), dtype=torch.float64, device=dev)
    recall = torch.empty((10, num_classes, 4, 3), dtype=torch.float64, device=dev)
    counts = torch.empty(3, dtype=torch.int32, device=dev)
    iou = (_lib.ctypes.c_double * 10)(*IOU_THRS.tolist())
    rec = (_lib.ctypes.c_double * 101)(*REC_THRS.tolist())
    with torch.cuda.device(dev):
        _lib.check(L.leod_coco_eval(_lib.ptr(gt.t), _lib.ptr(gt.xywh), _lib.ptr(gt.cls), _lib.ptr(gt.ptr), _lib.ptr(dt.t), _lib.ptr(dt.xywh),
                                    _lib.ptr(dt.cls), _lib.ptr(dt.score), _lib.ptr(dt.ptr), num_frames, num_classes, skip_ts, diag, side,
                                    only_class, iou, rec, _lib.ptr(ws), _lib.ptr(precision), _lib.ptr(recall), _lib.ptr(counts),
                                    _lib.stream_ptr(dev)), 'coco_eval')
    c = counts.cpu().numpy()
    if c[2] != 0:
        raise RuntimeError('leod_coco_eval: a frame holds more than 128 ground-truth boxes or 2048 detections of one class')
    return precision.cpu().numpy(), recall.cpu().numpy(), c


class PropheseeEvaluator:
    LABELS = 'lables'
    PREDICTIONS = 'predictions'

    def __init__(self, dataset: str, downsample_by_2: bool, device='cuda'):
        assert dataset in {'gen1', 'gen4'}
        self.dataset = dataset
        self.label_map = get_labelmap(dataset)
        self.downsample_by_2 = downsample_by_2
        self.device = torch.device(device)
        self._reset_buffer()

    def _reset_buffer(self):
        self._buffer_empty = True
        self._buffer = {self.LABELS: [], self.PREDICTIONS: []}

    def _add_to_buffer(self, key: str, value: List[np.ndarray]):
        assert isinstance(value, list) and all(isinstance(v, np.ndarray) for v in value)
        self._buffer_empty = False
        self._buffer[key].extend(value)

    def add_predictions(self, predictions: List[np.ndarray]):
        self._add_to_buffer(self.PREDICTIONS, predictions)

    def add_labels(self, labels: List[np.ndarray]):
        self._add_to_buffer(self.LABELS, labels)

    def reset_buffer(self) -> None:
        self._reset_buffer()

    def has_data(self):
        return not self._buffer_empty

    def evaluate_buffer(self, img_height: int, img_width: int, ret_pr_curve: bool = False) -> Optional[Dict[str, Any]]:
        """evaluator.py:73-110.  img_height / img_width only enter COCO's image records, never the numbers."""
        if self._buffer_empty:
            warn('Attempt to use prophesee evaluation buffer, but it is empty', UserWarning, stacklevel=2)
            return None
        labels, predictions = self._buffer[self.LABELS], self._buffer[self.PREDICTIONS]
        assert len(labels) == len(predictions)
        K = len(self.label_map)
        gt = FrameBoxes(labels, self.device, with_score=False)
        dt = FrameBoxes(predictions, self.device, with_score=True)
        metrics = {}
        for only, suffix in [(-1, '')] + [(k, f'_{name}') for k, name in enumerate(self.label_map)]:
            precision, recall, counts = coco_eval_device(gt, dt, len(labels), K, self.dataset, self.downsample_by_2, only_class=only)
            if counts[1] == 0:      # coco_eval.py:96-99: no detections -> zeros
                stats = np.zeros(12)
            else:
                stats = summarize(precision, recall)
            metrics.update({f'{k}{suffix}': float(stats[i]) for i, k in enumerate(OUT_KEYS)})
        return metrics
