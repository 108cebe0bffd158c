"""Oracle: confidence filter + class-aware NMS + pseudo-label filters, numpy float32 / int64.

Restates (paths relative to /root/reference):
  models/detection/yolox/utils/boxes.py:32-86            (postprocess)
  modules/utils/ssod.py:40-63, 90-110, 113-133, 136-188  (FOV clamp, size filters, thresholds, pred2label)
  modules/pseudo_labeler.py:37-91                        (TTA merge NMS)
and the third-party op the reference calls at boxes.py:73 — torchvision.ops.batched_nms
(torchvision pinned 0.15.2 in the reference's environment.yml:94; not vendored).  Published
algorithm restated here: boxes are shifted by class_id * (max_coordinate + 1) so classes never
overlap, candidates are visited in stable descending-score order, a candidate is dropped when its
IoU with an already-kept box is strictly greater than the threshold, kept indices are returned in
score order.  All arithmetic in float32, in the same operation order as the torchvision kernels
((x2-x1)*(y2-y1) areas, inter / (a_i + a_j - inter)).
TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

f32 = np.float32


def nms_indices(boxes: np.ndarray, scores: np.ndarray, thr: float) -> np.ndarray:
    """Greedy NMS on xyxy float32 boxes -> kept indices (int64) in descending-score order."""
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), np.int64)
    boxes = boxes.astype(f32, copy=False)
    order = np.argsort(-scores.astype(f32), kind='stable')
    x1, y1, x2, y2 = (boxes[:, i] for i in range(4))
    areas = (x2 - x1) * (y2 - y1)
    suppressed = np.zeros(n, bool)
    keep = []
    thr = f32(thr)
    for a in range(n):
        i = order[a]
        if suppressed[i]:
            continue
        keep.append(i)
        rest = order[a + 1:]
        xx1 = np.maximum(x1[i], x1[rest])
        yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest])
        yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(f32(0), xx2 - xx1)
        h = np.maximum(f32(0), yy2 - yy1)
        inter = w * h
        with np.errstate(divide='ignore', invalid='ignore'):
            iou = inter / (areas[i] + areas[rest] - inter)
        suppressed[rest[iou > thr]] = True
    return np.asarray(keep, np.int64)


def batched_nms_indices(boxes: np.ndarray, scores: np.ndarray, cls: np.ndarray, thr: float) -> np.ndarray:
    """torchvision.ops.batched_nms, coordinate-trick branch (taken for < 4000 boxes on CPU and
    < 20 000 (0.15) / 100 000 (0.26) on CUDA — always, for 1 680 / 5 040 anchors)."""
    if boxes.shape[0] == 0:
        return np.zeros((0,), np.int64)
    boxes = boxes.astype(f32, copy=False)
    max_coord = boxes.max()
    offsets = cls.astype(f32) * (max_coord + f32(1))
    return nms_indices(boxes + offsets[:, None], scores, thr)


def postprocess(prediction: np.ndarray, num_classes: int, conf_thre: float = 0.7, nms_thre: float = 0.45,
                class_agnostic: bool = False) -> List[np.ndarray]:
    """boxes.py:32-86.  prediction [B,A,5+C] (cx,cy,w,h,obj,cls...) float32, NOT modified here
    (the reference overwrites [..., :4] with corners in place; callers that rely on that see the
    same values in `corners`).  Returns per image [N_i,7] = x1,y1,x2,y2,obj,cls_conf,cls_idx in
    descending obj*cls_conf order; empty images give a [0,7] array (reference: `pad`)."""
    p = prediction.astype(f32, copy=True)
    half_w, half_h = p[..., 2] / f32(2), p[..., 3] / f32(2)
    corners = np.stack((p[..., 0] - half_w, p[..., 1] - half_h, p[..., 0] + half_w, p[..., 1] + half_h), -1)
    out = []
    for b in range(p.shape[0]):
        cls_scores = p[b, :, 5:5 + num_classes]
        cls_idx = cls_scores.argmax(1)            # first max on ties, as torch.max
        cls_conf = cls_scores[np.arange(len(cls_idx)), cls_idx]
        obj = p[b, :, 4]
        score = obj * cls_conf
        m = score >= f32(conf_thre)
        det = np.concatenate((corners[b][m], obj[m, None], cls_conf[m, None], cls_idx[m, None].astype(f32)), 1)
        if det.shape[0] == 0:
            out.append(np.zeros((0, 7), f32))
            continue
        s = det[:, 4] * det[:, 5]
        keep = nms_indices(det[:, :4], s, nms_thre) if class_agnostic else \
            batched_nms_indices(det[:, :4], s, det[:, 6], nms_thre)
        out.append(det[keep])
    return out


def _per_class_gt(scores, cls, thresh: Union[float, Sequence[float]]):
    """ssod.py:136-144 (strict >)."""
    if isinstance(thresh, float):
        return scores > f32(thresh)
    m = np.zeros(scores.shape, bool)
    for i, t in enumerate(thresh):
        m |= (cls == i) & (scores > f32(t))
    return m


def filter_pred_boxes(xyxy: np.ndarray, frame_hw: Tuple[int, int]):
    """ssod.py:113-133: clamp to [0,W-1]x[0,H-1]; keep w,h > 0 after the clamp, w,h >= 5 px,
    w <= (9*W)//10.  frame_hw is the dataset frame (240,304) gen1 / (360,640) gen4-downsampled."""
    H, W = frame_hw
    x1 = np.clip(xyxy[:, 0], f32(0), f32(W - 1))
    y1 = np.clip(xyxy[:, 1], f32(0), f32(H - 1))
    x2 = np.clip(xyxy[:, 2], f32(0), f32(W - 1))
    y2 = np.clip(xyxy[:, 3], f32(0), f32(H - 1))
    w, h = x2 - x1, y2 - y1
    keep = (w > 0) & (h > 0) & (w >= 5) & (h >= 5) & (w <= (9 * W) // 10)
    return np.stack((x1, y1, x2, y2), 1), keep


def pred2label(dets: List[np.ndarray], obj_thresh, cls_thresh, frame_hw: Optional[Tuple[int, int]] = None) \
        -> List[np.ndarray]:
    """ssod.py:147-188.  dets: per image [N,7] from postprocess.  Returns per image [M,8] =
    (t=0, x, y, w, h (corner format), cls_idx, cls_conf, obj_conf) — the ObjectLabels row layout."""
    out = []
    for d in dets:
        d = d.astype(f32, copy=True)
        sel = _per_class_gt(d[:, 4], d[:, 6], obj_thresh) & _per_class_gt(d[:, 5], d[:, 6], cls_thresh)
        if frame_hw is not None:
            d[:, :4], keep = filter_pred_boxes(d[:, :4], frame_hw)
            sel &= keep
        d = d[sel]
        lab = np.zeros((d.shape[0], 8), f32)
        lab[:, 1], lab[:, 2] = d[:, 0], d[:, 1]
        lab[:, 3], lab[:, 4] = d[:, 2] - d[:, 0], d[:, 3] - d[:, 1]
        lab[:, 5], lab[:, 6], lab[:, 7] = d[:, 6], d[:, 5], d[:, 4]
        out.append(lab)
    return out


def tta_merge(labels: np.ndarray, conf_thre: float, nms_thre: float) -> np.ndarray:
    """pseudo_labeler.py:37-91 for one frame of pseudo labels [N,8] (ObjectLabels rows, corner
    xywh): second NMS over the concatenated TTA views; rows returned in score order."""
    if labels.shape[0] == 0:
        return labels
    l = labels.astype(f32, copy=False)
    xyxy = np.stack((l[:, 1], l[:, 2], l[:, 1] + l[:, 3], l[:, 2] + l[:, 4]), 1)
    obj, cls_conf, cls = l[:, 7], l[:, 6], l[:, 5]
    m = (obj * cls_conf) >= f32(conf_thre)
    if not m.any():
        return l[:0]
    idx = np.nonzero(m)[0]
    keep = batched_nms_indices(xyxy[idx], (obj * cls_conf)[idx], cls[idx], nms_thre)
    sel = idx[keep]
    out = l[sel].copy()
    out[:, 3] = xyxy[sel, 2] - xyxy[sel, 0]
    out[:, 4] = xyxy[sel, 3] - xyxy[sel, 1]
    return out
