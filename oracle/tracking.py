"""Oracle: the sequence post-processing of the pseudo-label sweep — linear-velocity greedy-IoU
tracking of the per-frame boxes, removal ("ignore" class) of boxes on short tracks, in-painting of missed detections.
TEST INFRASTRUCTURE — see oracle/__init__.py.

Restates (paths relative to /root/reference):
  modules/tracking/linear.py:10-151   (LinearBoxTracker: state, predict, update, robust velocity, miss)
  modules/tracking/linear.py:154-193  (associate_tracking), modules/tracking/utils.py:7-18 (greedy_matching),
  modules/tracking/utils.py:21-49     (iou_batch_xywh), :72-96 (clamp_bbox)
  modules/tracking/linear.py:196-292  (LinearTracker.update / finish via tracker.py:26-41)
  modules/pseudo_labeler.py:201-266   (EventSeqData._track), :268-333 (EventSeqData._track_filter)
Pinned against tests/golden/tracking_cases.npz, produced by running the reference (make_golden.py: gen_tracking).

Box arithmetic is float32 exactly where the reference's numpy arrays are float32 (detections, predicted boxes, velocity
after the first update); track confidences are Python floats.  Tracks are plain dict records in a list that keeps the
reference's order (creation order, deletions remove in place) because the greedy association visits tracks by confidence
with numpy's argsort tie-breaking on that order.
"""
from typing import Dict, List, Sequence, Tuple

import numpy as np

f32 = np.float32


def _xyxy(b):
    """centre xywh -> corners, float32 (tracking/utils.py:62-69)."""
    x, y, w, h = b[0], b[1], b[2], b[3]
    return np.array([x - w / 2., y - h / 2., x + w / 2., y + h / 2.], dtype=f32)


def clamp_center_box(b: np.ndarray, hw: Tuple[int, int]):
    """tracking/utils.py:72-96 for format 'xywh': clamp the corners to the image, report which sides were clamped."""
    H, W = hw
    x1_, y1_, x2_, y2_ = _xyxy(b)
    x1, x2 = np.clip(x1_, 0., W - 1.), np.clip(x2_, 0., W - 1.)
    y1, y2 = np.clip(y1_, 0., H - 1.), np.clip(y2_, 0., H - 1.)
    out = np.array([(x1 + x2) / 2., (y1 + y2) / 2., x2 - x1, y2 - y1], dtype=f32)
    return out, bool(y1 != y1_), bool(y2 != y2_), bool(x1 != x1_), bool(x2 != x2_)   # top, down, left, right


def pairwise_iou_center(trk: np.ndarray, det: np.ndarray) -> np.ndarray:
    """tracking/utils.py:21-49: IoU of [n,5] x [m,5] centre boxes, 0 between different classes."""
    t, d = trk[:, None, :], det[None, :, :]
    xx1 = np.maximum(t[..., 0] - t[..., 2] / 2., d[..., 0] - d[..., 2] / 2.)
    yy1 = np.maximum(t[..., 1] - t[..., 3] / 2., d[..., 1] - d[..., 3] / 2.)
    xx2 = np.minimum(t[..., 0] + t[..., 2] / 2., d[..., 0] + d[..., 2] / 2.)
    yy2 = np.minimum(t[..., 1] + t[..., 3] / 2., d[..., 1] + d[..., 3] / 2.)
    wh = np.maximum(0., xx2 - xx1) * np.maximum(0., yy2 - yy1)
    with np.errstate(divide='ignore', invalid='ignore'):
        o = wh / (t[..., 2] * t[..., 3] + d[..., 2] * d[..., 3] - wh)
    o[np.broadcast_to(d[..., 4] != t[..., 4], o.shape)] = 0.
    return o


def greedy_assign(iou: np.ndarray, order: Sequence[int], thr: float) -> List[Tuple[int, int]]:
    """tracking/utils.py:7-18: tracks in `order` take their best remaining detection if its IoU reaches `thr`."""
    iou = iou.copy()
    pairs = []
    for i in order:
        if iou[i].max() < thr:
            continue
        j = int(np.argmax(iou[i]))
        iou[:, j] = -np.inf
        pairs.append((int(i), j))
    return pairs


def _new_track(det: np.ndarray, box_id: int, gt: bool, q: float) -> Dict:
    return dict(box=det[:4].astype(f32).copy(), cls=det[4], last=None, pred=None, v=np.zeros(2), clamp=(False, False, False, False),
                ids=[box_id], missed={}, cache={}, gt=bool(gt), conf=q, age=0, hits=1, done=False)


def _predict(tr: Dict, hw) -> np.ndarray:
    """linear.py:71-83 (+ get_state :59-69): advance by the velocity, report the clamped box, keep the raw one."""
    tr['age'] += 1
    tr['last'] = tr['box'].copy()
    tr['box'][:2] += tr['v']
    clamped, top, down, left, right = clamp_center_box(tr['box'], hw)
    tr['clamp'] = (top, down, left, right)
    pred = np.zeros(5, dtype=f32)
    pred[:4] = clamped
    pred[4] = tr['cls']
    tr['pred'] = pred
    return pred.copy()


def _velocity(tr: Dict, det: np.ndarray) -> np.ndarray:
    """linear.py:108-124: centre displacement, or the displacement of the free edge when the prediction was clamped."""
    v = (det[:2] - tr['last'][:2]).astype(f32)
    top, down, left, right = tr['clamp']
    if not (top or down or left or right):
        return v
    assert not (top and down) and not (left and right)
    ox1, oy1, ox2, oy2 = _xyxy(tr['last'])
    nx1, ny1, nx2, ny2 = _xyxy(det)
    if top:
        v[1] = ny2 - oy2
    if down:
        v[1] = ny1 - oy1
    if left:
        v[0] = nx2 - ox2
    if right:
        v[0] = nx1 - ox1
    return v


def _update(tr: Dict, det: np.ndarray, box_id: int, gt: bool, q: float) -> None:
    """linear.py:85-106."""
    assert det[4] == tr['cls']
    tr['hits'] = tr['age'] + 1
    tr['v'] = _velocity(tr, det)
    tr['box'] = det[:4].astype(f32).copy()
    tr['ids'].append(box_id)
    tr['gt'] = tr['gt'] or bool(gt)
    w = q * (1. - q ** tr['age']) / (1. - q)
    tr['conf'] = (w * tr['conf'] + 1.) / (w + 1.)
    tr['missed'].update(tr['cache'])
    tr['cache'] = {}


def linear_track(frames: List[np.ndarray], frame_idx: List[int], gt_flags: List[np.ndarray], hw: Tuple[int, int], q: float = 0.9,
                 min_conf: float = 0.55, iou_thr: float = 0.45):
    """LinearTracker over a sequence (linear.py:196-292, pseudo_labeler.py:211-225).  frames[k]: [n,5] float32 centre
    boxes (x, y, w, h, cls) of frame frame_idx[k].  Returns (box_id -> track record, finished tracks in deletion order)."""
    live: List[Dict] = []
    finished: List[Dict] = []
    owner: Dict[int, Dict] = {}
    n_boxes = 0

    def retire(i: int, done: bool):
        tr = live.pop(i)
        tr['done'] = done
        tr.pop('cache', None)
        finished.append(tr)
        for b in tr['ids']:
            owner[b] = tr

    pos = {f: k for k, f in enumerate(frame_idx)}
    for f in range(max(frame_idx) + 1):
        if f in pos:
            dets = np.asarray(frames[pos[f]], dtype=f32)
            gts = np.asarray(gt_flags[pos[f]], dtype=bool)
        else:
            dets, gts = np.empty((0, 5), f32), np.zeros((0,), bool)
        if len(dets) == 0 and len(live) == 0:
            continue
        preds, neg_conf, dead = [], [], []
        for i, tr in enumerate(live):
            if tr['box'][2] * tr['box'][3] <= 0.:
                dead.append(i)
                continue
            preds.append(_predict(tr, hw))
            neg_conf.append(-tr['conf'])
        for i in reversed(dead):
            retire(i, True)
        order = np.argsort(neg_conf)
        pairs: List[Tuple[int, int]] = []
        if len(preds) and len(dets):
            iou = pairwise_iou_center(np.stack(preds, 0), dets)
            if iou.max() > 0:
                pairs = greedy_assign(iou, order, iou_thr)
        hit_t = {p[0] for p in pairs}
        hit_d = {p[1] for p in pairs}
        for ti, di in pairs:
            _update(live[ti], dets[di], n_boxes + di, gts[di], q)
        for ti in range(len(preds)):
            if ti not in hit_t:                      # linear.py:126-134
                tr = live[ti]
                tr['conf'] *= q
                if not gts.any():
                    tr['cache'][f] = tr['pred'].copy()
        for di in range(len(dets)):
            if di not in hit_d:
                live.append(_new_track(dets[di], n_boxes + di, gts[di], q))
        for i in reversed(range(len(live))):
            if live[i]['conf'] < min_conf:
                retire(i, True)
        n_boxes += len(dets)
    for i in reversed(range(len(live))):             # tracker.py:34-39: unfinished tracks are never filtered
        retire(i, False)
    return owner, finished


def rows_to_center(rows: np.ndarray) -> np.ndarray:
    """ObjectLabels rows (t,x,y,w,h,cls,cls_conf,obj) -> [n,5] centre boxes with class (labels.py:521-531)."""
    rows = np.asarray(rows, f32)
    return np.stack((rows[:, 1] + f32(0.5) * rows[:, 3], rows[:, 2] + f32(0.5) * rows[:, 4], rows[:, 3], rows[:, 4], rows[:, 5]), -1)


def track(frames_rows: List[np.ndarray], frame_idx: List[int], hw: Tuple[int, int], min_track_len: int = 6, inpaint: bool = False):
    """EventSeqData._track (pseudo_labeler.py:201-266).  Returns (indices of boxes to ignore — numbered frame by frame
    in the given order —, {frame: [n,8] in-painted label rows})."""
    if len(frames_rows) == 0:
        return [], {}
    owner, finished = linear_track([rows_to_center(r) for r in frames_rows], list(frame_idx), [np.asarray(r)[:, 0] > 0 for r in frames_rows], hw)
    remove, b = [], 0
    for r in frames_rows:
        for _ in range(len(r)):
            tr = owner[b]
            if tr['done'] and not tr['gt'] and tr['hits'] < min_track_len:
                remove.append(b)
            b += 1
    if not inpaint:
        return remove, {}
    holes: Dict[int, List[np.ndarray]] = {}
    for tr in finished:
        if tr['done'] and not tr['gt'] and tr['hits'] < min_track_len:
            continue
        for f, box in tr['missed'].items():
            holes.setdefault(f, []).append(box)
    out = {}
    for f, boxes in holes.items():
        c = np.stack(boxes)
        rows = np.zeros((c.shape[0], 8), f32)
        rows[:, 1] = c[:, 0] - c[:, 2] / 2.
        rows[:, 2] = c[:, 1] - c[:, 3] / 2.
        rows[:, 3:6] = c[:, 2:5]
        out[f] = rows
    return remove, out


def track_filter(frames_rows: List[np.ndarray], frame_idx: List[int], hw: Tuple[int, int], min_track_len: int = 6,
                 method: str = 'forward or backward', inpaint: bool = True, ignore_label: int = 1024):
    """EventSeqData._track_filter (pseudo_labeler.py:268-333).  Returns (frame_idx, per-frame label rows) after marking
    boxes of short tracks as `ignore_label` and appending the in-painted boxes (also `ignore_label`)."""
    frames_rows = [np.asarray(r, f32).copy() for r in frames_rows]
    frame_idx = list(frame_idx)
    if len(frames_rows) == 0 or min_track_len <= 0:
        return frame_idx, frames_rows
    remove, holes = track(frames_rows, frame_idx, hw, min_track_len, inpaint)
    if 'backward' in method:
        rev_rows = [r[::-1] for r in frames_rows[::-1]]
        rev_idx = [max(frame_idx) - i for i in frame_idx[::-1]]
        back, _ = track(rev_rows, rev_idx, hw, min_track_len, False)
        n = sum(len(r) for r in frames_rows)
        back = [n - i - 1 for i in back[::-1]]
        remove = list(set(remove) & set(back))        # ignored only if both directions say so
    remove = set(remove)
    b = 0
    for r in frames_rows:
        for i in range(len(r)):
            if b in remove:
                assert r[i, 0] == 0, 'ground truth is never ignored'
                r[i, 5] = ignore_label
            b += 1
    for f in range(max(frame_idx) + 1):
        if f not in holes:
            continue
        rows = holes[f].copy()
        rows[:, 5] = ignore_label
        if f in frame_idx:
            k = frame_idx.index(f)
            frames_rows[k] = np.concatenate((frames_rows[k], rows), 0)
        else:
            frame_idx.append(f)
            frames_rows.append(rows)
    order = sorted(range(len(frame_idx)), key=lambda k: frame_idx[k])
    return [frame_idx[k] for k in order], [frames_rows[k] for k in order]
