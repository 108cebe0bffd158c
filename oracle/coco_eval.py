"""Oracle: Prophesee box filter + COCO bounding-box evaluation, numpy (float64 where pycocotools uses float64).

Restates, for the way LEOD calls it (one entry per labelled frame in the evaluator buffers, utils/evaluation/prophesee/evaluator.py:
73-110):
  utils/evaluation/prophesee/io/box_filtering.py:18-36 (filter_boxes), utils/evaluation/prophesee/evaluation.py:5-42 (thresholds
  per camera, halved when downsampled), utils/evaluation/prophesee/metrics/coco_eval.py:32-120 (evaluate_detection: a frame becomes an
  image only if a ground-truth box survives the filter; _match_times with one timestamp per frame; _to_coco_format: score =
  class_confidence, area = w * h in float32) and the third-party evaluator behind it, pycocotools.cocoeval.COCOeval (iouType 'bbox';
  pinned `pycocotools==2.0.6` by the reference's environment, NOT vendored and NOT installed in this image):
    evaluate/evaluateImg  — per (image, category, area range): detections by descending score (stable), top 100; ground truth with
                            out-of-range area ignored and sorted last; greedy matching per IoU threshold 0.50:0.05:0.95 with
                            IoU in float64 (maskApi bbIou); unmatched detections with out-of-range area ignored
    accumulate            — per (category, area range, maxDets 1/10/100): scores merged over images (stable), cumulative TP/FP,
                            precision envelope, precision sampled at 101 recall thresholds (searchsorted left)
    summarize             — AP, AP50, AP75, AP_S, AP_M, AP_L (mean of the precision entries > -1), AR@1/10/100, AR_S/M/L
Parity status of THIS file: the Prophesee half is pinned by fixtures made from the reference's own functions (tests/golden/
eval_cases.npz); the COCOeval half is PARITY UNPINNED — pycocotools cannot be run here; it is checked against analytic cases and
against a second, differently structured AP computation on seeded buffers (tests/test_eval_cpu.py).
TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
import numpy as np

IOU_THRS = np.linspace(.5, 0.95, int(np.round((0.95 - .5) / .05)) + 1, endpoint=True)
REC_THRS = np.linspace(.0, 1.00, int(np.round((1.00 - .0) / .01)) + 1, endpoint=True)
MAX_DETS = (1, 10, 100)
AREA_RNG = ((0 ** 2, 1e5 ** 2), (0 ** 2, 32 ** 2), (32 ** 2, 96 ** 2), (96 ** 2, 1e5 ** 2))


def filter_thresholds(camera: str, downsampled_by_2: bool):
    """evaluation.py:24-33 -> (skip_ts, min_box_diag, min_box_side)."""
    diag, side = (60, 20) if camera == 'gen4' else (30, 10)
    if downsampled_by_2:
        diag, side = diag // 2, side // 2
    return int(5e5), diag, side


def filter_mask(t, w, h, skip_ts, min_box_diag, min_box_side):
    """box_filtering.py:31-36 on float32 w/h (BBOX_DTYPE fields) and int64 t."""
    w, h = np.asarray(w, np.float32), np.asarray(h, np.float32)
    return (np.asarray(t) > skip_ts) & ((w ** 2 + h ** 2) >= min_box_diag ** 2) & (w >= min_box_side) & (h >= min_box_side)


def bbox_iou(d, g):
    """maskApi.c bbIou, iscrowd = 0: d [nd,4], g [ng,4] xywh (float64) -> [nd, ng]."""
    d, g = np.asarray(d, np.float64).reshape(-1, 4), np.asarray(g, np.float64).reshape(-1, 4)
    out = np.zeros((len(d), len(g)))
    for j in range(len(g)):
        ga = g[j, 2] * g[j, 3]
        for i in range(len(d)):
            da = d[i, 2] * d[i, 3]
            w = min(d[i, 2] + d[i, 0], g[j, 2] + g[j, 0]) - max(d[i, 0], g[j, 0])
            if w <= 0:
                continue
            h = min(d[i, 3] + d[i, 1], g[j, 3] + g[j, 1]) - max(d[i, 1], g[j, 1])
            if h <= 0:
                continue
            inter = w * h
            out[i, j] = inter / (da + ga - inter)
    return out


def evaluate_img(gt_xywh, gt_area, dt_xywh, dt_area, dt_score, max_det=100):
    """COCOeval.computeIoU + evaluateImg for one (image, category), all area ranges.
    -> dict(scores [D], dtm [A, T, D] bool, dt_ig [A, T, D] bool, gt_ig [A, G] bool (in the evaluator's sorted gt order))."""
    order = np.argsort(-np.asarray(dt_score, np.float64), kind='mergesort')[:max_det]
    dt_xywh, dt_area, dt_score = np.asarray(dt_xywh).reshape(-1, 4)[order], np.asarray(dt_area)[order], np.asarray(dt_score)[order]
    ious_all = bbox_iou(dt_xywh, gt_xywh)
    D, G, T = len(order), len(gt_area), len(IOU_THRS)
    dtm_all, dtig_all, gtig_all = [], [], []
    for lo, hi in AREA_RNG:
        g_ig = np.array([(a < lo or a > hi) for a in gt_area], bool)
        gtind = np.argsort(g_ig.astype(np.uint8), kind='mergesort')
        g_ig = g_ig[gtind]
        ious = ious_all[:, gtind] if G > 0 else ious_all
        gtm = np.zeros((T, G), bool)
        dtm = np.zeros((T, D), bool)
        dt_ig = np.zeros((T, D), bool)
        for ti, t in enumerate(IOU_THRS):
            for di in range(D):
                iou = min(t, 1 - 1e-10)
                m = -1
                for gi in range(G):
                    if gtm[ti, gi]:
                        continue
                    if m > -1 and not g_ig[m] and g_ig[gi]:
                        break
                    if ious[di, gi] < iou:
                        continue
                    iou = ious[di, gi]
                    m = gi
                if m == -1:
                    continue
                dt_ig[ti, di] = g_ig[m]
                dtm[ti, di] = True
                gtm[ti, m] = True
        a_out = np.array([(a < lo or a > hi) for a in dt_area], bool).reshape(1, D)
        dt_ig = dt_ig | (~dtm & np.repeat(a_out, T, 0))
        dtm_all.append(dtm)
        dtig_all.append(dt_ig)
        gtig_all.append(g_ig)
    return dict(scores=dt_score, dtm=np.stack(dtm_all), dt_ig=np.stack(dtig_all), gt_ig=np.stack(gtig_all))


def accumulate(per_img_cat, num_classes):
    """COCOeval.accumulate.  per_img_cat[k] = list (images in order) of evaluate_img results.
    -> precision [T, R, K, A, M] (-1 where undefined), recall [T, K, A, M]."""
    T, R, A, M = len(IOU_THRS), len(REC_THRS), len(AREA_RNG), len(MAX_DETS)
    precision = -np.ones((T, R, num_classes, A, M))
    recall = -np.ones((T, num_classes, A, M))
    for k in range(num_classes):
        E = per_img_cat[k]
        if len(E) == 0:
            continue
        for a in range(A):
            for m, max_det in enumerate(MAX_DETS):
                scores = np.concatenate([e['scores'][:max_det] for e in E]).astype(np.float64)
                inds = np.argsort(-scores, kind='mergesort')
                dtm = np.concatenate([e['dtm'][a][:, :max_det] for e in E], axis=1)[:, inds]
                dt_ig = np.concatenate([e['dt_ig'][a][:, :max_det] for e in E], axis=1)[:, inds]
                gt_ig = np.concatenate([e['gt_ig'][a] for e in E])
                npig = np.count_nonzero(gt_ig == 0)
                if npig == 0:
                    continue
                tps = np.logical_and(dtm, np.logical_not(dt_ig))
                fps = np.logical_and(np.logical_not(dtm), np.logical_not(dt_ig))
                tp_sum = np.cumsum(tps, axis=1).astype(dtype=float)
                fp_sum = np.cumsum(fps, axis=1).astype(dtype=float)
                for t, (tp, fp) in enumerate(zip(tp_sum, fp_sum)):
                    nd = len(tp)
                    rc = tp / npig
                    pr = tp / (fp + tp + np.spacing(1))
                    q = np.zeros((R,))
                    recall[t, k, a, m] = rc[-1] if nd else 0
                    pr = pr.tolist()
                    q = q.tolist()
                    for i in range(nd - 1, 0, -1):
                        if pr[i] > pr[i - 1]:
                            pr[i - 1] = pr[i]
                    idx = np.searchsorted(rc, REC_THRS, side='left')
                    try:
                        for ri, pi in enumerate(idx):
                            q[ri] = pr[pi]
                    except IndexError:
                        pass
                    precision[t, :, k, a, m] = np.array(q)
    return precision, recall


def summarize(precision, recall):
    """COCOeval.summarize (bbox): the 12 stats; the reference returns the first six (coco_eval.py:103-120)."""
    def _s(ap, iou=None, area=0, m=2):
        s = precision if ap else recall
        if iou is not None:
            s = s[np.where(iou == IOU_THRS)[0]]
        s = s[:, :, :, area, m] if ap else s[:, :, area, m]
        return -1 if len(s[s > -1]) == 0 else np.mean(s[s > -1])
    return np.array([_s(1), _s(1, .5), _s(1, .75), _s(1, area=1), _s(1, area=2), _s(1, area=3),
                     _s(0, m=0), _s(0, m=1), _s(0, m=2), _s(0, area=1), _s(0, area=2), _s(0, area=3)])


def evaluate_frames(gt_frames, dt_frames, camera='gen1', downsampled_by_2=False, num_classes=None, only_class=None):
    """evaluate_list + evaluate_detection over per-frame buffers.
    gt_frames[f] = dict(t int64 [n], xywh float32 [n,4], cls int [n]); dt_frames[f] = the same + score float32 [n].
    only_class: the per-category evaluation of evaluator.py:95-105 (boxes of other classes removed BEFORE everything else).
    -> (stats [12], precision, recall), or None when there is no detection at all (coco_eval.py:96-99 returns zeros)."""
    num_classes = num_classes or (3 if camera == 'gen4' else 2)
    thr = filter_thresholds(camera, downsampled_by_2)
    per = [[] for _ in range(num_classes)]
    n_dt = 0
    for g, d in zip(gt_frames, dt_frames):
        gk = filter_mask(g['t'], g['xywh'][:, 2], g['xywh'][:, 3], *thr)
        dk = filter_mask(d['t'], d['xywh'][:, 2], d['xywh'][:, 3], *thr)
        if only_class is not None:
            gk &= np.asarray(g['cls']) == only_class
            dk &= np.asarray(d['cls']) == only_class
        if not gk.any():          # np.unique(gt_boxes['t']) is empty: the frame contributes no image
            continue
        n_dt += int(dk.sum())
        for k in range(num_classes):
            gs, ds = gk & (np.asarray(g['cls']) == k), dk & (np.asarray(d['cls']) == k)
            gx, dx = g['xywh'][gs].astype(np.float32), d['xywh'][ds].astype(np.float32)
            per[k].append(evaluate_img(gx, (gx[:, 2] * gx[:, 3]).astype(np.float64), dx, (dx[:, 2] * dx[:, 3]).astype(np.float64),
                                       np.asarray(d['score'], np.float32)[ds]))
    if n_dt == 0:
        return None
    precision, recall = accumulate(per, num_classes)
    return summarize(precision, recall), precision, recall


def to_coco_records(gt_frames, dt_frames, camera='gen1', downsampled_by_2=False):
    """The Prophesee half alone (evaluation.py:35-38 filter, coco_eval.py:47-60 windows, :140-194 _to_coco_format): which frames become
    images and the annotation / result records COCOeval is fed.  -> (image frame indices, annotations [n, 7], results [m, 7]) with
    rows (image_id, category_id, x, y, w, h, area | score)."""
    thr = filter_thresholds(camera, downsampled_by_2)
    frames, ann, res = [], [], []
    for f, (g, d) in enumerate(zip(gt_frames, dt_frames)):
        gk = filter_mask(g['t'], g['xywh'][:, 2], g['xywh'][:, 3], *thr)
        dk = filter_mask(d['t'], d['xywh'][:, 2], d['xywh'][:, 3], *thr)
        if not gk.any():
            continue
        frames.append(f)
        im = len(frames)
        for b, c in zip(g['xywh'][gk].astype(np.float32), np.asarray(g['cls'])[gk]):
            ann.append([im, int(c) + 1, b[0], b[1], b[2], b[3], float(b[2] * b[3])])
        for b, c, s in zip(d['xywh'][dk].astype(np.float32), np.asarray(d['cls'])[dk], np.asarray(d['score'], np.float32)[dk]):
            res.append([im, int(c) + 1, b[0], b[1], b[2], b[3], float(s)])
    return np.array(frames, np.int64), np.array(ann, np.float64).reshape(-1, 7), np.array(res, np.float64).reshape(-1, 7)
