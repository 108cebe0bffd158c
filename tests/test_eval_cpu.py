"""Evaluation oracle.  The Prophesee half (box filter, frame -> image windows, COCO records) is pinned against
fixtures produced by the reference's own filter_boxes / _match_times / _to_coco_format (tests/golden/make_golden.py::gen_eval).
pycocotools (third party, not vendored, absent here) cannot be run: its restatement is checked on cases with known answers."""
import os

import numpy as np

from helpers import EVAL_CASES, eval_inputs
from oracle import coco_eval as oc

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'eval_cases.npz'))


def test_prophesee_half_matches_reference_functions():
    assert int(GOLD['n']) == len(EVAL_CASES)
    for ci, (camera, ds2, F, seed) in enumerate(EVAL_CASES):
        gts, dts = eval_inputs(camera, ds2, F, seed)
        frames, ann, res = oc.to_coco_records(gts, dts, camera, ds2)
        np.testing.assert_array_equal(frames, GOLD[f'{ci}/frames'])
        np.testing.assert_array_equal(ann, GOLD[f'{ci}/ann'])        # image id, category id, box, area (fp32 product)
        np.testing.assert_array_equal(res, GOLD[f'{ci}/res'])


def _frame(boxes, cls, score=None, t=10 ** 6):
    d = dict(t=np.full(len(cls), t, np.int64), xywh=np.asarray(boxes, np.float32).reshape(-1, 4), cls=np.asarray(cls, np.int64))
    if score is not None:
        d['score'] = np.asarray(score, np.float32)
    return d


def test_perfect_detections_score_one():
    gts, dts = eval_inputs('gen1', False, 16, 5, det_noise=0.0, fp_rate=0.0, miss_p=0.0)
    for g, d in zip(gts, dts):          # detections = the ground truth itself
        d['xywh'], d['cls'], d['score'], d['t'] = g['xywh'].copy(), g['cls'].copy(), np.full(len(g['cls']), 0.9, np.float32), g['t'].copy()
    stats, precision, recall = oc.evaluate_frames(gts, dts, 'gen1', False)
    assert all(abs(stats[i] - 1.0) < 1e-12 for i in (0, 1, 2, 8))
    p100 = precision[..., 0, 2]                      # all areas, maxDets 100: 1 / (1 + eps) everywhere
    assert ((p100 == -1.0) | (abs(p100 - 1.0) < 1e-12)).all() and (p100 > 0).any()


def test_worked_example_ap50():
    """One image, one class, two gt boxes; detections by descending score: TP, FP, TP.  Precision/recall points (1, .5), (.5, .5),
    (2/3, 1); envelope (1, 2/3, 2/3); 51 recall thresholds <= 0.5 read 1, the other 50 read 2/3."""
    g = _frame([[10, 10, 40, 40], [100, 100, 50, 50]], [0, 0])
    d = _frame([[10, 10, 40, 40], [200, 20, 40, 40], [100, 100, 50, 50]], [0, 0, 0], [0.9, 0.8, 0.7])
    stats, precision, recall = oc.evaluate_frames([g], [d], 'gen1', False)
    want = (51 * 1.0 + 50 * (2.0 / 3.0)) / 101
    assert abs(precision[0, :, 0, 0, 2].mean() - want) < 1e-12
    assert abs(stats[1] - want) < 1e-12 and abs(stats[0] - want) < 1e-12       # exact boxes: the same at every IoU threshold
    assert recall[0, 0, 0, 2] == 1.0 and recall[0, 0, 0, 0] == 0.5              # maxDets = 1 sees only the first detection
    assert (precision[:, :, 1] == -1).all()                                      # class without ground truth stays undefined


def test_iou_threshold_and_area_ranges():
    """A detection shifted so that IoU = 0.6 is a TP up to the 0.60 threshold only; a 25x25 gt is 'small', 50x50 'medium'."""
    g = _frame([[0, 0, 50, 50], [200, 100, 25, 25]], [0, 0])
    d = _frame([[12.5, 0, 50, 50], [200, 100, 25, 25]], [0, 0], [0.9, 0.8])     # IoU(first) = 37.5*50 / (2*2500 - 1875) = 0.6
    stats, precision, recall = oc.evaluate_frames([g], [d], 'gen1', False)
    assert abs(oc.bbox_iou([[12.5, 0, 50, 50]], [[0, 0, 50, 50]])[0, 0] - 0.6) < 1e-12
    np.testing.assert_array_equal(recall[:, 0, 2, 2], [1, 1, 1] + [0] * 7)       # medium: matched at 0.50, 0.55, 0.60
    np.testing.assert_array_equal(recall[:, 0, 1, 2], [1] * 10)                  # small: exact box
    assert abs(stats[3] - 1.0) < 1e-12 and abs(stats[4] - 0.3) < 1e-12 and stats[5] == -1    # AP_S, AP_M, AP_L (no large gt)


def test_filter_drops_frames_and_boxes():
    g_early = _frame([[0, 0, 50, 50]], [0], t=400000)                            # before 0.5 s: no image
    g_small = _frame([[0, 0, 8, 60]], [0])                                       # side < 10: no image
    g_ok = _frame([[0, 0, 50, 50], [60, 60, 9, 9]], [0, 1])                      # second box filtered
    d = _frame([[0, 0, 50, 50]], [0], [0.5])
    frames, ann, res = oc.to_coco_records([g_early, g_small, g_ok], [dict(d, t=np.full(1, 400000)), d, d], 'gen1', False)
    assert frames.tolist() == [2] and len(ann) == 1 and len(res) == 1
    assert oc.evaluate_frames([g_early], [dict(d, t=np.full(1, 400000))], 'gen1', False) is None


def test_per_class_evaluation_and_host_summarize():
    """evaluator.py:95-105: the per-class numbers come from buffers filtered to that class BEFORE everything else (frames whose only
    ground truth is of another class stop being images); the host mirror's summarize equals the oracle's on the same arrays."""
    from leod_b200.utils.evaluation.prophesee import evaluator as ev
    g0 = _frame([[10, 10, 40, 40]], [0])                      # frame with a class-0 box only
    g1 = _frame([[50, 60, 45, 35]], [1], t=2 * 10 ** 6)       # frame with a class-1 box only
    d0 = _frame([[10, 10, 40, 40], [100, 100, 40, 40]], [0, 1], [0.9, 0.8])            # class-1 false positive in frame 0
    d1 = _frame([[50, 60, 45, 35]], [1], [0.7], t=2 * 10 ** 6)
    all_stats, p_all, r_all = oc.evaluate_frames([g0, g1], [d0, d1], 'gen1', False)
    c1_stats, p1, r1 = oc.evaluate_frames([g0, g1], [d0, d1], 'gen1', False, only_class=1)
    # overall: class 1 sees its false positive (score 0.8) ranked above the true positive (0.7): AP50 of class 1 = 0.5
    assert abs(p_all[0, :, 1, 0, 2].mean() - 0.5) < 1e-12 and abs(p_all[0, :, 0, 0, 2].mean() - 1.0) < 1e-12
    # per class: frame 0 has no class-1 ground truth, so it is no image and its false positive is never seen
    assert abs(c1_stats[1] - 1.0) < 1e-12 and (p1[:, :, 0] == -1).all()
    np.testing.assert_array_equal(ev.summarize(p_all, r_all), all_stats)
    np.testing.assert_array_equal(ev.IOU_THRS, oc.IOU_THRS)
    np.testing.assert_array_equal(ev.REC_THRS, oc.REC_THRS)
    assert ev.filter_thresholds('gen4', True) == oc.filter_thresholds('gen4', True) == (500000, 30, 10)


def _ap_independent(gts, dts, k, thr):
    """A second, differently structured computation of one COCO number (class k, one IoU threshold, area 'all', maxDets 100):
    vectorised IoU, per-image greedy matching by masked argmax, precision envelope by a reversed running maximum."""
    scores, tps, n_gt = [], [], 0
    skip, diag, side = oc.filter_thresholds('gen1', False)
    for g, d in zip(gts, dts):
        gk = (g['t'] > skip) & (g['xywh'][:, 2] ** 2 + g['xywh'][:, 3] ** 2 >= diag ** 2) & (g['xywh'][:, 2] >= side) & (g['xywh'][:, 3] >= side)
        dk = (d['t'] > skip) & (d['xywh'][:, 2] ** 2 + d['xywh'][:, 3] ** 2 >= diag ** 2) & (d['xywh'][:, 2] >= side) & (d['xywh'][:, 3] >= side)
        if not gk.any():
            continue
        gb = g['xywh'][gk & (g['cls'] == k)].astype(np.float64)
        db = d['xywh'][dk & (d['cls'] == k)].astype(np.float64)
        ds = d['score'][dk & (d['cls'] == k)].astype(np.float64)
        order = np.argsort(-ds, kind='stable')[:100]
        db, ds = db[order], ds[order]
        n_gt += len(gb)
        if len(gb) and len(db):
            x1 = np.maximum(db[:, None, 0], gb[None, :, 0]); y1 = np.maximum(db[:, None, 1], gb[None, :, 1])
            x2 = np.minimum(db[:, None, 0] + db[:, None, 2], gb[None, :, 0] + gb[None, :, 2])
            y2 = np.minimum(db[:, None, 1] + db[:, None, 3], gb[None, :, 1] + gb[None, :, 3])
            w, h = x2 - x1, y2 - y1
            inter = np.where((w > 0) & (h > 0), w * h, 0.0)
            iou = inter / (db[:, None, 2] * db[:, None, 3] + gb[None, :, 2] * gb[None, :, 3] - inter)
        else:
            iou = np.zeros((len(db), len(gb)))
        free = np.ones(len(gb), bool)
        for i in range(len(db)):
            cand = np.where(free & (iou[i] >= min(thr, 1 - 1e-10)))[0] if len(gb) else []
            hit = len(cand) > 0
            if hit:
                best = cand[np.flatnonzero(iou[i, cand] == iou[i, cand].max())[-1]]      # equal IoU: the later ground-truth box wins
                free[best] = False
            scores.append(ds[i]); tps.append(hit)
    if n_gt == 0:
        return -1.0
    scores, tps = np.array(scores), np.array(tps, bool)
    o = np.argsort(-scores, kind='stable')
    tp = np.cumsum(tps[o]).astype(float); fp = np.cumsum(~tps[o]).astype(float)
    rc, pr = tp / n_gt, tp / (tp + fp + np.spacing(1))
    env = np.maximum.accumulate(pr[::-1])[::-1] if len(pr) else pr
    idx = np.searchsorted(rc, oc.REC_THRS, side='left')
    q = np.array([env[i] if i < len(env) else 0.0 for i in idx])
    return q.mean()


def test_oracle_agrees_with_independent_ap_computation():
    for seed in (31, 32, 33):
        gts, dts = eval_inputs('gen1', False, 40, seed, max_gt=6, fp_rate=4.0)
        _, precision, _ = oc.evaluate_frames(gts, dts, 'gen1', False)
        for k in (0, 1):
            for ti, thr in ((0, 0.5), (5, oc.IOU_THRS[5]), (9, oc.IOU_THRS[9])):
                want = _ap_independent(gts, dts, k, thr)
                got = precision[ti, :, k, 0, 2]
                got = -1.0 if (got == -1).all() else got.mean()
                assert abs(got - want) < 1e-12, (seed, k, thr, got, want)
