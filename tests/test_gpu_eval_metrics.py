"""Device evaluation (leod_coco_eval behind PropheseeEvaluator) against the oracle: the precision / recall arrays of COCOeval.eval
must be IDENTICAL (fp64, assert_array_equal) on seeded per-frame buffers — both datasets, score ties, frames the Prophesee filter
removes, frames without detections, per-class evaluation — and on a 3 000-frame buffer (sort across tiles, many chunks)."""
import numpy as np
import pytest
import torch

from helpers import EVAL_CASES, eval_inputs
from oracle import coco_eval as oc

pytestmark = pytest.mark.gpu


def _records(frames, with_score):
    """Buffers as to_prophesee builds them (io/box_loading.py:58-107): structured arrays per frame."""
    dt = np.dtype([('t', '<i8'), ('x', '<f4'), ('y', '<f4'), ('w', '<f4'), ('h', '<f4'), ('class_id', '<u4'), ('class_confidence', '<f4')])
    out = []
    for d in frames:
        r = np.zeros(len(d['cls']), dtype=dt)
        r['t'], r['x'], r['y'], r['w'], r['h'] = d['t'], d['xywh'][:, 0], d['xywh'][:, 1], d['xywh'][:, 2], d['xywh'][:, 3]
        r['class_id'] = d['cls']
        r['class_confidence'] = d['score'] if with_score else 1.0
        out.append(r)
    return out


def _device_arrays(gts, dts, camera, ds2, only=-1):
    from leod_b200.utils.evaluation.prophesee.evaluator import FrameBoxes, coco_eval_device
    K = 3 if camera == 'gen4' else 2
    g = FrameBoxes(_records(gts, False), torch.device('cuda'), with_score=False)
    d = FrameBoxes(_records(dts, True), torch.device('cuda'), with_score=True)
    return coco_eval_device(g, d, len(gts), K, camera, ds2, only_class=only)


@pytest.mark.parametrize('case', range(len(EVAL_CASES)))
def test_precision_recall_identical_to_oracle(case):
    camera, ds2, F, seed = EVAL_CASES[case]
    gts, dts = eval_inputs(camera, ds2, F, seed)
    K = 3 if camera == 'gen4' else 2
    for only in [None] + list(range(K)):
        stats, precision, recall = oc.evaluate_frames(gts, dts, camera, ds2, only_class=only)
        p, r, counts = _device_arrays(gts, dts, camera, ds2, -1 if only is None else only)
        np.testing.assert_array_equal(p, precision, err_msg=f'precision, only_class={only}')
        np.testing.assert_array_equal(r, recall, err_msg=f'recall, only_class={only}')
        assert counts[2] == 0


def test_large_buffer_and_evaluator_api():
    from leod_b200.utils.evaluation.prophesee.evaluator import OUT_KEYS, PropheseeEvaluator
    F = 3000
    gts, dts = eval_inputs('gen1', False, F, 11, max_gt=5, fp_rate=12.0)
    stats, precision, recall = oc.evaluate_frames(gts, dts, 'gen1', False)
    p, r, counts = _device_arrays(gts, dts, 'gen1', False)
    np.testing.assert_array_equal(p, precision)
    np.testing.assert_array_equal(r, recall)
    ev = PropheseeEvaluator('gen1', downsample_by_2=False)
    for i in range(0, F, 500):                                                       # buffers filled batch by batch (evaluator.py:63-67)
        ev.add_labels(_records(gts[i:i + 500], False))
        ev.add_predictions(_records(dts[i:i + 500], True))
    m = ev.evaluate_buffer(img_height=240, img_width=304)
    for i, k in enumerate(OUT_KEYS):
        assert m[k] == float(stats[i]), k
    for ci, name in enumerate(('car', 'ped')):
        s_c = oc.evaluate_frames(gts, dts, 'gen1', False, only_class=ci)[0]
        for i, k in enumerate(OUT_KEYS):
            assert m[f'{k}_{name}'] == float(s_c[i]), (k, name)


def test_no_detections_and_more_than_100_per_frame():
    from leod_b200.utils.evaluation.prophesee.evaluator import PropheseeEvaluator
    gts, dts = eval_inputs('gen1', False, 10, 21)
    empty = [dict(t=d['t'][:0], xywh=d['xywh'][:0], cls=d['cls'][:0], score=d['score'][:0]) for d in dts]
    ev = PropheseeEvaluator('gen1', downsample_by_2=False)
    ev.add_labels(_records(gts, False))
    ev.add_predictions(_records(empty, True))
    assert all(v == 0.0 for v in ev.evaluate_buffer(240, 304).values())              # coco_eval.py:96-99
    gts, dts = eval_inputs('gen1', False, 6, 22, fp_rate=260.0)                      # > 100 detections of a class in a frame: top-100 cut
    assert max(len(d['cls']) for d in dts) > 200
    stats, precision, recall = oc.evaluate_frames(gts, dts, 'gen1', False)
    p, r, _ = _device_arrays(gts, dts, 'gen1', False)
    np.testing.assert_array_equal(p, precision)
    np.testing.assert_array_equal(r, recall)
