"""Pins oracle/ (the CPU restatement) to fixtures produced by running the reference itself
(tests/golden/make_golden.py).  CPU-only."""
import os

import numpy as np
import pytest
import torch

from oracle import binning, optim, postprocess as opp, rvt, yolox
from oracle.config import ModelCfg
from helpers import GOLDEN, load_net_fixture, rel_err

TOL = 2e-5


@pytest.fixture(scope='module')
def net():
    return load_net_fixture()


def test_backbone_unroll_matches_reference(net):
    z, cfg, sd, d = net
    x = torch.from_numpy(z['x']).float()
    states = None
    with torch.no_grad():
        for t in range(d['T']):
            feats, states = rvt.backbone_forward(x[t], states, sd, cfg)
            for s in (1, 2, 3, 4):
                assert rel_err(feats[s], z[f'eval/feat{s}_t{t}']) < TOL, (s, t)
    for s in range(4):
        assert rel_err(states[s][1], z[f'eval/c{s}']) < TOL


def test_head_decode_and_postprocess_match_reference(net):
    z, cfg, sd, d = net
    T = d['T']
    feats = {s: torch.from_numpy(z[f'eval/feat{s}_t{T - 1}']) for s in (1, 2, 3, 4)}
    with torch.no_grad():
        preds, losses = yolox.detect_forward(feats, sd, cfg)
    assert losses is None
    assert rel_err(preds, z['eval/preds']) < TOL
    # NMS on the REFERENCE predictions so index lists must agree exactly
    dets = opp.postprocess(z['eval/preds'], cfg.num_classes, 0.001, 0.45)
    for b, det in enumerate(dets):
        ref = z[f'eval/det{b}']
        assert det.shape == ref.shape, (b, det.shape, ref.shape)
        np.testing.assert_array_equal(det, ref)


@pytest.mark.parametrize('tag,thr', [('plain', None), ('ignore', None), ('thresh', [0.7, 0.35])])
def test_train_losses_and_grads_match_reference(net, tag, thr):
    z, cfg, sd, d = net
    cfg = ModelCfg(**{**cfg.__dict__, 'ignore_bbox_thresh': thr})
    x = torch.from_numpy(z['x']).float()
    p = {k: v.clone().requires_grad_(v.is_floating_point() and 'running' not in k) for k, v in sd.items()}
    states = None
    for t in range(d['T']):
        feats, states = rvt.backbone_forward(x[t], states, p, cfg)
    bn_state = {}
    preds, losses = yolox.detect_forward(feats, p, cfg, targets=torch.from_numpy(z[f'train_{tag}/labels']),
                                         training=True, bn_state=bn_state)
    for k in ('loss', 'iou_loss', 'conf_loss', 'cls_loss', 'num_fg'):
        assert abs(float(losses[k]) - float(z[f'train_{tag}/{k}'])) < 2e-5 * max(1.0, abs(float(z[f'train_{tag}/{k}']))), k
    assert rel_err(preds.detach(), z[f'train_{tag}/preds']) < TOL
    losses['loss'].backward()
    for key in z.files:
        if key.startswith(f'train_{tag}/grad/'):
            name = key.split('/grad/')[1]
            assert rel_err(p[name].grad, z[key]) < 2e-4, name
    if tag == 'plain':
        for key in z.files:
            if key.startswith('train_plain/bn/'):
                assert rel_err(bn_state[key.split('/bn/')[1]], z[key]) < TOL


def test_nms_matches_torchvision_fixture():
    z = np.load(os.path.join(GOLDEN, 'nms_cases.npz'))
    for i in range(int(z['n'])):
        keep = opp.batched_nms_indices(z[f'{i}/boxes'], z[f'{i}/scores'], z[f'{i}/cls'], float(z[f'{i}/thr']))
        np.testing.assert_array_equal(keep, z[f'{i}/keep'], err_msg=f'case {i}')


def test_pred2label_and_tta_merge_match_reference():
    z = np.load(os.path.join(GOLDEN, 'pred2label_cases.npz'))
    for ci in range(int(z['n'])):
        dets = [z[f'{ci}/det{b}'] for b in range(4)]
        hw = tuple(int(v) for v in z[f'{ci}/hw'])
        labs = opp.pred2label(dets, [float(v) for v in z[f'{ci}/obj_thresh']],
                              [float(v) for v in z[f'{ci}/cls_thresh']], frame_hw=hw)
        for b, lab in enumerate(labs):
            ref = z[f'{ci}/label{b}']
            assert lab.shape == ref.shape, (ci, b)
            np.testing.assert_array_equal(lab, ref)
        merged = opp.tta_merge(z[f'{ci}/tta_in'], 0.1, 0.45)
        assert merged.shape == z[f'{ci}/tta_out'].shape
        np.testing.assert_allclose(merged, z[f'{ci}/tta_out'], rtol=0, atol=1e-4)


def test_binning_matches_reference():
    z = np.load(os.path.join(GOLDEN, 'binning_cases.npz'))
    for i in range(int(z['n'])):
        bins, H, W, cutoff, fast = (int(v) for v in z[f'{i}/cfg'])
        rep = binning.stacked_histogram(z[f'{i}/x'], z[f'{i}/y'], z[f'{i}/p'], z[f'{i}/t'], bins, H, W,
                                        None if cutoff < 0 else cutoff, bool(fast))
        np.testing.assert_array_equal(rep, z[f'{i}/rep'], err_msg=f'case {i}')


def test_ema_and_adamw_match_reference():
    z = np.load(os.path.join(GOLDEN, 'optim_cases.npz'))
    n = len([k for k in z.files if k.startswith('ema/student')])
    for step in (0, 5, 5000):
        student = [torch.from_numpy(z[f'ema/student{i}']) for i in range(n)]
        teacher = [torch.from_numpy(z[f'ema/teacher{i}']).clone() for i in range(n)]
        optim.ema_update(teacher, student, step, 0.999)
        for i in range(n):
            np.testing.assert_allclose(teacher[i].numpy(), z[f'ema/step{step}/teacher{i}'], rtol=1e-6, atol=1e-7)
    p = torch.from_numpy(z['adamw/p0']).clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for s in range(3):
        optim.adamw_step(p, torch.from_numpy(z[f'adamw/g{s}']), m, v, s + 1, lr=2e-4, clip_value=1.0)
        np.testing.assert_allclose(p.numpy(), z[f'adamw/p{s + 1}'], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize('tag', ['tiny_gen1_c10', 'small_gen1', 'base_gen4', 'small_gen1_l8'])
def test_full_size_oracle_matches_reference_run(tag):
    """BASELINE configs[0] (RVT-tiny, 10 input channels, 240x304, one frame) and the configs[1] model at batch 1: the
    oracle against outputs of the reference itself (tests/golden/make_golden.py: gen_fullsize, name-seeded weights)."""
    from helpers import FULLSIZE_CASES, canon_rows, det_events, det_state_value
    from leod_b200.config import DATASETS, make_model_cfg
    from leod_b200.models.detection.yolox_extension.models.detector import YoloXDetector
    from oracle import postprocess as opp, rvt, yolox
    from oracle.config import ModelCfg
    z = np.load(os.path.join(GOLDEN, 'fullsize_cases.npz'))
    size, dataset, inch, B, L = FULLSIZE_CASES[tag]
    torch.set_num_threads(8)
    ocfg = ModelCfg.named(size, dataset)
    ocfg.input_channels = inch
    # parameter names / shapes come from the product's (reference-compatible) state dict
    m = YoloXDetector(make_model_cfg(size=size, dataset=dataset, input_channels=inch, compute_dtype='fp32'))
    sd = {k: det_state_value(k, v.shape).to(v.dtype) for k, v in m.state_dict().items()}
    fh, fw = DATASETS[dataset]['frame_hw']
    x = det_events(7, (L, B, inch, fh, fw))
    states = None
    with torch.no_grad():
        for t in range(L):
            feats, states = rvt.backbone_forward(rvt.pad_input(x[t].float(), DATASETS[dataset]['in_res_hw']), states, sd, ocfg)
        preds = yolox.detect_forward(feats, sd, ocfg)[0]
    assert np.abs(feats[4].numpy() - z[f'{tag}/feat4']).max() < 2e-4 * np.abs(z[f'{tag}/feat4']).max()
    assert np.abs(states[3][1].numpy() - z[f'{tag}/c4']).max() < 2e-4 * max(1.0, np.abs(z[f'{tag}/c4']).max())
    np.testing.assert_allclose(preds.numpy(), z[f'{tag}/preds'], rtol=2e-4, atol=2e-3)
    dets = opp.postprocess(z[f'{tag}/preds'], ocfg.num_classes, 0.001, 0.45)
    for b in range(B):   # anchors over the zero-padded rows produce exactly equal scores: compare order-independently there
        np.testing.assert_allclose(canon_rows(dets[b]), canon_rows(z[f'{tag}/det{b}']), rtol=0, atol=1e-5)


def _split(rows, counts):
    out, o = [], 0
    for n in counts:
        out.append(rows[o:o + int(n)])
        o += int(n)
    return out


def test_tracking_filter_matches_reference():
    """oracle/tracking.py against EventSeqData._track / _track_filter + LinearTracker of the reference run on synthetic
    sequences (tests/golden/make_golden.py: gen_tracking): ignored boxes, in-painted boxes, final per-frame labels."""
    from oracle import tracking as otr
    z = np.load(os.path.join(GOLDEN, 'tracking_cases.npz'))
    for ci in range(int(z['n'])):
        hw = tuple(int(v) for v in z[f'{ci}/hw'])
        frame_idx = [int(v) for v in z[f'{ci}/frame_idx']]
        frames = _split(z[f'{ci}/rows'], z[f'{ci}/counts'])
        for inpaint in (False, True):
            remove, holes = otr.track(frames, frame_idx, hw, min_track_len=6, inpaint=inpaint)
            np.testing.assert_array_equal(np.array(sorted(remove), np.int64), z[f'{ci}/remove_idx_inpaint{int(inpaint)}'])
            if inpaint:
                keys = sorted(holes.keys())
                np.testing.assert_array_equal(np.array(keys, np.int64), z[f'{ci}/inpaint_frames'])
                np.testing.assert_array_equal(np.array([len(holes[k]) for k in keys], np.int64), z[f'{ci}/inpaint_counts'])
                got = np.concatenate([holes[k] for k in keys], 0) if keys else np.zeros((0, 8), np.float32)
                np.testing.assert_array_equal(got, z[f'{ci}/inpaint_rows'])
        for tag, method in (('f', 'forward'), ('fb', 'forward or backward')):
            fidx, rows = otr.track_filter(frames, frame_idx, hw, min_track_len=6, method=method, inpaint=True, ignore_label=1024)
            np.testing.assert_array_equal(np.array(fidx, np.int64), z[f'{ci}/final_{tag}_frame_idx'])
            np.testing.assert_array_equal(np.array([len(r) for r in rows], np.int64), z[f'{ci}/final_{tag}_counts'])
            np.testing.assert_array_equal(np.concatenate(rows, 0), z[f'{ci}/final_{tag}_rows'])


def test_label_records_match_reference_bytes():
    """oracle/labels_io.py: the 40-byte BBOX_DTYPE records and the two index arrays the reference writes for a sequence
    (EventSeqData._summarize), byte for byte."""
    from oracle import labels_io
    z = np.load(os.path.join(GOLDEN, 'tracking_cases.npz'))
    for ci in range(int(z['n'])):
        frames = _split(z[f'{ci}/final_fb_rows'], z[f'{ci}/final_fb_counts'])
        recs, lbl_idx, repr_idx = labels_io.summarize(z[f'{ci}/final_fb_frame_idx'], frames)
        assert recs.dtype.names == labels_io.BBOX_DTYPE.names and recs.dtype.itemsize in (36, 40)   # see labels_io.summarize
        np.testing.assert_array_equal(np.frombuffer(recs.tobytes(), np.uint8), z[f'{ci}/packed_bytes'])
        np.testing.assert_array_equal(lbl_idx, z[f'{ci}/objframe_idx_2_label_idx'])
        np.testing.assert_array_equal(repr_idx, z[f'{ci}/objframe_idx_2_repr_idx'])
