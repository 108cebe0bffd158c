"""Teacher pseudo-label sweep on the device (leod_b200/modules/pseudo_labeler.py) against the CPU oracle and the
fixtures generated from the reference (tests/golden/make_golden.py).  Kept indices / label rows are compared
bit-exactly where the arithmetic is fp32 box math; the network outputs feeding them are fp32 here (1e-3 gate)."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, load_net_fixture
from test_host_cpu import product_cfg

pytestmark = pytest.mark.gpu


def test_tta_merge_matches_reference_fixture_and_oracle():
    from oracle import postprocess as opp
    from leod_b200.modules.utils.ssod import tta_merge_packed
    z = np.load(os.path.join(GOLDEN, 'pred2label_cases.npz'))
    cases = [(z[f'{ci}/tta_in'], z[f'{ci}/tta_out']) for ci in range(int(z['n']))]
    # extra seeded cases checked against the oracle: ties, empty frames, a frame that carries ground truth
    rng = np.random.default_rng(5)
    for n in (0, 1, 7, 40, 200):
        lab = np.zeros((n, 8), np.float32)
        lab[:, 1] = rng.uniform(0, 250, n); lab[:, 2] = rng.uniform(0, 200, n)
        lab[:, 3] = rng.uniform(5, 60, n); lab[:, 4] = rng.uniform(5, 60, n)
        lab[:, 5] = rng.integers(0, 2, n); lab[:, 6] = rng.uniform(0.2, 1, n).round(1); lab[:, 7] = rng.uniform(0.2, 1, n).round(1)
        if n >= 7:
            lab[3] = lab[2]           # exact duplicate: equal score, IoU 1
        cases.append((lab, opp.tta_merge(lab, 0.1, 0.45)))
    gt = np.array([[1., 10, 10, 20, 20, 0, 1, 1], [1., 12, 12, 20, 20, 0, 1, 1]], np.float32)
    cases.append((gt, gt))            # GT frames pass through untouched (pseudo_labeler.py:50-53)
    nmax = max(1, max(len(c[0]) for c in cases))
    packed = np.zeros((len(cases), nmax, 8), np.float32)
    cnt = np.array([len(c[0]) for c in cases], np.int32)
    for i, (inp, _) in enumerate(cases):
        packed[i, :len(inp)] = inp
    out, n = tta_merge_packed(torch.from_numpy(packed).cuda(), torch.from_numpy(cnt).cuda(), 0.1, 0.45)
    for i, (_, ref) in enumerate(cases):
        assert int(n[i]) == ref.shape[0], (i, int(n[i]), ref.shape)
        np.testing.assert_array_equal(out[i, :int(n[i])].cpu().numpy(), ref.astype(np.float32))


def test_tta_postprocess_list_api():
    from leod_b200.data.labels import ObjectLabels
    from leod_b200.modules.pseudo_labeler import tta_postprocess
    a = torch.tensor([[0., 10, 10, 30, 30, 0, 0.9, 0.9], [0., 12, 11, 30, 30, 0, 0.8, 0.9], [0., 100, 100, 20, 20, 1, 0.9, 0.5]]).cuda()
    res = tta_postprocess([ObjectLabels(a, (240, 304)), ObjectLabels(a[:0], (240, 304))], conf_thre=0.1, nms_thre=0.45)
    assert len(res) == 2 and len(res[0]) == 2 and len(res[1]) == 0
    assert torch.equal(res[0].object_labels[:, 5].cpu(), torch.tensor([0., 1.]))


def _make_batch(x, gt_at, H, W, device):
    from leod_b200.data.labels import ObjectLabels, SparselyBatchedObjectLabels
    from leod_b200.data.utils.types import DataType
    L, B = x.shape[0], x.shape[1]
    obj, skipped = [], []
    for t in range(L):
        row = []
        for b in range(B):
            if (t, b) in gt_at:
                row.append(ObjectLabels(torch.tensor([[1., 10 + b, 8, 20, 16, b % 2, 1, 1]]), (H, W)))
            else:
                row.append(None)
        obj.append(SparselyBatchedObjectLabels(row))
        skipped.append(SparselyBatchedObjectLabels([None] * B))
    data = {DataType.EV_REPR: x.to(device), DataType.OBJLABELS_SEQ: obj, DataType.SKIPPED_OBJLABELS_SEQ: skipped,
            DataType.IS_FIRST_SAMPLE: torch.ones(B, dtype=torch.bool), DataType.IS_LAST_SAMPLE: torch.zeros(B, dtype=torch.bool),
            DataType.IS_REVERSED: torch.zeros(B, dtype=torch.bool),
            DataType.EV_IDX: [torch.full((B,), t) for t in range(L)],
            DataType.IS_PADDED_MASK: [torch.zeros(B, dtype=torch.bool) for _ in range(L)],
            DataType.PATH: [f'seq{b}' for b in range(B)]}
    return {'worker_id': 0, 'data': data}


def test_predict_step_matches_oracle_composition():
    """PseudoLabeler._predict_step_impl (hflip TTA on, GT frames skipped) against the oracle pieces composed the way
    modules/pseudo_labeler.py:622-770 composes the reference's: per-view backbone unroll, head, postprocess,
    pred2label.  fp32 compute; label rows must agree to 1e-3 and in count/order."""
    from oracle import postprocess as opp, rvt, yolox
    from leod_b200.config import Node
    from leod_b200.modules.pseudo_labeler import PseudoLabeler
    torch.backends.cudnn.allow_tf32 = False     # the neck/head convolutions are checked at fp32 accuracy here
    torch.backends.cuda.matmul.allow_tf32 = False
    z, cfg, sd, d = load_net_fixture()
    H, W, B, T = d['H'], d['W'], d['B'], d['T']
    mcfg = product_cfg(cfg, (H, W), compute_dtype='fp32')
    mcfg.postprocess.confidence_threshold = 0.001
    mcfg.pseudo_label = Node(skip_first_t=1, obj_thresh=[0.01] * cfg.num_classes, cls_thresh=[0.01] * cfg.num_classes)
    full = Node(model=mcfg, dataset=dict(sequence_length=T, name='gen1', downsample_by_factor_2=False),
                tta=dict(enable=True, hflip=True, tflip=False), use_gt=True)
    pl = PseudoLabeler(full)
    pl.mdl.load_state_dict(sd)
    pl.cuda().eval()
    x = torch.from_numpy(z['x'])                                  # [T,B,C,H,W] uint8
    gt_at = {(T - 1, 0)}
    out = pl.predict_step(_make_batch(x, gt_at, H, W, 'cuda'))
    all_labels, paths, ev_idx, first, last, padded, is_hflip, is_tflip = out
    assert len(all_labels) == 2 * B and list(is_hflip) == [False] * B + [True] * B and paths == [f'seq{b}' for b in range(B)] * 2
    # oracle: both views through the CPU restatement
    views = torch.cat((x, torch.flip(x, dims=[-1])), 1).float()
    states, feats_t = None, []
    for t in range(T):
        f, states = rvt.backbone_forward(views[t], states, sd, cfg)
        feats_t.append(f)
    n_checked = 0
    for b in range(2 * B):
        for t in range(T):
            lab = all_labels[b][t]
            # (t == 0 is predicted as well: the reference's GT loop overwrites the skip_first_t rows, pseudo_labeler.py:538)
            if (t, b % B) in gt_at:                               # GT frames are kept, for both views (flipped for the 2nd)
                assert lab is not None and bool((lab.object_labels[:, 0] > 0).all())
                if b >= B:
                    assert float(lab.object_labels[0, 1]) == W - 1 - (10 + b % B) - 20
                continue
            pred, _ = yolox.detect_forward({k: v[b:b + 1] for k, v in feats_t[t].items()}, sd, cfg, targets=None, training=False)
            dets = opp.postprocess(pred.detach().numpy(), cfg.num_classes, 0.001, 0.45)
            ref = opp.pred2label(dets, [0.01] * cfg.num_classes, [0.01] * cfg.num_classes, (240, 304))[0]
            got = lab.object_labels.cpu().numpy()
            assert got.shape == ref.shape, (b, t, got.shape, ref.shape)
            np.testing.assert_allclose(got, ref, rtol=1e-3, atol=2e-3)
            n_checked += got.shape[0]
    assert n_checked > 0, 'the fixture produced no pseudo labels: the comparison would be vacuous'
