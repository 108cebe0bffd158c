"""`Module._val_test_step_impl` (modules/detection.py:300-401) and `Module.predict_one_seq` (:520-581) on the device against the
composition of the oracle's pieces (backbone unroll, eval-mode neck/head, postprocess) on the reference-generated fixture network."""
import numpy as np
import pytest
import torch

from helpers import canon_rows, load_net_fixture
from test_host_cpu import product_cfg

pytestmark = pytest.mark.gpu


def _module(cfg, sd, hw, L):
    from leod_b200.config import Node
    from leod_b200.modules.detection import Module
    mc = product_cfg(cfg, hw, compute_dtype='fp32')
    mc.postprocess.confidence_threshold = 0.001
    mod = Module(Node(model=mc, dataset=dict(sequence_length=L, name='gen1')))
    mod.mdl.load_state_dict(sd)
    return mod.cuda().eval()


def _oracle_dets(x_seq, sd, cfg, frames):
    """frames: list of (t, b).  -> {(t, b): [N, 7] rows (x1, y1, x2, y2, obj, cls_conf, cls)}"""
    from oracle import postprocess as opp, rvt, yolox
    out, states = {}, None
    with torch.no_grad():
        for t in range(x_seq.shape[0]):
            f, states = rvt.backbone_forward(x_seq[t].float(), states, sd, cfg)
            for (tt, b) in frames:
                if tt == t:
                    pred, _ = yolox.detect_forward({k: v[b:b + 1] for k, v in f.items()}, sd, cfg, training=False)
                    out[(t, b)] = opp.postprocess(pred.numpy(), cfg.num_classes, 0.001, 0.45)[0]
    return out


def _same_detections(got, ref):
    got = canon_rows(np.zeros((0, 7), np.float32) if got is None else got.cpu().numpy())
    ref = canon_rows(np.zeros((0, 7), np.float32) if ref is None else ref)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    np.testing.assert_allclose(got, ref, rtol=1e-3, atol=2e-3)


def test_val_test_step_matches_oracle_composition():
    from leod_b200.data.labels import ObjectLabels, SparselyBatchedObjectLabels
    from leod_b200.data.utils.types import DataType
    from leod_b200.modules.utils.detection import Mode
    z, cfg, sd, d = load_net_fixture()
    T, B, H, W = d['T'], d['B'], d['H'], d['W']
    x = torch.from_numpy(z['x'])
    mod = _module(cfg, sd, (H, W), T)
    box = torch.tensor([[1., 10, 8, 20, 16, 0, 1, 1]])
    gt_at = [(1, 0), (T - 1, 2), (T - 1, 3)]
    labels = [SparselyBatchedObjectLabels([ObjectLabels(box.clone(), (H, W)) if (t, b) in gt_at else None for b in range(B)]) for t in range(T)]
    batch = {'worker_id': 0, 'data': {DataType.EV_REPR: [x[t].cuda() for t in range(T)], DataType.OBJLABELS_SEQ: labels,
                                      DataType.IS_FIRST_SAMPLE: torch.ones(B, dtype=torch.bool)}}
    out = mod._val_test_step_impl(batch, Mode.VAL)
    assert out['skip'] is False and len(out['labels']) == len(out['predictions']) == len(gt_at)
    ref = _oracle_dets(x, sd, cfg, gt_at)
    for (t, b), got in zip(sorted(gt_at), out['predictions']):       # frames are collected timestep by timestep (detection.py:345-362)
        _same_detections(got, ref[(t, b)])
    # a batch without any label is skipped (detection.py:365-368); the recurrent state is still advanced and saved
    none = [SparselyBatchedObjectLabels([None] * B) for _ in range(T)]
    out2 = mod._val_test_step_impl({'worker_id': 0, 'data': {DataType.EV_REPR: [x[t].cuda() for t in range(T)], DataType.OBJLABELS_SEQ: none,
                                                           DataType.IS_FIRST_SAMPLE: torch.zeros(B, dtype=torch.bool)}}, Mode.VAL)
    assert out2 == {'skip': True}
    assert mod.mode_2_rnn_states[Mode.VAL].get_states(worker_id=0) is not None


def test_predict_one_seq_matches_oracle_composition():
    from leod_b200.data.labels import SparselyBatchedObjectLabels
    from leod_b200.data.utils.types import DataType
    z, cfg, sd, d = load_net_fixture()
    T, H, W = d['T'], d['H'], d['W']
    x = torch.from_numpy(z['x'])[:, 1:2]                   # one sequence, [T, 1, C, H, W]
    x = torch.cat([x, x.flip(0)], 0)                       # 2T timesteps, so that chunk=4 splits the loop
    mod = _module(cfg, sd, (H, W), x.shape[0])
    labels = [SparselyBatchedObjectLabels([None]) for _ in range(x.shape[0])]
    batch = {'worker_id': 0, 'data': {DataType.EV_REPR: [x[t] for t in range(x.shape[0])], DataType.OBJLABELS_SEQ: labels,
                                      DataType.IS_FIRST_SAMPLE: torch.ones(1, dtype=torch.bool)}}
    preds, ev_seq, lbl = mod.predict_one_seq(batch, chunk=4)
    assert len(preds) == x.shape[0] == len(lbl) and tuple(ev_seq.shape) == (x.shape[0],) + tuple(x.shape[2:])
    ref = _oracle_dets(x, sd, cfg, [(t, 0) for t in range(x.shape[0])])
    n = 0
    for t in range(x.shape[0]):
        _same_detections(preds[t], ref[(t, 0)])
        n += 0 if preds[t] is None else preds[t].shape[0]
    assert n > 0
