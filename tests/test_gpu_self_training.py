"""The online self-training step (BASELINE configs[4]): teacher targets, ground-truth merge, student step, fused AdamW + EMA into the
teacher's parameters — checked against the composition of the already-tested pieces and against the EMA formula of
modules/utils/ssod.py:429-438."""
import pytest
import torch

from helpers import det_events, rel_err

pytestmark = pytest.mark.gpu


def _module(L=3, dtype='fp32'):
    from leod_b200.config import Node, make_model_cfg
    from leod_b200.modules.self_training import SelfTrainingModule
    torch.manual_seed(0)
    mc = make_model_cfg(embed_dim=16, dim_head=8, partition_size=(2, 3), num_classes=2, fpn_depth=0.33, input_channels=4, in_res_hw=(64, 96),
                        compute_dtype=dtype, conf_thre=0.01, pseudo_label=dict(skip_first_t=0, obj_thresh=[0.05, 0.05], cls_thresh=[0.05, 0.05]))
    full = Node(model=mc, dataset=dict(sequence_length=L, name='gen1', downsample_by_factor_2=False))
    m = SelfTrainingModule(full, max_labels_per_frame=8).cuda()
    with torch.no_grad():       # scores high enough for the teacher to emit boxes
        for p, _, _, _, name in m.student.mdl.detect_engine._param_views:
            if ('obj_preds' in name or 'cls_preds' in name) and name.endswith('bias'):
                p.fill_(0.5)
    m.student.mdl.detect_engine.mark_params_updated()
    m.sync_teacher_from_student()
    return m


def _batch(L, B, with_gt=True):
    from leod_b200.data.labels import ObjectLabels, SparselyBatchedObjectLabels
    from leod_b200.data.utils.types import DataType
    ev = det_events(3, (L, B, 4, 60, 90)).cuda()
    box = torch.tensor([[1., 10, 8, 30, 24, 1, 1, 1], [1., 40, 20, 20, 16, 0, 1, 1]])
    labels = [SparselyBatchedObjectLabels([ObjectLabels(box.clone(), (60, 90)) if (with_gt and t == L - 1 and b == 0) else None for b in range(B)])
              for t in range(L)]
    return {'worker_id': 0, 'data': {DataType.EV_REPR: ev, DataType.OBJLABELS_SEQ: labels, DataType.IS_FIRST_SAMPLE: torch.ones(B, dtype=torch.bool)}}


def test_teacher_targets_equal_the_pseudo_label_pipeline():
    from leod_b200.models.detection.yolox.utils.boxes import postprocess_packed
    from leod_b200.modules.utils.ssod import pred2label_packed
    L, B = 3, 2
    m = _module(L)
    batch = _batch(L, B)
    from leod_b200.data.utils.types import DataType, dget
    ev = dget(batch['data'], DataType.EV_REPR)
    tg = m.teacher_targets(ev, 0, torch.ones(B, dtype=torch.bool))
    assert tuple(tg.shape) == (L * B, 8, 7)
    # the same thing through the per-timestep reference call pattern
    t = m.teacher.eval()
    with torch.no_grad():
        states, rows = None, []
        for ti in range(L):
            feats, states = t.forward_backbone(ev[ti], states)
            preds, _ = t.forward_detect({k: feats[k] for k in (2, 3, 4)})
            dets, cnt = postprocess_packed(preds, 2, 0.01, 0.45, max_det=8)
            lab, n = pred2label_packed(dets, cnt, [0.05, 0.05], [0.05, 0.05], (240, 304))
            rows.append((lab, n))
    n_total = 0
    for ti in range(L):
        lab, n = rows[ti]
        for b in range(B):
            k = int(n[b])
            got = tg[ti * B + b]
            assert float(got[k:].abs().max() if k < 8 else 0.0) == 0.0
            if k:
                ref = torch.stack((lab[b, :k, 5], lab[b, :k, 1] + lab[b, :k, 3] / 2, lab[b, :k, 2] + lab[b, :k, 4] / 2, lab[b, :k, 3], lab[b, :k, 4],
                                   lab[b, :k, 7], lab[b, :k, 6]), -1)
                assert rel_err(got[:k], ref) < 1e-4
            n_total += k
    assert n_total > 0, 'the teacher produced no pseudo labels: the comparison would be vacuous'


def test_step_merges_ground_truth_and_updates_teacher_by_ema():
    from leod_b200.modules.utils.ssod import ema_alpha_at
    L, B = 3, 2
    m = _module(L)
    opt = m.make_optimizer(lr=1e-3)
    s_bb, t_bb = m.student.mdl.backbone, m.teacher.backbone
    s_de, t_de = m.student.mdl.detect_engine, m.teacher.detect_engine
    prev_teacher = None
    for step in range(3):
        opt.zero_grad()
        out = m.training_step(_batch(L, B))
        tg = out['targets']
        # the frame with ground truth carries it (yolox rows: cls, cx, cy, w, h, 1, 1), every other frame the teacher's labels
        f = (L - 1) * B + 0
        assert rel_err(tg[f, 0], torch.tensor([1., 25, 20, 30, 24, 1, 1], device='cuda')) < 1e-6
        assert rel_err(tg[f, 1], torch.tensor([0., 50, 28, 20, 16, 1, 1], device='cuda')) < 1e-6 and float(tg[f, 2:].abs().max()) == 0.0
        assert bool(torch.isfinite(out['loss']))
        teacher_before = (t_bb.flat_params.clone(), t_de.flat_params.clone())
        out['loss'].backward()
        opt.step()
        a = ema_alpha_at(step, 0.999)
        for tb, t_now, s_now in ((teacher_before[0], t_bb.flat_params, s_bb.flat_params), (teacher_before[1], t_de.flat_params, s_de.flat_params)):
            assert rel_err(t_now, a * tb + (1 - a) * s_now) < 1e-5, step
        assert rel_err(t_de.flat_buffers, s_de.flat_buffers) == 0.0            # BatchNorm statistics follow the student
        prev_teacher = t_bb.flat_params.clone()
    # the teacher is a different model from the student after a few steps, and its prepared weights are refreshed
    assert rel_err(t_bb.flat_params, s_bb.flat_params) > 0
