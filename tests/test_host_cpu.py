"""CPU-side checks of the product's host code: state_dict compatibility with the reference, the
batched (sync-free) head + SimOTA loss against the reference fixtures, and the C-ABI surface."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

from helpers import GOLDEN, load_net_fixture, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def product_cfg(cfg, hw, ignore_thresh=None, compute_dtype='fp32'):
    from leod_b200.config import make_model_cfg
    return make_model_cfg(embed_dim=cfg.embed_dim, dim_head=cfg.dim_head, partition_size=cfg.partition_size,
                          num_classes=cfg.num_classes, fpn_depth=cfg.fpn_depth, input_channels=cfg.input_channels,
                          in_res_hw=hw, ignore_bbox_thresh=ignore_thresh, compute_dtype=compute_dtype)


@pytest.fixture(scope='module')
def net():
    return load_net_fixture()


def test_state_dict_keys_and_shapes_match_reference(net):
    z, cfg, sd, d = net
    from leod_b200.models.detection.yolox_extension.models.detector import YoloXDetector
    m = YoloXDetector(product_cfg(cfg, (d['H'], d['W'])))
    msd = m.state_dict()
    assert set(msd.keys()) == set(sd.keys())
    for k in sd:
        assert tuple(msd[k].shape) == tuple(sd[k].shape), k
    m.load_state_dict(sd)
    # parameters alias one flat buffer, in place
    bb = m.backbone
    w = bb.stages[2].lstm.conv1x1.weight
    assert w.data_ptr() >= bb.flat_params.data_ptr()
    assert rel_err(w, sd['backbone.stages.2.lstm.conv1x1.weight']) == 0.0


def test_full_size_parameter_counts():
    """tiny 4 405 141, small 9 870 165, base 18 536 469 parameters (Gen1, 2 classes)."""
    from leod_b200.config import make_model_cfg
    from leod_b200.models.detection.yolox_extension.models.detector import YoloXDetector
    for size, n in (('tiny', 4405141), ('small', 9870165), ('base', 18536469)):
        m = YoloXDetector(make_model_cfg(size=size, dataset='gen1'))
        assert sum(p.numel() for p in m.parameters()) == n, size


def test_detect_refuses_cpu(net):
    """The neck/head/loss have no CPU or PyTorch fallback either: forward_detect on CPU tensors must raise."""
    z, cfg, sd, d = net
    from leod_b200.models.detection.yolox_extension.models.detector import YoloXDetector
    m = YoloXDetector(product_cfg(cfg, (d['H'], d['W'])))
    m.load_state_dict(sd)
    T = d['T']
    feats = {s: torch.from_numpy(z[f'eval/feat{s}_t{T - 1}']) for s in (1, 2, 3, 4)}
    with pytest.raises(RuntimeError, match='CUDA only'):
        m.eval().forward_detect(feats)
    with pytest.raises(RuntimeError, match='CUDA only'):
        m.train().forward_detect(feats, targets=torch.from_numpy(z['train_plain/labels']))
    # the parameters of the neck / head alias one flat buffer under the reference's names
    e = m.detect_engine
    w = m.fpn.C3_p4.m[0].conv2.conv.weight
    assert e.flat_params.data_ptr() <= w.data_ptr() < e.flat_params.data_ptr() + 4 * e.flat_params.numel()
    assert rel_err(w, sd['fpn.C3_p4.m.0.conv2.conv.weight']) == 0.0
    assert rel_err(m.yolox_head.stems[1].bn.running_var, sd['yolox_head.stems.1.bn.running_var']) == 0.0


def test_backbone_refuses_cpu(net):
    z, cfg, sd, d = net
    from leod_b200.models.detection.yolox_extension.models.detector import YoloXDetector
    m = YoloXDetector(product_cfg(cfg, (d['H'], d['W'])))
    with pytest.raises(RuntimeError, match='CUDA only'):
        m.forward_backbone(torch.zeros(1, cfg.input_channels, d['H'], d['W']))


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'leod_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(leod_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 20
    from leod_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f'{name} declared in include/leod_b200.h but not exported'
    assert set(_lib.EXPORTED_SYMBOLS) == declared
    assert _lib.lib().leod_abi_version() == 1


def test_layout_query_needs_no_gpu():
    from leod_b200 import _lib
    l = _lib.lib()
    cfg = _lib.BackboneCfg(20, 48, 24, 8, 10, 4, 256, 320, _lib.LEOD_BF16, 1e-5)
    h = ctypes.c_void_p()
    assert l.leod_backbone_layout_only(ctypes.byref(cfg), ctypes.byref(h)) == 0
    assert l.leod_backbone_param_info(h, -1, None, 0, None, None, None) == 124
    assert l.leod_backbone_save_bytes(h, 8) > 0
    l.leod_backbone_destroy(h)
    bad = _lib.BackboneCfg(20, 48, 24, 8, 10, 4, 250, 320, _lib.LEOD_BF16, 1e-5)
    assert l.leod_backbone_layout_only(ctypes.byref(bad), ctypes.byref(h)) != 0
    assert b'multiple of 32' in l.leod_last_error()


def test_pseudo_labeler_masks_and_hflip_batch_doubling(net):
    """Host logic of modules/pseudo_labeler.py:458-495 and :514-547 (no device work): TTA batch doubling duplicates
    the metadata and mirrors the labels; frames with GT, padded frames and the first `skip_first_t` frames of a fresh
    sequence are not predicted."""
    from leod_b200.config import Node
    from leod_b200.data.labels import ObjectLabels, SparselyBatchedObjectLabels
    from leod_b200.data.utils.types import DataType
    from leod_b200.modules.pseudo_labeler import PseudoLabeler
    z, cfg, sd, d = net
    H, W = d['H'], d['W']
    mcfg = product_cfg(cfg, (H, W))
    mcfg.pseudo_label = Node(skip_first_t=2, obj_thresh=[0.6, 0.3], cls_thresh=[0.6, 0.3])
    full = Node(model=mcfg, dataset=dict(sequence_length=4, name='gen1', downsample_by_factor_2=False),
                tta=dict(enable=True, hflip=True, tflip=True), use_gt=True)
    pl = PseudoLabeler(full)
    L, B = 4, 2
    box = torch.tensor([[1., 10, 8, 20, 16, 0, 1, 1]])
    obj = [SparselyBatchedObjectLabels([ObjectLabels(box.clone(), (H, W)) if (t, b) == (3, 1) else None for b in range(B)]) for t in range(L)]
    skipped = [SparselyBatchedObjectLabels([None] * B) for _ in range(L)]
    padded = [torch.tensor([False, False]) for _ in range(L)]
    padded[2] = torch.tensor([True, False])
    data = {DataType.EV_REPR: torch.zeros(L, B, cfg.input_channels, H, W, dtype=torch.uint8), DataType.OBJLABELS_SEQ: obj,
            DataType.SKIPPED_OBJLABELS_SEQ: skipped, DataType.IS_FIRST_SAMPLE: torch.tensor([True, False]),
            DataType.IS_LAST_SAMPLE: torch.tensor([False, False]), DataType.IS_REVERSED: torch.tensor([False, False]),
            DataType.EV_IDX: [torch.full((B,), t) for t in range(L)], DataType.IS_PADDED_MASK: padded, DataType.PATH: ['a', 'b']}
    out = pl.get_data_from_batch({'worker_id': 0, 'data': data})
    assert out['EV_REPR'].shape[1] == 2 * B and out['EV_REPR'].dtype == torch.uint8
    assert out['PATH'] == ['a', 'b', 'a', 'b'] and out['IS_FIRST_SAMPLE'].tolist() == [True, False, True, False]
    assert list(out['is_hflip']) == [False, False, True, True]
    flipped = out['OBJLABELS_SEQ'][3][3]
    assert float(flipped.object_labels[0, 1]) == W - 1 - 10 - 20 and float(out['OBJLABELS_SEQ'][3][1].object_labels[0, 1]) == 10
    # sequence 0 (and its mirrored view) is fresh, sequence 1 has 5 frames of history
    pl.mode_2_seq_lens.lens[0] = torch.tensor([0, 5, 0, 5])
    pse, gt, skipped_gt = pl._get_pred_mask(worker_id=0, data=out)
    exp = np.ones((L, 2 * B), bool)
    # the reference overwrites the skip_first_t rows in its GT loop (modules/pseudo_labeler.py:538 `skip_mask[tidx, bidx] = has_gt`),
    # so the first frames of a fresh sequence ARE predicted: only padding and ground truth switch a frame off
    exp[2, 0] = exp[2, 2] = False             # padded
    exp[3, 1] = exp[3, 3] = False             # ground truth present
    np.testing.assert_array_equal(pse, exp)
    assert gt.sum() == 2 and gt[3, 1] and gt[3, 3] and skipped_gt.sum() == 0


def test_model_update_modes():
    """modules/utils/ssod.py:429-460: EMA with the true-average warm-up, and the 'every-N' hard copy (parameters and buffers)."""
    from leod_b200.modules.utils.ssod import ema_alpha_at, model_update
    torch.manual_seed(0)
    student = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.BatchNorm1d(3))
    teacher = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.BatchNorm1d(3))
    s0 = [p.detach().clone() for p in student.parameters()]
    t0 = [p.detach().clone() for p in teacher.parameters()]
    assert ema_alpha_at(0) == 0.0 and ema_alpha_at(1) == 0.5 and ema_alpha_at(10 ** 6) == 0.999
    model_update(student, teacher, global_step=1, method='ema')
    for t, a, b in zip(teacher.parameters(), t0, s0):
        assert rel_err(t, 0.5 * a + 0.5 * b) < 1e-6
    student[1].running_mean.fill_(3.0)
    model_update(student, teacher, global_step=0, method='every-2')       # (0 + 1) % 2 != 0: untouched
    assert float(teacher[1].running_mean.abs().max()) == 0.0
    model_update(student, teacher, global_step=1, method='every-2')
    assert float(teacher[1].running_mean.min()) == 3.0
    for t, b in zip(teacher.parameters(), s0):
        assert rel_err(t, b) == 0.0
    with pytest.raises(NotImplementedError):
        model_update(student, teacher, 0, method='nope')


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the oracle port on the host cores; runs without a GPU) prints ONE JSON line with the
    keys the driver reads."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'event-frames/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['steps'] == 1 and d['n_gpus'] == 1 and d['vs_baseline'] is None
    assert d['cpu_baseline']['kind'] in ('reference', 'port') and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'event-frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config']


@pytest.mark.parametrize('total,pct,lr', [(1000, 0.005, 2e-4), (400, 0.1, 3.46e-4), (50, 0.3, 1e-3)])
def test_one_cycle_lr_matches_torch_scheduler(total, pct, lr):
    """The closed-form schedule used with FlatOptimizer equals torch's OneCycleLR configured as the reference does
    (modules/detection.py:495-510: linear anneal, final_div_factor / div_factor, no momentum cycling)."""
    from leod_b200.modules.detection import one_cycle_lr
    div, final_div = 20.0, 10000.0
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([p], lr=lr)
    sch = torch.optim.lr_scheduler.OneCycleLR(optimizer=opt, max_lr=lr, div_factor=div, final_div_factor=final_div / div, total_steps=total,
                                              pct_start=pct, cycle_momentum=False, anneal_strategy='linear')
    for step in range(total):
        assert abs(opt.param_groups[0]['lr'] - one_cycle_lr(step, lr, total, pct, div, final_div)) < 1e-12 + 1e-9 * lr, step
        opt.step()
        if step + 1 < total:
            sch.step()


def test_merge_mixed_batches_contract():
    """modules/utils/detection.py:226-240: stream rows first, random-access rows after; worker id of the stream loader;
    tensors / per-timestep lists / label containers / paths all concatenated along the batch axis."""
    from leod_b200.data.labels import ObjectLabels, SparselyBatchedObjectLabels
    from leod_b200.data.utils.types import DatasetSamplingMode, DataType
    from leod_b200.modules.utils.detection import merge_mixed_batches
    L = 3

    def half(B, base, wid):
        lab = [SparselyBatchedObjectLabels([ObjectLabels(torch.full((1, 8), float(base + b)), (240, 304)) if t == L - 1 else None
                                            for b in range(B)]) for t in range(L)]
        return {'worker_id': wid, 'data': {DataType.EV_REPR: [torch.full((B, 2, 4, 4), float(base + t)) for t in range(L)],
                                           DataType.OBJLABELS_SEQ: lab, DataType.IS_FIRST_SAMPLE: torch.tensor([base == 100] * B),
                                           DataType.PATH: [f'p{base + b}' for b in range(B)]}}
    merged = merge_mixed_batches({DatasetSamplingMode.RANDOM: half(2, 100, 7), DatasetSamplingMode.STREAM: half(2, 0, 4)})
    assert merged['worker_id'] == 4
    d = merged['data']
    assert [tuple(x.shape) for x in d[DataType.EV_REPR]] == [(4, 2, 4, 4)] * L
    assert d[DataType.EV_REPR][1][:, 0, 0, 0].tolist() == [1.0, 1.0, 101.0, 101.0]
    assert d[DataType.IS_FIRST_SAMPLE].tolist() == [False, False, True, True]
    assert d[DataType.PATH] == ['p0', 'p1', 'p100', 'p101']
    last = d[DataType.OBJLABELS_SEQ][L - 1]
    assert len(last) == 4 and [float(last[b].object_labels[0, 0]) for b in range(4)] == [0.0, 1.0, 100.0, 101.0]
    assert merge_mixed_batches(merged) is merged          # already merged batches pass through
