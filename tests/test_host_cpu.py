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


@pytest.mark.parametrize('tag,thr', [('plain', None), ('ignore', None), ('thresh', [0.7, 0.35])])
def test_batched_head_loss_matches_reference(net, tag, thr):
    z, cfg, sd, d = net
    from leod_b200.models.detection.yolox_extension.models.detector import YoloXDetector
    m = YoloXDetector(product_cfg(cfg, (d['H'], d['W']), ignore_thresh=thr))
    m.load_state_dict(sd)
    m.train()
    T = d['T']
    feats = {s: torch.from_numpy(z[f'eval/feat{s}_t{T - 1}']) for s in (1, 2, 3, 4)}
    preds, losses = m.forward_detect(feats, targets=torch.from_numpy(z[f'train_{tag}/labels']))
    for k in ('loss', 'iou_loss', 'conf_loss', 'cls_loss', 'num_fg'):
        ref = float(z[f'train_{tag}/{k}'])
        assert abs(float(losses[k]) - ref) < 2e-5 * max(1.0, abs(ref)), (k, float(losses[k]), ref)
    assert rel_err(preds.detach(), z[f'train_{tag}/preds']) < 2e-5
    losses['loss'].backward()
    grads = dict(m.named_parameters())
    for key in z.files:
        if key.startswith(f'train_{tag}/grad/') and not key.split('/grad/')[1].startswith('backbone'):
            name = key.split('/grad/')[1]
            assert rel_err(grads[name].grad, z[key]) < 2e-4, name


def test_eval_head_matches_reference(net):
    z, cfg, sd, d = net
    from leod_b200.models.detection.yolox_extension.models.detector import YoloXDetector
    m = YoloXDetector(product_cfg(cfg, (d['H'], d['W'])))
    m.load_state_dict(sd)
    m.eval()
    T = d['T']
    feats = {s: torch.from_numpy(z[f'eval/feat{s}_t{T - 1}']) for s in (1, 2, 3, 4)}
    with torch.no_grad():
        preds, losses = m.forward_detect(feats)
    assert losses is None
    assert rel_err(preds, z['eval/preds']) < 2e-5


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
