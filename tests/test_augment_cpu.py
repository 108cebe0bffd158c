"""Augmentation: the oracle restatement against fixtures produced by RUNNING the reference's augmentor
(tests/golden/make_golden.py::gen_augment), and the host-side state sampler against the reference's RNG consumption."""
import os
import zlib

import numpy as np
import torch

from helpers import AUGM_CFG, AUGMENT_CASES, augment_inputs
from oracle import augment as oa

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'augment_cases.npz'))


def golden_state(ci):
    s, f = GOLD[f'{ci}/state'], GOLD[f'{ci}/factors']
    return {'h_flip': {'active': bool(s[0])},
            'zoom_in': {'active': bool(s[1]), 'x0': int(s[2]), 'y0': int(s[3]), 'factor': float(f[0])},
            'zoom_out': {'active': bool(s[4]), 'x0': int(s[5]), 'y0': int(s[6]), 'factor': float(f[1])}}


def golden_labels(ci, L):
    return [GOLD[f'{ci}/label{t}'] if f'{ci}/label{t}' in GOLD.files else None for t in range(L)]


def test_oracle_matches_reference_augmentor():
    assert int(GOLD['n']) == len(AUGMENT_CASES)
    for ci, (H, W, L, C, seed, tflip, store) in enumerate(AUGMENT_CASES):
        ev, labels = augment_inputs(H, W, L, C, seed)
        st = golden_state(ci)
        out = oa.augment_ev_repr(ev, st, is_reversed=tflip)
        assert zlib.crc32(out.tobytes()) == int(GOLD[f'{ci}/crc']), f'case {ci}'
        assert int((out != 0).sum()) == int(GOLD[f'{ci}/nnz'])
        if store:
            np.testing.assert_array_equal(out, GOLD[f'{ci}/ev'])
        if tflip:
            labels = labels[::-1]
        want = golden_labels(ci, L)
        for t in range(L):
            if labels[t] is None:
                assert want[t] is None
                continue
            got = oa.augment_labels(labels[t], (H, W), st)
            np.testing.assert_array_equal(got, want[t], err_msg=f'case {ci} frame {t}')     # bit-exact fp32


def test_nearest_exact_matches_torch_interpolate():
    for insz, outsz in ((240, 160), (160, 240), (304, 203), (203, 304), (360, 341), (640, 549), (45, 33), (33, 45)):
        x = torch.arange(insz, dtype=torch.float32).view(1, 1, 1, insz)
        want = torch.nn.functional.interpolate(x, size=(1, outsz), mode='nearest-exact').view(-1).numpy().astype(np.int64)
        np.testing.assert_array_equal(oa.nearest_exact_index(outsz, insz), want)


def test_host_sampler_draws_the_reference_states():
    """Same torch seed -> same flips / factors / windows as data/utils/augmentor.py (fixture states)."""
    from leod_b200.data.labels import ObjectLabels
    from leod_b200.data.utils.augmentor import RandomSpatialAugmentorGenX
    for ci, (H, W, L, C, seed, tflip, store) in enumerate(AUGMENT_CASES):
        _, labels = augment_inputs(H, W, L, C, seed)
        if tflip:
            labels = labels[::-1]
        smp = RandomSpatialAugmentorGenX((H, W), False, AUGM_CFG)
        torch.manual_seed(seed)
        smp.randomize_augmentation()
        if smp.augm_state.zoom_in.active:
            smp.choose_zoom_in_window([None if r is None else ObjectLabels(torch.from_numpy(r), (H, W)) for r in labels])
        got, want = smp.augm_state.to_dict(), golden_state(ci)
        for k in ('h_flip', 'zoom_in', 'zoom_out'):
            assert got[k] == want[k], f'case {ci} {k}: {got[k]} vs {want[k]}'
