"""Device augmentation (leod_augment_ev_repr, leod_augment_labels) and leod_upload_small — BIT-EXACT against the fixtures produced by
running the reference's RandomSpatialAugmentorGenX / time_flip_data (tests/golden/make_golden.py::gen_augment) and against the
oracle at the full Gen1 / Gen4 frame sizes; cases of one geometry go through the kernels as ONE batch (one sequence per case)."""
import os
import zlib

import numpy as np
import pytest
import torch

from helpers import AUGM_CFG, AUGMENT_CASES, augment_inputs
from oracle import augment as oa
from test_augment_cpu import GOLD, golden_labels, golden_state

pytestmark = pytest.mark.gpu


def _groups():
    g = {}
    for ci, c in enumerate(AUGMENT_CASES):
        g.setdefault(c[:4], []).append(ci)
    return list(g.items())


def _state(d):
    from leod_b200.data.utils.augmentor import AugmentationState, ZoomState
    z = lambda s: ZoomState(active=s['active'], x0=s['x0'], y0=s['y0'], factor=s['factor'])  # noqa: E731
    return AugmentationState(apply_h_flip=d['h_flip']['active'], zoom_in=z(d['zoom_in']), zoom_out=z(d['zoom_out']))


@pytest.mark.parametrize('geom,cases', _groups(), ids=lambda v: 'x'.join(map(str, v)) if isinstance(v, tuple) else None)
def test_batch_augmentor_bit_exact_vs_reference_fixture(geom, cases):
    from leod_b200.data.labels import ObjectLabels, SparselyBatchedObjectLabels
    from leod_b200.data.utils.augmentor import BatchSpatialAugmentor
    H, W, L, C = geom
    B = len(cases)
    evs, labs, states, rev = [], [], [], []
    for ci in cases:
        _, _, _, _, seed, tflip, _ = AUGMENT_CASES[ci]
        ev, rows = augment_inputs(H, W, L, C, seed)
        evs.append(ev)
        labs.append(rows)
        states.append(_state(golden_state(ci)))
        rev.append(tflip)
    ev = torch.from_numpy(np.stack(evs, 1)).cuda()            # [L, B, C, H, W]
    labels = [SparselyBatchedObjectLabels([None if labs[b][t] is None else ObjectLabels(torch.from_numpy(labs[b][t]), (H, W))
                                           for b in range(B)]) for t in range(L)]
    aug = BatchSpatialAugmentor((H, W), AUGM_CFG, B)
    out, new_labels, _ = aug(ev, labels, states=states, is_reversed=rev)
    out = out.cpu().numpy()
    for b, ci in enumerate(cases):
        got = np.ascontiguousarray(out[:, b])
        assert zlib.crc32(got.tobytes()) == int(GOLD[f'{ci}/crc']), f'case {ci}'
        if f'{ci}/ev' in GOLD.files:
            np.testing.assert_array_equal(got, GOLD[f'{ci}/ev'])
        np.testing.assert_array_equal(got, oa.augment_ev_repr(evs[b], golden_state(ci), is_reversed=rev[b]))
        want = golden_labels(ci, L)
        for t in range(L):
            if want[t] is None:
                assert new_labels[t][b] is None
            else:
                np.testing.assert_array_equal(new_labels[t][b].object_labels.numpy(), want[t], err_msg=f'case {ci} frame {t}')


def test_sampled_batch_matches_oracle_at_bench_shape():
    """B = 8 sequences of the Gen1 frame, states drawn by the host sampler (seeded): device == oracle for every sequence; the
    inverse property flip(flip(x)) == x on the device."""
    from leod_b200.data.labels import ObjectLabels, SparselyBatchedObjectLabels
    from leod_b200.data.utils.augmentor import AugmentationState, BatchSpatialAugmentor
    H, W, L, C, B = 240, 304, 3, 20, 8
    ins = [augment_inputs(H, W, L, C, 100 + b) for b in range(B)]
    ev = torch.from_numpy(np.stack([i[0] for i in ins], 1)).cuda()
    labels = [SparselyBatchedObjectLabels([None if ins[b][1][t] is None else ObjectLabels(torch.from_numpy(ins[b][1][t]), (H, W))
                                           for b in range(B)]) for t in range(L)]
    aug = BatchSpatialAugmentor((H, W), AUGM_CFG, B)
    torch.manual_seed(5)
    out, new_labels, states = aug(ev, labels)
    assert len({(s.apply_h_flip, s.zoom_in.active, s.zoom_out.active) for s in states}) >= 3      # the draw covers several modes
    out = out.cpu().numpy()
    for b in range(B):
        sd = states[b].to_dict()
        np.testing.assert_array_equal(out[:, b], oa.augment_ev_repr(ins[b][0], sd))
        for t in range(L):
            if ins[b][1][t] is not None:
                np.testing.assert_array_equal(new_labels[t][b].object_labels.numpy(), oa.augment_labels(ins[b][1][t], (H, W), sd))
    flip = [AugmentationState(apply_h_flip=True) for _ in range(B)]
    once, _, _ = aug(ev, None, states=flip)
    twice, _, _ = aug(once, None, states=flip)
    assert torch.equal(twice, ev) and not torch.equal(once, ev)


def test_upload_small_roundtrip():
    from leod_b200 import _lib
    g = torch.Generator().manual_seed(0)
    for t in (torch.randn(7, 5, generator=g), torch.arange(13, dtype=torch.int64), torch.tensor([True, False, True]),
              torch.randint(0, 255, (70001,), generator=g).to(torch.uint8), torch.zeros(0, 8), torch.randn(300, 300, generator=g)):
        d = _lib.upload_small(t, 'cuda')
        assert d.dtype == t.dtype and d.shape == t.shape and torch.equal(d.cpu(), t)
    many = [torch.randn(64, generator=g) for _ in range(100)]            # ring reuse
    dev = [_lib.upload_small(t, 'cuda') for t in many]
    assert all(torch.equal(d.cpu(), t) for d, t in zip(dev, many))
    with pytest.raises(RuntimeError):
        _lib.check(_lib.lib().leod_upload_small(_lib.ptr(dev[0]), many[0].data_ptr(), 256, None), 'pageable source')


def test_pinned_batch_feeder_double_buffering():
    from leod_b200.data.feeder import PinnedBatchFeeder
    g = torch.Generator().manual_seed(1)
    shape = (5, 2, 4, 48, 64)
    host = [torch.randint(0, 255, shape, generator=g).to(torch.uint8).pin_memory() for _ in range(5)]
    f = PinnedBatchFeeder(shape, 'cuda', n_buffers=2, n_streams=3)
    f.submit(host[0])
    sums = []
    for i in range(5):
        ev = f.acquire()
        if i + 1 < 5:
            f.submit(host[i + 1])                # overlaps the "step" below
        sums.append(ev.long().sum())             # the step: reads the buffer on the current stream
        assert torch.equal(ev.cpu(), host[i])
        f.release()
    assert [int(s) for s in sums] == [int(h.long().sum()) for h in host]
    with pytest.raises(RuntimeError):
        f.submit(torch.zeros(shape, dtype=torch.uint8))      # pageable memory is refused


def test_rnn_states_reset_with_host_mask():
    """modules/utils/detection.py:95-157: `reset(worker_id, indices_or_bool_tensor)` with the collate's HOST bool mask zeroes the
    selected rows of every state tensor on the device (one leod_upload_small for all of them, no pageable cudaMemcpy)."""
    from leod_b200.modules.utils.detection import RNNStates
    st = RNNStates()
    g = torch.Generator().manual_seed(3)
    states = [(torch.randn(4, 6, 3, 5, generator=g).cuda(), torch.randn(4, 6, 3, 5, generator=g).cuda()) for _ in range(4)]
    ref = [(h.clone(), c.clone()) for h, c in states]
    st.save_states_and_detach(worker_id=0, states=states)
    mask = torch.tensor([True, False, False, True])
    st.reset(worker_id=0, indices_or_bool_tensor=mask)
    got = st.get_states(worker_id=0)
    for (h, c), (rh, rc) in zip(got, ref):
        rh[mask] = 0
        rc[mask] = 0
        assert torch.equal(h.cpu(), rh.cpu()) and torch.equal(c.cpu(), rc.cpu())
