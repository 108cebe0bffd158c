"""Device tracker / short-track filter / in-painting / label records (leod_track_filter, leod_pack_bbox) — BIT-EXACT against the
fixtures produced by running the reference's EventSeqData._track_filter / _summarize (tests/golden/make_golden.py: gen_tracking) and
against the oracle on extra seeded sequences, all sequences of a test in ONE call."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN
from oracle import labels_io, tracking as otr

pytestmark = pytest.mark.gpu


def _cases():
    z = np.load(os.path.join(GOLDEN, 'tracking_cases.npz'))
    out = []
    for ci in range(int(z['n'])):
        counts = z[f'{ci}/counts']
        rows = np.split(z[f'{ci}/rows'], np.cumsum(counts)[:-1])
        out.append(dict(hw=tuple(int(v) for v in z[f'{ci}/hw']), frame_idx=z[f'{ci}/frame_idx'].tolist(), rows=rows, z=z, ci=ci))
    return out


def _split(z, ci, tag):
    counts = z[f'{ci}/final_{tag}_counts']
    return z[f'{ci}/final_{tag}_frame_idx'].tolist(), np.split(z[f'{ci}/final_{tag}_rows'], np.cumsum(counts)[:-1])


@pytest.mark.parametrize('method,tag', [('forward', 'f'), ('forward or backward', 'fb')])
def test_track_filter_bit_exact_vs_reference_fixture(method, tag):
    from leod_b200.modules.tracking import track_filter_sequences
    cases = _cases()
    res = track_filter_sequences([(c['frame_idx'], [torch.from_numpy(r) for r in c['rows']]) for c in cases], [c['hw'] for c in cases],
                                 min_track_len=6, track_method=method, inpaint=True, ignore_label=1024)
    for c, (fi, rows) in zip(cases, res):
        ref_fi, ref_rows = _split(c['z'], c['ci'], tag)
        assert fi == ref_fi, c['ci']
        assert len(rows) == len(ref_rows)
        for k, (a, b) in enumerate(zip(rows, ref_rows)):
            np.testing.assert_array_equal(a.cpu().numpy(), b, err_msg=f'case {c["ci"]} frame {fi[k]}')


def test_label_records_bit_exact_vs_reference_fixture():
    from leod_b200.modules.tracking import BBOX_DTYPE, pack_labels, summarize, track_filter_sequences
    cases = _cases()
    res = track_filter_sequences([(c['frame_idx'], [torch.from_numpy(r) for r in c['rows']]) for c in cases], [c['hw'] for c in cases])
    for c, (fi, rows) in zip(cases, res):
        z, ci = c['z'], c['ci']
        labels, lbl_idx, repr_idx = summarize(fi, rows)
        assert labels.tobytes() == z[f'{ci}/packed_bytes'].tobytes()
        np.testing.assert_array_equal(lbl_idx, z[f'{ci}/objframe_idx_2_label_idx'])
        np.testing.assert_array_equal(repr_idx, z[f'{ci}/objframe_idx_2_repr_idx'])
        # the declared 40-byte BBOX_DTYPE layout carries the same field values
        rec40 = pack_labels(torch.cat(rows), packed=False)
        assert rec40.dtype == BBOX_DTYPE and rec40.dtype.itemsize == 40
        ref = labels_io.pack_rows(torch.cat(rows).cpu().numpy())
        for name in BBOX_DTYPE.names:
            np.testing.assert_array_equal(rec40[name], ref[name])


def _synth(seed, n_frames, hw, n_obj, miss_p, fp_rate, gt_every):
    rng = np.random.default_rng(seed)
    H, W = hw
    objs = [dict(x=rng.uniform(20, W - 60), y=rng.uniform(20, H - 60), w=rng.uniform(15, 60), h=rng.uniform(15, 50), vx=rng.uniform(-4, 4),
                 vy=rng.uniform(-3, 3), cls=int(rng.integers(0, 2)), t0=int(rng.integers(0, n_frames // 2)), t1=int(rng.integers(n_frames // 2, n_frames)))
            for _ in range(n_obj)]
    frames, idx = [], []
    for f in range(n_frames):
        rows = []
        for o in objs:
            if not (o['t0'] <= f < o['t1']) or rng.uniform() < miss_p:
                continue
            x, y = o['x'] + o['vx'] * (f - o['t0']) + rng.normal(0, 0.7), o['y'] + o['vy'] * (f - o['t0']) + rng.normal(0, 0.7)
            rows.append([1.0 if (gt_every and f % gt_every == 0) else 0.0, x, y, o['w'], o['h'], o['cls'], rng.uniform(0.5, 1), rng.uniform(0.5, 1)])
        for _ in range(rng.poisson(fp_rate)):
            rows.append([0.0, rng.uniform(0, W - 30), rng.uniform(0, H - 30), rng.uniform(8, 40), rng.uniform(8, 40), int(rng.integers(0, 2)),
                         rng.uniform(0.3, 0.9), rng.uniform(0.3, 0.9)])
        if rows:
            frames.append(np.asarray(rows, np.float32))
            idx.append(f)
    return idx, frames


def test_many_sequences_one_call_vs_oracle():
    """40 seeded sequences (objects leaving the frame -> clamped predictions, misses, false positives, ground-truth frames, empty
    sequences) post-processed in one launch, every output row compared bit-exactly with the oracle."""
    from leod_b200.modules.tracking import track_filter_sequences
    seqs, hws = [], []
    for s in range(40):
        hw = (240, 304) if s % 3 else (360, 640)
        idx, frames = _synth(100 + s, 20 + 3 * (s % 17), hw, 2 + s % 9, 0.1 + 0.05 * (s % 7), 0.3 * (s % 4), 10 if s % 5 == 0 else 0)
        if s == 7:
            idx, frames = [], []
        seqs.append((idx, frames))
        hws.append(hw)
    res = track_filter_sequences([(i, [torch.from_numpy(r) for r in f]) for i, f in seqs], hws)
    n_ignored = n_inpainted = 0
    for s, ((idx, frames), (fi, rows)) in enumerate(zip(seqs, res)):
        ref_fi, ref_rows = otr.track_filter(frames, idx, hws[s], 6, 'forward or backward', True, 1024)
        assert fi == ref_fi, s
        for k, (a, b) in enumerate(zip(rows, ref_rows)):
            np.testing.assert_array_equal(a.cpu().numpy(), np.asarray(b, np.float32), err_msg=f'sequence {s} frame {fi[k]}')
            n_ignored += int((np.asarray(b)[:, 5] == 1024).sum())
        n_inpainted += sum(len(r) for r in ref_rows) - sum(len(r) for r in frames)
    assert n_ignored > 50 and n_inpainted > 20, (n_ignored, n_inpainted)


def test_event_seq_data_update_aggregate_track_save(tmp_path):
    """EventSeqData (modules/pseudo_labeler.py:94-400): chunks of a normal, an hflip and a tflip view collected per frame, merged by the
    TTA NMS, tracked / filtered / in-painted, written as labels.npz — against the composition of the oracle's pieces."""
    from oracle import postprocess as opp
    from leod_b200.config import Node
    from leod_b200.data.labels import ObjectLabels
    from leod_b200.modules.tracking import EventSeqData, finalize_sequences
    hw, n_frames, off = (240, 304), 30, 1
    idx, frames = _synth(9, n_frames, hw, 5, 0.2, 0.4, 0)
    fc = Node(min_track_len=6, track_method='forward or backward', inpaint=True, ignore_label=1024)
    pc = Node(confidence_threshold=0.1, nms_threshold=0.45)
    seq = EventSeqData('/data/gen1/train/seq_a', 1, fc, pc, hw)
    per_frame = {}
    rng = np.random.default_rng(0)
    for view in ('plain', 'hflip', 'tflip'):
        rows_v = {f: r + (rng.normal(0, 0.5, r.shape).astype(np.float32) * np.array([0, 1, 1, 1, 1, 0, 0, 0], np.float32)) for f, r in zip(idx, frames)}
        for c0 in range(0, n_frames, 10):                   # three chunks of 10 timesteps
            ts = list(range(c0, c0 + 10))
            labels, ev_idx = [], []
            for t in ts:
                r = rows_v.get(t)
                if view == 'tflip':
                    ev_idx.append(t - off)                  # the time-flipped stream reports indices shifted by tflip_offset
                else:
                    ev_idx.append(t)
                if r is None:
                    labels.append(None)
                    continue
                lab = ObjectLabels(torch.from_numpy(r.copy()).cuda(), hw)
                if view == 'hflip':
                    lab.flip_lr_()                          # what the model saw; update() flips it back
                labels.append(lab)
                per_frame.setdefault(t, []).append(r)
            seq.update(labels, ev_idx, is_last_sample=c0 + 10 >= n_frames, is_padded_mask=[False] * 10, is_hflip=view == 'hflip',
                       is_tflip=view == 'tflip', tflip_offset=off)
    finalize_sequences([seq], [n_frames])
    # oracle composition
    fi_ref = sorted(per_frame)
    merged = [opp.tta_merge(np.concatenate(per_frame[f]), 0.1, 0.45) for f in fi_ref]
    fi_ref, rows_ref = otr.track_filter(merged, fi_ref, hw, 6, 'forward or backward', True, 1024)
    assert seq.frame_idx == fi_ref
    for a, b in zip(seq.labels, rows_ref):
        got, ref = a.cpu().numpy(), np.asarray(b, np.float32)
        assert got.shape == ref.shape
        np.testing.assert_allclose(got, ref, rtol=0, atol=2e-5)     # the hflip round trip (W-1-x-w twice) costs an ulp
    d = seq.save(str(tmp_path), n_frames)
    z = np.load(os.path.join(d, 'labels_v2', 'labels.npz'))
    rec, lbl_idx, _ = labels_io.summarize(seq.frame_idx, [r.cpu().numpy() for r in seq.labels])
    assert z['labels'].tobytes() == rec.tobytes() and z['labels'].dtype.names == rec.dtype.names
    np.testing.assert_array_equal(z['objframe_idx_2_label_idx'], lbl_idx)
    np.testing.assert_array_equal(np.load(os.path.join(d, 'event_representations_v2', 'objframe_idx_2_repr_idx.npy')), np.asarray(seq.frame_idx))
