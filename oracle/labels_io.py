"""Oracle: the on-disk pseudo-label records.  TEST INFRASTRUCTURE — see oracle/__init__.py.

Restates (paths relative to /root/reference):
  data/genx_utils/labels.py:12-16     BBOX_DTYPE — 40-byte little-endian records
                                       (t i8 @0, x f4 @8, y f4 @12, w f4 @16, h f4 @20, class_id u4 @24,
                                        class_confidence f4 @28, objectness f4 @32)
  data/genx_utils/labels.py:312-325   ObjectLabels.to_structured_array
  modules/pseudo_labeler.py:179-199   EventSeqData._summarize (labels.npz: `labels`, `objframe_idx_2_label_idx`;
                                       objframe_idx_2_repr_idx.npy)
Pinned against tests/golden/tracking_cases.npz (packed_bytes / objframe_idx_*), produced by running the reference.
"""
from typing import List, Sequence, Tuple

import numpy as np

BBOX_DTYPE = np.dtype({'names': ['t', 'x', 'y', 'w', 'h', 'class_id', 'class_confidence', 'objectness'],
                       'formats': ['<i8', '<f4', '<f4', '<f4', '<f4', '<u4', '<f4', '<f4'],
                       'offsets': [0, 8, 12, 16, 20, 24, 28, 32], 'itemsize': 40})


def pack_rows(rows: np.ndarray) -> np.ndarray:
    """[n,8] float32 ObjectLabels rows (t,x,y,w,h,cls,cls_conf,obj) -> n BBOX_DTYPE records."""
    rows = np.asarray(rows, np.float32)
    rec = np.zeros((rows.shape[0],), dtype=BBOX_DTYPE)
    for k, name in enumerate(BBOX_DTYPE.names):
        rec[name] = np.asarray(rows[:, k], dtype=BBOX_DTYPE[name])
    return rec


def summarize(frame_idx: Sequence[int], frames_rows: List[np.ndarray]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """-> (records of all frames concatenated, first record index of every labelled frame, its frame index).
    Like the reference, the per-frame record arrays go through np.concatenate, which (numpy >= 1.?/2.x) returns the
    PACKED form of the dtype: same fields and values, 36-byte records without the 4 padding bytes of BBOX_DTYPE.  np.savez
    stores the dtype description with the data, so readers see identical fields either way."""
    recs, starts, n = [], [], 0
    for rows in frames_rows:
        starts.append(n)
        n += len(rows)
        recs.append(pack_rows(rows))
    labels = np.concatenate(recs) if recs else np.zeros((0,), dtype=BBOX_DTYPE)
    return labels, np.asarray(starts, np.int64), np.asarray(list(frame_idx), np.int64)
