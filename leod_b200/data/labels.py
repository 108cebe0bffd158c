"""Minimal tensor-backed label containers with the methods the hot path calls on the reference's
`ObjectLabels` / `SparselyBatchedObjectLabels` (data/genx_utils/labels.py:249-603, 606-749).  The
training / predict steps are duck-typed: the reference's own classes can be passed instead.

Row layout (ObjectLabels, labels.py:19-60): (t, x, y, w, h, class_id, class_confidence, objectness)
with (x, y) the top-left corner.  t > 0 marks ground truth, t == 0 a pseudo label."""
from typing import List, Optional, Tuple

import torch


class ObjectLabels:
    def __init__(self, object_labels: torch.Tensor, input_size_hw: Tuple[int, int]):
        assert object_labels.dim() == 2 and object_labels.shape[1] == 8
        self.object_labels = object_labels
        self.input_size_hw = tuple(input_size_hw)

    def __len__(self):
        return self.object_labels.shape[0]

    def to(self, *a, **k):
        return ObjectLabels(self.object_labels.to(*a, **k), self.input_size_hw)

    def is_gt_label(self):
        return self.object_labels[:, 0] > 0

    def clone(self):
        return ObjectLabels(self.object_labels.clone(), self.input_size_hw)

    def flip_lr_(self) -> None:
        """labels.py:506-509: mirror the boxes horizontally, in place (x <- W - 1 - x - w)."""
        if len(self) == 0:
            return
        l = self.object_labels
        l[:, 1] = self.input_size_hw[1] - 1 - l[:, 1] - l[:, 3]

    def get_labels_as_tensors(self, format_: str = 'yolox') -> torch.Tensor:
        """labels.py:543-571.  'yolox': [N,7] = (cls, cx, cy, w, h, obj_conf, cls_conf)."""
        l = self.object_labels
        if format_ == 'yolox':
            return torch.stack((l[:, 5], l[:, 1] + 0.5 * l[:, 3], l[:, 2] + 0.5 * l[:, 4], l[:, 3], l[:, 4], l[:, 7], l[:, 6]), 1)
        if format_ == 'prophesee':   # (x1, y1, x2, y2, obj, cls_conf, cls)
            return torch.stack((l[:, 1], l[:, 2], l[:, 1] + l[:, 3], l[:, 2] + l[:, 4], l[:, 7], l[:, 6], l[:, 5]), 1)
        raise NotImplementedError(format_)

    @staticmethod
    def get_labels_as_batched_tensor(obj_label_list: List['ObjectLabels'], format_: str = 'yolox') -> torch.Tensor:
        """labels.py:573-582: zero-pad to the longest entry -> [B, Nmax, 7]."""
        ts = [o.get_labels_as_tensors(format_) for o in obj_label_list]
        n = max(1, max(t.shape[0] for t in ts))
        out = ts[0].new_zeros(len(ts), n, 7)
        for b, t in enumerate(ts):
            out[b, :t.shape[0]] = t
        return out


class SparselyBatchedObjectLabels:
    """`B`-len list of ObjectLabels or None (labels.py:606-749)."""

    def __init__(self, sparse_object_labels_batch: List[Optional[ObjectLabels]]):
        self.sparse_object_labels_batch = list(sparse_object_labels_batch)

    def __len__(self):
        return len(self.sparse_object_labels_batch)

    def __getitem__(self, i):
        return self.sparse_object_labels_batch[i]

    def __iter__(self):
        return iter(self.sparse_object_labels_batch)

    def __add__(self, other: 'SparselyBatchedObjectLabels'):
        """labels.py:~640: concatenate along the batch (used by the hflip TTA batch doubling)."""
        return SparselyBatchedObjectLabels(self.sparse_object_labels_batch + other.sparse_object_labels_batch)

    def is_empty(self):
        return all(l is None for l in self.sparse_object_labels_batch)

    def set_non_gt_labels_to_none_(self) -> None:
        """data/genx_utils/labels.py:645-648: drop frames that carry pseudo labels only (t == 0 in every row)."""
        for i, l in enumerate(self.sparse_object_labels_batch):
            if l is not None and bool((l.object_labels[:, 0] == 0).all()):
                self.sparse_object_labels_batch[i] = None

    def flip_lr_(self) -> None:
        for l in self.sparse_object_labels_batch:
            if l is not None:
                l.flip_lr_()

    def clone(self):
        return SparselyBatchedObjectLabels([None if l is None else l.clone() for l in self.sparse_object_labels_batch])

    def get_valid_labels_and_batch_indices(self, ignore: bool = False, ignore_label: int = 1024):
        """data/genx_utils/labels.py:716-729: every non-None entry — INCLUDING labels with zero boxes (frames emptied by the
        zoom/crop augmentation train as background-only frames) — optionally skipping frames whose boxes all carry
        `ignore_label` (an empty label counts as all-ignore there too, `.all()` of nothing)."""
        if ignore:
            assert ignore_label is not None, 'ignore_label must be provided'
        labels, idx = [], []
        for i, l in enumerate(self.sparse_object_labels_batch):
            if l is None:
                continue
            if ignore and bool((l.object_labels[:, 5] == ignore_label).all()):
                continue
            labels.append(l)
            idx.append(i)
        return labels, idx
