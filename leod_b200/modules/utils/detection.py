"""State / feature book-keeping of the step loops.  This file is a TRANSCRIPTION of the host-side contract in the reference's
modules/utils/detection.py:11-14 (Mode), :27-58 (BackboneFeatureSelector), :95-157 (RNNStates), :160-192 (SeqLens), :195-223
(mixed_collate_fn): the callers keep their names and semantics (states keyed by dataloader worker id, partial reset in place,
detach between batches), so the structure follows the reference statement for statement.  What is original here: the
sync-free masked reset (no device->host round trip of boolean indexing) and the single small upload of a host mask
(leod_upload_small) for all state tensors."""
from enum import Enum, auto
from typing import Dict, List, Optional, Union

import torch as th

WORKER_ID_KEY = 'worker_id'
DATA_KEY = 'data'


class Mode(Enum):
    TRAIN = auto()
    VAL = auto()
    TEST = auto()


mode_2_string = {Mode.TRAIN: 'train', Mode.VAL: 'val', Mode.TEST: 'test'}


class BackboneFeatureSelector:
    """Collect the rows of the per-timestep features that have labels; concatenate over time."""

    def __init__(self):
        self.features: Optional[Dict[int, List[th.Tensor]]] = None

    def reset(self):
        self.features = None

    def add_backbone_features(self, backbone_features: Dict[int, th.Tensor], selected_indices: Optional[List[int]] = None):
        if selected_indices is not None:
            assert len(selected_indices) > 0
        if self.features is None:
            self.features = {k: [] for k in backbone_features}
        for k, v in backbone_features.items():
            self.features[k].append(v[selected_indices] if selected_indices is not None else v)

    def get_batched_backbone_features(self) -> Optional[Dict[int, th.Tensor]]:
        if self.features is None:
            return None
        return {k: th.cat(v, dim=0) for k, v in self.features.items()}


class RNNStates:
    def __init__(self):
        self.states = {}

    @classmethod
    def recursive_detach(cls, inp):
        if isinstance(inp, th.Tensor):
            return inp.detach()
        if isinstance(inp, (list, tuple)):
            return type(inp)(cls.recursive_detach(x) for x in inp)
        if isinstance(inp, dict):
            return {k: cls.recursive_detach(v) for k, v in inp.items()}
        raise NotImplementedError

    @classmethod
    def recursive_reset(cls, inp, indices_or_bool_tensor=None):
        if isinstance(inp, th.Tensor):
            assert inp.requires_grad is False
            if indices_or_bool_tensor is None:
                inp[:] = 0
            else:
                assert len(indices_or_bool_tensor) > 0
                idx = indices_or_bool_tensor
                if th.is_tensor(idx) and idx.dtype == th.bool:
                    # same rows zeroed as `inp[mask] = 0`, without the device->host round trip of boolean indexing
                    # (which would stall the host on the previous step's GPU work at the start of every step)
                    inp.masked_fill_(idx.to(inp.device).view(-1, *([1] * (inp.dim() - 1))), 0)
                else:
                    inp[idx] = 0
            return inp
        if isinstance(inp, (list, tuple)):
            return type(inp)(cls.recursive_reset(x, indices_or_bool_tensor) for x in inp)
        if isinstance(inp, dict):
            return {k: cls.recursive_reset(v, indices_or_bool_tensor) for k, v in inp.items()}
        raise NotImplementedError

    def save_states_and_detach(self, worker_id: int, states) -> None:
        self.states[worker_id] = self.recursive_detach(states)

    def get_states(self, worker_id: int):
        return self.states.get(worker_id, None)

    def reset(self, worker_id: int, indices_or_bool_tensor: Optional[Union[List[int], th.Tensor]] = None):
        if worker_id in self.states:
            idx = indices_or_bool_tensor
            if th.is_tensor(idx) and idx.dtype == th.bool and not idx.is_cuda:
                dev = self._first_device(self.states[worker_id])
                if dev is not None and dev.type == 'cuda':      # one upload for all state tensors, off the copy engines
                    from leod_b200 import _lib
                    idx = _lib.upload_small(idx, dev)
            self.states[worker_id] = self.recursive_reset(self.states[worker_id], idx)

    @classmethod
    def _first_device(cls, inp):
        if isinstance(inp, th.Tensor):
            return inp.device
        if isinstance(inp, dict):
            inp = list(inp.values())
        for x in inp or ():
            d = cls._first_device(x)
            if d is not None:
                return d
        return None


class SeqLens:
    """How many timesteps of each streamed sequence have been seen (modules/utils/detection.py:160-192)."""

    def __init__(self):
        self.lens = {}

    def update_lens(self, worker_id: int, lens: th.Tensor) -> None:
        if worker_id not in self.lens:
            self.lens[worker_id] = lens
        else:
            self.lens[worker_id] += lens

    def get_lens(self, worker_id: int) -> Optional[th.Tensor]:
        return self.lens.get(worker_id, None)

    def reset(self, worker_id: int, indices_or_bool_tensor: Optional[Union[List[int], th.Tensor]] = None):
        if worker_id not in self.lens:
            self.lens[worker_id] = th.zeros(len(indices_or_bool_tensor)).long()
            return
        if indices_or_bool_tensor is None:
            self.lens[worker_id] = th.zeros_like(self.lens[worker_id])
        else:
            assert len(indices_or_bool_tensor) > 0
            idx = indices_or_bool_tensor.cpu() if th.is_tensor(indices_or_bool_tensor) else indices_or_bool_tensor
            self.lens[worker_id][idx] = 0


def _by_name(d: dict, name: str):
    for k, v in d.items():
        if k == name or getattr(k, 'name', None) == name or getattr(k, 'value', None) == name.lower():
            return v
    raise KeyError(name)


def mixed_collate_fn(x1, x2):
    """Concatenate the stream half and the random-access half of a mixed batch along the batch axis
    (modules/utils/detection.py:195-223): tensors, per-timestep lists, label containers, string lists, nested dicts."""
    if isinstance(x1, th.Tensor):
        assert isinstance(x2, th.Tensor)
        return th.cat((x1, x2))
    if hasattr(x1, 'sparse_object_labels_batch'):          # SparselyBatchedObjectLabels (ours or the reference's)
        return x1 + x2
    if isinstance(x1, list):
        assert isinstance(x2, list) and len(x1) == len(x2)
        if len(x1) and isinstance(x1[0], str):
            return x1 + x2
        return [mixed_collate_fn(a, b) for a, b in zip(x1, x2)]
    if isinstance(x1, dict):
        assert isinstance(x2, dict)
        out = {}
        for k in x1:
            if isinstance(x1[k], dict):
                out[k] = mixed_collate_fn(x1[k], x2[k])
            elif isinstance(x1[k], list):
                out[k] = x1[k] + x2[k]
            else:
                raise NotImplementedError(f'{type(x1[k])=}, {type(x2[k])=}')
        return out
    raise NotImplementedError(f'{type(x1)=}, {type(x2)=}')


def merge_mixed_batches(batch: dict):
    """modules/utils/detection.py:226-240: {RANDOM: batch, STREAM: batch} -> one batch, stream rows first; the worker id
    (key of the recurrent-state store) is the streaming loader's."""
    if DATA_KEY in batch:
        return batch
    rnd_data = _by_name(batch, 'RANDOM')[DATA_KEY]
    stream_batch = _by_name(batch, 'STREAM')
    stream_data = stream_batch[DATA_KEY]
    assert rnd_data.keys() == stream_data.keys(), f'{rnd_data.keys()=}, {stream_data.keys()=}'
    return {WORKER_ID_KEY: stream_batch[WORKER_ID_KEY],
            DATA_KEY: {k: mixed_collate_fn(stream_data[k], rnd_data[k]) for k in rnd_data.keys()}}
