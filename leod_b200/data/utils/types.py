"""Batch-dict contract.  Mirror of data/utils/types.py:15-32, 46-49 (same member names, so batches
produced by the reference's dataloaders index correctly when the reference's own enum is passed, and
synthetic batches built here look the same)."""
from enum import Enum, auto


class DataType(Enum):
    PATH = auto()
    EV_IDX = auto()
    EV_REPR = auto()
    FLOW = auto()
    IMAGE = auto()
    OBJLABELS = auto()
    OBJLABELS_SEQ = auto()
    SKIPPED_OBJLABELS_SEQ = auto()
    IS_PADDED_MASK = auto()
    IS_FIRST_SAMPLE = auto()
    IS_LAST_SAMPLE = auto()
    IS_REVERSED = auto()
    TOKEN_MASK = auto()
    PRED_MASK = auto()
    GT_MASK = auto()
    PRED_PROBS = auto()
    AUGM_STATE = auto()


class DatasetSamplingMode(Enum):
    """data/utils/types.py:46-49 (the reference derives from StrEnum; values are the same strings)."""
    RANDOM = 'random'
    STREAM = 'stream'
    MIXED = 'mixed'


def dget(data: dict, key: DataType, default=None):
    """Look a DataType up by NAME so dicts keyed by the reference's own enum class work too."""
    if key in data:
        return data[key]
    for k, v in data.items():
        if getattr(k, 'name', None) == key.name:
            return v
    return default
