"""Oracle: event -> stacked-histogram voxel tensor, numpy integer arithmetic.

Restates data/utils/representations.py:78-123 (StackedHistogram.construct), both modes:
fastmode=True accumulates in uint8 (wraps mod 256, then clamps to count_cutoff), fastmode=False
accumulates in int16 and clamps.  Time normalisation follows the reference literally:
float32 true-division of int64 (t - t0) by max(t1 - t0, 1), times bins, floor, clamp to bins-1.
TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
import numpy as np


def time_bin(t: np.ndarray, bins: int) -> np.ndarray:
    """representations.py:103-109.  torch int64 / int64 -> float32 true division."""
    t = t.astype(np.int64)
    t0, t1 = t[0], t[-1]
    tn = (t - t0).astype(np.float32) / np.float32(max(int(t1 - t0), 1))
    tn = tn * np.float32(bins)
    return np.minimum(np.floor(tn), np.float32(bins - 1)).astype(np.int64)


def stacked_histogram(x, y, pol, t, bins: int, height: int, width: int, count_cutoff=None, fastmode: bool = True):
    """-> uint8 [2*bins, H, W], channel = pol*bins + bin (polarity-major)."""
    cutoff = 255 if count_cutoff is None else min(int(count_cutoff), 255)
    rep = np.zeros(2 * bins * height * width, np.int64)
    if len(x) > 0:
        idx = x.astype(np.int64) + width * y.astype(np.int64) + height * width * time_bin(t, bins) \
            + bins * height * width * pol.astype(np.int64)
        np.add.at(rep, idx, 1)
    if fastmode:
        rep = rep % 256           # uint8 accumulation wraps
    else:
        rep = ((rep + 32768) % 65536) - 32768   # int16 accumulation wraps (never reached in practice)
    rep = np.clip(rep, 0, cutoff)
    return rep.astype(np.uint8).reshape(2 * bins, height, width)
