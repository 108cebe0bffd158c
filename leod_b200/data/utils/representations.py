"""Event representations.  Mirror of data/utils/representations.py:37-123 (StackedHistogram): same
constructor, `construct(x, y, pol, time)`, `get_shape`, dtype helpers; the scatter-add runs in the
CUDA library (leod_voxel_bin), reproducing the uint8 wrap-around of `fastmode` and the cutoff clamp."""
from typing import Optional, Tuple

import numpy as np
import torch as th

from leod_b200 import _lib


class StackedHistogram:
    def __init__(self, bins: int, height: int, width: int, count_cutoff: Optional[int] = None, fastmode: bool = True):
        assert bins >= 1 and height >= 1 and width >= 1
        self.bins, self.height, self.width = bins, height, width
        self.count_cutoff = 255 if count_cutoff is None else min(int(count_cutoff), 255)
        assert self.count_cutoff >= 1
        self.fastmode = fastmode
        self.channels = 2

    @staticmethod
    def get_numpy_dtype() -> np.dtype:
        return np.dtype('uint8')

    @staticmethod
    def get_torch_dtype() -> th.dtype:
        return th.uint8

    @property
    def dtype(self) -> th.dtype:
        return th.uint8

    def get_shape(self) -> Tuple[int, int, int]:
        return 2 * self.bins, self.height, self.width

    def construct(self, x: th.Tensor, y: th.Tensor, pol: th.Tensor, time: th.Tensor) -> th.Tensor:
        """-> uint8 [2*bins, H, W], channel = pol*bins + time_bin.  `time` must be sorted (as in the
        reference, which reads time[0] / time[-1])."""
        for t in (x, y, pol, time):
            assert not th.is_floating_point(t) and not th.is_complex(t)
        if not x.is_cuda:
            raise RuntimeError('leod_b200 StackedHistogram runs on CUDA tensors only (no CPU fallback)')
        n = x.numel()
        assert y.numel() == n and pol.numel() == n and time.numel() == n
        dev = x.device
        out = th.empty(self.get_shape(), dtype=th.uint8, device=dev)
        xi, yi, pi = (t.to(th.int32).contiguous() for t in (x, y, pol))
        ti = time.to(th.int64).contiguous()
        with th.cuda.device(dev):
            _lib.check(_lib.lib().leod_voxel_bin(_lib.ptr(xi), _lib.ptr(yi), _lib.ptr(pi), _lib.ptr(ti), n, self.bins, self.height,
                                                 self.width, self.count_cutoff, int(self.fastmode), _lib.ptr(out),
                                                 _lib.stream_ptr(dev)), 'voxel_bin')
        return out
