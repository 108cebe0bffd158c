"""Pinned, double-buffered host -> device feeding of event batches (the loader side of the path).

The reference hands every batch to the GPU with a blocking `.to(device)` inside the step (modules/detection.py:129-148 after
pytorch-lightning's transfer hook) — 245 MB of uint8 per RVT-S/Gen1 training step, 490 MB per sweep chunk.  Here the upload of
batch i+1 runs on copy streams while batch i computes:

    feeder = PinnedBatchFeeder(shape, device)
    feeder.submit(host_batch_0)
    for i in ...:
        feeder.submit(host_batch_{i+1})          # returns at once; the copy waits (on the device) for its buffer to be free
        ev = feeder.acquire()                    # device tensor of batch i; the CURRENT stream waits for its upload
        ... step on ev ...
        feeder.release()                         # the current stream is done with the buffer of batch i

The bulk copy is split along the first (time) axis over several streams: one DMA engine does not fill the link.  Small per-step
tensors (labels, index lists, masks) must NOT use cudaMemcpy while such a bulk copy is in flight — they would queue behind it on the
host->device engine — which is what leod_upload_small (leod_b200/_lib.py: upload_small) is for.
The HDF5 / blosc-zstd decoding of the reference's datapipes (data/genx_utils/sequence_base.py:184-193) stays on host worker
processes (h5py and hdf5plugin are not in this image); this class starts where they hand over a pinned uint8 batch."""
from collections import deque
from typing import Sequence

import torch


class PinnedBatchFeeder:
    def __init__(self, shape: Sequence[int], device, dtype=torch.uint8, n_buffers: int = 2, n_streams: int = 4):
        self.device = torch.device(device)
        assert self.device.type == 'cuda' and n_buffers >= 2 and n_streams >= 1
        self.buf = [torch.empty(tuple(shape), dtype=dtype, device=self.device) for _ in range(n_buffers)]
        self.streams = [torch.cuda.Stream(self.device) for _ in range(n_streams)]
        n0 = shape[0]
        self.bounds = [n0 * c // n_streams for c in range(n_streams + 1)]
        self.uploaded = [[torch.cuda.Event() for _ in range(n_streams)] for _ in range(n_buffers)]
        self.consumed = [torch.cuda.Event() for _ in range(n_buffers)]
        with torch.cuda.device(self.device):
            for e in self.consumed:
                e.record()
        self.n_submitted = 0
        self.pending = deque()      # slots submitted and not yet acquired
        self.in_use = deque()       # slots acquired and not yet released

    def submit(self, host: torch.Tensor) -> None:
        """Start the upload of the next batch.  `host` must be pinned and must stay valid until the matching acquire()."""
        assert host.shape == self.buf[0].shape and host.dtype == self.buf[0].dtype
        if not host.is_pinned():
            raise RuntimeError('PinnedBatchFeeder.submit needs a pinned host tensor (a pageable copy blocks the host and cannot overlap)')
        assert len(self.pending) + len(self.in_use) < len(self.buf), 'every device buffer is in flight: release() one first'
        slot = self.n_submitted % len(self.buf)
        self.n_submitted += 1
        for c, cs in enumerate(self.streams):
            a, b = self.bounds[c], self.bounds[c + 1]
            if a == b:
                continue
            with torch.cuda.stream(cs):
                cs.wait_event(self.consumed[slot])         # the step that last read this buffer has finished
                self.buf[slot][a:b].copy_(host[a:b], non_blocking=True)
                self.uploaded[slot][c].record(cs)
        self.pending.append(slot)

    def acquire(self) -> torch.Tensor:
        """Device tensor of the oldest submitted batch; the current stream waits for its upload (the host does not)."""
        slot = self.pending.popleft()
        cur = torch.cuda.current_stream(self.device)
        for c in range(len(self.streams)):
            if self.bounds[c] != self.bounds[c + 1]:
                cur.wait_event(self.uploaded[slot][c])
        self.in_use.append(slot)
        return self.buf[slot]

    def release(self) -> None:
        """The current stream has enqueued everything that reads the oldest acquired buffer."""
        slot = self.in_use.popleft()
        self.consumed[slot].record(torch.cuda.current_stream(self.device))
