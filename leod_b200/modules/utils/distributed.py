"""Multi-GPU plumbing of the hot path: one process per GPU, torch.distributed (NCCL over NVLink on the B200 box, gloo in
the CPU tests).  The path is data-parallel over event sequences:

  * training: per-rank batches, ONE all-reduce (mean) per flat gradient buffer per step — the reference gets the same
    result from DDP's 25 MB buckets (train.py:126-133) — and identical initial replicas by broadcast from rank 0;
  * teacher sweep: sequences are dealt to the `world_size x num_workers` global workers in the reference's pyramid
    order over the length-sorted list (data/utils/stream_sharded_datapipe.py:31-57), no data-path collective;
  * teacher weights: every rank holds identical students after the all-reduce, so the EMA teacher is rank-local; a
    broadcast of the flat buffers from rank 0 re-synchronises on demand (drift guard / checkpoint load on rank 0).
"""
from typing import Iterable, List, Sequence, TypeVar

import torch
import torch.distributed as dist

T = TypeVar('T')


def yield_pyramid_indices(start_idx: int, end_idx: int):
    """0,1,..,n-1,n-1,..,1,0,0,1,...  (stream_sharded_datapipe.py:31-38)."""
    while True:
        for idx in range(start_idx, end_idx):
            yield idx
        for idx in range(end_idx - 1, start_idx - 1, -1):
            yield idx


def assign_sequences_to_worker(lengths: Sequence[int], total_num_workers: int, global_worker_id: int) -> List[int]:
    """Indices (into `lengths`) of the sequences global worker `global_worker_id` processes: sort long -> short
    (stable, stream_sharded_datapipe.py:27), then deal in pyramid order (:40-57)."""
    n = len(lengths)
    assert n >= total_num_workers > global_worker_id >= 0, f'{n=}, {total_num_workers=}, {global_worker_id=}'
    order = sorted(range(n), key=lambda i: lengths[i], reverse=True)
    gen = yield_pyramid_indices(0, total_num_workers)
    return [i for i in order if next(gen) == global_worker_id]


def shard_for_rank(lengths: Sequence[int], num_workers_per_rank: int = 1, rank: int = None, world_size: int = None) -> List[List[int]]:
    """Per local worker, the sequence indices of this rank (global_worker_id = rank * num_workers + local id,
    stream_sharded_datapipe.py:92-99)."""
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    total = world_size * num_workers_per_rank
    return [assign_sequences_to_worker(lengths, total, rank * num_workers_per_rank + w) for w in range(num_workers_per_rank)]


def broadcast_flat(buffers: Iterable[torch.Tensor], src: int = 0) -> None:
    """Make every rank's flat parameter / optimizer / teacher buffers equal to rank `src`'s."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    for b in buffers:
        dist.broadcast(b, src)


def allreduce_mean_(flat_grads: Iterable[torch.Tensor]) -> None:
    """Gradient synchronisation: one all-reduce per flat buffer, averaged over ranks (DDP semantics)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    ws = dist.get_world_size()
    for g in flat_grads:
        if dist.get_backend() == 'nccl':
            dist.all_reduce(g, op=dist.ReduceOp.AVG)
        else:                      # gloo has no AVG
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
            g.div_(ws)
