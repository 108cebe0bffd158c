"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sequence sharding for the teacher sweep, gradient
all-reduce over flat buffers, replica broadcast (leod_b200/modules/utils/distributed.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, lengths, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from leod_b200.modules.utils.distributed import allreduce_mean_, broadcast_flat, shard_for_rank
        shards = shard_for_rank(lengths, num_workers_per_rank=2)
        # gradient sync: rank r holds r+1 everywhere -> mean 1.5
        g = [torch.full((1000,), float(rank + 1)), torch.full((7,), float(10 * (rank + 1)))]
        allreduce_mean_(g)
        # replica broadcast
        p = [torch.arange(5, dtype=torch.float32) * (rank + 1)]
        broadcast_flat(p, src=0)
        out.put((rank, shards, [float(t[0]) for t in g], p[0].tolist()))
    finally:
        dist.destroy_process_group()


def test_pyramid_assignment_matches_reference_rule():
    from leod_b200.modules.utils.distributed import assign_sequences_to_worker
    lengths = [5, 40, 12, 33, 7, 21, 18, 9, 30, 2]
    # reference rule (stream_sharded_datapipe.py:27, 40-57): sort long->short, deal 0,1,2,2,1,0,0,1,2,...
    order = sorted(range(len(lengths)), key=lambda i: lengths[i], reverse=True)
    pattern = [0, 1, 2, 2, 1, 0, 0, 1, 2, 2]
    for w in range(3):
        exp = [i for i, p in zip(order, pattern) if p == w]
        assert assign_sequences_to_worker(lengths, 3, w) == exp
    allw = sorted(i for w in range(3) for i in assign_sequences_to_worker(lengths, 3, w))
    assert allw == list(range(len(lengths)))          # a partition: every sequence exactly once
    with pytest.raises(AssertionError):
        assign_sequences_to_worker([1, 2], 3, 0)      # fewer sequences than workers (reference asserts too)


def test_world_size_2_gloo_sharding_allreduce_broadcast():
    lengths = [60, 3, 45, 45, 12, 8, 31, 27, 19, 5, 50, 2]
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lengths, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    seen = []
    for rank, shards, g, p in res:
        assert len(shards) == 2
        seen += [i for s in shards for i in s]
        assert g == [1.5, 15.0]
        assert p == [0.0, 1.0, 2.0, 3.0, 4.0]           # rank 0's values everywhere
    assert sorted(seen) == list(range(len(lengths)))    # the two ranks x two workers partition the sequence list
    # balance: total frames per rank within the longest sequence of each other
    tot = [sum(lengths[i] for s in shards for i in s) for _, shards, _, _ in res]
    assert abs(tot[0] - tot[1]) <= max(lengths)
