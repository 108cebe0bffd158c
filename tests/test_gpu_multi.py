"""Two ranks over NCCL against ONE process holding both ranks' data (needs >= 2 GPUs; skipped otherwise):

  * neck + head with SyncBatchNorm semantics (train.py:247): the statistics exchange of `leod_detect_set_allreduce` (one packed
    collective per dependency level, forward AND backward) makes every rank's outputs equal the corresponding rows of the
    single-process 2x-batch run, the BatchNorm running statistics equal, and — with the flat-gradient all-reduce (mean) — every
    parameter gradient equal to the single-process gradient of the averaged loss;
  * backbone: flat gradients after the all-reduce equal the single-process gradients of the averaged loss.
The backward goes through a loss that is linear in the raw head outputs / the features, so that "averaged loss" means the same
thing in both runs (the SimOTA loss normalises by the per-rank foreground count, yolo_head.py:563, which no single-process run of
the concatenated batch reproduces)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _model(dtype):
    from leod_b200.config import make_model_cfg
    from leod_b200.models.detection.yolox_extension.models.detector import YoloXDetector
    torch.manual_seed(0)
    m = YoloXDetector(make_model_cfg(embed_dim=16, dim_head=8, partition_size=(2, 3), num_classes=2, fpn_depth=0.33, input_channels=4,
                                     in_res_hw=(64, 96), compute_dtype=dtype))
    with torch.no_grad():
        for k, p in m.named_parameters():
            if k.endswith('gamma'):
                p.fill_(0.5)
    return m


def _data(Bt):
    g = torch.Generator().manual_seed(3)
    x = ((torch.rand(2, Bt, 4, 60, 90, generator=g) < 0.2).float() * torch.randint(1, 5, (2, Bt, 4, 60, 90), generator=g)).to(torch.uint8)
    feats = {s: torch.randn(Bt, c, 64 // st, 96 // st, generator=g) for s, c, st in zip((2, 3, 4), (32, 64, 128), (8, 16, 32))}
    labels = torch.zeros(Bt, 2, 7)
    labels[:, 0] = torch.tensor([0, 30., 30., 20., 16., 1., 1.])
    return x, feats, labels, g


def _run(m, x, feats, labels, R, Rf, scale, sync):
    """One forward/backward of the detect engine (smooth loss) and of the backbone (linear loss); returns everything compared."""
    dev = x.device
    B = x.shape[1]
    e, bb = m.detect_engine, m.backbone
    m.train()
    e.prepare()
    e._ensure_grad_buffer()
    e.flat_grads.zero_()
    fl = [feats[s] for s in (2, 3, 4)]
    preds, losses, _ = e._forward(fl, labels, training=True)
    e.backward_from_raw_grad(fl, R * scale)
    if sync:
        sync(e.flat_grads)
    f, _ = bb.forward_sequence(x, None)
    bb.flat_grads.zero_()
    (sum((f[s].float() * Rf[s]).sum() for s in Rf) * scale).backward()
    torch.cuda.synchronize()
    return dict(preds=preds.cpu(), det_grads=e.flat_grads.cpu().clone(), bn=e.flat_buffers.cpu().clone(), bb_grads=bb.flat_grads.cpu().clone(),
                loss=losses.cpu())


def _worker(rank, world, port, dtype, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        from leod_b200.modules.utils.distributed import allreduce_mean_
        Bt = 4
        x, feats, labels, g = _data(Bt)
        A = 126
        R = torch.randn(Bt, A, 8, generator=g) / A ** 0.5
        R[..., 7] = 0
        Rf = {s: torch.randn(2, Bt, c, 64 // st, 96 // st, generator=g) for s, c, st in zip((1, 2, 3, 4), (16, 32, 64, 128), (4, 8, 16, 32))}
        half = slice(rank * Bt // world, (rank + 1) * Bt // world)
        m = _model(dtype).to(dev)
        m.detect_engine.set_sync_batchnorm()
        m.backbone.grad_sync = lambda gflat: allreduce_mean_([gflat])
        res = _run(m, x[:, half].to(dev), {s: v[half].to(dev) for s, v in feats.items()}, labels[half].to(dev), R[half].to(dev),
                   {s: v[:, half].to(dev) for s, v in Rf.items()}, 1.0, lambda gflat: allreduce_mean_([gflat]))
        dist.barrier()
        single = None
        if rank == 0:     # the same data in one process, loss = mean over the two ranks' losses
            m1 = _model(dtype).to(dev)
            single = _run(m1, x.to(dev), {s: v.to(dev) for s, v in feats.items()}, labels.to(dev), R.to(dev), {s: v.to(dev) for s, v in Rf.items()},
                          1.0 / world, None)
        out.put((rank, res, single))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('dtype', ['fp32', 'bf16'])
def test_two_ranks_equal_single_process_double_batch(dtype):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (run with gpurun --gpus 2)')
    from helpers import rel_err
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, dtype, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted((out.get(timeout=300) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    single = got[0][2]
    tol = 1e-4 if dtype == 'fp32' else 3e-2
    for rank, res, _ in got:
        rows = slice(rank * 2, rank * 2 + 2)
        assert rel_err(res['preds'], single['preds'][rows]) < tol, ('preds', rank)
        assert rel_err(res['bn'], single['bn']) < tol, ('bn running statistics', rank)
        assert rel_err(res['det_grads'], single['det_grads']) < (1e-3 if dtype == 'fp32' else 6e-2), ('neck/head grads', rank)
        assert rel_err(res['bb_grads'], single['bb_grads']) < (1e-3 if dtype == 'fp32' else 6e-2), ('backbone grads', rank)
    assert rel_err(got[0][1]['det_grads'], got[1][1]['det_grads']) == 0.0      # identical replicas after the all-reduce
    print(f'[2 ranks/{dtype}] preds err {rel_err(got[1][1]["preds"], single["preds"][2:4]):.2e}, neck/head grad err '
          f'{rel_err(got[0][1]["det_grads"], single["det_grads"]):.2e}, backbone grad err {rel_err(got[0][1]["bb_grads"], single["bb_grads"]):.2e}')
