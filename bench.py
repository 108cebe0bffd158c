#!/usr/bin/env python
"""Benchmark of the LEOD hot path on B200 (contract: see the task statement / DESIGN.md §Measurement).

    python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU port of the reference path

One "step" = one training step of RVT-small on a Gen1-shaped batch: 8 sequences x 21 event frames
(uint8 voxel tensors 20x240x304), recurrent backbone forward + backward through all 21 timesteps,
YOLOX neck/head + SimOTA loss on the labelled frames, gradient clip + AdamW.  metric = event-frames/s.
Prints ONE JSON line on rank 0.
"""
import argparse
import datetime
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'event-frames/s (fwd+bwd) RVT-S Gen1 seq-len 21'
B, L, CIN, FH, FW = 8, 21, 20, 240, 304
LABEL_T = (10, 20)                       # labelled timesteps of every sequence (sparse GT, ~2 Hz)
GFLOP_BACKBONE, GFLOP_HEAD = 4.115, 1.760   # per frame forward, BASELINE.md §2


def synth_batch(seed, device=None, pin=False):
    """Seeded synthetic Gen1-shaped batch: ~90% zeros, counts 1+Poisson(1.5)."""
    import torch
    from leod_b200.data.labels import ObjectLabels, SparselyBatchedObjectLabels
    from leod_b200.data.utils.types import DataType
    g = torch.Generator().manual_seed(seed)
    ev = (torch.rand(L, B, CIN, FH, FW, generator=g) < 0.1)
    ev = (ev * (1 + torch.poisson(torch.full((L, B, CIN, FH, FW), 1.5), generator=g)).clamp(max=255)).to(torch.uint8)
    if pin:
        ev = ev.pin_memory()
    labels = []
    for t in range(L):
        row = []
        for b in range(B):
            if t not in LABEL_T:
                row.append(None)
                continue
            n = int(torch.randint(1, 9, (1,), generator=g))
            w = torch.rand(n, generator=g) * 110 + 10
            h = torch.rand(n, generator=g) * 90 + 10
            x = torch.rand(n, generator=g) * (FW - w)
            y = torch.rand(n, generator=g) * (FH - h)
            cls = torch.randint(0, 2, (n,), generator=g).float()
            lab = torch.stack((torch.ones(n), x, y, w, h, cls, torch.ones(n), torch.ones(n)), 1)
            row.append(ObjectLabels(lab, (FH, FW)))
        labels.append(SparselyBatchedObjectLabels(row))
    # mixed sampling (modules/data/genx.py:120-144): the random-access half restarts every step, the
    # streaming half carries its recurrent state over
    first = torch.arange(B) < B // 2
    return ev, labels, first


def make_batch(ev_dev, labels, first):
    from leod_b200.data.utils.types import DataType
    return {'worker_id': 0, 'data': {DataType.EV_REPR: ev_dev,
                                     DataType.OBJLABELS_SEQ: labels, DataType.IS_FIRST_SAMPLE: first}}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith('active')})
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(sm)}


def cpu_port_step(nb, nl, threads, seed=0):
    """The oracle (CPU port of the reference path) on a bounded sample: nb sequences x nl frames,
    forward + loss + backward.  Returns seconds per step."""
    import torch
    from oracle import rvt, yolox
    from oracle.config import ModelCfg
    from leod_b200.config import make_model_cfg
    from leod_b200.models.detection.yolox_extension.models.detector import YoloXDetector
    torch.set_num_threads(threads)
    ocfg = ModelCfg.named('small', 'gen1')
    if not hasattr(cpu_port_step, 'sd'):
        torch.manual_seed(0)
        m = YoloXDetector(make_model_cfg(size='small', dataset='gen1'))
        cpu_port_step.sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and 'running' not in k) for k, v in cpu_port_step.sd.items()}
    g = torch.Generator().manual_seed(seed)
    x = ((torch.rand(nl, nb, CIN, FH, FW, generator=g) < 0.1).float() * 2)
    labels = torch.zeros(nb, 2, 7)
    labels[:, 0] = torch.tensor([0, 100., 100., 40., 30., 1., 1.])
    labels[:, 1] = torch.tensor([1, 200., 150., 60., 50., 1., 1.])
    t0 = time.perf_counter()
    states = None
    for t in range(nl):
        feats, states = rvt.backbone_forward(rvt.pad_input(x[t], (256, 320)), states, sd, ocfg)
    _, losses = yolox.detect_forward(feats, sd, ocfg, targets=labels, training=True)
    losses['loss'].backward()
    return time.perf_counter() - t0


def run_reference(args, rank, world):
    """--impl reference: the reference path's CPU implementation (the oracle port: /root/reference is
    not present on the GPU box) on the host cores, same metric/config, bounded sample per step."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    nb, nl = 2, 4
    for _ in range(args.warmup):
        cpu_port_step(nb, nl, threads)
    t0 = time.perf_counter()
    for i in range(args.steps):
        cpu_port_step(nb, nl, threads, seed=i)
    dt = (time.perf_counter() - t0) / args.steps
    fps = nb * nl / dt
    sample = f'{nb} sequences x {nl} frames (of the 8x21 workload) per step, fp32 torch eager, fwd+loss+bwd, no optimizer'
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': 'event-frames/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'RVT-small Gen1 240x304 bins=10, batch 8, seq-len 21, fwd+bwd (bounded sample)', 'sample': sample},
        'cpu_baseline': {'value': fps, 'unit': 'event-frames/s', 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': fps, 'unit': 'event-frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


def run_sweep(args, rank, local_rank, world):
    """Secondary workload (BASELINE configs[3] shape): the teacher pseudo-label sweep — PseudoLabeler.predict_step on
    chunks of 16 Gen1 sequences x 21 frames with hflip TTA (32 views), head + NMS + label filters on every frame.
    Sequences are sharded over ranks (no data-path collective); metric = view-frames/s through backbone+head+NMS."""
    import torch
    import torch.distributed as dist
    from leod_b200 import _lib
    from leod_b200.config import Node, make_model_cfg
    from leod_b200.data.labels import SparselyBatchedObjectLabels
    from leod_b200.data.utils.types import DataType
    from leod_b200.modules.pseudo_labeler import PseudoLabeler
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=180))
    torch.manual_seed(0)
    SB = 16
    mcfg = make_model_cfg(size='small', dataset='gen1', compute_dtype=args.dtype, conf_thre=0.01)
    full = Node(model=mcfg, dataset=dict(sequence_length=L, name='gen1', downsample_by_factor_2=False),
                tta=dict(enable=True, hflip=True, tflip=False), use_gt=True)
    pl = PseudoLabeler(full).to(dev).eval()
    g = torch.Generator().manual_seed(100 + rank)
    host = []
    for i in range(3):
        ev = (torch.rand(L, SB, CIN, FH, FW, generator=g) < 0.1)
        ev = (ev * (1 + torch.poisson(torch.full((L, SB, CIN, FH, FW), 1.5), generator=g)).clamp(max=255)).to(torch.uint8).pin_memory()
        host.append(ev)
    none_labels = [SparselyBatchedObjectLabels([None] * SB) for _ in range(L)]

    def batch_of(ev_dev, first):
        return {'worker_id': 0, 'data': {DataType.EV_REPR: ev_dev, DataType.OBJLABELS_SEQ: none_labels,
                                         DataType.SKIPPED_OBJLABELS_SEQ: none_labels,
                                         DataType.IS_FIRST_SAMPLE: torch.full((SB,), first, dtype=torch.bool)}}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    resident = [e.to(dev) for e in host]
    lib = _lib.lib()
    for i in range(args.warmup):
        pl.predict_step(batch_of(resident[i % 3], i == 0))
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = lib.leod_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        out = pl.predict_step(batch_of(resident[i % 3], False))
    e1.record()
    barrier()
    launches = lib.leod_launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    frames = world * 2 * SB * L * args.steps
    barrier()
    e0.record()
    for i in range(args.steps):
        out = pl.predict_step(batch_of(host[i % 3].to(dev, non_blocking=True), False))
        n_boxes = sum(len(l) for row in out[0] for l in row if l is not None)    # labels read back on the host
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = frames / (float(t) * 1e-3)
    if rank == 0:
        print(json.dumps({
            'metric': 'event-frames/s (teacher sweep: backbone + head + NMS + label filters) RVT-S Gen1 seq-len 21', 'value': frames / (ms * 1e-3),
            'unit': 'event-frames/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
            'config': {'workload': 'teacher pseudo-label sweep, RVT-small Gen1 240x304, 16 sequences x 21 frames per step per GPU, hflip TTA '
                                   '(32 views), head+NMS on every frame, conf 0.01, nms 0.45, thresholds (0.6, 0.3) (BASELINE configs[3] shape)',
                       'note': 'random-init weights: almost no box passes the confidence filter, NMS work is minimal',
                       'parallelism': f'dp{world} (sequences sharded, no collective)'},
            'e2e': {'value': e2e, 'unit': 'event-frames/s', 'h2d_bytes_per_step': host[0].numel(), 'd2h_bytes_per_step': 4 * 2 * SB * L},
            'gpu_launches': int(launches), 'roofline': None, 'cpu_baseline': None, 'clocks': clocks, 'labels_last_step': n_boxes}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='leod_b200', choices=['leod_b200', 'reference'])
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--gemm-impl', type=int, default=None, help='0 SIMT, 1 tcgen05 (debug)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--profile-kinds', action='store_true', help='print the per-kernel-class breakdown to stderr')
    ap.add_argument('--profile-csv', default=None, help='write one line per launch of the roofline pass to this CSV')
    ap.add_argument('--workload', default='train', choices=['train', 'sweep'],
                    help='train: BASELINE configs[1] (the headline metric); sweep: teacher pseudo-label sweep, configs[3] shape')
    ap.add_argument('--e2e-lag', type=int, default=2, help='the loss of step i is read back on the host after step i+lag was enqueued')
    ap.add_argument('--copy-streams', type=int, default=4, help='streams the per-step host->device upload is split over (e2e leg)')
    ap.add_argument('--phases', action='store_true', help='after the timed regions, time the phases of 3 extra steps (stderr)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
        return run_reference(args, rank, world)
    if args.workload == 'sweep':
        return run_sweep(args, rank, local_rank, world)
    assert args.warmup >= 3, 'timing rules: at least 3 warm-up steps'

    import torch
    import torch.distributed as dist
    from leod_b200 import _lib
    from leod_b200.config import make_model_cfg, Node
    from leod_b200.modules.detection import FlatOptimizer, Module

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=180))
    torch.manual_seed(0)
    full_cfg = Node(model=make_model_cfg(size='small', dataset='gen1', compute_dtype=args.dtype),
                    dataset=dict(sequence_length=L, name='gen1'))
    module = Module(full_cfg)
    module.to(dev).train()
    if world > 1:  # the reference trains with sync_batchnorm under DDP (train.py:247): one statistics exchange per dependency level
        module.mdl.detect_engine.set_sync_batchnorm()
    bb = module.mdl.backbone
    if args.gemm_impl is not None:
        bb.set_gemm_impl(args.gemm_impl)
    opt = FlatOptimizer(module.mdl, lr=2e-4, weight_decay=0.0, clip_value=1.0)
    if world > 1:
        from leod_b200.modules.utils.distributed import allreduce_mean_, broadcast_flat
        broadcast_flat([p for p, _ in opt.bufs], src=0)      # identical replicas, as DDP's initial broadcast
        bb.mark_params_updated()
        module.mdl.detect_engine.mark_params_updated()
        bb.grad_sync = lambda flat_grad: allreduce_mean_([flat_grad])
        module.mdl.detect_engine.grad_sync = lambda flat_grad: allreduce_mean_([flat_grad])

    # several distinct batches so consecutive steps do not re-read the same 25 MB of input
    n_batches = 4
    host = [synth_batch(1000 * rank + i, pin=True) for i in range(n_batches)]
    resident = [(ev.to(dev), lab, first.to(dev)) for ev, lab, first in host]

    def train_step(ev_dev, labels, first):
        opt.zero_grad()
        out = module.training_step(make_batch(ev_dev, labels, first))
        out['loss'].backward()
        opt.step()
        return out['loss']

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    lib = _lib.lib()
    for i in range(args.warmup):
        train_step(*resident[i % n_batches])
    barrier()

    # ---- timed region 1: inputs resident in HBM -> `value`
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.leod_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    h0 = time.perf_counter()
    for i in range(args.steps):
        loss = train_step(*resident[i % n_batches])
    host_enqueue_ms = (time.perf_counter() - h0) * 1e3 / args.steps     # host time to enqueue one step (no sync inside)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.leod_launch_count() - launches0
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    frames = world * B * L * args.steps
    value = frames / (ms * 1e-3)

    # ---- timed region 2: end to end through the public API with HOST buffers -> `e2e`
    # Every step's uint8 batch comes from pinned host memory inside the timed region; the copy of step i+1 is issued
    # on a copy stream before step i computes (double buffering, as a prefetching input pipeline does), and the loss
    # of every step is read back to the host.
    h2d = host[0][0].numel()
    NCOPY = args.copy_streams     # the upload is split along L over several streams (one DMA engine does not fill the link)
    copy_streams = [torch.cuda.Stream() for _ in range(NCOPY)]
    dev_buf = [torch.empty_like(host[0][0], device=dev) for _ in range(2)]
    copied = [[torch.cuda.Event() for _ in range(NCOPY)] for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    bounds = [L * c // NCOPY for c in range(NCOPY + 1)]

    def issue_copy(i):
        slot = i % 2
        src = host[i % n_batches][0]
        for c, cs in enumerate(copy_streams):
            with torch.cuda.stream(cs):
                cs.wait_event(consumed[slot])             # the step that last read this buffer has finished
                dev_buf[slot][bounds[c]:bounds[c + 1]].copy_(src[bounds[c]:bounds[c + 1]], non_blocking=True)
                copied[slot][c].record(cs)
    for c in consumed:
        c.record()
    first_dev = [h[2].to(dev) for h in host]
    lag = max(0, args.e2e_lag)
    NLAG = lag + 1
    loss_pinned = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(NLAG)]
    read_ev = [torch.cuda.Event() for _ in range(NLAG)]
    barrier()
    e0.record()
    issue_copy(0)
    for i in range(args.steps):
        slot = i % 2
        for ev_c in copied[slot]:
            torch.cuda.current_stream().wait_event(ev_c)
        _, lab, first = host[i % n_batches]
        loss = train_step(dev_buf[slot], lab, first_dev[i % n_batches])
        consumed[slot].record()
        # the next batch's bulk upload is issued AFTER this step has been enqueued: the step's own small uploads (label
        # and index tensors) would otherwise queue behind 245 MB on the host->device copy engine and stall the stream
        if i + 1 < args.steps:
            issue_copy(i + 1)
        # device -> host read of EVERY step's result, `lag` steps late (as an asynchronous logger does) so that the host
        # keeps enqueuing work while the step runs; the outstanding ones are read before the region closes
        loss_pinned[i % NLAG].copy_(loss.detach(), non_blocking=True)
        read_ev[i % NLAG].record()
        if i >= lag:
            read_ev[(i - lag) % NLAG].synchronize()
            loss_host = float(loss_pinned[(i - lag) % NLAG])
    for i in range(max(0, args.steps - lag), args.steps):
        read_ev[i % NLAG].synchronize()
        loss_host = float(loss_pinned[i % NLAG])
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = frames / (float(t) * 1e-3)
    clocks = sampler.stop() if rank == 0 else None      # sampled over both timed regions (value and e2e)

    # ---- roofline pass (not timed into value): CUDA events around every launch of each kernel class
    roof = None
    if rank == 0:
        lib.leod_profile_enable(1)
        if args.profile_csv:
            lib.leod_profile_csv(args.profile_csv.encode())
    train_step(*resident[0])          # every rank takes part (the step contains collectives); only rank 0 records
    torch.cuda.synchronize()
    if rank == 0:
        kinds = _lib.profile_collect()
        lib.leod_profile_enable(0)
        lib.leod_profile_csv(None)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        hbm_peak = peaks.get('hbm_gbs', 6650.0)
        tf_peak = peaks.get('bf16_tflops_sustained', 1400.0)
        src = 'measured' if peaks else 'fallback'
        dom = max(kinds, key=lambda k: kinds[k]['ms'])
        d = kinds[dom]
        avg_ms = d['ms'] / max(d['launches'], 1)
        gbs = d['bytes'] / d['launches'] / (avg_ms * 1e-3) / 1e9 if d['launches'] else 0.0
        tfs = d['flops'] / d['launches'] / (avg_ms * 1e-3) / 1e12 if d['launches'] else 0.0
        if gbs / hbm_peak >= tfs / tf_peak:
            roof = {'bound': 'hbm', 'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': gbs / hbm_peak}
        else:
            roof = {'bound': 'tensor', 'achieved': tfs, 'peak': tf_peak, 'unit': 'TFLOP/s', 'frac': tfs / tf_peak}
        traffic, traffic_src = None, None
        try:   # DRAM bytes per launch of the same kernel class from the committed ncu capture (profiles/)
            tj = json.load(open(os.path.join(ROOT, 'profiles', 'r01_gemm_nt_traffic.json')))
            if dom == 'gemm_nt':
                traffic = tj['dram_bytes_read_per_launch'] + tj['dram_bytes_write_per_launch']
                traffic_src = 'profiles/r01_gemm_nt_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch)'
        except Exception:
            pass
        roof.update({'traffic': traffic, 'traffic_source': traffic_src, 'algorithmic_bytes_per_launch': d['bytes'] / max(d['launches'], 1),
                     'kernel': dom, 'launches_per_step': d['launches'], 'avg_launch_us': avg_ms * 1e3,
                     'peak_source': src, 'share_of_kernel_time': d['ms'] / max(sum(k['ms'] for k in kinds.values()), 1e-9),
                     'classes_ms_per_step': {k: round(v['ms'], 3) for k, v in kinds.items()}})
        if args.profile_kinds:
            for k, v in kinds.items():
                print(f'  {k:14s} launches {v["launches"]:5d}  {v["ms"]:8.3f} ms  {v["flops"] / 1e9:9.1f} GF  {v["bytes"] / 1e6:9.1f} MB',
                      file=sys.stderr)

    if rank == 0 and args.phases:
        marks = []

        def mark(name):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

        def wrap(obj, attr, name):
            fn = getattr(obj, attr)

            def inner(*a, **k):
                mark('pre_' + name)
                r = fn(*a, **k)
                mark(name)
                return r
            setattr(obj, attr, inner)
            return fn
        o1 = wrap(bb, 'forward_sequence', 'backbone_fwd')
        o2 = wrap(module.mdl, 'forward_detect', 'neck_head_loss_fwd')
        acc = {}
        for i in range(3):
            marks.clear()
            mark('start')
            opt.zero_grad()
            out = module.training_step(make_batch(*resident[i % n_batches]))
            mark('host_glue_fwd')
            out['loss'].backward()
            mark('backward(head+backbone)')
            opt.step()
            mark('optimizer')
            torch.cuda.synchronize()
            for (n0, e0_), (n1, e1_) in zip(marks[:-1], marks[1:]):
                acc[n1] = acc.get(n1, 0.0) + e0_.elapsed_time(e1_) / 3
        bb.forward_sequence, module.mdl.forward_detect = o1, o2
        for k, v in acc.items():
            print(f'  phase {k:28s} {v:8.3f} ms', file=sys.stderr)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        nb, nl = 2, 4
        cpu_port_step(nb, nl, threads)
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 12.0:
            cpu_port_step(nb, nl, threads, seed=n)
            n += 1
        dt = (time.perf_counter() - t0) / n
        cpu = {'value': nb * nl / dt, 'unit': 'event-frames/s', 'cores': threads, 'kind': 'port',
               'sample': f'{n} steps of {nb} sequences x {nl} frames (of the 8x21 workload), fp32 torch eager oracle, fwd+loss+bwd'}

    if rank == 0:
        algo_tflop = B * L * 3 * GFLOP_BACKBONE / 1e3 + B * len(LABEL_T) * 3 * GFLOP_HEAD / 1e3
        print(json.dumps({
            'metric': METRIC, 'value': value, 'unit': 'event-frames/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': args.dtype, 'data': 'synthetic',
            'config': {'workload': 'RVT-small Gen1 240x304 bins=10, batch 8 per GPU, seq-len 21, fwd+bwd + AdamW (BASELINE configs[1])',
                       'labelled_frames_per_step': B * len(LABEL_T), 'l2_policy': f'{n_batches} rotating input batches + ~6 GB of '
                       'saved activations per step (>> 126 MB L2)', 'parallelism': f'dp{world}',
                       'algorithmic_tflop_per_step': algo_tflop, 'achieved_tflops': algo_tflop / (ms / args.steps * 1e-3) * world},
            'e2e': {'value': e2e_value, 'unit': 'event-frames/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4},
            'gpu_launches': int(launches), 'host_enqueue_ms_per_step': host_enqueue_ms, 'roofline': roof, 'cpu_baseline': cpu, 'clocks': clocks, 'loss': loss_host}))
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == '__main__':
    main()
