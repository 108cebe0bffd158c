"""Training / evaluation step loops.  Mirror of modules/detection.py:24 (class Module):
`training_step` (:150-298), `_val_test_step_impl` (:300-401), `configure_optimizers` (:485-518),
`load_weight` (:583-594), on the same batch-dict contract (data/utils/types.py:15-32).

pytorch_lightning is optional: when it is importable the class derives from pl.LightningModule and
plugs into the reference's train.py / val.py unchanged; otherwise it is a plain nn.Module driven by
bench.py / the tests.  Logging, visualisation and the Prophesee evaluator are out of scope
 — `training_step` returns the same {'loss': ...} (+ 'log_dict').
"""
from typing import Any, Dict, Optional

import torch
import torch.nn as nn

from leod_b200.data.utils.types import DataType, dget
from leod_b200.models.detection.yolox.utils.boxes import postprocess
from leod_b200.models.detection.yolox_extension.models.detector import YoloXDetector
from .utils.detection import DATA_KEY, WORKER_ID_KEY, BackboneFeatureSelector, Mode, RNNStates, merge_mixed_batches, mode_2_string
from .utils.ssod import fused_adamw_ema

try:  # pragma: no cover - not installed in the build image
    import pytorch_lightning as pl
    _Base = pl.LightningModule
except Exception:  # noqa: BLE001
    _Base = nn.Module


def _h2d_async(t: torch.Tensor, device) -> torch.Tensor:
    """Small host tensor -> device without blocking the host and without touching the copy engines (leod_upload_small: an SM
    kernel reads a pinned staging buffer).  A cudaMemcpyAsync here would queue behind the input pipeline's bulk upload of the
    next batch on the host->device engine and stall the compute stream at the start of every step."""
    if t.is_cuda or torch.device(device).type != 'cuda':
        return t.to(device)
    from leod_b200 import _lib
    return _lib.upload_small(t, device)


class _SelectFrames(torch.autograd.Function):
    """`features[ti, bi]` for the labelled (timestep, sequence) pairs, in the memory layout both neighbours use: the backbone's
    feature maps are NCHW-shaped views of channels-last storage and the neck reads channels-last, but plain advanced indexing
    returns an NCHW-contiguous copy (one transposing copy per pyramid level on the way in) and its backward builds an
    NCHW-contiguous dense gradient (a second, window-sized transposing copy on the way back).  Here the gather and the scatter of
    the gradient both happen on the channels-last tensors."""

    @staticmethod
    def forward(ctx, v, ti, bi):
        vn = v.permute(0, 1, 3, 4, 2)                       # [L, B, h, w, C]; contiguous for the library's outputs
        out = vn[ti, bi]                                    # [n, h, w, C]
        ctx.save_for_backward(ti, bi)
        ctx.full_shape = tuple(vn.shape)
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        ti, bi = ctx.saved_tensors
        full = torch.zeros(ctx.full_shape, dtype=g.dtype, device=g.device)
        full[ti, bi] = g.permute(0, 2, 3, 1)                # the (ti, bi) pairs are distinct: a plain scatter
        return full.permute(0, 1, 4, 2, 3), None, None


def get_subsample_label_idx(L: int, use_every: int = -1, remove_every: int = -1):
    """modules/utils/ssod.py:19-37."""
    assert use_every == -1 or remove_every == -1
    all_idx = list(range(L))
    if use_every == 1:
        return tuple(all_idx)
    if use_every > 0:
        use_idx = all_idx[1::use_every]
    elif remove_every > 0:
        use_idx = sorted(set(all_idx) - set(all_idx[::remove_every]))
    else:
        raise ValueError('Either use_every or remove_every must be > 0')
    if L - 1 not in use_idx:
        use_idx.append(L - 1)
    return tuple(use_idx)


class Module(_Base):
    def __init__(self, full_config, ssod: bool = False):
        super().__init__()
        self.full_config = full_config
        self.mdl_config = full_config.model
        self.num_classes = self.mdl_config.head.num_classes
        self.mdl = YoloXDetector(self.mdl_config, ssod=ssod)
        self.dst_config = full_config.get('dataset', None) if hasattr(full_config, 'get') else None
        L = self.dst_config.sequence_length if self.dst_config is not None else 1
        self.label_subsample_idx = get_subsample_label_idx(L=L, use_every=self.mdl_config.get('use_label_every', 1))
        self.mode_2_rnn_states: Dict[Mode, RNNStates] = {m: RNNStates() for m in Mode}
        self._opt_state = None
        # True: one library call per BPTT window (forward_sequence); False: the reference's per-timestep calls
        self.use_sequence_kernel = True

    # ------------------------------------------------------------------ data
    def get_data_from_batch(self, batch: Any):
        """modules/detection.py:129-148.  The uint8 -> float cast and the zero padding to `in_res_hw` are fused into the
        stem's patch loader, so the event tensors pass through untouched.  In training the label sub-sampling of :137-147
        (`model.use_label_every`) is applied: timesteps outside `label_subsample_idx` keep ground truth only."""
        data = batch[DATA_KEY]
        if not self.training:
            return data
        sparse_obj_labels = dget(data, DataType.OBJLABELS_SEQ)
        if len(self.label_subsample_idx) < len(sparse_obj_labels):
            for tidx in range(len(sparse_obj_labels)):
                if tidx not in self.label_subsample_idx:
                    sparse_obj_labels[tidx].set_non_gt_labels_to_none_()
        return data

    # ------------------------------------------------------------------ train
    def training_step(self, batch: Any, batch_idx: int = 0, log: bool = False):
        batch = merge_mixed_batches(batch)      # {RANDOM: ..., STREAM: ...} of the mixed sampler -> one batch (:152)
        data = self.get_data_from_batch(batch)
        worker_id = batch[WORKER_ID_KEY]
        mode = Mode.TRAIN
        ev_seq = dget(data, DataType.EV_REPR)
        sparse_obj_labels = dget(data, DataType.OBJLABELS_SEQ)
        is_first_sample = dget(data, DataType.IS_FIRST_SAMPLE)
        self.mode_2_rnn_states[mode].reset(worker_id=worker_id, indices_or_bool_tensor=is_first_sample)
        L = len(ev_seq)
        assert L > 0
        B = len(sparse_obj_labels[0])
        prev_states = self.mode_2_rnn_states[mode].get_states(worker_id=worker_id)
        obj_labels = []
        ignore = self.mdl_config.get('ignore_image', False)
        ignore_label = self.mdl_config.head.get('ignore_label', 1024)
        if self.use_sequence_kernel:
            # the time loop of modules/detection.py:188-224 runs inside the library (one call per window).  Everything
            # that touches the host (label lists, index tensors, their host->device copies) is prepared BEFORE the
            # window is enqueued: a pageable copy issued afterwards would block the host until the backbone has finished
            # and leave the GPU idle while the head is being set up.
            ev = ev_seq if torch.is_tensor(ev_seq) else torch.stack(list(ev_seq))
            t_idx, b_idx = [], []
            for tidx in range(L):
                current_labels, valid_idx = sparse_obj_labels[tidx].get_valid_labels_and_batch_indices(
                    ignore=ignore, ignore_label=ignore_label)
                if len(current_labels) > 0:
                    obj_labels.extend(current_labels)
                    t_idx.extend([tidx] * len(valid_idx))
                    b_idx.extend(valid_idx)
            assert len(obj_labels) > 0
            ti = _h2d_async(torch.as_tensor(t_idx), ev.device)
            bi = _h2d_async(torch.as_tensor(b_idx), ev.device)
            labels_yolox = type(obj_labels[0]).get_labels_as_batched_tensor(obj_label_list=obj_labels, format_='yolox')
            labels_yolox = _h2d_async(labels_yolox.to(torch.float32), ev.device)
            feats_all, prev_states = self.mdl.backbone.forward_sequence(ev, prev_states)
            sel = {k: _SelectFrames.apply(v, ti, bi) for k, v in feats_all.items() if k in self.mdl.fpn.in_features}
        else:
            selector = BackboneFeatureSelector()
            for tidx in range(L):
                feats, states = self.mdl.forward_backbone(x=ev_seq[tidx], previous_states=prev_states, token_mask=None)
                prev_states = states
                current_labels, valid_idx = sparse_obj_labels[tidx].get_valid_labels_and_batch_indices(
                    ignore=ignore, ignore_label=ignore_label)
                if len(current_labels) > 0:
                    selector.add_backbone_features(backbone_features=feats, selected_indices=valid_idx)
                    obj_labels.extend(current_labels)
            sel = selector.get_batched_backbone_features()
        self.mode_2_rnn_states[mode].save_states_and_detach(worker_id=worker_id, states=prev_states)
        assert len(obj_labels) > 0
        if not self.use_sequence_kernel:
            labels_yolox = type(obj_labels[0]).get_labels_as_batched_tensor(obj_label_list=obj_labels, format_='yolox')
            labels_yolox = labels_yolox.to(device=ev_seq[0].device, dtype=torch.float32)
        predictions, losses = self.mdl.forward_detect(backbone_features=sel, targets=labels_yolox)
        assert losses is not None and 'loss' in losses
        out = {'loss': losses['loss']}
        prefix = f'{mode_2_string[mode]}/'
        out['log_dict'] = {f'{prefix}{k}': v for k, v in losses.items()}
        out['predictions'] = predictions
        return out

    # ------------------------------------------------------------------ eval
    @torch.inference_mode()
    def _val_test_step_impl(self, batch: Any, mode: Mode = Mode.VAL):
        """modules/detection.py:300-401 without the evaluator buffer: returns the post-processed
        predictions of the labelled frames and the labels."""
        data = self.get_data_from_batch(batch)
        worker_id = batch[WORKER_ID_KEY]
        ev_seq = dget(data, DataType.EV_REPR)
        sparse_obj_labels = dget(data, DataType.OBJLABELS_SEQ)
        is_first_sample = dget(data, DataType.IS_FIRST_SAMPLE)
        self.mode_2_rnn_states[mode].reset(worker_id=worker_id, indices_or_bool_tensor=is_first_sample)
        prev_states = self.mode_2_rnn_states[mode].get_states(worker_id=worker_id)
        selector = BackboneFeatureSelector()
        obj_labels = []
        for tidx in range(len(ev_seq)):
            feats, states = self.mdl.forward_backbone(x=ev_seq[tidx], previous_states=prev_states)
            prev_states = states
            current_labels, valid_idx = sparse_obj_labels[tidx].get_valid_labels_and_batch_indices()
            if len(current_labels) > 0:
                selector.add_backbone_features(backbone_features=feats, selected_indices=valid_idx)
                obj_labels.extend(current_labels)
        self.mode_2_rnn_states[mode].save_states_and_detach(worker_id=worker_id, states=prev_states)
        if len(obj_labels) == 0:
            return {'skip': True}
        predictions, _ = self.mdl.forward_detect(backbone_features=selector.get_batched_backbone_features())
        pp = self.mdl_config.postprocess
        pred_processed = postprocess(prediction=predictions, num_classes=self.num_classes,
                                     conf_thre=pp.confidence_threshold, nms_thre=pp.nms_threshold)
        return {'labels': obj_labels, 'predictions': pred_processed, 'skip': False}

    @torch.inference_mode()
    def predict_one_seq(self, batch: Any, chunk: int = 128):
        """modules/detection.py:520-581: run the model over ONE full event sequence (batch size 1) and return
        (per-timestep detections `L`-len list of [N_t, 7] | None, the event tensors [L, C, H, W], the `L`-len label list).
        The time loop runs in the library `chunk` timesteps at a time (the reference batches the head over 128 timesteps, :545)."""
        mode = Mode.TEST
        data = self.get_data_from_batch(batch)
        worker_id = batch[WORKER_ID_KEY]
        ev_seq = dget(data, DataType.EV_REPR)
        ev = (ev_seq if torch.is_tensor(ev_seq) else torch.stack(list(ev_seq), dim=0)).cuda()
        sparse_obj_labels = dget(data, DataType.OBJLABELS_SEQ)
        is_first_sample = dget(data, DataType.IS_FIRST_SAMPLE)
        assert ev.shape[1] == len(sparse_obj_labels[0]) == is_first_sample.shape[0] == 1
        self.mode_2_rnn_states[mode].reset(worker_id=worker_id, indices_or_bool_tensor=is_first_sample)
        L = ev.shape[0]
        assert L > 0
        pp = self.mdl_config.postprocess
        prev_states, all_preds = None, []       # like the reference, the sequence starts from a zero state (:543)
        for t0 in range(0, L, chunk):
            feats_all, prev_states = self.mdl.backbone.forward_sequence(ev[t0:t0 + chunk], prev_states)
            sel = {k: v[:, 0] for k, v in feats_all.items() if k in self.mdl.fpn.in_features}
            predictions, _ = self.mdl.forward_detect(backbone_features=sel)
            all_preds.extend(postprocess(prediction=predictions, num_classes=self.num_classes, conf_thre=pp.confidence_threshold,
                                         nms_thre=pp.nms_threshold))
        return all_preds, ev[:, 0], [lbl[0] for lbl in sparse_obj_labels]

    def validation_step(self, batch, batch_idx=0):
        return self._val_test_step_impl(batch, Mode.VAL)

    def test_step(self, batch, batch_idx=0):
        return self._val_test_step_impl(batch, Mode.TEST)

    # ------------------------------------------------------------------ optimisation
    def configure_optimizers(self):
        """modules/detection.py:485-518: AdamW(lr, weight_decay) + OneCycleLR (linear anneal)."""
        tc = self.full_config.training
        opt = torch.optim.AdamW(self.mdl.parameters(), lr=tc.learning_rate, weight_decay=tc.get('weight_decay', 0))
        sched = tc.get('lr_scheduler', None)
        if sched is None or not sched.get('use', False):
            return opt
        total = sched.total_steps if hasattr(sched, 'total_steps') else tc.max_steps
        # the reference interprets its final_div_factor as max_lr / final_lr (modules/detection.py:499-501)
        s = torch.optim.lr_scheduler.OneCycleLR(optimizer=opt, max_lr=tc.learning_rate, div_factor=sched.div_factor,
                                                final_div_factor=sched.final_div_factor / sched.div_factor, total_steps=total,
                                                pct_start=sched.pct_start, cycle_momentum=False, anneal_strategy='linear')
        return {'optimizer': opt, 'lr_scheduler': {'scheduler': s, 'interval': 'step', 'frequency': 1, 'strict': True}}

    def load_weight(self, ckpt_path: str, strict: bool = True):
        """modules/detection.py:583-594: raw state dict or {'state_dict': ...}."""
        ckpt = torch.load(ckpt_path, map_location='cpu')
        if 'state_dict' in ckpt:
            ckpt = ckpt['state_dict']
        self.load_state_dict(ckpt, strict=strict)


def one_cycle_lr(step: int, max_lr: float, total_steps: int, pct_start: float = 0.005, div_factor: float = 20.0,
                 final_div_factor: float = 10000.0) -> float:
    """Learning rate of optimizer step `step` (0-based) under the reference's schedule (modules/detection.py:495-510):
    torch OneCycleLR, two linear phases, no momentum cycling, with the reference's reading of `final_div_factor` as
    max_lr / final_lr.  Closed form for `FlatOptimizer.step(lr=...)`, which has no param_groups for a torch scheduler."""
    initial = max_lr / div_factor
    final = initial / (final_div_factor / div_factor)
    up_end = float(pct_start * total_steps) - 1
    down_end = total_steps - 1
    if step <= up_end:
        pct = step / up_end if up_end > 0 else 1.0
        return (max_lr - initial) * pct + initial
    pct = (step - up_end) / (down_end - up_end)
    return (final - max_lr) * pct + max_lr


class FlatOptimizer:
    """AdamW + clip-by-value (+ optional teacher EMA) over the model's flat parameter buffers: one
    kernel launch per buffer per step (leod_adamw_ema) instead of ~4 launches x 376 tensors."""

    def __init__(self, detector: YoloXDetector, lr: float, weight_decay: float = 0.0, clip_value: float = 1.0,
                 betas=(0.9, 0.999), eps: float = 1e-8, ema: bool = False, ema_alpha: float = 0.999, ema_buffers=None):
        self.detector = detector
        self.lr, self.wd, self.clip, self.betas, self.eps = lr, weight_decay, clip_value, betas, eps
        self.step_count = 0
        self.ema_alpha = ema_alpha
        bb, de = detector.backbone, detector.detect_engine
        # two flat fp32 buffers hold every parameter: the backbone's and the neck+head's (the nn.Parameters are views)
        self.bufs = [(bb.flat_params, bb.flat_grads), (de.flat_params, de.flat_grads)]
        self.m = [torch.zeros_like(p) for p, _ in self.bufs]
        self.v = [torch.zeros_like(p) for p, _ in self.bufs]
        # teacher EMA in the same launch: into private clones, or straight into a teacher model's flat parameter buffers
        if ema_buffers is not None:
            assert len(ema_buffers) == 2 and all(e.shape == p.shape and e.device == p.device for e, (p, _) in zip(ema_buffers, self.bufs))
            self.ema = list(ema_buffers)
        else:
            self.ema = [p.clone() for p, _ in self.bufs] if ema else [None, None]

    def zero_grad(self):
        for _, g in self.bufs:
            g.zero_()

    def step(self, lr: Optional[float] = None):
        from .utils.ssod import ema_alpha_at
        self.step_count += 1
        a = ema_alpha_at(self.step_count - 1, self.ema_alpha)
        for (p, g), m, v, e in zip(self.bufs, self.m, self.v, self.ema):
            fused_adamw_ema(p, g, m, v, step=self.step_count, lr=self.lr if lr is None else lr, beta1=self.betas[0],
                            beta2=self.betas[1], eps=self.eps, weight_decay=self.wd, clip_value=self.clip, ema=e, ema_alpha=a)
        self.detector.backbone.mark_params_updated()
        self.detector.detect_engine.mark_params_updated()
