"""Online self-training step (BASELINE configs[4]): EMA-teacher forward -> pseudo labels -> student forward/backward ->
AdamW + teacher EMA.

The reference ships the pieces but not the loop: `ema_model_update` / `model_update` (modules/utils/ssod.py:429-460) are never
called, its self-training is offline (predict.py writes pseudo labels with a frozen teacher, train.py then trains on them with
`use_label_every: 1`, config/model/rnndet-soft.yaml:23).  This module composes the same pieces online, every step on the device:

  teacher  (eval, no grad):  backbone over the window (modules/pseudo_labeler.py:676-704), head on every frame, `postprocess` +
                             `pred2label` thresholds (:565-589, ssod.py:147-188)  ->  label rows per frame, on the device
  merge:                     frames that carry ground truth keep it (pseudo_labeler.py:706-770); the others take the pseudo labels
  student  (train):          the dense-label training step of modules/detection.py:150-298 on those targets
  update:                    clip + AdamW + EMA in one launch per flat buffer, the EMA written straight into the teacher's parameters
                             (ssod.py:429-438: parameters only); the teacher's BatchNorm running statistics follow the student's
                             (`copy_bn_buffers`, a deviation the dead reference code never had to decide: an EMA over parameters with
                             frozen initial statistics would leave the teacher's BatchNorm unusable).
Nothing here synchronises with the host: label counts stay on the device, the targets have a fixed capacity per frame.
"""
from typing import Any, Dict

import torch
import torch.nn as nn

from leod_b200.data.utils.types import DataType, dget
from leod_b200.models.detection.yolox.utils.boxes import postprocess_packed
from leod_b200.models.detection.yolox_extension.models.detector import YoloXDetector
from .detection import FlatOptimizer, Module, _h2d_async
from .utils.detection import WORKER_ID_KEY, Mode, RNNStates
from .utils.ssod import frame_hw, pred2label_packed


class SelfTrainingModule(nn.Module):
    def __init__(self, full_config, max_labels_per_frame: int = 32, copy_bn_buffers: bool = True):
        super().__init__()
        self.full_config = full_config
        self.student = Module(full_config)
        self.teacher = YoloXDetector(full_config.model)
        for p in self.teacher.parameters():
            p.requires_grad_(False)
        self.teacher_states = RNNStates()
        self.max_labels = int(max_labels_per_frame)
        self.copy_bn_buffers = copy_bn_buffers
        mc = full_config.model
        pl = mc.get('pseudo_label', None) or {}
        n = mc.head.num_classes
        self.obj_thresh = list(pl.get('obj_thresh', [0.9] * n))
        self.cls_thresh = list(pl.get('cls_thresh', [0.9] * n))
        ds = full_config.get('dataset', None)
        name = ds.get('name', 'gen1') if ds is not None else 'gen1'
        # Gen4 trains at half resolution (config/dataset/gen4.yaml:8-9)
        self.hw = frame_hw(name, downsampled_by_2=bool(ds.get('downsample_by_factor_2', name == 'gen4')) if ds is not None else False)
        self.optimizer = None
        self.sync_teacher_from_student()

    # ------------------------------------------------------------------ teacher <- student
    @torch.no_grad()
    def sync_teacher_from_student(self):
        s, t = self.student.mdl, self.teacher
        t.backbone.flat_params.copy_(s.backbone.flat_params.to(t.backbone.flat_params.device))
        t.detect_engine.flat_params.copy_(s.detect_engine.flat_params.to(t.detect_engine.flat_params.device))
        t.detect_engine.flat_buffers.copy_(s.detect_engine.flat_buffers.to(t.detect_engine.flat_buffers.device))
        t.backbone.mark_params_updated()
        t.detect_engine.mark_params_updated()

    def make_optimizer(self, lr: float, weight_decay: float = 0.0, clip_value: float = 1.0, ema_alpha: float = 0.999):
        t = self.teacher
        self.optimizer = _SelfTrainOptimizer(self, FlatOptimizer(
            self.student.mdl, lr=lr, weight_decay=weight_decay, clip_value=clip_value, ema=True, ema_alpha=ema_alpha,
            ema_buffers=[t.backbone.flat_params, t.detect_engine.flat_params]))
        return self.optimizer

    # ------------------------------------------------------------------ teacher pass
    @torch.no_grad()
    def teacher_targets(self, ev: torch.Tensor, worker_id: int, is_first_sample) -> torch.Tensor:
        """ev [L,B,C,H,W] -> yolox target rows [L*B, max_labels, 7] (cls, cx, cy, w, h, obj_conf, cls_conf), zero padded, frame order (t, b)."""
        t = self.teacher
        t.eval()
        L, B = ev.shape[0], ev.shape[1]
        self.teacher_states.reset(worker_id=worker_id, indices_or_bool_tensor=is_first_sample)
        prev = self.teacher_states.get_states(worker_id=worker_id)
        feats_all, states = t.backbone.forward_sequence(ev, prev)
        self.teacher_states.save_states_and_detach(worker_id=worker_id, states=states)
        sel = {k: v.reshape(L * B, *v.shape[2:]) for k, v in feats_all.items() if k in t.fpn.in_features}
        preds, _ = t.forward_detect(backbone_features=sel)
        pp = self.full_config.model.postprocess
        dets, cnt = postprocess_packed(preds, num_classes=t.yolox_head.num_classes, conf_thre=pp.confidence_threshold,
                                       nms_thre=pp.nms_threshold, max_det=self.max_labels)
        lab, n = pred2label_packed(dets, cnt, self.obj_thresh, self.cls_thresh, self.hw)     # rows (t=0, x, y, w, h, cls, cls_conf, obj_conf)
        keep = (torch.arange(self.max_labels, device=lab.device)[None, :] < n[:, None]).unsqueeze(-1)
        tg = torch.stack((lab[..., 5], lab[..., 1] + lab[..., 3] / 2, lab[..., 2] + lab[..., 4] / 2, lab[..., 3], lab[..., 4],
                          lab[..., 7], lab[..., 6]), -1)
        return torch.where(keep, tg, torch.zeros_like(tg))

    # ------------------------------------------------------------------ the step
    def training_step(self, batch: Any, batch_idx: int = 0) -> Dict[str, Any]:
        st = self.student
        data = st.get_data_from_batch(batch)
        worker_id = batch[WORKER_ID_KEY]
        ev_seq = dget(data, DataType.EV_REPR)
        gt_seq = dget(data, DataType.OBJLABELS_SEQ)
        is_first_sample = dget(data, DataType.IS_FIRST_SAMPLE)
        ev = ev_seq if torch.is_tensor(ev_seq) else torch.stack(list(ev_seq))
        L, B = ev.shape[0], ev.shape[1]
        # ground truth (host lists) -> one fixed-shape tensor + mask, uploaded asynchronously before any device work is queued
        gt = torch.zeros(L * B, self.max_labels, 7)
        has_gt = torch.zeros(L * B, dtype=torch.bool)
        for t in range(L):
            for b in range(B):
                lab = gt_seq[t][b] if gt_seq is not None else None
                if lab is not None and bool((lab.object_labels[:, 0] > 0).any()):
                    rows = lab.get_labels_as_tensors('yolox')[:self.max_labels]
                    gt[t * B + b, :rows.shape[0]] = rows
                    has_gt[t * B + b] = True
        gt, has_gt = _h2d_async(gt, ev.device), _h2d_async(has_gt, ev.device)
        pseudo = self.teacher_targets(ev, worker_id, is_first_sample)
        targets = torch.where(has_gt[:, None, None], gt, pseudo)
        # student: dense-label step (modules/detection.py:150-298 with a label on every frame)
        mode = Mode.TRAIN
        st.mode_2_rnn_states[mode].reset(worker_id=worker_id, indices_or_bool_tensor=is_first_sample)
        prev = st.mode_2_rnn_states[mode].get_states(worker_id=worker_id)
        feats_all, states = st.mdl.backbone.forward_sequence(ev, prev)
        st.mode_2_rnn_states[mode].save_states_and_detach(worker_id=worker_id, states=states)
        sel = {k: v.reshape(L * B, *v.shape[2:]) for k, v in feats_all.items() if k in st.mdl.fpn.in_features}
        predictions, losses = st.mdl.forward_detect(backbone_features=sel, targets=targets)
        return {'loss': losses['loss'], 'log_dict': {f'train/{k}': v for k, v in losses.items()}, 'predictions': predictions,
                'targets': targets}

    @torch.no_grad()
    def after_optimizer_step(self):
        """The EMA has been written into the teacher's flat parameters by the optimizer launch."""
        t, s = self.teacher, self.student.mdl
        if self.copy_bn_buffers:
            t.detect_engine.flat_buffers.copy_(s.detect_engine.flat_buffers)
        t.backbone.mark_params_updated()
        t.detect_engine.mark_params_updated()


class _SelfTrainOptimizer:
    """FlatOptimizer whose step also refreshes the teacher."""

    def __init__(self, module: SelfTrainingModule, opt: FlatOptimizer):
        self.module, self.opt = module, opt
        self.bufs = opt.bufs

    def zero_grad(self):
        self.opt.zero_grad()

    def step(self, lr=None):
        self.opt.step(lr)
        self.module.after_optimizer_step()
