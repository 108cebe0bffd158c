"""Teacher pseudo-label sweep.  Mirror of modules/pseudo_labeler.py:410-796 (class PseudoLabeler): the predict step
over unlabelled chunks — hflip-TTA batch doubling (`get_data_from_batch` :458-495), the frames-to-predict masks
(`_get_pred_mask` :514-547), head + NMS on the selected frames (`_predict_bbox` :565-589), threshold / box filters
(`pred2label`, modules/utils/ssod.py:147-188) and the re-assembly into `B x L` label lists (`_predict_step_impl`
:622-770) — plus `tta_postprocess` (:37-91), the per-frame NMS over merged TTA views.

Differences by design (results identical):
  * the time loop runs inside the library (`forward_sequence`, one call per chunk) and the uint8 -> float cast and
    padding are fused into the stem, so `get_data_from_batch` keeps the stacked uint8 tensor;
  * confidence filter + NMS (`leod_postprocess`) and the pseudo-label filters (`leod_pred2label`) run batched over
    all selected frames with no per-image host loop; `tta_postprocess` batches its frames the same way.
Out of scope here: the numpy tracker and the on-disk writer (`EventSeqData`), the label-quality
metrics and the Prophesee evaluator; `predict_step` hands its per-sequence results to an `on_sequence_labels`
callback instead.
"""
import copy
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np
import torch
import torch as th

from leod_b200.data.labels import ObjectLabels, SparselyBatchedObjectLabels
from leod_b200.data.utils.types import DataType, dget
from leod_b200.models.detection.yolox.utils.boxes import postprocess_packed
from .detection import Module
from .utils.detection import DATA_KEY, WORKER_ID_KEY, Mode, SeqLens
from .utils.ssod import frame_hw, pred2label_packed, tta_merge_packed


def tta_postprocess(preds: List[ObjectLabels], conf_thre: float = 0.7, nms_thre: float = 0.45,
                    class_agnostic: bool = False) -> List[ObjectLabels]:
    """pseudo_labeler.py:37-91: NMS over the (concatenated TTA views of the) boxes of every frame; frames that carry
    ground truth pass through, empty frames give empty labels.  One kernel launch for all frames."""
    if len(preds) == 0:
        return preds
    dev = None
    for p in preds:
        if p.object_labels.is_cuda:
            dev = p.object_labels.device
            break
    if dev is None:
        raise RuntimeError('tta_postprocess runs on CUDA label tensors only (no CPU fallback)')
    nmax = max(1, max(len(p) for p in preds))
    packed = torch.zeros(len(preds), nmax, 8, dtype=torch.float32, device=dev)
    cnt = torch.tensor([len(p) for p in preds], dtype=torch.int32, device=dev)
    for i, p in enumerate(preds):
        if len(p):
            packed[i, :len(p)] = p.object_labels.to(device=dev, dtype=torch.float32)
    merged, n = tta_merge_packed(packed, cnt, conf_thre, nms_thre, class_agnostic)
    return [ObjectLabels(merged[i, :k], preds[i].input_size_hw) for i, k in enumerate(n.tolist())]


class PseudoLabeler(Module):
    """Generate pseudo labels on training data (pseudo_labeler.py:410)."""

    def __init__(self, full_config, ssod: bool = False, on_sequence_labels: Optional[Callable] = None):
        super().__init__(full_config, ssod=ssod)
        self.mode_2_seq_lens = SeqLens()
        self.dst_name = self.dst_config.name
        self.ds_by2 = bool(self.dst_config.get('downsample_by_factor_2', False))
        assert self.dst_name in ('gen1', 'gen4'), f'Unknown dataset {self.dst_name}'
        self.use_gt = full_config.get('use_gt', True)
        self.tta_cfg = full_config.get('tta', None)
        self.pl_cfg = self.mdl_config.pseudo_label
        self.on_sequence_labels = on_sequence_labels
        self.mode_2_batch_size: Dict[Mode, Optional[int]] = {m: None for m in Mode}

    # ------------------------------------------------------------------ data
    def _tta(self, key: str) -> bool:
        return bool(self.tta_cfg is not None and self.tta_cfg.get('enable', False) and self.tta_cfg.get(key, False))

    def get_data_from_batch(self, batch: Any):
        """pseudo_labeler.py:458-495.  Returns a dict keyed by DataType NAME -> value (the reference's enum or ours)."""
        src = batch[DATA_KEY]
        data = {getattr(k, 'name', k): v for k, v in src.items()}
        assert 'AUGM_STATE' not in data, 'should not apply data augmentation in testing'
        ev = data['EV_REPR']
        ev = ev if th.is_tensor(ev) else th.stack(list(ev))           # [L, B, C, H, W], dtype untouched (uint8 stays uint8)
        B = ev.shape[1]
        data['is_hflip'] = np.array([False] * B, dtype=bool)
        if self._tta('hflip'):
            ev = th.cat([ev, th.flip(ev, dims=[-1])], dim=1)          # 2B
            for k in ('IS_FIRST_SAMPLE', 'IS_LAST_SAMPLE', 'IS_REVERSED'):
                if k in data:
                    data[k] = th.cat([data[k]] * 2, dim=-1)
            for k in ('EV_IDX', 'IS_PADDED_MASK'):
                if k in data:
                    data[k] = [th.cat([d] * 2, dim=-1) for d in data[k]]
            if 'PATH' in data:
                data['PATH'] = list(data['PATH']) * 2
            for k in ('OBJLABELS_SEQ', 'SKIPPED_OBJLABELS_SEQ'):
                if k in data:
                    labels = list(data[k])
                    flipped = [copy.deepcopy(l) for l in labels]
                    for i, (lbl, lbl_flip) in enumerate(zip(labels, flipped)):
                        lbl_flip.flip_lr_()
                        labels[i] = lbl + lbl_flip
                    data[k] = labels
            data['is_hflip'] = np.array([False] * B + [True] * B, dtype=bool)
        data['EV_REPR'] = ev
        return data

    def collect_data(self, data):
        """pseudo_labeler.py:497-512."""
        ev_paths = data.get('PATH', [''] * len(data['is_hflip']))
        L = data['EV_REPR'].shape[0]
        B = data['EV_REPR'].shape[1]
        ev_idx = th.stack(list(data['EV_IDX'])).transpose(1, 0).cpu().numpy().tolist() if 'EV_IDX' in data else \
            [[-1] * L for _ in range(B)]
        first = data['IS_FIRST_SAMPLE'].cpu().numpy().tolist()
        last = data['IS_LAST_SAMPLE'].cpu().numpy().tolist() if 'IS_LAST_SAMPLE' in data else [False] * B
        padding = th.stack(list(data['IS_PADDED_MASK'])).transpose(1, 0).cpu().numpy().tolist() if 'IS_PADDED_MASK' in data \
            else [[False] * L for _ in range(B)]
        is_tflip = data['IS_REVERSED'].cpu().numpy().tolist() if 'IS_REVERSED' in data else [False] * B
        return ev_paths, ev_idx, first, last, padding, data['is_hflip'], is_tflip

    def _get_pred_mask(self, worker_id: int, data: Dict) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """pseudo_labeler.py:514-547: [L,B] masks (predict here, has GT, has skipped GT)."""
        obj_labels = data['OBJLABELS_SEQ']
        skipped = data.get('SKIPPED_OBJLABELS_SEQ', None)
        L, B = len(obj_labels), len(obj_labels[0])
        skip_mask = np.zeros((L, B), dtype=bool)
        gt_mask = np.zeros((L, B), dtype=bool)
        skipped_gt_mask = np.zeros((L, B), dtype=bool)
        skip_len = max(int(self.pl_cfg.get('skip_first_t', 0)), 1)
        prev_lens = self.mode_2_seq_lens.get_lens(worker_id=worker_id)
        for b in range(B):
            seen = int(prev_lens[b]) if prev_lens is not None else 0
            if seen < skip_len:
                skip_mask[:skip_len - seen, b] = True
        for t in range(L):
            for b in range(B):
                has_gt = (obj_labels[t][b] is not None) and self.use_gt
                has_skipped = skipped is not None and skipped[t][b] is not None
                assert not (has_gt and has_skipped)
                gt_mask[t, b] = has_gt
                # the reference OVERWRITES the mask here (pseudo_labeler.py:538), which voids step 1: the first frame(s) of a
                # new sequence are predicted too.  Reproduced as is — the label sets must be identical.
                skip_mask[t, b] = has_gt
                skipped_gt_mask[t, b] = has_skipped
        if 'IS_PADDED_MASK' in data:
            padded = th.stack(list(data['IS_PADDED_MASK'])).cpu().numpy().astype(bool)
            skip_mask[padded] = True
        return ~skip_mask, gt_mask, skipped_gt_mask

    # ------------------------------------------------------------------ head + filters, batched
    def _predict_bbox(self, feats: Optional[Dict[int, th.Tensor]]) -> Optional[Tuple[th.Tensor, th.Tensor]]:
        """pseudo_labeler.py:565-589: forward_detect + postprocess on the selected frames.  Returns the packed result
        (dets [B',max_det,7], count [B']) instead of a Python list (same rows, same order)."""
        if feats is None:
            return None
        predictions, _ = self.mdl.forward_detect(backbone_features=feats)
        pp = self.mdl_config.postprocess
        return postprocess_packed(predictions, self.num_classes, pp.confidence_threshold, pp.nms_threshold)

    @torch.inference_mode()
    def _predict_step_impl(self, batch: Any, mode: Mode = Mode.TEST):
        """pseudo_labeler.py:622-770 (prediction branch)."""
        data = self.get_data_from_batch(batch)
        worker_id = batch[WORKER_ID_KEY]
        ev = data['EV_REPR']
        obj_labels = data['OBJLABELS_SEQ']
        skipped_obj_labels = data.get('SKIPPED_OBJLABELS_SEQ', None)
        is_first_sample = data['IS_FIRST_SAMPLE']
        L, B = len(obj_labels), len(obj_labels[0])
        assert L > 0 and B > 0 and ev.shape[0] == L and ev.shape[1] == B
        if self.mode_2_batch_size[mode] is None:
            self.mode_2_batch_size[mode] = B
        else:
            assert self.mode_2_batch_size[mode] == B
        self.mode_2_rnn_states[mode].reset(worker_id=worker_id, indices_or_bool_tensor=is_first_sample)
        prev_states = self.mode_2_rnn_states[mode].get_states(worker_id=worker_id)
        self.mode_2_seq_lens.reset(worker_id=worker_id, indices_or_bool_tensor=is_first_sample)
        pse_mask, gt_mask, skipped_gt_mask = self._get_pred_mask(worker_id=worker_id, data=data)
        if not self.use_gt:
            assert gt_mask.sum() == 0, 'should not use GT labels'

        # the time loop of :676-704 in one library call; features of every timestep come back
        feats_all, states = self.mdl.backbone.forward_sequence(ev, prev_states)
        self.mode_2_rnn_states[mode].save_states_and_detach(worker_id=worker_id, states=states)
        self.mode_2_seq_lens.update_lens(worker_id=worker_id, lens=torch.ones(B).long() * L)

        gt_obj_labels, skipped_gt_obj_labels = [], []
        for t in range(L):
            if self.use_gt:
                gt_obj_labels.extend(obj_labels[t].get_valid_labels_and_batch_indices()[0])
            if skipped_obj_labels is not None:
                skipped_gt_obj_labels.extend(skipped_obj_labels[t].get_valid_labels_and_batch_indices()[0])
        t_idx, b_idx = np.nonzero(pse_mask)            # row-major: time-major then batch, as the reference gathers them
        pse_labels: List[ObjectLabels] = []
        if len(t_idx) > 0:
            ti = th.as_tensor(t_idx, device=ev.device)
            bi = th.as_tensor(b_idx, device=ev.device)
            sel = {k: v[ti, bi] for k, v in feats_all.items() if k in self.mdl.fpn.in_features}
            dets, cnt = self._predict_bbox(sel)
            hw = frame_hw(self.dst_name, self.ds_by2)
            labels, n = pred2label_packed(dets, cnt, self.pl_cfg.obj_thresh, self.pl_cfg.cls_thresh, hw)
            pse_labels = [ObjectLabels(labels[i, :k], hw) for i, k in enumerate(n.tolist())]

        all_labels = [[None] * L for _ in range(B)]
        skipped_gt_pse_labels = []
        gt_cnt = pse_cnt = 0
        for t in range(L):
            for b in range(B):
                is_pse, is_gt, is_skipped = pse_mask[t, b], gt_mask[t, b], skipped_gt_mask[t, b]
                if is_skipped:
                    assert is_pse, 'should predict on skipped GT frames'
                    skipped_gt_pse_labels.append(pse_labels[pse_cnt])
                assert not (is_pse and is_gt), 'do not predict on GT frames'
                if is_pse:
                    all_labels[b][t] = pse_labels[pse_cnt]
                    pse_cnt += 1
                elif is_gt:
                    all_labels[b][t] = gt_obj_labels[gt_cnt]
                    gt_cnt += 1
        assert pse_cnt == pse_mask.sum() and gt_cnt == gt_mask.sum() and len(skipped_gt_pse_labels) == skipped_gt_mask.sum()
        self.last_eval_pairs = (skipped_gt_obj_labels, skipped_gt_pse_labels)   # for the (out-of-scope) quality metrics
        return (all_labels,) + tuple(self.collect_data(data))

    def predict_step(self, batch: Any, batch_idx: int = 0):
        """pseudo_labeler.py:772-796 without the EventSeqData accumulation: every (sequence, view) row of the batch is
        handed to `on_sequence_labels(labels=[L], ev_path, ev_idx, is_first, is_last, padded, is_hflip, is_tflip)`."""
        out = self._predict_step_impl(batch=batch, mode=Mode.TEST)
        if self.on_sequence_labels is not None:
            for row in zip(*out):
                if not row[1]:
                    continue
                self.on_sequence_labels(*row)
        return out
