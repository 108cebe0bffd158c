"""Pseudo-label filters, EMA and the fused optimizer step.

Mirror of modules/utils/ssod.py: `pred2label` (:147-188) with `filter_pred_boxes` (:113-133) and
`filter_w_thresh` (:136-144) folded into one kernel over the packed NMS output; `ema_model_update`
(:429-438) and `model_update` (:441-460); plus `fused_adamw_ema`, the single-launch replacement of
torch.optim.AdamW + clip_grad_value_ (modules/detection.py:485-518, train.py:236-237)."""
import ctypes
from typing import List, Optional, Sequence, Tuple, Union

import torch

from leod_b200 import _lib

DATASET2HEIGHT = {'gen1': 240, 'gen4': 720}
DATASET2WIDTH = {'gen1': 304, 'gen4': 1280}


def frame_hw(dataset_name: str = 'gen1', downsampled_by_2: bool = False) -> Tuple[int, int]:
    h, w = DATASET2HEIGHT[dataset_name], DATASET2WIDTH[dataset_name]
    return (h // 2, w // 2) if downsampled_by_2 else (h, w)


def _thr_list(t: Union[float, Sequence[float]], n: int):
    return [float(t)] * n if isinstance(t, (int, float)) else [float(v) for v in t]


def pred2label_packed(dets: torch.Tensor, count: torch.Tensor, obj_thresh, cls_thresh, hw: Optional[Tuple[int, int]]):
    """dets [B,max_det,7] fp32 + count [B] int32 (from postprocess_packed) -> (labels [B,max_det,8], n [B]).
    Rows: (t=0, x, y, w, h (corner format), cls_idx, cls_conf, obj_conf) — the ObjectLabels layout."""
    B, md, _ = dets.shape
    ncls = max(len(obj_thresh) if not isinstance(obj_thresh, (int, float)) else 1,
               len(cls_thresh) if not isinstance(cls_thresh, (int, float)) else 1)
    if isinstance(obj_thresh, (int, float)) and isinstance(cls_thresh, (int, float)):
        ncls = 16
    o = (ctypes.c_float * ncls)(*_thr_list(obj_thresh, ncls))
    c = (ctypes.c_float * ncls)(*_thr_list(cls_thresh, ncls))
    labels = torch.empty(B, md, 8, dtype=torch.float32, device=dets.device)
    n = torch.empty(B, dtype=torch.int32, device=dets.device)
    fh, fw = hw if hw is not None else (-1, -1)
    with torch.cuda.device(dets.device):
        _lib.check(_lib.lib().leod_pred2label(_lib.ptr(dets.contiguous()), _lib.ptr(count), B, md, ncls, o, c, int(fh), int(fw),
                                              _lib.ptr(labels), _lib.ptr(n), _lib.stream_ptr(dets.device)), 'pred2label')
    return labels, n


def pred2label(pred: List[Optional[torch.Tensor]], obj_thresh=0.9, cls_thresh=0.9, frame_hw_: Optional[Tuple[int, int]] = None):
    """List API of ssod.py:147-188: `B`-len list of [N_i,7] -> `B`-len list of [M_i,8] label rows."""
    dev = next(p.device for p in pred if p is not None)
    md = max([1] + [p.shape[0] for p in pred if p is not None])
    dets = torch.zeros(len(pred), md, 7, dtype=torch.float32, device=dev)
    cnt = torch.zeros(len(pred), dtype=torch.int32, device=dev)
    for b, p in enumerate(pred):
        if p is not None and p.shape[0]:
            dets[b, :p.shape[0]] = p.float()
            cnt[b] = p.shape[0]
    labels, n = pred2label_packed(dets, cnt, obj_thresh, cls_thresh, frame_hw_)
    return [labels[b, :k] for b, k in enumerate(n.tolist())]


def tta_merge_packed(labels: torch.Tensor, count: torch.Tensor, conf_thre: float, nms_thre: float, class_agnostic: bool = False):
    """labels [F,nmax,8] fp32 ObjectLabels rows + count [F] int32 -> (merged [F,nmax,8], n [F]): the second NMS of
    modules/pseudo_labeler.py:37-91 for F frames in one launch (leod_tta_merge)."""
    F, nmax, _ = labels.shape
    out = torch.zeros_like(labels)
    n = torch.zeros(F, dtype=torch.int32, device=labels.device)
    with torch.cuda.device(labels.device):
        _lib.check(_lib.lib().leod_tta_merge(_lib.ptr(labels.contiguous()), _lib.ptr(count), F, nmax, float(conf_thre), float(nms_thre),
                                             int(class_agnostic), _lib.ptr(out), _lib.ptr(n), _lib.stream_ptr(labels.device)), 'tta_merge')
    return out, n


def ema_alpha_at(global_step: int, alpha: float = 0.999) -> float:
    """ssod.py:435: the true average until the exponential average is more correct."""
    return min(1. - 1. / (global_step + 1.), alpha)


@torch.no_grad()
def ema_model_update(params: Sequence[torch.Tensor], ema_params: Sequence[torch.Tensor], global_step: int, alpha: float = 0.999):
    """ssod.py:429-438 over parameter lists (parameters only; BN buffers are not averaged)."""
    a = ema_alpha_at(global_step, alpha)
    torch._foreach_mul_(list(ema_params), a)
    torch._foreach_add_(list(ema_params), list(params), alpha=1. - a)


@torch.no_grad()
def model_update(student_model: torch.nn.Module, teacher_model: torch.nn.Module, global_step: int, method: str = 'ema', alpha: float = 0.999):
    """ssod.py:441-460: 'ema' -> ema_model_update over the parameters; 'every-N' -> hard copy of the student's state_dict
    (parameters AND buffers) into the teacher whenever (global_step + 1) % N == 0."""
    method = method.lower()
    if method == 'ema':
        ema_model_update([p.data for p in student_model.parameters()], [p.data for p in teacher_model.parameters()], global_step, alpha)
    elif 'every-' in method:
        num_step = int(method.split('-')[-1])
        if (global_step + 1) % num_step == 0:
            teacher_model.load_state_dict(student_model.state_dict())
    else:
        raise NotImplementedError(f'Unknown model update method: {method}')
    for m in teacher_model.modules():       # the library keeps operand-typed weight copies: refresh them before the next forward
        if hasattr(m, 'mark_params_updated'):
            m.mark_params_updated()
    if hasattr(teacher_model, 'detect_engine'):
        teacher_model.detect_engine.mark_params_updated()


@torch.no_grad()
def fused_adamw_ema(p, g, m, v, step: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, clip_value=0.0,
                    ema: Optional[torch.Tensor] = None, ema_alpha: float = 0.999):
    """One launch over flat fp32 buffers: clip-by-value, AdamW, optional teacher EMA."""
    assert p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()
    for t in (g, m, v):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == p.numel()
    with torch.cuda.device(p.device):
        _lib.check(_lib.lib().leod_adamw_ema(_lib.ptr(p), _lib.ptr(g), _lib.ptr(m), _lib.ptr(v), _lib.ptr(ema), p.numel(), int(step),
                                             float(lr), float(beta1), float(beta2), float(eps), float(weight_decay),
                                             float(clip_value), float(ema_alpha), _lib.stream_ptr(p.device)), 'adamw_ema')
