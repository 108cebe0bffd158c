"""Mirror of models/detection/yolox/utils/boxes.py:32-113 (postprocess, bboxes_iou).

`postprocess` keeps the reference's signature and return value but runs as ONE kernel launch for the
whole batch (leod_postprocess: confidence filter, stable score order, class-aware NMS that reproduces
torchvision.ops.batched_nms) instead of a Python loop with a torchvision call per image.
"""
from typing import List, Optional

import torch

from leod_b200 import _lib


def postprocess_packed(prediction: torch.Tensor, num_classes: int, conf_thre: float = 0.7, nms_thre: float = 0.45,
                       class_agnostic: bool = False, max_det: Optional[int] = None):
    """Device-resident result: (dets [B, max_det, 7] fp32, count [B] int32), no host synchronisation."""
    if not prediction.is_cuda:
        raise RuntimeError('leod_b200 postprocess runs on CUDA tensors only (no CPU fallback)')
    B, A, D = prediction.shape
    assert D == 5 + num_classes, (D, num_classes)
    pred = prediction.detach()
    if pred.dtype != torch.float32 or not pred.is_contiguous():
        pred = pred.float().contiguous()
    max_det = A if max_det is None else min(int(max_det), A)
    dets = torch.empty(B, max_det, 7, dtype=torch.float32, device=pred.device)
    count = torch.empty(B, dtype=torch.int32, device=pred.device)
    with torch.cuda.device(pred.device):
        _lib.check(_lib.lib().leod_postprocess(_lib.ptr(pred), B, A, num_classes, float(conf_thre), float(nms_thre),
                                               int(bool(class_agnostic)), _lib.ptr(dets), _lib.ptr(count), max_det,
                                               _lib.stream_ptr(pred.device)), 'postprocess')
    return dets, count


def postprocess(prediction, num_classes, conf_thre=0.7, nms_thre=0.45, class_agnostic=False, pad=None) -> List[Optional[torch.Tensor]]:
    """boxes.py:32-86.  Returns a `B`-len list of [N_i,7] = (x1,y1,x2,y2,obj_conf,cls_conf,cls_idx)
    sorted by descending score, `pad` where nothing survives.  Like the reference it rewrites
    prediction[..., :4] in place from (cx,cy,w,h) to corners."""
    dets, count = postprocess_packed(prediction, num_classes, conf_thre, nms_thre, class_agnostic)
    with torch.no_grad():
        half = prediction[..., 2:4] / 2
        xy = prediction[..., 0:2].clone()
        prediction[..., 0:2] = xy - half
        prediction[..., 2:4] = xy + half
    counts = count.tolist()  # the one device->host read of the step
    out = []
    for b, n in enumerate(counts):
        out.append(dets[b, :n].to(prediction.dtype) if n > 0 else pad)
    return out


def bboxes_iou(bboxes_a, bboxes_b, xyxy=True):
    """boxes.py:89-113: pairwise IoU [M,4] x [N,4] -> [M,N]."""
    if bboxes_a.shape[1] != 4 or bboxes_b.shape[1] != 4:
        raise IndexError
    if xyxy:
        tl = torch.max(bboxes_a[:, None, :2], bboxes_b[:, :2])
        br = torch.min(bboxes_a[:, None, 2:], bboxes_b[:, 2:])
        area_a = torch.prod(bboxes_a[:, 2:] - bboxes_a[:, :2], 1)
        area_b = torch.prod(bboxes_b[:, 2:] - bboxes_b[:, :2], 1)
    else:
        tl = torch.max(bboxes_a[:, None, :2] - bboxes_a[:, None, 2:] / 2, bboxes_b[:, :2] - bboxes_b[:, 2:] / 2)
        br = torch.min(bboxes_a[:, None, :2] + bboxes_a[:, None, 2:] / 2, bboxes_b[:, :2] + bboxes_b[:, 2:] / 2)
        area_a = torch.prod(bboxes_a[:, 2:], 1)
        area_b = torch.prod(bboxes_b[:, 2:], 1)
    en = (tl < br).type(tl.type()).prod(dim=2)
    area_i = torch.prod(br - tl, 2) * en
    return area_i / (area_a[:, None] + area_b - area_i)
