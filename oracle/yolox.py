"""Oracle: YOLOX PAFPN neck, decoupled head, decode and SimOTA loss, functional over a state_dict.

Restates (paths relative to /root/reference):
  models/detection/yolox/models/network_blocks.py:29-54, 79-101, 104-142 (Conv-BN-SiLU, Bottleneck, CSP)
  models/detection/yolox_extension/models/yolo_pafpn.py:109-140         (top-down + bottom-up PAN)
  models/detection/yolox/models/yolo_head.py:195-287, 289-332            (head forward, grids, decode)
  models/detection/yolox/models/yolo_head.py:382-401                     (_ignore_bbox)
  models/detection/yolox/models/yolo_head.py:403-597, 776-972            (losses, with/without ignore)
  models/detection/yolox/models/yolo_head.py:606-774, 974-1148           (SimOTA assignment, geometry)
  models/detection/yolox/models/losses.py:18-66                          (IoU loss)
  models/detection/yolox/utils/boxes.py:89-113                           (pairwise IoU)
Options that every shipped config leaves off (obj_focal_loss, bbox_loss_weighting, ignore_bg_k,
use_l1, depthwise) are not restated.
TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from .config import ModelCfg


# ----------------------------------------------------------------------------- blocks
def conv_bn_silu(x, sd, prefix, stride=1, training=False, bn_state=None):
    """network_blocks.py:29-54.  training=True uses batch statistics; running stats are updated in
    `bn_state` (dict name -> tensor) when given, never in `sd`."""
    w = sd[prefix + '.conv.weight']
    y = F.conv2d(x, w, None, stride=stride, padding=(w.shape[-1] - 1) // 2)
    rm, rv = sd[prefix + '.bn.running_mean'], sd[prefix + '.bn.running_var']
    if training:
        rm, rv = rm.clone(), rv.clone()
    y = F.batch_norm(y, rm, rv, sd[prefix + '.bn.weight'], sd[prefix + '.bn.bias'],
                     training=training, momentum=0.1, eps=1e-5)
    if training and bn_state is not None:
        bn_state[prefix + '.bn.running_mean'] = rm
        bn_state[prefix + '.bn.running_var'] = rv
    return F.silu(y)


def csp_layer(x, sd, prefix, n, **kw):
    """network_blocks.py:104-142 with shortcut=False (as the PAFPN builds it)."""
    a = conv_bn_silu(x, sd, prefix + '.conv1', **kw)
    b = conv_bn_silu(x, sd, prefix + '.conv2', **kw)
    for i in range(n):
        a = conv_bn_silu(conv_bn_silu(a, sd, f'{prefix}.m.{i}.conv1', **kw), sd, f'{prefix}.m.{i}.conv2', **kw)
    return conv_bn_silu(torch.cat((a, b), 1), sd, prefix + '.conv3', **kw)


def _up2(x):
    return F.interpolate(x, scale_factor=2, mode='nearest-exact')


def pafpn_forward(feats: Dict[int, torch.Tensor], sd, cfg: ModelCfg, **kw):
    """yolo_pafpn.py:109-140."""
    n = round(3 * cfg.fpn_depth)
    x2, x1, x0 = (feats[s] for s in cfg.in_stages)
    fpn0 = conv_bn_silu(x0, sd, 'fpn.lateral_conv0', **kw)
    f0 = csp_layer(torch.cat((_up2(fpn0), x1), 1), sd, 'fpn.C3_p4', n, **kw)
    fpn1 = conv_bn_silu(f0, sd, 'fpn.reduce_conv1', **kw)
    pan2 = csp_layer(torch.cat((_up2(fpn1), x2), 1), sd, 'fpn.C3_p3', n, **kw)
    p1 = conv_bn_silu(pan2, sd, 'fpn.bu_conv2', stride=2, **kw)
    pan1 = csp_layer(torch.cat((p1, fpn1), 1), sd, 'fpn.C3_n3', n, **kw)
    p0 = conv_bn_silu(pan1, sd, 'fpn.bu_conv1', stride=2, **kw)
    pan0 = csp_layer(torch.cat((p0, fpn0), 1), sd, 'fpn.C3_n4', n, **kw)
    return pan2, pan1, pan0


def head_raw(fpn_feats: Sequence[torch.Tensor], sd, **kw) -> List[torch.Tensor]:
    """yolo_head.py:209-222: per level [B, 4+1+C, h, w] = (reg, obj logit, cls logits)."""
    outs = []
    for k, x in enumerate(fpn_feats):
        p = 'yolox_head'
        x = conv_bn_silu(x, sd, f'{p}.stems.{k}', **kw)
        c = conv_bn_silu(conv_bn_silu(x, sd, f'{p}.cls_convs.{k}.0', **kw), sd, f'{p}.cls_convs.{k}.1', **kw)
        r = conv_bn_silu(conv_bn_silu(x, sd, f'{p}.reg_convs.{k}.0', **kw), sd, f'{p}.reg_convs.{k}.1', **kw)
        cls = F.conv2d(c, sd[f'{p}.cls_preds.{k}.weight'], sd[f'{p}.cls_preds.{k}.bias'])
        reg = F.conv2d(r, sd[f'{p}.reg_preds.{k}.weight'], sd[f'{p}.reg_preds.{k}.bias'])
        obj = F.conv2d(r, sd[f'{p}.obj_preds.{k}.weight'], sd[f'{p}.obj_preds.{k}.bias'])
        outs.append(torch.cat((reg, obj, cls), 1))
    return outs


def anchor_grid(hws: Sequence[Tuple[int, int]], strides: Sequence[int], dtype=torch.float32):
    """Anchor order = level-major, row-major (y, x) (yolo_head.py:297-299, 318-325).
    Returns x_shift, y_shift, stride, each [A]."""
    xs, ys, ss = [], [], []
    for (h, w), s in zip(hws, strides):
        yv, xv = torch.meshgrid(torch.arange(h, dtype=dtype), torch.arange(w, dtype=dtype), indexing='ij')
        xs.append(xv.reshape(-1))
        ys.append(yv.reshape(-1))
        ss.append(torch.full((h * w,), float(s), dtype=dtype))
    return torch.cat(xs), torch.cat(ys), torch.cat(ss)


def flatten_decode(raw: Sequence[torch.Tensor], strides: Sequence[int], sigmoid_scores: bool):
    """Flatten levels to [B, A, 5+C] and decode boxes: xy = (xy + grid) * s, wh = exp(wh) * s
    (yolo_head.py:303-308 for the training branch — logits kept; :249-251, 310-332 for inference)."""
    hws = [tuple(r.shape[-2:]) for r in raw]
    gx, gy, gs = anchor_grid(hws, strides, raw[0].dtype)
    flat = torch.cat([r.flatten(2) for r in raw], 2).permute(0, 2, 1)  # [B, A, 5+C]
    xy = (flat[..., 0:2] + torch.stack((gx, gy), -1)) * gs[:, None]
    wh = torch.exp(flat[..., 2:4]) * gs[:, None]
    rest = flat[..., 4:].sigmoid() if sigmoid_scores else flat[..., 4:]
    return torch.cat((xy, wh, rest), -1), (gx, gy, gs)


# ----------------------------------------------------------------------------- assignment + loss
def pairwise_iou_cxcywh(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """boxes.py:89-113 with xyxy=False: [M,4] x [N,4] -> [M,N]."""
    tl = torch.max(a[:, None, :2] - a[:, None, 2:] / 2, b[:, :2] - b[:, 2:] / 2)
    br = torch.min(a[:, None, :2] + a[:, None, 2:] / 2, b[:, :2] + b[:, 2:] / 2)
    area_a = a[:, 2] * a[:, 3]
    area_b = b[:, 2] * b[:, 3]
    en = (tl < br).to(tl.dtype).prod(2)
    inter = (br - tl).prod(2) * en
    return inter / (area_a[:, None] + area_b - inter)


def apply_ignore_thresh(labels: torch.Tensor, cfg: ModelCfg) -> torch.Tensor:
    """yolo_head.py:382-401: class id -> ignore_label where obj/cls confidence < per-class threshold."""
    if not cfg.ignore_bbox_thresh:
        return labels
    labels = labels.clone()
    cls, obj_c, cls_c = labels[..., 0], labels[..., 5], labels[..., 6]
    ign = torch.zeros_like(cls, dtype=torch.bool)
    for idx, th in enumerate(cfg.ignore_bbox_thresh):
        ign |= (cls == idx) & ((obj_c < th) | (cls_c < th))
    ign &= labels.sum(2) > 0
    labels[..., 0] = torch.where(ign, torch.full_like(cls, float(cfg.ignore_label)), cls)
    return labels


@torch.no_grad()
def simota_assign(rows: torch.Tensor, pred: torch.Tensor, grid, cfg: ModelCfg):
    """One image.  rows: [n,7] non-padded labels (cls, cx, cy, w, h, obj_conf, cls_conf);
    pred: [A, 5+C] decoded boxes + raw logits.  Returns (fg_mask[A], matched_row[A] (index into
    `rows`, -1 if bg), matched_iou[A], ignore_mask[A]).
    yolo_head.py:606-774 (plain) and :974-1148 (ignore-aware); both collapse to: candidates =
    anchors inside the 1.5-stride centre box of some VALID gt; ignore = inside only ignored gts."""
    gx, gy, gs = grid
    A = pred.shape[0]
    valid = rows[:, 0] != cfg.ignore_label
    cx, cy = (gx + 0.5) * gs, (gy + 0.5) * gs
    r = gs * 1.5
    deltas = torch.stack((cx - (rows[:, 1:2] - r), cy - (rows[:, 2:3] - r),
                          (rows[:, 1:2] + r) - cx, (rows[:, 2:3] + r) - cy), 2)
    inside = deltas.min(-1).values > 0.0                       # [n, A]
    any_all = inside.any(0)
    any_valid = inside[valid].any(0) if valid.any() else torch.zeros(A, dtype=torch.bool)
    ignore = any_all & ~any_valid
    matched = torch.full((A,), -1, dtype=torch.long)
    miou = torch.zeros(A)
    if not valid.any() or not any_valid.any():
        return torch.zeros(A, dtype=torch.bool), matched, miou, ignore
    vidx = valid.nonzero()[:, 0]
    cand = any_valid.nonzero()[:, 0]
    gt = rows[vidx]
    geom = inside[vidx][:, cand]
    ious = pairwise_iou_cxcywh(gt[:, 1:5].float(), pred[cand, :4].float())
    onehot = F.one_hot(gt[:, 0].long(), cfg.num_classes).float()
    p = (pred[cand, 5:].float().sigmoid() * pred[cand, 4:5].float().sigmoid()).sqrt()
    cls_cost = F.binary_cross_entropy(p[None].expand(len(gt), -1, -1), onehot[:, None].expand(-1, len(cand), -1),
                                      reduction='none').sum(-1)
    cost = cls_cost + 3.0 * (-torch.log(ious + 1e-8)) + 1e6 * (~geom)
    kk = min(10, len(cand))
    dyn_k = torch.clamp(torch.topk(ious, kk, dim=1).values.sum(1).int(), min=1)
    match = torch.zeros_like(cost, dtype=torch.bool)
    for g in range(len(gt)):
        # k smallest costs; ties resolved towards the lower anchor index (stable sort)
        order = torch.sort(cost[g], stable=True).indices[:int(dyn_k[g])]
        match[g, order] = True
    multi = match.sum(0) > 1
    if multi.any():
        best = cost[:, multi].argmin(0)
        match[:, multi] = False
        match[best, multi.nonzero()[:, 0]] = True
    fg_c = match.any(0)
    which = match.float().argmax(0)
    fg = torch.zeros(A, dtype=torch.bool)
    fg[cand[fg_c]] = True
    matched[cand[fg_c]] = vidx[which[fg_c]]
    miou[cand[fg_c]] = (match * ious).sum(0)[fg_c]
    return fg, matched, miou, ignore


def iou_loss_elem(pred: torch.Tensor, tgt: torch.Tensor) -> torch.Tensor:
    """losses.py:18-44, loss_type 'iou': 1 - iou^2 per row, boxes cxcywh."""
    tl = torch.max(pred[:, :2] - pred[:, 2:] / 2, tgt[:, :2] - tgt[:, 2:] / 2)
    br = torch.min(pred[:, :2] + pred[:, 2:] / 2, tgt[:, :2] + tgt[:, 2:] / 2)
    en = (tl < br).to(tl.dtype).prod(1)
    inter = (br - tl).prod(1) * en
    union = pred[:, 2] * pred[:, 3] + tgt[:, 2] * tgt[:, 3] - inter
    return 1 - (inter / (union + 1e-16)) ** 2


def yolox_losses(train_out: torch.Tensor, grid, labels: torch.Tensor, cfg: ModelCfg,
                 forced_assign: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """yolo_head.py:403-597 / 776-972.  train_out [B,A,5+C]: decoded boxes, obj/cls LOGITS.
    labels [B,N,7] zero-padded at the end of dim 1.
    forced_assign (test hook, not in the reference): [B,A] label-row index per anchor (-1 = background) that replaces the
    SimOTA matching, so a reduced-precision implementation can be compared through the smooth part of the loss with its
    own (possibly flipped) discrete assignment."""
    labels = apply_ignore_thresh(labels, cfg)
    B, A, _ = train_out.shape
    C = cfg.num_classes
    fg_all = torch.zeros(B, A, dtype=torch.bool)
    ign_all = torch.zeros(B, A, dtype=torch.bool)
    matched_all = torch.full((B, A), -1, dtype=torch.long)
    reg_t = torch.zeros(B, A, 4)
    cls_t = torch.zeros(B, A, C)
    num_gts = 0
    for b in range(B):
        n_all = int((labels[b].sum(1) > 0).sum())
        rows = labels[b, :n_all]
        num_gts += int((rows[:, 0] != cfg.ignore_label).sum())
        if n_all == 0:
            continue
        fg, matched, miou, ign = simota_assign(rows, train_out[b].detach(), grid, cfg)
        if forced_assign is not None:
            matched = forced_assign[b].long()
            fg = matched >= 0
            miou = torch.zeros(A)
            if fg.any():
                pi = pairwise_iou_cxcywh(rows[matched[fg], 1:5].float(), train_out[b, fg, :4].detach().float())
                miou[fg] = pi.diagonal()
        fg_all[b], ign_all[b] = fg, ign
        matched_all[b] = torch.where(fg, matched, torch.full_like(matched, -1))
        if fg.any():
            m = matched[fg]
            reg_t[b, fg] = rows[m, 1:5].float()
            cls_t[b, fg] = F.one_hot(rows[m, 0].long(), C).float() * miou[fg][:, None]
    n_fg = int(fg_all.sum())
    num_fg = max(n_fg, 1)
    box, obj, cls = train_out[..., :4], train_out[..., 4], train_out[..., 5:]
    loss_iou = iou_loss_elem(box[fg_all], reg_t[fg_all]).mean() if n_fg > 0 else train_out.new_zeros(())
    valid = ~ign_all
    loss_obj = F.binary_cross_entropy_with_logits(obj[valid], fg_all[valid].to(obj.dtype), reduction='sum') / num_fg
    loss_cls = F.binary_cross_entropy_with_logits(cls[fg_all], cls_t[fg_all], reduction='sum') / num_fg
    loss_iou = cfg.reg_weight * loss_iou
    loss_obj = cfg.obj_weight * loss_obj
    loss_cls = cfg.cls_weight * loss_cls
    return {'loss': loss_iou + loss_obj + loss_cls, 'iou_loss': loss_iou, 'conf_loss': loss_obj,
            'cls_loss': loss_cls, 'l1_loss': 0.0, 'num_fg': num_fg / max(num_gts, 1),
            '_fg_mask': fg_all, '_ignore_mask': ign_all, '_matched': matched_all}


def detect_forward(feats: Dict[int, torch.Tensor], sd, cfg: ModelCfg, targets: Optional[torch.Tensor] = None,
                   training: bool = False, bn_state=None, forced_assign: Optional[torch.Tensor] = None):
    """detector.py:55-77 + yolo_head.py:195-287.  Returns (decoded predictions [B,A,5+C] with
    sigmoid scores, losses or None)."""
    kw = dict(training=training, bn_state=bn_state)
    strides = tuple(cfg.strides[s - 1] for s in cfg.in_stages)
    raw = head_raw(pafpn_forward(feats, sd, cfg, **kw), sd, **kw)
    losses = None
    if training:
        assert targets is not None
        train_out, grid = flatten_decode(raw, strides, sigmoid_scores=False)
        losses = yolox_losses(train_out, grid, targets, cfg, forced_assign=forced_assign)
    preds, _ = flatten_decode(raw, strides, sigmoid_scores=True)
    return preds, losses
