"""YOLOX decoupled head with SimOTA loss.

Mirror of models/detection/yolox/models/yolo_head.py:21-182 (constructor, sub-module names and the
bias initialisation), :195-287 (forward), :310-332 (decode) and the loss of :403-597 / :776-1148
(plain and ignore-label variants).  The loss here is a fixed-shape, batched formulation of the same
assignment: no Python loop over images or ground-truth boxes, no `.item()` / `int()` host reads, no
`empty_cache()`, so a whole training step can be captured in a CUDA graph.  Selection ties (equal
costs) resolve towards the lower anchor index (the reference leaves them to torch.topk).
Options left off by every shipped config (obj_focal_loss, bbox_loss_weighting, ignore_bg_k, use_l1,
depthwise) raise NotImplementedError.
"""
import math
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .network_blocks import BaseConv


class YOLOXHead(nn.Module):
    def __init__(self, num_classes=80, strides=(8, 16, 32), in_channels=(256, 512, 1024), act='silu', depthwise=False,
                 compile_cfg: Optional[Dict] = None, obj_focal_loss=False, bbox_loss_weighting='', ignore_bg_k=-1,
                 reg_weight=5.0, obj_weight=1.0, cls_weight=1.0, ignore_bbox_thresh=None, ignore_label=1024):
        super().__init__()
        if depthwise or obj_focal_loss or bbox_loss_weighting or (ignore_bg_k is not None and ignore_bg_k > 0):
            raise NotImplementedError('depthwise / focal / bbox_loss_weighting / ignore_bg_k are off in every shipped config')
        if compile_cfg is not None and compile_cfg.get('enable', False):
            raise NotImplementedError('torch.compile is not part of the B200-native path')
        self.num_classes = num_classes
        self.decode_in_inference = True
        self.strides = tuple(strides)
        hidden = int(256 * in_channels[-1] / 1024)
        self.hidden_dim = hidden
        self.stems = nn.ModuleList()
        self.cls_convs = nn.ModuleList()
        self.reg_convs = nn.ModuleList()
        self.cls_preds = nn.ModuleList()
        self.reg_preds = nn.ModuleList()
        self.obj_preds = nn.ModuleList()
        for c in in_channels:
            self.stems.append(BaseConv(c, hidden, 1, 1, act))
            self.cls_convs.append(nn.Sequential(BaseConv(hidden, hidden, 3, 1, act), BaseConv(hidden, hidden, 3, 1, act)))
            self.reg_convs.append(nn.Sequential(BaseConv(hidden, hidden, 3, 1, act), BaseConv(hidden, hidden, 3, 1, act)))
            self.cls_preds.append(nn.Conv2d(hidden, num_classes, 1, 1, 0))
            self.reg_preds.append(nn.Conv2d(hidden, 4, 1, 1, 0))
            self.obj_preds.append(nn.Conv2d(hidden, 1, 1, 1, 0))
        self.reg_weight, self.obj_weight, self.cls_weight = reg_weight, obj_weight, cls_weight
        self.ignore_bbox_thresh = list(ignore_bbox_thresh) if ignore_bbox_thresh else None
        self.ignore_label = ignore_label
        self.use_l1 = False
        self._grid_cache = {}
        prior = -math.log((1 - 0.01) / 0.01)  # yolo_head.py:183-193
        for conv in list(self.cls_preds) + list(self.obj_preds):
            nn.init.constant_(conv.bias, prior)

    def _has_sync_bn(self):
        # SyncBatchNorm layers issue collectives: keep them in one stream so every rank captures the same order
        if not hasattr(self, '_sync_bn'):
            self._sync_bn = any(isinstance(m, nn.SyncBatchNorm) for m in self.modules())
        return self._sync_bn

    # ------------------------------------------------------------------ grids / decode
    def _grids(self, hws, device, dtype):
        key = (tuple(hws), str(device), dtype)
        g = self._grid_cache.get(key)
        if g is None:
            xs, ys, ss = [], [], []
            for (h, w), s in zip(hws, self.strides):
                yv, xv = torch.meshgrid(torch.arange(h, device=device, dtype=dtype),
                                        torch.arange(w, device=device, dtype=dtype), indexing='ij')
                xs.append(xv.reshape(-1))
                ys.append(yv.reshape(-1))
                ss.append(torch.full((h * w,), float(s), device=device, dtype=dtype))
            g = (torch.cat(xs), torch.cat(ys), torch.cat(ss))
            self._grid_cache[key] = g
        return g

    def forward(self, xin, labels=None, pred_probs=None):
        assert pred_probs is None
        capturing = xin[0].is_cuda and torch.cuda.is_current_stream_capturing() and not self._has_sync_bn()

        def level(k, x, side=None):
            x = self.stems[k](x)
            if side is None:
                cls_feat = self.cls_convs[k](x)
                reg_feat = self.reg_convs[k](x)
            else:       # the classification and regression towers of a level are independent too
                here = torch.cuda.current_stream()
                side.wait_stream(here)
                with torch.cuda.stream(side):
                    cls_feat = self.cls_convs[k](x)
                    cls_out = self.cls_preds[k](cls_feat)
                    cls_out.record_stream(here)
                reg_feat = self.reg_convs[k](x)
                reg_out, obj_out = self.reg_preds[k](reg_feat), self.obj_preds[k](reg_feat)
                here.wait_stream(side)
                return torch.cat((reg_out, obj_out, cls_out), 1)
            return torch.cat((self.reg_preds[k](reg_feat), self.obj_preds[k](reg_feat), self.cls_preds[k](cls_feat)), 1)

        if capturing:
            # CUDA-graph capture (detector.py, _GraphedDetect): the three pyramid levels (and the two towers of each) are
            # independent chains of small kernels; fork them onto side streams so the captured graph (forward and
            # backward) runs them side by side
            cur = torch.cuda.current_stream()
            n = len(xin)
            if not hasattr(self, '_side_streams'):
                self._side_streams = [torch.cuda.Stream() for _ in range(2 * n - 1)]
            raw = [None] * n
            for k in range(1, n):
                st = self._side_streams[k - 1]
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    raw[k] = level(k, xin[k], side=self._side_streams[n - 1 + k])
                    raw[k].record_stream(cur)
            raw[0] = level(0, xin[0], side=self._side_streams[n - 1])
            for st in self._side_streams[:n - 1]:
                cur.wait_stream(st)
        else:
            raw = [level(k, x) for k, x in enumerate(xin)]
        hws = [tuple(r.shape[-2:]) for r in raw]
        self.hw = hws
        flat = torch.cat([r.flatten(2) for r in raw], 2).permute(0, 2, 1).float()  # [B, A, 5+C], level-major anchors
        gx, gy, gs = self._grids(hws, flat.device, flat.dtype)
        xy = (flat[..., 0:2] + torch.stack((gx, gy), -1)) * gs[:, None]
        wh = torch.exp(flat[..., 2:4]) * gs[:, None]
        losses = None
        if self.training:
            assert labels is not None
            train_out = torch.cat((xy, wh, flat[..., 4:]), -1)
            losses = self.get_losses(train_out, (gx, gy, gs), labels.to(flat.dtype))
        outputs = torch.cat((xy, wh, flat[..., 4:].sigmoid()), -1)
        return outputs, losses

    # ------------------------------------------------------------------ loss
    @torch.no_grad()
    def _ignore_bbox(self, labels):
        """yolo_head.py:382-401."""
        if not self.ignore_bbox_thresh:
            return labels
        labels = labels.clone()
        cls, obj_c, cls_c = labels[..., 0], labels[..., 5], labels[..., 6]
        ign = torch.zeros_like(cls, dtype=torch.bool)
        for idx, th in enumerate(self.ignore_bbox_thresh):
            ign |= (cls == idx) & ((obj_c < th) | (cls_c < th))
        ign &= labels.sum(2) > 0
        labels[..., 0] = torch.where(ign, torch.full_like(cls, float(self.ignore_label)), cls)
        return labels

    @torch.no_grad()
    def assign(self, out, grid, labels):
        """Batched SimOTA (yolo_head.py:606-774, 974-1148).  out [B,A,5+C] (decoded boxes, logits),
        labels [B,N,7].  Returns fg [B,A] bool, ignore [B,A] bool, reg_target [B,A,4], cls_target [B,A,C],
        num_gts (0-dim)."""
        gx, gy, gs = grid
        B, A, _ = out.shape
        C = self.num_classes
        nonzero = labels.sum(2) > 0                                    # [B,N]
        valid = nonzero & (labels[..., 0] != self.ignore_label)
        cx, cy, r = (gx + 0.5) * gs, (gy + 0.5) * gs, gs * 1.5
        lx, ly = labels[..., 1:2], labels[..., 2:3]                    # [B,N,1]
        deltas = torch.stack((cx - (lx - r), cy - (ly - r), (lx + r) - cx, (ly + r) - cy), -1)
        inside = (deltas.min(-1).values > 0.0) & nonzero[..., None]    # [B,N,A]
        inside_v = inside & valid[..., None]
        cand = inside_v.any(1)                                         # [B,A] anchors that can become fg
        ignore = inside.any(1) & ~cand
        # pairwise IoU of gt boxes vs predicted boxes (boxes.py:89-113, centre format)
        gt = labels[..., 1:5]
        pb = out[..., :4]
        tl = torch.max(gt[:, :, None, :2] - gt[:, :, None, 2:] / 2, pb[:, None, :, :2] - pb[:, None, :, 2:] / 2)
        br = torch.min(gt[:, :, None, :2] + gt[:, :, None, 2:] / 2, pb[:, None, :, :2] + pb[:, None, :, 2:] / 2)
        en = (tl < br).to(tl.dtype).prod(-1)
        inter = (br - tl).prod(-1) * en
        iou = inter / ((gt[..., 2] * gt[..., 3])[:, :, None] + (pb[..., 2] * pb[..., 3])[:, None, :] - inter)
        pair_ok = valid[..., None] & cand[:, None, :]                   # [B,N,A]
        iou = torch.where(pair_ok, iou, torch.zeros_like(iou))
        # classification cost: BCE(sqrt(sig(cls)*sig(obj)), onehot) summed over classes, log clamped at -100
        p = (out[..., 5:].sigmoid() * out[..., 4:5].sigmoid()).sqrt()   # [B,A,C]
        logp = torch.log(p).clamp_min(-100.0)
        log1mp = torch.log(1 - p).clamp_min(-100.0)
        gcls = labels[..., 0].long().clamp(0, C - 1)                    # [B,N] (ignored rows masked below)
        sum_log1mp = log1mp.sum(-1)                                     # [B,A]
        idx = gcls[:, :, None].expand(B, labels.shape[1], A)            # [B,N,A]
        lp = torch.gather(logp.transpose(1, 2), 1, idx)                 # logp[b,a,cls_n]
        l1 = torch.gather(log1mp.transpose(1, 2), 1, idx)
        cls_cost = -(lp + (sum_log1mp[:, None, :] - l1))
        cost = cls_cost + 3.0 * (-torch.log(iou + 1e-8)) + 1e6 * (~inside_v).to(iou.dtype)
        cost = torch.where(pair_ok, cost, torch.full_like(cost, float('inf')))
        kk = min(10, A)
        dyn_k = torch.clamp(torch.topk(iou, kk, dim=2).values.sum(2).to(torch.int32), min=1)   # [B,N]
        order = torch.sort(cost, dim=2, stable=True).indices
        rank = torch.empty_like(order)
        rank.scatter_(2, order, torch.arange(A, device=out.device).expand_as(order))
        match = (rank < dyn_k[..., None]) & pair_ok
        multi = match.sum(1) > 1                                        # [B,A]
        best = cost.argmin(1)                                           # [B,A] first minimum over gts
        one = F.one_hot(best, labels.shape[1]).permute(0, 2, 1).bool()  # [B,N,A]
        match = torch.where(multi[:, None, :], one & pair_ok, match)
        fg = match.any(1)
        which = match.to(torch.float32).argmax(1)                       # [B,A]
        miou = (match.to(iou.dtype) * iou).sum(1)
        reg_t = torch.gather(gt, 1, which[..., None].expand(B, A, 4))
        cls_t = F.one_hot(torch.gather(gcls, 1, which), C).to(iou.dtype) * miou[..., None]
        return fg, ignore, reg_t, cls_t, valid.sum()

    def get_losses(self, out, grid, labels):
        labels = self._ignore_bbox(labels)
        fg, ignore, reg_t, cls_t, num_gts = self.assign(out.detach(), grid, labels)
        fgf = fg.to(out.dtype)
        n_fg = fgf.sum()
        num_fg = n_fg.clamp_min(1.0)
        box, obj, cls = out[..., :4], out[..., 4], out[..., 5:]
        # IoU loss (losses.py:18-44), mean over foreground anchors
        tl = torch.max(box[..., :2] - box[..., 2:] / 2, reg_t[..., :2] - reg_t[..., 2:] / 2)
        br = torch.min(box[..., :2] + box[..., 2:] / 2, reg_t[..., :2] + reg_t[..., 2:] / 2)
        en = ((tl[..., 0] < br[..., 0]) & (tl[..., 1] < br[..., 1])).to(tl.dtype)
        d = br - tl                      # explicit product: prod()'s backward reads a zero-check back to the host
        inter = d[..., 0] * d[..., 1] * en
        union = box[..., 2] * box[..., 3] + reg_t[..., 2] * reg_t[..., 3] - inter
        iou = inter / (union + 1e-16)
        loss_iou = (torch.where(fg, 1 - iou ** 2, torch.zeros_like(iou))).sum() / num_fg   # == mean over fg; 0 if none
        obj_l = F.binary_cross_entropy_with_logits(obj, fgf, reduction='none')
        loss_obj = torch.where(ignore, torch.zeros_like(obj_l), obj_l).sum() / num_fg
        cls_l = F.binary_cross_entropy_with_logits(cls, cls_t, reduction='none').sum(-1)
        loss_cls = torch.where(fg, cls_l, torch.zeros_like(cls_l)).sum() / num_fg
        loss_iou = self.reg_weight * loss_iou
        loss_obj = self.obj_weight * loss_obj
        loss_cls = self.cls_weight * loss_cls
        return {'loss': loss_iou + loss_obj + loss_cls, 'iou_loss': loss_iou, 'conf_loss': loss_obj, 'cls_loss': loss_cls,
                'l1_loss': 0.0, 'num_fg': num_fg / num_gts.clamp_min(1).to(out.dtype)}
