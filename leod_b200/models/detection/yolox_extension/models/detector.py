"""Model facade — the drop-in boundary.  Mirror of
models/detection/yolox_extension/models/detector.py:18-91 (YoloXDetector): same constructor, same
`forward_backbone` / `forward_detect` / `forward` signatures and return values, same state_dict keys
(`backbone.*`, `fpn.*`, `yolox_head.*`)."""
from typing import Dict, Optional, Tuple, Union

import torch as th

from ...recurrent_backbone import build_recurrent_backbone
from .build import build_yolox_fpn, build_yolox_head


class _GraphedDetectFn(th.autograd.Function):
    """Autograd node in front of a captured neck+head+loss forward/backward (see _GraphedDetect)."""

    @staticmethod
    def forward(ctx, runner, labels, *feats):
        ctx.runner = runner
        ctx.set_materialize_grads(False)
        out = runner.replay(feats, labels)
        ctx.mark_non_differentiable(*out[1:])
        return out

    @staticmethod
    def backward(ctx, g_loss, *unused):
        if g_loss is None:
            return (None, None) + (None,) * len(ctx.runner.s_feats)
        return (None, None) + ctx.runner.grads(g_loss)


class _GraphedDetect:
    """Training-mode neck + head + SimOTA loss, forward AND backward, captured once per input shape in a CUDA graph
    and replayed: ~800 small library launches per step become one graph launch (the loss has no host
    synchronisation, yolo_head.py in this package).  Parameter gradients are produced by the graph into static
    buffers and added to `.grad` when autograd reaches the node; feature gradients flow on to the backbone."""

    def __init__(self, det, feats, labels):
        self.det = det
        self.params = [p for m in (det.fpn, det.yolox_head) for p in m.parameters() if p.requires_grad]
        self.s_feats = [f.detach().clone().requires_grad_(True) for f in feats]
        self.s_labels = labels.detach().clone()
        buffers = [b for m in (det.fpn, det.yolox_head) for b in m.buffers()]
        saved = [b.clone() for b in buffers]     # the warm-up passes must not advance the BatchNorm running statistics
        side = th.cuda.Stream()
        side.wait_stream(th.cuda.current_stream())
        with th.cuda.stream(side):       # warm-up outside capture (cuDNN plans, BN buffers, grid cache)
            for _ in range(3):
                _, losses = det._detect_eager(self.s_feats, self.s_labels)
                th.autograd.grad(losses['loss'], self.s_feats + self.params, allow_unused=True)
        th.cuda.current_stream().wait_stream(side)
        with th.no_grad():
            for b, v in zip(buffers, saved):
                b.copy_(v)
        self.graph = th.cuda.CUDAGraph()
        with th.cuda.graph(self.graph):
            preds, losses = det._detect_eager(self.s_feats, self.s_labels)
            grads = th.autograd.grad(losses['loss'], self.s_feats + self.params, allow_unused=True)
        self.s_preds, self.s_losses = preds.detach(), {k: (v.detach() if th.is_tensor(v) else v) for k, v in losses.items()}
        nf = len(self.s_feats)
        self.s_dfeats = grads[:nf]
        self.pg = [(p, g) for p, g in zip(self.params, grads[nf:]) if g is not None]

    def replay(self, feats, labels):
        # Multi-tensor copy KERNELS, not Tensor.copy_/clone: dense same-layout device copies go to a copy engine, where
        # they would queue behind the bulk host->device upload of the next step's input batch.
        th._foreach_copy_([s.data for s in self.s_feats], [f.detach() for f in feats])
        th._foreach_copy_([self.s_labels], [labels.detach()])
        self.graph.replay()
        scal = [k for k, v in self.s_losses.items() if th.is_tensor(v)]
        self.keys = scal
        src = [self.s_losses['loss'], self.s_preds] + [self.s_losses[k] for k in scal if k != 'loss']
        out = [th.empty_like(t) for t in src]
        by_dtype = {}
        for o, t in zip(out, src):
            by_dtype.setdefault(t.dtype, ([], []))
            by_dtype[t.dtype][0].append(o)
            by_dtype[t.dtype][1].append(t)
        for dst, srcs in by_dtype.values():
            th._foreach_copy_(dst, srcs)
        return tuple(out)

    def grads(self, g_loss):
        ps = [p for p, _ in self.pg]
        gs = th._foreach_mul([g for _, g in self.pg], g_loss.to(th.float32))
        for p in ps:
            if p.grad is None:
                p.grad = th.zeros_like(p)
        th._foreach_add_([p.grad for p in ps], gs)
        return tuple((d * g_loss).to(d.dtype) for d in self.s_dfeats)


class YoloXDetector(th.nn.Module):
    """RNN-based MaxViT backbone (CUDA library) + YOLOX PAFPN/head."""

    def __init__(self, model_cfg, ssod: bool = False):
        super().__init__()
        backbone_cfg, fpn_cfg, head_cfg = model_cfg.backbone, model_cfg.fpn, model_cfg.head
        self.backbone = build_recurrent_backbone(backbone_cfg)
        in_channels = self.backbone.get_stage_dims(tuple(fpn_cfg.in_stages))
        self.fpn = build_yolox_fpn(fpn_cfg, in_channels=in_channels)
        strides = self.backbone.get_strides(tuple(fpn_cfg.in_stages))
        self.yolox_head = build_yolox_head(head_cfg, in_channels=in_channels, strides=strides, ssod=ssod)
        # True: training-mode forward_detect replays a CUDA graph of neck+head+loss fwd/bwd (one per input shape)
        self.graph_detect = False
        self._detect_graphs = {}

    def forward_backbone(self, x: th.Tensor, previous_states=None, token_mask: Optional[th.Tensor] = None):
        """-> ({stage: [B,C,h,w]}, [(h,c)]*4)   (detector.py:35-53)"""
        return self.backbone(x, previous_states, token_mask)

    def forward_detect(self, backbone_features: Dict[int, th.Tensor], targets: Optional[th.Tensor] = None,
                       soft_targets: Optional[th.Tensor] = None) -> Tuple[th.Tensor, Union[Dict[str, th.Tensor], None]]:
        """-> (predictions [B,A,4+1+num_cls] decoded, losses dict | None)   (detector.py:55-77)"""
        if self.training and self.graph_detect and th.is_grad_enabled() and targets is not None and soft_targets is None:
            return self._detect_graphed(backbone_features, targets)
        keys = self.fpn.in_features
        return self._detect_eager([backbone_features[k] for k in keys], targets, soft_targets)

    def _detect_eager(self, feats, targets=None, soft_targets=None):
        feats = {k: v.float() for k, v in zip(self.fpn.in_features, feats)}  # neck/head run in fp32 (TF32 convs)
        fpn_features = self.fpn(feats)
        if self.training:
            assert targets is not None
            return self.yolox_head(fpn_features, targets, soft_targets)
        outputs, losses = self.yolox_head(fpn_features)
        assert losses is None
        return outputs, losses

    def _detect_graphed(self, backbone_features, targets):
        feats = [backbone_features[k] for k in self.fpn.in_features]
        n = targets.shape[1]
        npad = max(16, -(-n // 16) * 16)       # static label capacity: zero rows are "no label" (yolo_head.py:410)
        if npad != n:
            targets = th.cat((targets, targets.new_zeros(targets.shape[0], npad - n, targets.shape[2])), 1)
        key = (tuple(tuple(f.shape) for f in feats), tuple(f.dtype for f in feats), tuple(targets.shape))
        runner = self._detect_graphs.get(key)
        if runner is None:
            runner = self._detect_graphs[key] = _GraphedDetect(self, feats, targets)
        out = _GraphedDetectFn.apply(runner, targets, *feats)
        losses = {'loss': out[0]}
        others = [k for k in runner.keys if k != 'loss']
        for k, v in zip(others, out[2:]):
            losses[k] = v.detach()
        for k, v in runner.s_losses.items():
            if not th.is_tensor(v):
                losses[k] = v
        return out[1].detach(), losses

    def forward(self, x: th.Tensor, previous_states=None, retrieve_detections: bool = True,
                targets: Optional[th.Tensor] = None):
        backbone_features, states = self.forward_backbone(x, previous_states)
        if not retrieve_detections:
            assert targets is None
            return None, None, states
        outputs, losses = self.forward_detect(backbone_features=backbone_features, targets=targets)
        return outputs, losses, states
