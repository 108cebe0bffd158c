"""Host binding of the CUDA neck + head + loss executor (include/leod_b200.h: leod_detect_*, leod_fpn_head_*,
leod_simota_loss_*).

`YOLOPAFPN` and `YOLOXHead` (the reference's two modules, models/detection/yolox_extension/models/yolo_pafpn.py:18 and
models/detection/yolox/models/yolo_head.py:21) are parameter name-spaces here: their `nn.Parameter`s / BatchNorm buffers
are views into the flat buffers this engine owns, under the reference's state_dict keys.  The arithmetic of
`YoloXDetector.forward_detect` (detector.py:55-77) — neck, head, decode, SimOTA loss, and the whole backward — runs in
the library; there is no PyTorch fallback.
"""
import ctypes
import math
from typing import List, Optional

import torch
import torch.nn as nn

from leod_b200 import _lib


class _Node(nn.Module):
    """Name-space module; indexable like the reference's nn.ModuleList / nn.Sequential."""

    def __getitem__(self, i):
        return getattr(self, str(i))

    def __len__(self):
        return len(self._modules)


def _walk(root: nn.Module, parts):
    mod = root
    for part in parts:
        if not hasattr(mod, part):
            mod.add_module(part, _Node())
        mod = getattr(mod, part)
    return mod


class _DetectFn(torch.autograd.Function):
    """neck + head + loss forward; backward returns the feature gradients and leaves the parameter gradients in the
    engine's flat gradient buffer."""

    @staticmethod
    def forward(ctx, engine, anchor, labels, *feats):
        preds, losses, keep = engine._forward(feats, labels, training=True)
        ctx.engine = engine
        ctx.keep = keep
        ctx.gen = engine._gen
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(preds)
        out = (losses[0], preds, losses[1], losses[2], losses[3], losses[5])
        ctx.mark_non_differentiable(*out[2:])
        return out

    @staticmethod
    def backward(ctx, g_loss, *unused):
        n = len(ctx.keep[0])
        if g_loss is None:
            return (None, None, None) + (None,) * n
        dfeats = ctx.engine._backward(ctx.keep, g_loss, ctx.gen)
        ctx.keep = None
        return (None, None, None) + tuple(dfeats)


class DetectEngine:
    def __init__(self, fpn: nn.Module, head: nn.Module, in_channels, strides, in_res_hw, compute_dtype: torch.dtype):
        assert compute_dtype in (torch.bfloat16, torch.float32)
        self.fpn, self.head = fpn, head
        self.in_channels = tuple(int(c) for c in in_channels)
        self.strides = tuple(int(s) for s in strides)
        self.in_res_hw = (int(in_res_hw[0]), int(in_res_hw[1]))
        self.compute_dtype = compute_dtype
        self.num_classes = int(head.num_classes)
        self._handle = None
        self._handle_device = None
        self._layout = self._query_layout()
        self.num_anchors = self._layout['anchors']
        self._flat = torch.zeros(self._layout['n_params'], dtype=torch.float32)
        self._buf = torch.zeros(self._layout['n_buffers'], dtype=torch.float32)
        self._cnt = torch.zeros(self._layout['n_counters'], dtype=torch.int64)
        self._flat_grad = None
        self._anchor = None
        self._prepared_version = None
        self._bwd_pending = False
        self._gen = 0
        self.grad_sync = None            # optional callable(flat_grad): e.g. the data-parallel all-reduce
        self._allreduce_cb = None        # keeps the ctypes callback object alive
        self._register()
        self.reset_parameters()

    # ------------------------------------------------------------------ layout / parameters
    def _make_cfg(self):
        c = _lib.DetectCfg()
        for i in range(3):
            c.in_channels[i] = self.in_channels[i]
            c.strides[i] = self.strides[i]
        c.num_classes = self.num_classes
        c.n_bottleneck = int(self.fpn.n_bottleneck)
        c.in_h, c.in_w = self.in_res_hw
        c.dtype = _lib.leod_dtype(self.compute_dtype)
        c.bn_eps, c.bn_momentum = 1e-5, 0.1
        c.ignore_label = float(self.head.ignore_label)
        th = self.head.ignore_bbox_thresh or []
        assert len(th) <= 8
        c.n_ignore_thresh = len(th)
        for i, t in enumerate(th):
            c.ignore_thresh[i] = float(t)
        c.reg_weight, c.obj_weight, c.cls_weight = float(self.head.reg_weight), float(self.head.obj_weight), float(self.head.cls_weight)
        return c

    def _query_layout(self):
        l = _lib.lib()
        h = ctypes.c_void_p()
        cfg = self._make_cfg()
        _lib.check(l.leod_detect_layout_only(ctypes.byref(cfg), ctypes.byref(h)), 'detect_layout_only')
        try:
            name = ctypes.create_string_buffer(256)
            off, nd, shp = ctypes.c_int64(), ctypes.c_int32(), (ctypes.c_int64 * 4)()

            def entries(fn):
                out = []
                for i in range(fn(h, -1, None, 0, None, None, None)):
                    _lib.check(fn(h, i, name, 256, ctypes.byref(off), ctypes.byref(nd), ctypes.byref(shp)))
                    out.append((name.value.decode(), int(off.value), tuple(int(shp[k]) for k in range(nd.value))))
                return out

            params = entries(l.leod_detect_param_info)
            buffers = entries(l.leod_detect_buffer_info)
            counters = []
            for i in range(l.leod_detect_counter_info(h, -1, None, 0, None)):
                _lib.check(l.leod_detect_counter_info(h, i, name, 256, ctypes.byref(off)))
                counters.append((name.value.decode(), int(off.value)))
            return dict(params=params, buffers=buffers, counters=counters, n_params=int(l.leod_detect_param_count(h)),
                        n_buffers=int(l.leod_detect_buffer_count(h)), n_counters=int(l.leod_detect_counter_count(h)),
                        anchors=int(l.leod_detect_num_anchors(h)))
        finally:
            l.leod_detect_destroy(h)

    def _root(self, name):
        top, rest = name.split('.', 1)
        return {'fpn': self.fpn, 'yolox_head': self.head}[top], rest.split('.')

    def _register(self):
        self._param_views, self._buffer_views, self._counter_views = [], [], []
        for name, off, shape in self._layout['params']:
            n = math.prod(shape)
            root, parts = self._root(name)
            p = nn.Parameter(self._flat[off:off + n].view(shape))
            _walk(root, parts[:-1]).register_parameter(parts[-1], p)
            self._param_views.append((p, off, n, shape, name))
        for name, off, shape in self._layout['buffers']:
            n = math.prod(shape)
            root, parts = self._root(name)
            mod = _walk(root, parts[:-1])
            mod.register_buffer(parts[-1], self._buf[off:off + n].view(shape))
            self._buffer_views.append((mod, parts[-1], off, n, shape))
        for name, off in self._layout['counters']:
            root, parts = self._root(name)
            mod = _walk(root, parts[:-1])
            mod.register_buffer(parts[-1], self._cnt[off:off + 1].view(()))
            self._counter_views.append((mod, parts[-1], off))

    def reset_parameters(self):
        """Same distributions as the reference's constructors: nn.Conv2d default init, BatchNorm2d (1, 0 / 0, 1), and the
        prior-probability bias of the objectness / class predictors (yolo_head.py:183-193)."""
        prior = -math.log((1 - 0.01) / 0.01)
        shapes = {name: shape for _, _, _, shape, name in self._param_views}
        with torch.no_grad():
            for p, _, _, shape, name in self._param_views:
                if name.endswith('bn.weight'):
                    p.fill_(1.0)
                elif name.endswith('bn.bias'):
                    p.zero_()
                elif p.dim() >= 2:
                    nn.init.kaiming_uniform_(p, a=math.sqrt(5))
                elif 'cls_preds' in name or 'obj_preds' in name:
                    p.fill_(prior)
                else:   # reg_preds bias: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
                    w = shapes[name[:-len('bias')] + 'weight']
                    bound = 1.0 / math.sqrt(math.prod(w[1:]))
                    p.uniform_(-bound, bound)
            for mod, key, off, n, shape in self._buffer_views:
                getattr(mod, key).fill_(1.0 if key == 'running_var' else 0.0)
            self._cnt.zero_()

    def reflatten(self, device=None):
        """Re-establish the aliasing after nn.Module._apply (.cuda() / .to()) or load_state_dict(assign=True) replaced
        the storage of the parameters / buffers."""
        if device is None:
            device = self._param_views[0][0].device
        if any(p.dtype != torch.float32 for p, *_ in self._param_views):
            raise RuntimeError('leod_b200 neck/head parameters must stay fp32 (the compute dtype is chosen by compute_dtype)')
        flat = torch.zeros(self._layout['n_params'], dtype=torch.float32, device=device)
        for p, off, n, shape, _ in self._param_views:
            flat[off:off + n].copy_(p.data.reshape(-1).to(device=device, dtype=torch.float32))
            p.data = flat[off:off + n].view(shape)
            p.grad = None
        buf = torch.zeros(self._layout['n_buffers'], dtype=torch.float32, device=device)
        for mod, key, off, n, shape in self._buffer_views:
            buf[off:off + n].copy_(getattr(mod, key).reshape(-1).to(device=device, dtype=torch.float32))
            mod._buffers[key] = buf[off:off + n].view(shape)
        cnt = torch.zeros(self._layout['n_counters'], dtype=torch.int64, device=device)
        for mod, key, off in self._counter_views:
            cnt[off:off + 1].copy_(getattr(mod, key).reshape(-1).to(device=device, dtype=torch.int64))
            mod._buffers[key] = cnt[off:off + 1].view(())
        self._flat, self._buf, self._cnt = flat, buf, cnt
        self._flat_grad = None
        self._prepared_version = None
        if self._handle is not None and self._handle_device == flat.device:
            self._bind()

    @property
    def flat_params(self) -> torch.Tensor:
        return self._flat

    @property
    def flat_grads(self) -> torch.Tensor:
        self._ensure_grad_buffer()
        return self._flat_grad

    @property
    def flat_buffers(self) -> torch.Tensor:
        return self._buf

    # ------------------------------------------------------------------ device handle
    def _ensure_handle(self):
        dev = self._flat.device
        if dev.type != 'cuda':
            raise RuntimeError('leod_b200 neck/head run on CUDA only (no CPU fallback): move the model to a B200 first')
        if self._handle is not None and self._handle_device == dev:
            return
        self.destroy()
        h = ctypes.c_void_p()
        cfg = self._make_cfg()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().leod_detect_create(ctypes.byref(cfg), ctypes.byref(h)), 'detect_create')
        self._handle, self._handle_device = h, dev
        self._anchor = torch.zeros(1, device=dev, requires_grad=True)
        self._bind()
        if self._allreduce_cb is not None:
            _lib.check(_lib.lib().leod_detect_set_allreduce(self._handle, self._allreduce_cb, None), 'set_allreduce')

    def destroy(self):
        if self._handle is not None:
            _lib.lib().leod_detect_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:   # noqa: BLE001  (interpreter shutdown)
            pass

    def _ensure_grad_buffer(self):
        if self._flat_grad is None or self._flat_grad.device != self._flat.device:
            self._flat_grad = torch.zeros_like(self._flat)
            if self._handle is not None:
                self._bind()

    def _bind(self):
        _lib.check(_lib.lib().leod_detect_bind(self._handle, _lib.ptr(self._flat), _lib.ptr(self._flat_grad), _lib.ptr(self._buf),
                                               _lib.ptr(self._cnt)), 'detect_bind')
        self._prepared_version = None

    def _params_version(self):
        return sum(p._version for p, *_ in self._param_views)

    def prepare(self, force: bool = False):
        self._ensure_handle()
        # parameters / buffers replaced behind our back (load_state_dict(assign=True), manual .data assignment)
        p0 = self._param_views[0][0]
        if p0.data_ptr() != self._flat.data_ptr() + 4 * self._param_views[0][1]:
            self.reflatten(p0.device)
        v = self._params_version()
        if force or v != self._prepared_version:
            with torch.cuda.device(self._handle_device):
                _lib.check(_lib.lib().leod_detect_prepare(self._handle, _lib.stream_ptr(self._handle_device)), 'detect_prepare')
            self._prepared_version = v

    def mark_params_updated(self):
        """For optimizers that write `flat_params` directly."""
        self._prepared_version = None

    def set_sync_batchnorm(self, process_group=None):
        """SyncBatchNorm semantics (train.py:247) over torch.distributed: the library calls back once per dependency level
        with a device buffer of doubles to be summed over the ranks."""
        import torch.distributed as dist

        def cb(ctx, buf, n, stream):
            try:
                t = _lib.tensor_from_ptr(buf, int(n), torch.float64, self._flat.device)
                dist.all_reduce(t, group=process_group)
                return 0
            except Exception as e:   # noqa: BLE001  (must not propagate through the C frame)
                self._cb_error = e
                return -1

        self._allreduce_cb = _lib.ALLREDUCE_FN(cb)
        if self._handle is not None:
            _lib.check(_lib.lib().leod_detect_set_allreduce(self._handle, self._allreduce_cb, None), 'set_allreduce')

    # ------------------------------------------------------------------ execution
    def _nhwc(self, t: torch.Tensor) -> torch.Tensor:
        u = t.permute(0, 2, 3, 1)
        if u.dtype != self.compute_dtype:
            u = u.to(self.compute_dtype)
        return u if u.is_contiguous() else u.contiguous()

    def _forward(self, feats, labels: Optional[torch.Tensor], training: bool):
        l = _lib.lib()
        dev = self._flat.device
        B = feats[0].shape[0]
        H, W = self.in_res_hw
        for f, c, s in zip(feats, self.in_channels, self.strides):
            assert tuple(f.shape) == (B, c, H // s, W // s), (tuple(f.shape), (B, c, H // s, W // s))
        x = [self._nhwc(f) for f in feats]
        preds = torch.empty((B, self.num_anchors, 5 + self.num_classes), dtype=torch.float32, device=dev)
        arr = (ctypes.c_void_p * 3)(*[t.data_ptr() for t in x])
        st = _lib.stream_ptr(dev)
        self._gen += 1
        with torch.cuda.device(dev):
            _lib.check(l.leod_fpn_head_fwd(self._handle, arr, B, 1 if training else 0, _lib.ptr(preds), st), 'fpn_head_fwd')
            losses = None
            if training:
                assert labels is not None and labels.dim() == 3 and labels.shape[0] == B and labels.shape[2] == 7, labels.shape
                labels = labels.to(device=dev, dtype=torch.float32).contiguous()
                losses = torch.empty(6, dtype=torch.float32, device=dev)
                _lib.check(l.leod_simota_loss_fwd(self._handle, _lib.ptr(labels), labels.shape[1], _lib.ptr(losses), st), 'simota_loss_fwd')
        return preds, losses, (x, labels)

    def _backward(self, keep, g_loss, gen):
        if gen != self._gen:
            raise RuntimeError('leod_b200 neck/head: another forward_detect ran between this forward and its backward; the library '
                               'keeps the activations of ONE training forward')
        l = _lib.lib()
        x, labels = keep
        dev = self._flat.device
        self._begin_backward_pass()
        g = g_loss.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        dx = [torch.empty_like(t) for t in x]
        arr = (ctypes.c_void_p * 3)(*[t.data_ptr() for t in dx])
        st = _lib.stream_ptr(dev)
        with torch.cuda.device(dev):
            _lib.check(l.leod_simota_loss_bwd(self._handle, _lib.ptr(labels), labels.shape[1], _lib.ptr(g), st), 'simota_loss_bwd')
            _lib.check(l.leod_fpn_head_bwd(self._handle, arr, st), 'fpn_head_bwd')
        return [t.permute(0, 3, 1, 2) for t in dx]

    def raw_outputs(self, B: int) -> torch.Tensor:
        """Diagnostics: undecoded head outputs [B, A, 8] (reg x4, obj logit, class logits, zero padding) of the last forward."""
        out = torch.empty((B, self.num_anchors, 8), dtype=torch.float32, device=self._flat.device)
        with torch.cuda.device(out.device):
            _lib.check(_lib.lib().leod_detect_get_raw(self._handle, _lib.ptr(out), _lib.stream_ptr(out.device)), 'get_raw')
        return out

    def raw_grad(self, B: int) -> torch.Tensor:
        """Diagnostics: d loss / d raw outputs [B, A, 8] as left by the last loss backward."""
        out = torch.empty((B, self.num_anchors, 8), dtype=torch.float32, device=self._flat.device)
        with torch.cuda.device(out.device):
            _lib.check(_lib.lib().leod_detect_get_raw_grad(self._handle, _lib.ptr(out), _lib.stream_ptr(out.device)), 'get_raw_grad')
        return out

    def backward_from_raw_grad(self, feats, draw: torch.Tensor):
        """Back-propagate a caller-supplied gradient w.r.t. the raw head outputs ([B, A, 8] fp32) through head + neck of the
        last training-mode forward.  Parameter gradients accumulate in `flat_grads`; returns the feature gradients (NCHW views)."""
        l = _lib.lib()
        dev = self._flat.device
        self._ensure_grad_buffer()
        draw = draw.to(device=dev, dtype=torch.float32).contiguous()
        dx = [torch.empty_like(self._nhwc(f)) for f in feats]
        arr = (ctypes.c_void_p * 3)(*[t.data_ptr() for t in dx])
        with torch.cuda.device(dev):
            _lib.check(l.leod_detect_set_raw_grad(self._handle, _lib.ptr(draw), _lib.stream_ptr(dev)), 'set_raw_grad')
            _lib.check(l.leod_fpn_head_bwd(self._handle, arr, _lib.stream_ptr(dev)), 'fpn_head_bwd')
        return [t.permute(0, 3, 1, 2) for t in dx]

    def last_assignment(self, B: int):
        """Diagnostics: (matched label row per anchor [B, A] int32, -1 = background; IoU of the matched pair [B, A]) of the last
        training-mode forward (yolo_head.py:768-774)."""
        dev = self._flat.device
        a = torch.empty((B, self.num_anchors), dtype=torch.int32, device=dev)
        m = torch.empty((B, self.num_anchors), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().leod_simota_assignment(self._handle, _lib.ptr(a), _lib.ptr(m), _lib.stream_ptr(dev)), 'simota_assignment')
        return a, m

    def detect(self, feats: List[torch.Tensor], targets: Optional[torch.Tensor], training: bool):
        """-> (predictions [B, A, 5 + C] fp32, losses dict | None)  (detector.py:55-77, yolo_head.py:269-276)"""
        self.prepare()
        if not training:
            preds, _, _ = self._forward(feats, None, training=False)
            return preds, None
        assert targets is not None, 'training-mode forward_detect needs targets (detector.py:69-70)'
        if torch.is_grad_enabled():
            self._ensure_grad_buffer()
            loss, preds, iou_l, obj_l, cls_l, nfg = _DetectFn.apply(self, self._anchor, targets, *feats)
        else:
            preds, ls, _ = self._forward(feats, targets, training=True)
            loss, iou_l, obj_l, cls_l, nfg = ls[0], ls[1], ls[2], ls[3], ls[5]
        return preds, {'loss': loss, 'iou_loss': iou_l, 'conf_loss': obj_l, 'cls_loss': cls_l, 'l1_loss': 0.0, 'num_fg': nfg}

    # ------------------------------------------------------------------ gradient hand-over (as the backbone's)
    def _begin_backward_pass(self):
        if self._bwd_pending:
            return
        self._bwd_pending = True
        self._ensure_grad_buffer()
        if self._param_views[0][0].grad is None:     # zero_grad(set_to_none=True) dropped the views: start from zero
            self._flat_grad.zero_()
        torch.autograd.Variable._execution_engine.queue_callback(self._end_backward_pass)

    def _end_backward_pass(self):
        self._bwd_pending = False
        if self.grad_sync is not None:
            self.grad_sync(self._flat_grad)
        for p, off, n, shape, _ in self._param_views:
            if p.grad is None:
                p.grad = self._flat_grad[off:off + n].view(shape)
