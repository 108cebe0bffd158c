"""B200-native RVT recurrent backbone behind the reference's interface.

Mirror of models/detection/recurrent_backbone/maxvit_rnn.py:23-115 (class RNNDetector): same
constructor argument (the `model.backbone` config node), same `forward(x, prev_states, token_mask)`
signature and return value, same `get_stage_dims` / `get_strides`, same parameter names and shapes
(state_dict compatible with the released checkpoints).  The arithmetic runs in the CUDA library
(include/leod_b200.h, `leod_backbone_*`); there is no PyTorch fallback.

Differences that are visible to a caller, all by design:
  * parameters are views into ONE flat fp32 buffer (`flat_params`) and their gradients views into
    one flat buffer (`flat_grads`), so the optimizer / EMA / all-reduce can each be a single launch;
  * returned features/states have NCHW *shape* but channels-last memory, in the compute dtype;
  * parameter gradients are written by the library at the end of the backward pass (an autograd
    engine callback), not by autograd accumulation nodes.
"""
from typing import Dict, List, Optional, Tuple

import ctypes
import torch
import torch.nn as nn

from leod_b200 import _lib

LstmState = Tuple[torch.Tensor, torch.Tensor]


def _cfg_get(node, key, default=None):
    if hasattr(node, 'get'):
        return node.get(key, default)
    return getattr(node, key, default)


class _Node(nn.Module):
    """Name-space module of the parameter tree; indexable like the reference's nn.ModuleList."""

    def __getitem__(self, i):
        return getattr(self, str(i))

    def __len__(self):
        return len(self._modules)


class _BackboneStep(torch.autograd.Function):
    """One timestep through the four stages (leod_backbone_step_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, bb, anchor, x, *states):
        outs, save = bb._step_forward(x, states, need_grad=True)
        ctx.bb = bb
        ctx.save = save
        ctx.xinfo = (x, )
        ctx.save_for_backward(*[s for s in states if s is not None], *outs)
        ctx.state_mask = [s is not None for s in states]
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        bb = ctx.bb
        saved = list(ctx.saved_tensors)
        n_in = sum(ctx.state_mask)
        it = iter(saved[:n_in])
        states = [next(it) if m else None for m in ctx.state_mask]
        outs = saved[n_in:]
        dstates = bb._step_backward(ctx.xinfo[0], states, outs, ctx.save, grads)
        ctx.save = None
        return (None, None, None) + tuple(d if m else None for d, m in zip(dstates, ctx.state_mask))


class _BackboneSeq(torch.autograd.Function):
    """A whole BPTT window (leod_backbone_seq_fwd / _bwd): x [L,B,C,H,W] -> h of every timestep, final c."""

    @staticmethod
    def forward(ctx, bb, anchor, x, *states):
        outs, keep = bb._seq_forward(x, states)
        ctx.bb = bb
        ctx.keep = keep
        ctx.gen = bb._seq_gen
        ctx.state_mask = [s is not None for s in states]
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(*outs[:4])
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        if ctx.gen != ctx.bb._seq_gen:
            raise RuntimeError('leod_b200 backbone: another forward_sequence ran between this forward and its backward; the library '
                               'keeps the activations of ONE window (leod_backbone_seq_arena_bytes) — run backward first')
        dstates = ctx.bb._seq_backward(ctx.keep, ctx.saved_tensors, grads)
        ctx.keep = None
        return (None, None, None) + tuple(d if m else None for d, m in zip(dstates, ctx.state_mask))


class RNNDetector(nn.Module):
    """RNN-based backbone with MaxViT blocks (reference: maxvit_rnn.py:23)."""

    def __init__(self, mdl_config, compute_dtype: Optional[torch.dtype] = None):
        super().__init__()
        self.in_channels = int(mdl_config.input_channels)
        self.embed_dim = int(mdl_config.embed_dim)
        dim_multiplier = tuple(mdl_config.dim_multiplier)
        num_blocks = tuple(mdl_config.num_blocks)
        assert len(num_blocks) == 4 and len(dim_multiplier) == 4
        if tuple(dim_multiplier) != (1, 2, 4, 8) or tuple(num_blocks) != (1, 1, 1, 1):
            raise NotImplementedError('leod_b200 backbone supports dim_multiplier (1,2,4,8), num_blocks (1,1,1,1) '
                                      '(every shipped RVT config)')
        if bool(mdl_config.enable_masking):
            raise NotImplementedError('token masking is disabled in every shipped config and not implemented')
        assert int(mdl_config.stem.patch_size) == 4
        stage_cfg = mdl_config.stage
        att = stage_cfg.attention
        if bool(att.use_torch_mha) or bool(att.mlp_gated) or att.mlp_activation != 'gelu':
            raise NotImplementedError('only SelfAttentionCl + non-gated GELU MLP (the shipped configuration)')
        if bool(stage_cfg.lstm.dws_conv):
            raise NotImplementedError('dws_conv LSTM is not used by RVT')
        self.dim_head = int(_cfg_get(att, 'dim_head', 32))
        part = att.partition_size
        self.partition_size = (int(part[0]), int(part[1])) if not isinstance(part, int) else (part, part)
        self.mlp_ratio = int(_cfg_get(att, 'mlp_ratio', 4))
        self.norm_eps = float(_cfg_get(att, 'norm_eps', 1e-5))
        in_res = _cfg_get(mdl_config, 'in_res_hw', None)
        # config/modifier.py:49-64 always sets in_res_hw; derive it from the partition otherwise
        self.in_res_hw = (int(in_res[0]), int(in_res[1])) if in_res is not None else \
            (32 * self.partition_size[0], 32 * self.partition_size[1])
        if compute_dtype is None:
            compute_dtype = _cfg_get(mdl_config, 'compute_dtype', 'bf16')
        if isinstance(compute_dtype, str):
            compute_dtype = {'bf16': torch.bfloat16, 'fp32': torch.float32, 'bfloat16': torch.bfloat16,
                             'float32': torch.float32}[compute_dtype]
        assert compute_dtype in (torch.bfloat16, torch.float32)
        self.compute_dtype = compute_dtype
        self.stage_dims = [self.embed_dim * m for m in dim_multiplier]
        self.strides = [4, 8, 16, 32]
        self.num_stages = 4

        self._handle = None
        self._handle_device = None
        self._layout = self._query_layout()
        n = self._layout['count']
        self._flat = torch.zeros(n, dtype=torch.float32)
        self._flat_grad = None
        self._anchor = None
        self._prepared_version = None
        self._bwd_pending = False
        self._seq_gen = 0              # forward_sequence calls so far: a backward must belong to the latest one (shared arena)
        self.grad_sync = None          # optional callable(flat_grad) -> None, e.g. an all-reduce
        self._register_params()
        self.reset_parameters()

    # ------------------------------------------------------------------ parameters
    def _make_cfg(self):
        c = _lib.BackboneCfg()
        c.in_channels, c.embed_dim, c.dim_head = self.in_channels, self.embed_dim, self.dim_head
        c.part_h, c.part_w = self.partition_size
        c.mlp_ratio = self.mlp_ratio
        c.in_h, c.in_w = self.in_res_hw
        c.dtype = _lib.leod_dtype(self.compute_dtype)
        c.ln_eps = self.norm_eps
        return c

    def _query_layout(self):
        """Ask the library for the flat-buffer layout (names follow the reference state_dict)."""
        l = _lib.lib()
        h = ctypes.c_void_p()
        cfg = self._make_cfg()
        # creating a handle allocates device memory; on a CPU-only box fall back to a layout-only
        # query, which the library serves without touching the device when no GPU is present
        _lib.check(l.leod_backbone_layout_only(ctypes.byref(cfg), ctypes.byref(h)), 'backbone_layout_only')
        try:
            n = l.leod_backbone_param_info(h, -1, None, 0, None, None, None)
            entries = []
            name = ctypes.create_string_buffer(256)
            off = ctypes.c_int64()
            nd = ctypes.c_int32()
            shp = (ctypes.c_int64 * 4)()
            for i in range(n):
                _lib.check(l.leod_backbone_param_info(h, i, name, 256, ctypes.byref(off), ctypes.byref(nd), ctypes.byref(shp)))
                entries.append((name.value.decode(), int(off.value), tuple(int(shp[k]) for k in range(nd.value))))
            count = int(l.leod_backbone_param_count(h))
        finally:
            l.leod_backbone_destroy(h)
        return dict(entries=entries, count=count)

    def _register_params(self):
        """Create nn.Parameters that alias slices of the flat buffer, under the reference's names
        (stages.{i}.downsample_cf2cl.conv.weight, ...att_window.self_attn.qkv.weight, ...)."""
        self._param_views = []
        for name, off, shape in self._layout['entries']:
            numel = 1
            for s in shape:
                numel *= s
            p = nn.Parameter(self._flat[off:off + numel].view(shape))
            mod = self
            parts = name.split('.')
            for part in parts[:-1]:
                if not hasattr(mod, part):
                    mod.add_module(part, _Node())
                mod = getattr(mod, part)
            mod.register_parameter(parts[-1], p)
            self._param_views.append((p, off, numel, shape))

    def _reflatten(self, device):
        """Re-establish the aliasing after nn.Module._apply (.cuda(), .to()) replaced param storage."""
        flat = torch.zeros(self._layout['count'], dtype=torch.float32, device=device)
        for p, off, numel, shape in self._param_views:
            flat[off:off + numel].copy_(p.data.reshape(-1).to(device=device, dtype=torch.float32))
            p.data = flat[off:off + numel].view(shape)
            p.grad = None
        self._flat = flat
        self._flat_grad = None
        self._prepared_version = None

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        dev = self._param_views[0][0].device
        ok = all(p.dtype == torch.float32 for p, *_ in self._param_views)
        if not ok:
            raise RuntimeError('leod_b200 backbone parameters must stay fp32 (compute dtype is chosen by compute_dtype)')
        self._reflatten(dev)
        return self

    def reset_parameters(self):
        """Same distributions as the reference's default constructors (nn.Conv2d / nn.Linear /
        LayerNorm / LayerScale 1e-5, maxvit.py:45-53)."""
        import math
        with torch.no_grad():
            for p, off, numel, shape in self._param_views:
                name = [n for n, o, s in self._layout['entries'] if o == off][0]
                if name.endswith('gamma'):
                    p.fill_(1e-5)
                elif name.endswith('norm.weight') or name.endswith('norm1.weight') or name.endswith('norm2.weight'):
                    p.fill_(1.0)
                elif name.endswith('.bias') and ('norm' in name):
                    p.zero_()
                elif p.dim() >= 2:
                    nn.init.kaiming_uniform_(p, a=math.sqrt(5))
                else:  # linear / conv bias: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
                    wname = name[:-len('bias')] + 'weight'
                    wshape = [s for n, o, s in self._layout['entries'] if n == wname][0]
                    fan_in = 1
                    for s in wshape[1:]:
                        fan_in *= s
                    bound = 1.0 / math.sqrt(fan_in)
                    p.uniform_(-bound, bound)

    @property
    def flat_params(self) -> torch.Tensor:
        return self._flat

    @property
    def flat_grads(self) -> torch.Tensor:
        self._ensure_grad_buffer()
        return self._flat_grad

    def get_stage_dims(self, stages: Tuple[int, ...]) -> Tuple[int, ...]:
        idx = [s - 1 for s in stages]
        assert min(idx) >= 0 and max(idx) < 4, idx
        return tuple(self.stage_dims[i] for i in idx)

    def get_strides(self, stages: Tuple[int, ...]) -> Tuple[int, ...]:
        idx = [s - 1 for s in stages]
        assert min(idx) >= 0 and max(idx) < 4, idx
        return tuple(self.strides[i] for i in idx)

    # ------------------------------------------------------------------ device handle
    def _ensure_handle(self):
        dev = self._flat.device
        if dev.type != 'cuda':
            raise RuntimeError('leod_b200 backbone runs on CUDA only (no CPU fallback): move the module to a B200 first')
        if self._handle is not None and self._handle_device == dev:
            return
        self._destroy_handle()
        l = _lib.lib()
        h = ctypes.c_void_p()
        cfg = self._make_cfg()
        with torch.cuda.device(dev):
            _lib.check(l.leod_backbone_create(ctypes.byref(cfg), ctypes.byref(h)), 'backbone_create')
        self._handle, self._handle_device = h, dev
        self._anchor = torch.zeros(1, device=dev, requires_grad=True)
        self._bind()

    def _destroy_handle(self):
        if self._handle is not None:
            _lib.lib().leod_backbone_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._destroy_handle()
        except Exception:
            pass

    def _ensure_grad_buffer(self):
        if self._flat_grad is None or self._flat_grad.device != self._flat.device:
            self._flat_grad = torch.zeros_like(self._flat)
            if self._handle is not None:
                self._bind()

    def _bind(self):
        g = self._flat_grad
        _lib.check(_lib.lib().leod_backbone_bind(self._handle, _lib.ptr(self._flat), _lib.ptr(g)), 'backbone_bind')
        self._prepared_version = None

    def set_gemm_impl(self, impl: int):
        self._ensure_handle()
        _lib.check(_lib.lib().leod_backbone_set_gemm_impl(self._handle, int(impl)), 'set_gemm_impl')

    def _params_version(self):
        return sum(p._version for p, *_ in self._param_views)

    def prepare(self, force: bool = False):
        """Refresh the operand-typed weight copies if any parameter changed since the last call."""
        self._ensure_handle()
        v = self._params_version()
        if force or v != self._prepared_version:
            with torch.cuda.device(self._handle_device):
                _lib.check(_lib.lib().leod_backbone_prepare(self._handle, _lib.stream_ptr(self._handle_device)), 'prepare')
            self._prepared_version = v

    def mark_params_updated(self):
        """For optimizers that write `flat_params` directly (bypassing the Parameter objects)."""
        self._prepared_version = None

    # ------------------------------------------------------------------ one timestep
    def _nhwc(self, t: Optional[torch.Tensor], like_shape) -> Optional[torch.Tensor]:
        """[B,C,h,w]-shaped tensor (any strides/dtype) -> contiguous channels-last storage, compute dtype."""
        if t is None:
            return None
        assert tuple(t.shape) == tuple(like_shape), (t.shape, like_shape)
        u = t.permute(0, 2, 3, 1)
        if u.dtype != self.compute_dtype:
            u = u.to(self.compute_dtype)
        return u if u.is_contiguous() else u.contiguous()

    def _state_shapes(self, B):
        H, W = self.in_res_hw
        return [(B, self.stage_dims[s], H // self.strides[s], W // self.strides[s]) for s in range(4)]

    def _step_forward(self, x, states, need_grad):
        l = _lib.lib()
        B = x.shape[0]
        dev = x.device
        shapes = self._state_shapes(B)
        assert x.dim() == 4 and x.shape[1] == self.in_channels, x.shape
        if x.dtype not in (torch.float32, torch.bfloat16, torch.uint8):
            x = x.float()
        x = x.contiguous()
        st = [self._nhwc(states[i], shapes[i // 2]) for i in range(8)]
        h_prev, c_prev = [st[2 * s] for s in range(4)], [st[2 * s + 1] for s in range(4)]
        outs_nhwc = [torch.empty((shp[0], shp[2], shp[3], shp[1]), dtype=self.compute_dtype, device=dev)
                     for shp in shapes for _ in range(2)]
        h_out, c_out = [outs_nhwc[2 * s] for s in range(4)], [outs_nhwc[2 * s + 1] for s in range(4)]
        save = None
        if need_grad:
            nbytes = int(l.leod_backbone_save_bytes(self._handle, B))
            save = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(l.leod_backbone_step_fwd(self._handle, _lib.ptr(x), _lib.leod_dtype(x.dtype), x.shape[2], x.shape[3], B,
                                                _lib.vp4(h_prev), _lib.vp4(c_prev), _lib.vp4(h_out), _lib.vp4(c_out),
                                                _lib.ptr(save), _lib.stream_ptr(dev)), 'backbone_step_fwd')
        outs = [t.permute(0, 3, 1, 2) for t in outs_nhwc]
        if need_grad:
            # keep the tensors the kernels read alive and in the exact layout they were passed in
            save = (save, x, h_prev, c_prev)
        return outs, save

    def _step_backward(self, x_unused, states, outs, save, grads):
        l = _lib.lib()
        save_buf, x, h_prev, c_prev = save
        B = x.shape[0]
        dev = x.device
        shapes = self._state_shapes(B)
        self._begin_backward_pass()
        h_out = [outs[2 * s].permute(0, 2, 3, 1) for s in range(4)]
        c_out = [outs[2 * s + 1].permute(0, 2, 3, 1) for s in range(4)]
        g = [self._nhwc(grads[i], shapes[i // 2]) for i in range(8)]
        dh_out, dc_out = [g[2 * s] for s in range(4)], [g[2 * s + 1] for s in range(4)]
        dprev = [torch.empty((shp[0], shp[2], shp[3], shp[1]), dtype=self.compute_dtype, device=dev)
                 for shp in shapes for _ in range(2)]
        dh_prev = [dprev[2 * s] if h_prev[s] is not None else None for s in range(4)]
        dc_prev = [dprev[2 * s + 1] for s in range(4)]
        with torch.cuda.device(dev):
            _lib.check(l.leod_backbone_step_bwd(self._handle, _lib.ptr(x), _lib.leod_dtype(x.dtype), x.shape[2], x.shape[3], B,
                                                _lib.vp4(h_prev), _lib.vp4(c_prev), _lib.vp4(h_out), _lib.vp4(c_out),
                                                _lib.ptr(save_buf), _lib.vp4(dh_out), _lib.vp4(dc_out), _lib.vp4(dh_prev),
                                                _lib.vp4(dc_prev), _lib.stream_ptr(dev)), 'backbone_step_bwd')
        res = []
        for s in range(4):
            res.append(dprev[2 * s].permute(0, 3, 1, 2) if h_prev[s] is not None else None)
            res.append(dprev[2 * s + 1].permute(0, 3, 1, 2) if c_prev[s] is not None else None)
        return res

    # ------------------------------------------------------------------ whole-window fast path
    def _seq_forward(self, x, states):
        l = _lib.lib()
        L, B = x.shape[0], x.shape[1]
        dev = x.device
        shapes = self._state_shapes(B)
        self._seq_gen += 1
        if x.dtype not in (torch.float32, torch.bfloat16, torch.uint8):
            x = x.float()
        x = x.contiguous()
        st = [self._nhwc(states[i], shapes[i // 2]) for i in range(8)]
        h0, c0 = [st[2 * s] for s in range(4)], [st[2 * s + 1] for s in range(4)]
        h_all = [torch.empty((L, shp[0], shp[2], shp[3], shp[1]), dtype=self.compute_dtype, device=dev) for shp in shapes]
        c_last = [torch.empty((shp[0], shp[2], shp[3], shp[1]), dtype=self.compute_dtype, device=dev) for shp in shapes]
        with torch.cuda.device(dev):
            _lib.check(l.leod_backbone_seq_fwd(self._handle, _lib.ptr(x), _lib.leod_dtype(x.dtype), x.shape[3], x.shape[4], B, L,
                                               _lib.vp4(h0), _lib.vp4(c0), _lib.vp4(h_all), _lib.vp4(c_last),
                                               _lib.stream_ptr(dev)), 'backbone_seq_fwd')
        outs = [t.permute(0, 1, 4, 2, 3) for t in h_all] + [t.permute(0, 3, 1, 2) for t in c_last]
        return outs, (x, h0, c0)

    def _seq_backward(self, keep, h_all_views, grads):
        l = _lib.lib()
        x, h0, c0 = keep
        L, B = x.shape[0], x.shape[1]
        dev = x.device
        shapes = self._state_shapes(B)
        self._begin_backward_pass()
        h_all = [v.permute(0, 1, 3, 4, 2) for v in h_all_views]
        dh_all = []
        for s in range(4):
            g = grads[s]
            if g is not None:
                g = g.permute(0, 1, 3, 4, 2)
                if g.dtype != self.compute_dtype:
                    g = g.to(self.compute_dtype)
                g = g if g.is_contiguous() else g.contiguous()
            dh_all.append(g)
        dc_last = [self._nhwc(grads[4 + s], shapes[s]) for s in range(4)]
        dh0 = [torch.empty_like(h0[s]) if h0[s] is not None else None for s in range(4)]
        dc0 = [torch.empty_like(c0[s]) if c0[s] is not None else None for s in range(4)]
        with torch.cuda.device(dev):
            _lib.check(l.leod_backbone_seq_bwd(self._handle, _lib.ptr(x), _lib.leod_dtype(x.dtype), x.shape[3], x.shape[4], B, L,
                                               _lib.vp4(h0), _lib.vp4(c0), _lib.vp4(h_all), _lib.vp4(dh_all), _lib.vp4(dc_last),
                                               _lib.vp4(dh0), _lib.vp4(dc0), _lib.stream_ptr(dev)), 'backbone_seq_bwd')
        res = []
        for s in range(4):
            res.append(dh0[s].permute(0, 3, 1, 2) if dh0[s] is not None else None)
            res.append(dc0[s].permute(0, 3, 1, 2) if dc0[s] is not None else None)
        return res

    def forward_sequence(self, x: torch.Tensor, prev_states: Optional[List[Optional[LstmState]]] = None) \
            -> Tuple[Dict[int, torch.Tensor], List[LstmState]]:
        """All L timesteps of a batch in one call (the loop of modules/detection.py:188-224 moved into the
        library).  x: [L,B,C,H,W] (uint8 / float).  Returns ({stage: [L,B,C,h,w] features of every timestep},
        [(h_L, c_L)]*4).  Numerically identical to calling `forward` L times; one backward per forward."""
        assert x.dim() == 5 and x.shape[2] == self.in_channels, x.shape
        self.prepare()
        if prev_states is None:
            prev_states = [None] * 4
        flat_states = []
        for st in prev_states:
            flat_states.extend((None, None) if st is None else (st[0], st[1]))
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p, *_ in self._param_views)
        if need_grad:
            self._ensure_grad_buffer()
            outs = _BackboneSeq.apply(self, self._anchor, x, *flat_states)
        else:
            outs, _ = self._seq_forward(x, flat_states)
        feats = {s + 1: outs[s] for s in range(4)}
        states = [(outs[s][-1], outs[4 + s]) for s in range(4)]
        return feats, states

    # ------------------------------------------------------------------ gradient hand-over
    def _begin_backward_pass(self):
        if self._bwd_pending:
            return
        self._bwd_pending = True
        self._ensure_grad_buffer()
        # zero_grad(set_to_none=True) drops the views: start this pass from zero in that case
        if self._param_views[0][0].grad is None:
            self._flat_grad.zero_()
        torch.autograd.Variable._execution_engine.queue_callback(self._end_backward_pass)

    def _end_backward_pass(self):
        self._bwd_pending = False
        dev = self._handle_device
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().leod_backbone_grads_finalize(self._handle, _lib.stream_ptr(dev)), 'grads_finalize')
        if self.grad_sync is not None:
            self.grad_sync(self._flat_grad)
        for p, off, numel, shape in self._param_views:
            if p.grad is None:
                p.grad = self._flat_grad[off:off + numel].view(shape)

    # ------------------------------------------------------------------ reference interface
    def forward(self, x: torch.Tensor, prev_states: Optional[List[Optional[LstmState]]] = None,
                token_mask: Optional[torch.Tensor] = None) -> Tuple[Dict[int, torch.Tensor], List[LstmState]]:
        """maxvit_rnn.py:97-115.  Returns ({1..4: [B,C,h,w]}, [(h,c)]*4)."""
        if token_mask is not None:
            raise NotImplementedError('token_mask is unused by every shipped config (enable_masking: False)')
        self.prepare()
        if prev_states is None:
            prev_states = [None] * 4
        assert len(prev_states) == 4
        flat_states = []
        for st in prev_states:
            flat_states.extend((None, None) if st is None else (st[0], st[1]))
        need_grad = torch.is_grad_enabled() and (self._flat.requires_grad or any(p.requires_grad for p, *_ in self._param_views))
        if need_grad:
            self._ensure_grad_buffer()
            outs = _BackboneStep.apply(self, self._anchor, x, *flat_states)
        else:
            outs, _ = self._step_forward(x, flat_states, need_grad=False)
        states = [(outs[2 * s], outs[2 * s + 1]) for s in range(4)]
        feats = {s + 1: outs[2 * s] for s in range(4)}
        return feats, states
