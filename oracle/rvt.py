"""Oracle: RVT recurrent MaxViT backbone, functional over a reference-layout state_dict.

Restates (paths relative to /root/reference):
  models/detection/recurrent_backbone/maxvit_rnn.py:97-115  (backbone: 4 stages, feats + states)
  models/detection/recurrent_backbone/maxvit_rnn.py:182-201 (stage: downsample, attention pair, lstm)
  models/layers/maxvit/maxvit.py:174-178                    (strided conv -> channels-last -> LN)
  models/layers/maxvit/maxvit.py:252-270                    (pre-norm attention + MLP residual block)
  models/layers/maxvit/maxvit.py:273-304                    (window / grid token grouping)
  models/layers/maxvit/maxvit.py:343-354                    (multi-head self attention)
  models/layers/maxvit/maxvit.py:85-118, 45-53              (MLP, LayerScale)
  models/layers/rnn.py:37-70                                (1x1-conv LSTM cell)
TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from .config import ModelCfg

State = Tuple[torch.Tensor, torch.Tensor]


def _ln(x, sd, prefix, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + '.weight'], sd[prefix + '.bias'], eps)


def group_tokens(x: torch.Tensor, part: Tuple[int, int], window: bool) -> torch.Tensor:
    """[B,H,W,C] -> [B*groups, ph*pw, C].  window: contiguous ph x pw tiles (maxvit.py:273-278);
    grid: token (i,j) of group (p,q) is pixel (i*H/ph + p, j*W/pw + q) (maxvit.py:290-295)."""
    B, H, W, C = x.shape
    ph, pw = part
    if window:
        t = x.reshape(B, H // ph, ph, W // pw, pw, C).permute(0, 1, 3, 2, 4, 5)
    else:
        t = x.reshape(B, ph, H // ph, pw, W // pw, C).permute(0, 2, 4, 1, 3, 5)
    return t.reshape(-1, ph * pw, C)


def ungroup_tokens(t: torch.Tensor, part: Tuple[int, int], hw: Tuple[int, int], window: bool) -> torch.Tensor:
    """Inverse of group_tokens (maxvit.py:281-287, 298-304)."""
    H, W = hw
    ph, pw = part
    C = t.shape[-1]
    if window:
        x = t.reshape(-1, H // ph, W // pw, ph, pw, C).permute(0, 1, 3, 2, 4, 5)
    else:
        x = t.reshape(-1, H // ph, W // pw, ph, pw, C).permute(0, 3, 1, 4, 2, 5)
    return x.reshape(-1, H, W, C)


def self_attention(t: torch.Tensor, sd, prefix: str, dim_head: int) -> torch.Tensor:
    """maxvit.py:343-354.  qkv columns are laid out per head as [q | k | v] (chunk on the last dim
    after viewing [.., heads, 3*dh])."""
    G, T, C = t.shape
    nh = C // dim_head
    qkv = F.linear(t, sd[prefix + '.qkv.weight'], sd[prefix + '.qkv.bias'])
    qkv = qkv.reshape(G, T, nh, 3, dim_head)
    q, k, v = (qkv[:, :, :, i].transpose(1, 2) for i in range(3))  # [G, nh, T, dh]
    att = torch.softmax((q @ k.transpose(-2, -1)) * dim_head ** -0.5, dim=-1)
    o = (att @ v).transpose(1, 2).reshape(G, T, C)
    return F.linear(o, sd[prefix + '.proj.weight'], sd[prefix + '.proj.bias'])


def attention_block(x: torch.Tensor, sd, prefix: str, cfg: ModelCfg, window: bool) -> torch.Tensor:
    """maxvit.py:267-270; norm1 is absent for the first block of a stage (maxvit_rnn.py:165)."""
    B, H, W, C = x.shape
    y = _ln(x, sd, prefix + '.norm1', cfg.norm_eps) if (prefix + '.norm1.weight') in sd else x
    t = group_tokens(y, cfg.partition_size, window)
    t = self_attention(t, sd, prefix + '.self_attn', cfg.dim_head)
    y = ungroup_tokens(t, cfg.partition_size, (H, W), window)
    x = x + y * sd[prefix + '.ls1.gamma']
    y = _ln(x, sd, prefix + '.norm2', cfg.norm_eps)
    y = F.linear(y, sd[prefix + '.mlp.net.0.0.weight'], sd[prefix + '.mlp.net.0.0.bias'])
    y = F.gelu(y)  # exact erf (layers/activations.py:138-145)
    y = F.linear(y, sd[prefix + '.mlp.net.2.weight'], sd[prefix + '.mlp.net.2.bias'])
    return x + y * sd[prefix + '.ls2.gamma']


def lstm_cell(x: torch.Tensor, state: Optional[State], sd, prefix: str) -> State:
    """models/layers/rnn.py:37-70 with dws_conv=False: gates = conv1x1(cat(x,h)); order f,i,o | g."""
    if state is None:
        state = (torch.zeros_like(x), torch.zeros_like(x))
    h0, c0 = state
    C = x.shape[1]
    mix = F.conv2d(torch.cat((x, h0), 1), sd[prefix + '.conv1x1.weight'], sd[prefix + '.conv1x1.bias'])
    f, i, o = (torch.sigmoid(mix[:, k * C:(k + 1) * C]) for k in range(3))
    g = torch.tanh(mix[:, 3 * C:])
    c1 = f * c0 + i * g
    h1 = o * torch.tanh(c1)
    return h1, c1


def stage_forward(x: torch.Tensor, state: Optional[State], sd, si: int, cfg: ModelCfg) -> State:
    """maxvit_rnn.py:182-201 (token masking disabled)."""
    p = f'backbone.stages.{si}'
    w = sd[p + '.downsample_cf2cl.conv.weight']
    k = w.shape[-1]
    stride = 4 if si == 0 else 2
    y = F.conv2d(x, w, None, stride=stride, padding=k // 2).permute(0, 2, 3, 1)
    y = _ln(y, sd, p + '.downsample_cf2cl.norm', 1e-5)
    y = attention_block(y, sd, p + '.att_blocks.0.att_window', cfg, True)
    y = attention_block(y, sd, p + '.att_blocks.0.att_grid', cfg, False)
    y = y.permute(0, 3, 1, 2).contiguous()
    return lstm_cell(y, state, sd, p + '.lstm')


def backbone_forward(x: torch.Tensor, states: Optional[List[Optional[State]]], sd, cfg: ModelCfg) \
        -> Tuple[Dict[int, torch.Tensor], List[State]]:
    """maxvit_rnn.py:97-115: returns ({1..4: h_t NCHW}, [(h,c)]*4)."""
    if states is None:
        states = [None] * 4
    feats, out_states = {}, []
    for si in range(4):
        h, c = stage_forward(x, states[si], sd, si, cfg)
        out_states.append((h, c))
        feats[si + 1] = h
        x = h
    return feats, out_states


def pad_input(ev: torch.Tensor, hw: Tuple[int, int]) -> torch.Tensor:
    """utils/padding.py:33-58 + modules/detection.py:132: cast to float, zero-pad bottom/right."""
    H, W = ev.shape[-2:]
    return F.pad(ev.float(), (0, hw[1] - W, 0, hw[0] - H))


def init_state_dict(cfg: ModelCfg, input_channels: int = 20, seed: int = 0) -> Dict[str, torch.Tensor]:
    """A state_dict with the reference's parameter names / shapes and its constructors' distributions (nn.Conv2d / nn.Linear default
    init, LayerNorm and BatchNorm (1, 0), LayerScale 1e-5 — maxvit.py:45-53 —, prior-probability head biases — yolo_head.py:183-193),
    for the oracle-port arm of bench.py when oracle/_ref is absent.  Shapes: maxvit_rnn.py:23-95, maxvit.py:143-182, 185-354,
    models/layers/rnn.py:7-36, yolo_pafpn.py:18-107, network_blocks.py:29-142, yolo_head.py:21-182."""
    import math
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def uni(shape, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * b

    def linear(p, cout, cin, bias=True):
        sd[p + '.weight'] = uni((cout, cin), cin)
        if bias:
            sd[p + '.bias'] = uni((cout,), cin)

    def norm(p, c):
        sd[p + '.weight'], sd[p + '.bias'] = torch.ones(c), torch.zeros(c)

    dims = cfg.stage_dims
    cin = input_channels
    for si, c in enumerate(dims):
        p = f'backbone.stages.{si}'
        k = 7 if si == 0 else 3
        sd[p + '.downsample_cf2cl.conv.weight'] = uni((c, cin, k, k), cin * k * k)
        norm(p + '.downsample_cf2cl.norm', c)
        for bi, blk in enumerate(('att_window', 'att_grid')):
            q = f'{p}.att_blocks.0.{blk}'
            if bi > 0:
                norm(q + '.norm1', c)
            linear(q + '.self_attn.qkv', 3 * c, c)
            linear(q + '.self_attn.proj', c, c)
            sd[q + '.ls1.gamma'] = torch.full((c,), 1e-5)
            norm(q + '.norm2', c)
            linear(q + '.mlp.net.0.0', cfg.mlp_ratio * c, c)
            linear(q + '.mlp.net.2', c, cfg.mlp_ratio * c)
            sd[q + '.ls2.gamma'] = torch.full((c,), 1e-5)
        sd[p + '.lstm.conv1x1.weight'] = uni((4 * c, 2 * c, 1, 1), 2 * c)
        sd[p + '.lstm.conv1x1.bias'] = uni((4 * c,), 2 * c)
        cin = c

    def cbs(p, co, ci, k):
        sd[p + '.conv.weight'] = uni((co, ci, k, k), ci * k * k)
        sd[p + '.bn.weight'], sd[p + '.bn.bias'] = torch.ones(co), torch.zeros(co)
        sd[p + '.bn.running_mean'], sd[p + '.bn.running_var'] = torch.zeros(co), torch.ones(co)
        sd[p + '.bn.num_batches_tracked'] = torch.zeros((), dtype=torch.long)

    def csp(p, ci, co, n):
        h = co // 2
        cbs(p + '.conv1', h, ci, 1)
        cbs(p + '.conv2', h, ci, 1)
        cbs(p + '.conv3', co, 2 * h, 1)
        for i in range(n):
            cbs(f'{p}.m.{i}.conv1', h, h, 1)
            cbs(f'{p}.m.{i}.conv2', h, h, 3)

    c0, c1, c2 = (dims[s - 1] for s in cfg.in_stages)
    n = round(3 * cfg.fpn_depth)
    cbs('fpn.lateral_conv0', c1, c2, 1)
    csp('fpn.C3_p4', 2 * c1, c1, n)
    cbs('fpn.reduce_conv1', c0, c1, 1)
    csp('fpn.C3_p3', 2 * c0, c0, n)
    cbs('fpn.bu_conv2', c0, c0, 3)
    csp('fpn.C3_n3', 2 * c0, c1, n)
    cbs('fpn.bu_conv1', c1, c1, 3)
    csp('fpn.C3_n4', 2 * c1, c2, n)
    hid = int(256 * c2 / 1024)
    prior = -math.log((1 - 0.01) / 0.01)
    for k, c in enumerate((c0, c1, c2)):
        cbs(f'yolox_head.stems.{k}', hid, c, 1)
        for tower in ('cls_convs', 'reg_convs'):
            cbs(f'yolox_head.{tower}.{k}.0', hid, hid, 3)
            cbs(f'yolox_head.{tower}.{k}.1', hid, hid, 3)
        for name, co in (('cls_preds', cfg.num_classes), ('reg_preds', 4), ('obj_preds', 1)):
            sd[f'yolox_head.{name}.{k}.weight'] = uni((co, hid, 1, 1), hid)
            sd[f'yolox_head.{name}.{k}.bias'] = uni((co,), hid) if name == 'reg_preds' else torch.full((co,), prior)
    return sd
