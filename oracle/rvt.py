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
