"""Oracle: teacher EMA rule and the optimizer step, torch fp32 on CPU.

Restates modules/utils/ssod.py:429-438 (ema_model_update: alpha = min(1 - 1/(step+1), alpha);
ema = ema*alpha + p*(1-alpha), parameters only — BN buffers are not averaged) and the update the
reference's optimizer performs (modules/detection.py:485-518: torch.optim.AdamW, weight_decay 0;
train.py:236-237: gradient clip-by-VALUE 1.0 before the step).
TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
import math
from typing import Sequence

import torch


@torch.no_grad()
def ema_update(ema_params: Sequence[torch.Tensor], params: Sequence[torch.Tensor], global_step: int, alpha: float = 0.999):
    a = min(1. - 1. / (global_step + 1.), alpha)
    for e, p in zip(ema_params, params):
        e.mul_(a).add_(p, alpha=1. - a)


@torch.no_grad()
def adamw_step(p, g, m, v, step: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, clip_value=None):
    """One AdamW step on one tensor, torch.optim.AdamW (non-amsgrad) semantics; `step` is 1-based."""
    if clip_value is not None:
        g = g.clamp(-clip_value, clip_value)
    p.mul_(1 - lr * weight_decay)
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)
