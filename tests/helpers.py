"""Shared helpers for the test-suite (fixtures loading, synthetic inputs)."""
import os

import numpy as np
import torch

from oracle.config import ModelCfg

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_net_fixture():
    z = np.load(os.path.join(GOLDEN, 'net_small.npz'))
    embed, dh, p0, p1, ncls, inch, H, W, B, T = (int(v) for v in z['meta'])
    cfg = ModelCfg(input_channels=inch, embed_dim=embed, dim_head=dh, partition_size=(p0, p1), num_classes=ncls,
                   fpn_depth=float(z['fpn_depth']))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('sd/')}
    return z, cfg, sd, dict(H=H, W=W, B=B, T=T)


def rel_err(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def det_state_value(name: str, shape, dtype=torch.float32) -> torch.Tensor:
    """Deterministic O(1) value for a state_dict entry, seeded by its NAME: the reference model (make_golden.py) and the
    models under test are filled with identical weights without storing multi-megabyte state dicts."""
    import zlib
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7fffffff)
    shape = tuple(shape)
    if name.endswith('num_batches_tracked'):
        return torch.zeros(shape, dtype=torch.long)
    if name.endswith('running_var'):
        return torch.rand(shape, generator=g) + 0.5
    if name.endswith('gamma'):
        return torch.rand(shape, generator=g) * 0.5 + 0.25
    if name.endswith('.norm.weight') or name.endswith('norm1.weight') or name.endswith('norm2.weight') or name.endswith('bn.weight'):
        return torch.rand(shape, generator=g) + 0.5
    if len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return torch.randn(shape, generator=g) * (1.0 / fan_in ** 0.5)
    return torch.randn(shape, generator=g) * 0.2


def det_fill_(model) -> None:
    with torch.no_grad():
        for k, v in model.state_dict().items():
            v.copy_(det_state_value(k, v.shape).to(v.dtype))


def det_events(seed: int, shape, density=0.1, vmax=6) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return ((torch.rand(shape, generator=g) < density).float() * torch.randint(1, vmax, shape, generator=g)).to(torch.uint8)


FULLSIZE_CASES = {   # tag -> (size, dataset, input_channels, B, L): BASELINE configs[0] (plumbing case) and the configs[1] model
    'tiny_gen1_c10': ('tiny', 'gen1', 10, 1, 1),
    'small_gen1': ('small', 'gen1', 20, 1, 2),
}


def canon_rows(d):
    """Detections [N,7] in a canonical order: descending score (obj * cls_conf), ties broken by the remaining columns."""
    d = np.asarray(d, np.float32)
    if d.shape[0] == 0:
        return d
    score = d[:, 4] * d[:, 5]
    order = np.lexsort((d[:, 3], d[:, 2], d[:, 1], d[:, 0], d[:, 6], -score))
    return d[order]
