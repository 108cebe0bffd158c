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
