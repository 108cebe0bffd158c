"""Times leod_stem_conv_fwd / leod_stem_conv_wgrad alone at the bench shape (168 frames of 20 x 240 x 304 uint8 -> 64 x 80 x 48),
CUDA events on the launching stream, rotating inputs (4 x 245 MB > L2).  LEOD_STEM_DEBUG bits are honoured (see kernels_stem.cu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from leod_b200 import _lib as L

nimg, Cin, xh, xw, Ho, Wo, C = 168, 20, 240, 304, 64, 80, 48
xs = [(torch.rand(nimg, Cin, xh, xw, device='cuda') < 0.1).to(torch.uint8) * 3 for _ in range(4)]
W = (torch.randn(C, Cin * 56, device='cuda') * 0.05).half()
y = torch.empty(nimg * Ho * Wo, C, dtype=torch.bfloat16, device='cuda')
dY = torch.randn(nimg * Ho * Wo, C, device='cuda').bfloat16()
dW = torch.zeros(C, Cin * 56, device='cuda')
lib, st = L.lib(), L.stream_ptr()


def timed(fn, n=20):
    for i in range(3):
        fn(xs[i % 4])
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn(xs[i % 4])
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
    return ts[n // 2] * 1e3


f = timed(lambda x: L.check(lib.leod_stem_conv_fwd(L.ptr(x), nimg, Cin, xh, xw, Ho, Wo, C, L.ptr(W), W.stride(0), L.ptr(y), st)))
b = timed(lambda x: L.check(lib.leod_stem_conv_wgrad(L.ptr(x), nimg, Cin, xh, xw, Ho, Wo, C, L.ptr(dY), L.ptr(dW), Cin * 56, st)))
alg_f = nimg * Cin * xh * xw + nimg * Ho * Wo * C * 2
alg_b = nimg * Cin * xh * xw + nimg * Ho * Wo * C * 2
print(f'dbg={os.environ.get("LEOD_STEM_DEBUG", "0")}  stem fwd {f:7.1f} us ({alg_f / f / 1e3:6.0f} GB/s algorithmic)   wgrad {b:7.1f} us ({alg_b / b / 1e3:6.0f} GB/s)')
