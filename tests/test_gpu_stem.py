"""The implicit-GEMM stem (leod_stem_conv_fwd / leod_stem_conv_wgrad: maxvit.py:143-182 on the uint8 event tensor, no patch matrix)
against torch's conv2d in fp64 on the same inputs.  Inputs are exact in the kernel's arithmetic (uint8 counts; weights and gradients
rounded to bf16 first), products are accumulated in fp32 on the tensor cores: forward <= 1e-2 of the largest output after the bf16
store (observed 2e-3), weight gradient <= 2e-3 (fp32 accumulation order only)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_err

pytestmark = pytest.mark.gpu


def _patch_order(W):
    """[C, Cin, 7, 7] -> [C, Cin*56]: k = (cin*7 + ky)*8 + slot, slot 0 = 0, slot 1+kx = W[..., ky, kx]"""
    C, Cin = W.shape[:2]
    Wp = torch.zeros(C, Cin, 7, 8, dtype=W.dtype, device=W.device)
    Wp[..., 1:] = W
    return Wp.reshape(C, Cin * 56).contiguous()


def _case(nimg, Cin, xh, xw, Ho, Wo, C, seed):
    g = torch.Generator(device='cuda').manual_seed(seed)
    x = (torch.rand(nimg, Cin, xh, xw, device='cuda', generator=g) < 0.15).to(torch.uint8) * \
        torch.randint(1, 256, (nimg, Cin, xh, xw), device='cuda', generator=g, dtype=torch.int32).to(torch.uint8)
    W = (torch.randn(C, Cin, 7, 7, device='cuda', generator=g) * 0.05).bfloat16()
    dY = (torch.randn(nimg * Ho * Wo, C, device='cuda', generator=g) * 1e-3).bfloat16()
    return x, W, dY


SHAPES = [
    # nimg, Cin, xh, xw, Ho, Wo, C
    (3, 20, 240, 304, 64, 80, 48),     # RVT-S Gen1: frame smaller than the padded 256 x 320 input
    (2, 20, 256, 320, 64, 80, 64),     # RVT-B width
    (1, 10, 240, 304, 64, 80, 32),     # RVT-T, BASELINE configs[0]: 10 input channels
    (2, 20, 384, 640, 96, 160, 64),    # Gen4 at half resolution
    (5, 3, 32, 64, 8, 16, 16),         # a single tile per image, smallest channel count
]


@pytest.mark.parametrize('shape', SHAPES)
def test_stem_conv_forward_matches_conv2d(shape):
    from leod_b200 import _lib as L
    nimg, Cin, xh, xw, Ho, Wo, C = shape
    x, W, _ = _case(*shape, seed=1)
    Wp = _patch_order(W).half()                      # the library's own conversion: exact unless |w| < 6.1e-5 (half subnormals: abs error <= 3e-8)
    assert float((Wp.float() - _patch_order(W).float()).abs().max()) <= 3e-8
    y = torch.full((nimg * Ho * Wo, C), float('nan'), dtype=torch.bfloat16, device='cuda')
    L.check(L.lib().leod_stem_conv_fwd(L.ptr(x), nimg, Cin, xh, xw, Ho, Wo, C, L.ptr(Wp), Wp.stride(0), L.ptr(y), L.stream_ptr()), 'stem fwd')
    torch.cuda.synchronize()
    xp = F.pad(x.double(), (0, Wo * 4 - xw, 0, Ho * 4 - xh))
    ref = F.conv2d(xp, W.double(), stride=4, padding=3).permute(0, 2, 3, 1).reshape(nimg * Ho * Wo, C)
    assert torch.isfinite(y.float()).all()
    e = rel_err(y.double().cpu(), ref.cpu())
    assert e < 1e-2, (shape, e)


@pytest.mark.parametrize('shape', SHAPES)
def test_stem_conv_weight_gradient_matches_autograd(shape):
    from leod_b200 import _lib as L
    nimg, Cin, xh, xw, Ho, Wo, C = shape
    x, W, dY = _case(*shape, seed=2)
    ldw = Cin * 56
    dW = torch.zeros(C, ldw, dtype=torch.float32, device='cuda')
    for _ in range(2):                               # the kernel accumulates: two calls = twice the gradient
        L.check(L.lib().leod_stem_conv_wgrad(L.ptr(x), nimg, Cin, xh, xw, Ho, Wo, C, L.ptr(dY), L.ptr(dW), ldw, L.stream_ptr()), 'stem wgrad')
    torch.cuda.synchronize()
    xp = F.pad(x.double(), (0, Wo * 4 - xw, 0, Ho * 4 - xh))
    Wd = W.double().requires_grad_(True)
    out = F.conv2d(xp, Wd, stride=4, padding=3)
    out.backward(dY.double().reshape(nimg, Ho, Wo, C).permute(0, 3, 1, 2))
    # slot 0 of every (cin, ky) group is the column x-4: not a tap of the 7x7 kernel (its weight is zero), the kernel accumulates a
    # value there that callers drop (backbone.cu: stem_grad_unpermute_kernel); slots 1..7 are the gradient
    got = dW.view(C, Cin, 7, 8)[..., 1:]
    e = rel_err(got.double().cpu(), (2 * Wd.grad).cpu())
    assert e < 2e-3, (shape, e)


def test_stem_conv_rejects_unsupported_geometry():
    from leod_b200 import _lib as L
    x = torch.zeros(1, 20, 240, 300, dtype=torch.uint8, device='cuda')      # width not a multiple of 16
    W = torch.zeros(48, 1120, dtype=torch.float16, device='cuda')
    y = torch.zeros(64 * 80, 48, dtype=torch.bfloat16, device='cuda')
    rc = L.lib().leod_stem_conv_fwd(L.ptr(x), 1, 20, 240, 300, 64, 80, 48, L.ptr(W), 1120, L.ptr(y), L.stream_ptr())
    assert rc != 0 and b'unsupported geometry' in L.lib().leod_last_error()
