"""Parity of the individual CUDA building blocks (called through the C ABI) against plain torch
fp32 references / the numpy oracle.  Tolerances: fp32 path 1e-4 relative-to-max, bf16 path 1e-2
(BASELINE.json north_star: 1e-3 fp32 / 1e-2 bf16)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import GOLDEN, rel_err

pytestmark = pytest.mark.gpu

F32, BF16 = 0, 1


def _lib():
    from leod_b200 import _lib
    return _lib


def gemm_nt(impl, A, B, bias=None, epi=0, R=None, aux=None, A2=None):
    L = _lib()
    M, K1 = A.shape
    K = K1 + (A2.shape[1] if A2 is not None else 0)
    N = B.shape[0]
    dt = BF16 if A.dtype == torch.bfloat16 else F32
    C = torch.empty(M, N, dtype=A.dtype, device=A.device)
    if epi == 1:
        aux = torch.empty(M, N, dtype=A.dtype, device=A.device)
    L.check(L.lib().leod_gemm_nt(impl, dt, L.ptr(A), A.stride(0), L.ptr(A2), A2.stride(0) if A2 is not None else 0, K1,
                                 L.ptr(B), B.stride(0), L.ptr(C), C.stride(0), M, N, K, L.ptr(bias), epi, L.ptr(R),
                                 R.stride(0) if R is not None else 0, L.ptr(aux), aux.stride(0) if aux is not None else 0,
                                 L.stream_ptr()), 'gemm_nt')
    torch.cuda.synchronize()
    return C, aux


def ref_nt(A, B, bias, epi, R, aux, A2=None):
    A = A.float()
    if A2 is not None:
        A = torch.cat((A, A2.float()), 1)
    v = A @ B.float().t()
    if bias is not None:
        v = v + bias
    if epi == 1:     # aux = gelu'(v): the backward epilogue (3) multiplies by it
        x = v.detach().clone().requires_grad_(True)
        g, = torch.autograd.grad(F.gelu(x).sum(), x)
        return F.gelu(v), g
    if epi == 2:
        return v + R.float(), None
    if epi == 3:
        return v * aux.float(), None
    return v, None


SHAPES = [(300, 144, 48), (1024, 192, 48), (2560, 192, 768), (640, 1536, 384), (129, 48, 984), (4096, 96, 432),
          (77, 16, 24), (640, 288, 96)]


@pytest.mark.parametrize('M,N,K', SHAPES)
@pytest.mark.parametrize('epi', [0, 1, 2, 3])
def test_gemm_nt_simt_fp32(M, N, K, epi):
    g = torch.Generator(device='cuda').manual_seed(M + N + K + epi)
    A = torch.randn(M, K, device='cuda', generator=g)
    B = torch.randn(N, K, device='cuda', generator=g) / K ** 0.5
    bias = torch.randn(N, device='cuda', generator=g) if epi != 3 else None
    R = torch.randn(M, N, device='cuda', generator=g) if epi == 2 else None
    aux = torch.randn(M, N, device='cuda', generator=g) if epi == 3 else None
    torch.backends.cuda.matmul.allow_tf32 = False
    C, aux_out = gemm_nt(0, A, B, bias, epi, R, aux)
    ref, ref_aux = ref_nt(A, B, bias, epi, R, aux)
    assert rel_err(C, ref) < 1e-4
    if epi == 1:
        assert rel_err(aux_out, ref_aux) < 1e-4


@pytest.mark.parametrize('M,N,K', SHAPES)
@pytest.mark.parametrize('epi', [0, 1, 2, 3])
def test_gemm_nt_tcgen05_bf16(M, N, K, epi):
    if K % 8:
        pytest.skip('TMA needs 16-byte row pitch')
    g = torch.Generator(device='cuda').manual_seed(M + N + K + epi)
    A = torch.randn(M, K, device='cuda', generator=g).bfloat16()
    B = (torch.randn(N, K, device='cuda', generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device='cuda', generator=g) if epi != 3 else None
    R = torch.randn(M, N, device='cuda', generator=g).bfloat16() if epi == 2 else None
    aux = torch.randn(M, N, device='cuda', generator=g).bfloat16() if epi == 3 else None
    torch.backends.cuda.matmul.allow_tf32 = False
    C, aux_out = gemm_nt(1, A, B, bias, epi, R, aux)
    ref, ref_aux = ref_nt(A, B, bias, epi, R, aux)
    # inputs are identical bf16 values, accumulation is fp32: only the output rounding differs
    assert rel_err(C.float(), ref) < 6e-3, (rel_err(C.float(), ref))
    if epi == 1:
        assert rel_err(aux_out.float(), ref_aux) < 6e-3
    # and the SIMT kernel on the same bf16 operands must agree to output rounding
    C2, _ = gemm_nt(0, A, B, bias, epi, R, aux)
    assert rel_err(C.float(), C2.float()) < 6e-3


@pytest.mark.parametrize('impl,dtype', [(0, torch.float32), (1, torch.bfloat16)])
def test_gemm_nt_two_sources(impl, dtype):
    g = torch.Generator(device='cuda').manual_seed(5)
    for (M, C) in ((640, 384), (1000, 48), (256, 96)):
        x = torch.randn(M, C, device='cuda', generator=g).to(dtype)
        h = torch.randn(M, C, device='cuda', generator=g).to(dtype)
        W = (torch.randn(4 * C, 2 * C, device='cuda', generator=g) / (2 * C) ** 0.5).to(dtype)
        bias = torch.randn(4 * C, device='cuda', generator=g)
        out, _ = gemm_nt(impl, x, W, bias, 0, None, None, A2=h)
        ref, _ = ref_nt(x, W, bias, 0, None, None, A2=h)
        assert rel_err(out.float(), ref) < (1e-4 if dtype == torch.float32 else 6e-3)


@pytest.mark.parametrize('impl,dtype', [(0, torch.float32), (0, torch.bfloat16), (1, torch.bfloat16)])
@pytest.mark.parametrize('M,N,K', [(1000, 144, 48), (40960, 48, 192), (640, 1536, 384), (2560, 96, 864), (333, 20, 980)])
def test_gemm_tn(impl, dtype, M, N, K):
    L = _lib()
    g = torch.Generator(device='cuda').manual_seed(M + N)
    dY = torch.randn(M, N, device='cuda', generator=g).to(dtype)
    ldx = (K + 7) // 8 * 8
    Xfull = torch.randn(M, ldx, device='cuda', generator=g).to(dtype)
    X = Xfull[:, :K]
    dW = torch.full((N, K), 0.5, device='cuda')
    db = torch.full((N,), -1.0, device='cuda')
    L.check(L.lib().leod_gemm_tn(impl, BF16 if dtype == torch.bfloat16 else F32, L.ptr(dY), N, L.ptr(Xfull), ldx, L.ptr(dW), K,
                                 L.ptr(db), M, N, K, L.stream_ptr()), 'gemm_tn')
    torch.cuda.synchronize()
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = 0.5 + dY.float().t() @ X.float()
    refb = -1.0 + dY.float().sum(0)
    assert rel_err(dW, ref) < 2e-4
    assert rel_err(db, refb) < 2e-4


def _attn_ref(qkv, B, H, W, C, dh, part, window):
    from oracle.rvt import group_tokens, ungroup_tokens
    nh = C // dh
    t = group_tokens(qkv.view(B, H, W, 3 * C), part, window)
    G, T, _ = t.shape
    t = t.reshape(G, T, nh, 3, dh)
    q, k, v = (t[:, :, :, i].transpose(1, 2) for i in range(3))
    att = torch.softmax((q @ k.transpose(-2, -1)) * dh ** -0.5, -1)
    o = (att @ v).transpose(1, 2).reshape(G, T, C)
    return ungroup_tokens(o, part, (H, W), window).reshape(B * H * W, C)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('window', [1, 0])
@pytest.mark.parametrize('B,H,W,C,dh,part', [(2, 16, 20, 48, 24, (8, 10)), (1, 12, 20, 64, 32, (6, 10)), (3, 8, 10, 384, 24, (8, 10)),
                                             (2, 4, 6, 16, 4, (2, 3))])
def test_attention_fwd_bwd(dtype, window, B, H, W, C, dh, part):
    L = _lib()
    g = torch.Generator(device='cuda').manual_seed(B * H + C)
    M = B * H * W
    qkv = torch.randn(M, 3 * C, device='cuda', generator=g).to(dtype)
    dout = torch.randn(M, C, device='cuda', generator=g).to(dtype)
    out = torch.empty(M, C, device='cuda', dtype=dtype)
    dqkv = torch.empty(M, 3 * C, device='cuda', dtype=dtype)
    dt = BF16 if dtype == torch.bfloat16 else F32
    L.check(L.lib().leod_attention_fwd(dt, L.ptr(qkv), L.ptr(out), B, H, W, C, dh, part[0], part[1], window, L.stream_ptr()))
    L.check(L.lib().leod_attention_bwd(dt, L.ptr(qkv), L.ptr(dout), L.ptr(dqkv), B, H, W, C, dh, part[0], part[1], window,
                                       L.stream_ptr()))
    torch.cuda.synchronize()
    torch.backends.cuda.matmul.allow_tf32 = False
    q32 = qkv.float().requires_grad_(True)
    ref = _attn_ref(q32, B, H, W, C, dh, part, bool(window))
    ref.backward(dout.float())
    tol = 1e-4 if dtype == torch.float32 else 1e-2
    assert rel_err(out.float(), ref.detach()) < tol
    assert rel_err(dqkv.float(), q32.grad) < tol


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('M,C', [(1000, 48), (4099, 96), (640, 384), (77, 512), (300, 16)])
def test_layernorm_fwd_bwd(dtype, M, C):
    L = _lib()
    g = torch.Generator(device='cuda').manual_seed(M + C)
    x = (torch.randn(M, C, device='cuda', generator=g) * 2 + 0.5).to(dtype)
    dy = torch.randn(M, C, device='cuda', generator=g).to(dtype)
    dres = torch.randn(M, C, device='cuda', generator=g).to(dtype)
    w = torch.rand(C, device='cuda', generator=g) + 0.5
    b = torch.randn(C, device='cuda', generator=g)
    y = torch.empty_like(x)
    dx = torch.empty_like(x)
    dw = torch.full((C,), 0.25, device='cuda')
    db = torch.full((C,), -0.5, device='cuda')
    dt = BF16 if dtype == torch.bfloat16 else F32
    L.check(L.lib().leod_layernorm_fwd(dt, L.ptr(x), L.ptr(w), L.ptr(b), L.ptr(y), M, C, 1e-5, L.stream_ptr()))
    L.check(L.lib().leod_layernorm_bwd(dt, L.ptr(x), L.ptr(w), L.ptr(dy), L.ptr(dres), L.ptr(dx), L.ptr(dw), L.ptr(db), M, C, 1e-5,
                                       L.stream_ptr()))
    torch.cuda.synchronize()
    x32 = x.float().requires_grad_(True)
    w32 = w.clone().requires_grad_(True)
    b32 = b.clone().requires_grad_(True)
    ref = F.layer_norm(x32, (C,), w32, b32, 1e-5)
    ref.backward(dy.float())
    tol = 1e-4 if dtype == torch.float32 else 1e-2
    assert rel_err(y.float(), ref.detach()) < tol
    assert rel_err(dx.float(), x32.grad + dres.float()) < tol
    assert rel_err(dw, 0.25 + w32.grad) < 1e-3
    assert rel_err(db, -0.5 + b32.grad) < 1e-3


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('M,C,with_prev', [(1000, 48, True), (640, 384, True), (333, 96, False)])
def test_lstm_gates_fwd_bwd(dtype, M, C, with_prev):
    L = _lib()
    g = torch.Generator(device='cuda').manual_seed(M + C)
    pre = torch.randn(M, 4 * C, device='cuda', generator=g).to(dtype)
    c0 = torch.randn(M, C, device='cuda', generator=g).to(dtype) if with_prev else None
    dh = torch.randn(M, C, device='cuda', generator=g).to(dtype)
    dh2 = torch.randn(M, C, device='cuda', generator=g).to(dtype)
    dc = torch.randn(M, C, device='cuda', generator=g).to(dtype)
    gates = pre.clone()
    h1, c1 = torch.empty(M, C, device='cuda', dtype=dtype), torch.empty(M, C, device='cuda', dtype=dtype)
    dgates = torch.empty(M, 4 * C, device='cuda', dtype=dtype)
    dc0 = torch.empty(M, C, device='cuda', dtype=dtype)
    dt = BF16 if dtype == torch.bfloat16 else F32
    L.check(L.lib().leod_lstm_gates_fwd(dt, L.ptr(gates), L.ptr(c0), L.ptr(h1), L.ptr(c1), M, C, L.stream_ptr()))
    L.check(L.lib().leod_lstm_gates_bwd(dt, L.ptr(gates), L.ptr(c0), L.ptr(c1), L.ptr(dh), L.ptr(dh2), L.ptr(dc), L.ptr(dgates),
                                        L.ptr(dc0), M, C, L.stream_ptr()))
    torch.cuda.synchronize()
    p32 = pre.float().requires_grad_(True)
    c32 = (c0.float() if with_prev else torch.zeros(M, C, device='cuda')).requires_grad_(True)
    f, i, o = (torch.sigmoid(p32[:, k * C:(k + 1) * C]) for k in range(3))
    gg = torch.tanh(p32[:, 3 * C:])
    cr = f * c32 + i * gg
    hr = o * torch.tanh(cr)
    (hr * (dh.float() + dh2.float())).sum().add((cr * dc.float()).sum()).backward()
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    assert rel_err(h1.float(), hr.detach()) < tol
    assert rel_err(c1.float(), cr.detach()) < tol
    assert rel_err(dgates.float(), p32.grad) < tol
    assert rel_err(dc0.float(), c32.grad) < tol


def test_postprocess_matches_oracle_and_fixture():
    from oracle import postprocess as opp
    from leod_b200.models.detection.yolox.utils.boxes import postprocess, postprocess_packed
    z = np.load(os.path.join(GOLDEN, 'net_small.npz'))
    pred = torch.from_numpy(z['eval/preds']).cuda()
    dets = postprocess(pred.clone(), 2, 0.001, 0.45)
    for b, d in enumerate(dets):
        ref = z[f'eval/det{b}']
        got = np.zeros((0, 7), np.float32) if d is None else d.cpu().numpy()
        assert got.shape == ref.shape, (b, got.shape, ref.shape)
        np.testing.assert_array_equal(got, ref)
    # random, tie-heavy predictions at the full Gen1 / Gen4 anchor counts: bit-exact against the oracle
    rng = np.random.default_rng(0)
    for (B, A, C, conf) in ((5, 1680, 2, 0.1), (3, 5040, 3, 0.05), (2, 1680, 2, 0.9999), (2, 300, 1, 0.0)):
        p = np.zeros((B, A, 5 + C), np.float32)
        p[..., 0] = rng.uniform(0, 300, (B, A))
        p[..., 1] = rng.uniform(0, 240, (B, A))
        p[..., 2:4] = rng.uniform(2, 90, (B, A, 2))
        p[..., 4:] = np.round(rng.uniform(0, 1, (B, A, 1 + C)) * 16) / 16   # many equal scores
        p[:, A // 2:, :4] = p[:, :A - A // 2, :4]                             # duplicated boxes
        ref = opp.postprocess(p, C, conf, 0.45)
        got = postprocess(torch.from_numpy(p).cuda(), C, conf, 0.45, pad=torch.zeros(0, 7))
        for b in range(B):
            g = got[b].cpu().numpy()
            assert g.shape == ref[b].shape, (A, b, g.shape, ref[b].shape)
            np.testing.assert_array_equal(g, ref[b])
    # idempotence at full size: NMS of an NMS output keeps everything
    d, n = postprocess_packed(torch.from_numpy(p).cuda(), C, 0.0, 0.45)
    assert int(n.max()) <= 300


def test_nms_fixture_cases():
    """tests/golden/nms_cases.npz (torchvision.ops.batched_nms outputs) through leod_postprocess."""
    from leod_b200.models.detection.yolox.utils.boxes import postprocess_packed
    z = np.load(os.path.join(GOLDEN, 'nms_cases.npz'))
    for i in range(int(z['n'])):
        boxes, scores, cls, thr = z[f'{i}/boxes'], z[f'{i}/scores'], z[f'{i}/cls'].astype(np.int64), float(z[f'{i}/thr'])
        n = boxes.shape[0]
        if n == 0:
            continue
        C = int(cls.max()) + 1
        # encode as predictions: obj = score, one-hot class confidence 1 -> score product is exact
        p = np.zeros((1, n, 5 + C), np.float32)
        p[0, :, 0] = (boxes[:, 0] + boxes[:, 2]) / 2
        p[0, :, 1] = (boxes[:, 1] + boxes[:, 3]) / 2
        p[0, :, 2] = boxes[:, 2] - boxes[:, 0]
        p[0, :, 3] = boxes[:, 3] - boxes[:, 1]
        # only use cases whose corners survive the centre/size round trip exactly
        hw, hh = p[0, :, 2] / 2, p[0, :, 3] / 2
        exact = np.array_equal(p[0, :, 0] - hw, boxes[:, 0]) and np.array_equal(p[0, :, 0] + hw, boxes[:, 2]) and \
            np.array_equal(p[0, :, 1] - hh, boxes[:, 1]) and np.array_equal(p[0, :, 1] + hh, boxes[:, 3])
        if not exact:
            continue
        p[0, :, 4] = scores
        p[0, np.arange(n), 5 + cls] = 1.0
        d, cnt = postprocess_packed(torch.from_numpy(p).cuda(), C, -1.0, thr)
        k = int(cnt[0])
        keep = z[f'{i}/keep']
        assert k == len(keep), (i, k, len(keep))
        np.testing.assert_array_equal(d[0, :k, :4].cpu().numpy(), boxes[keep])


def test_pred2label_matches_fixture():
    from leod_b200.modules.utils.ssod import pred2label_packed
    z = np.load(os.path.join(GOLDEN, 'pred2label_cases.npz'))
    for ci in range(int(z['n'])):
        hw = tuple(int(v) for v in z[f'{ci}/hw'])
        dets = [z[f'{ci}/det{b}'] for b in range(4)]
        md = max(max(len(d) for d in dets), 1)
        packed = np.zeros((4, md, 7), np.float32)
        cnt = np.zeros(4, np.int32)
        for b, d in enumerate(dets):
            packed[b, :len(d)] = d
            cnt[b] = len(d)
        lab, n = pred2label_packed(torch.from_numpy(packed).cuda(), torch.from_numpy(cnt).cuda(),
                                   [float(v) for v in z[f'{ci}/obj_thresh']], [float(v) for v in z[f'{ci}/cls_thresh']], hw)
        for b in range(4):
            ref = z[f'{ci}/label{b}']
            assert int(n[b]) == ref.shape[0], (ci, b)
            np.testing.assert_array_equal(lab[b, :int(n[b])].cpu().numpy(), ref)


def test_voxel_binning_matches_fixture_and_oracle():
    from oracle import binning
    from leod_b200.data.utils.representations import StackedHistogram
    z = np.load(os.path.join(GOLDEN, 'binning_cases.npz'))
    for i in range(int(z['n'])):
        bins, H, W, cutoff, fast = (int(v) for v in z[f'{i}/cfg'])
        sh = StackedHistogram(bins=bins, height=H, width=W, count_cutoff=None if cutoff < 0 else cutoff, fastmode=bool(fast))
        x, y, p, t = (torch.from_numpy(z[f'{i}/{k}']).cuda() for k in 'xypt')
        rep = sh.construct(x, y, p, t.long())
        np.testing.assert_array_equal(rep.cpu().numpy(), z[f'{i}/rep'], err_msg=f'case {i}')
    # full Gen1 size: 2M events, hot pixels; linearity / checksum properties + oracle on a sample
    g = torch.Generator(device='cuda').manual_seed(0)
    n, bins, H, W = 2_000_000, 10, 240, 304
    x = torch.randint(0, W, (n,), device='cuda', generator=g, dtype=torch.int32)
    y = torch.randint(0, H, (n,), device='cuda', generator=g, dtype=torch.int32)
    p = torch.randint(0, 2, (n,), device='cuda', generator=g, dtype=torch.int32)
    t = torch.sort(torch.randint(0, 50_000, (n,), device='cuda', generator=g, dtype=torch.int64)).values
    x[: n // 100] = 7
    y[: n // 100] = 9
    sh = StackedHistogram(bins=bins, height=H, width=W, fastmode=False)
    rep = sh.construct(x, y, p, t)
    assert rep.shape == (2 * bins, H, W) and rep.dtype == torch.uint8
    ref = binning.stacked_histogram(x.cpu().numpy(), y.cpu().numpy(), p.cpu().numpy(), t.cpu().numpy(), bins, H, W, None, False)
    np.testing.assert_array_equal(rep.cpu().numpy(), ref)
    assert int(rep.max()) == 255   # the hot pixel saturates at the cutoff


def test_adamw_ema_matches_fixture():
    from leod_b200.modules.utils.ssod import fused_adamw_ema
    z = np.load(os.path.join(GOLDEN, 'optim_cases.npz'))
    p = torch.from_numpy(z['adamw/p0']).cuda().clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    ema = p.clone()
    for s in range(3):
        gr = torch.from_numpy(z[f'adamw/g{s}']).cuda()
        ema_before = ema.clone()
        fused_adamw_ema(p, gr, m, v, step=s + 1, lr=2e-4, clip_value=1.0, ema=ema, ema_alpha=0.75)
        np.testing.assert_allclose(p.cpu().numpy(), z[f'adamw/p{s + 1}'], rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(ema.cpu().numpy(), (0.75 * ema_before + 0.25 * p).cpu().numpy(), rtol=1e-6, atol=1e-7)
    # the alpha schedule of ema_model_update (ssod.py:429-438) against the reference fixture
    from leod_b200.modules.utils.ssod import ema_alpha_at, ema_model_update
    n = len([k for k in z.files if k.startswith('ema/student')])
    for step in (0, 5, 5000):
        student = [torch.from_numpy(z[f'ema/student{i}']).cuda() for i in range(n)]
        teacher = [torch.from_numpy(z[f'ema/teacher{i}']).cuda().clone() for i in range(n)]
        ema_model_update(student, teacher, step, 0.999)
        assert ema_alpha_at(step, 0.999) == min(1. - 1. / (step + 1.), 0.999)
        for i in range(n):
            np.testing.assert_allclose(teacher[i].cpu().numpy(), z[f'ema/step{step}/teacher{i}'], rtol=1e-6, atol=1e-7)
