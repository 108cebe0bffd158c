"""Micro-benchmark of the tcgen05 NT GEMM at the stage-1 shapes of the bench workload: median of 20 L2-flushed launches each."""
import os, sys, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from leod_b200 import _lib
L = _lib.lib()
dev = 'cuda'
def run(M, N, K, epi, reps=20):
    A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
    C = torch.empty(M, N, device=dev, dtype=torch.bfloat16); bias = torch.randn(N, device=dev)
    R = torch.randn(M, N, device=dev).bfloat16() if epi == 2 else None
    aux = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if epi in (1, 3) else None
    st = _lib.stream_ptr()
    def call():
        _lib.check(L.leod_gemm_nt(1, _lib.LEOD_BF16, _lib.ptr(A), K, None, 0, 0, _lib.ptr(B), K, _lib.ptr(C), N, M, N, K, _lib.ptr(bias), epi,
                                  _lib.ptr(R), N, _lib.ptr(aux), N, st), 'gemm_nt')
    for _ in range(3): call()
    big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2 * reps)]
    for i in range(reps):
        big.zero_()          # flush L2
        e[2 * i].record(); call(); e[2 * i + 1].record()
    torch.cuda.synchronize()
    t = sorted(e[2 * i].elapsed_time(e[2 * i + 1]) for i in range(reps))[reps // 2] * 1e3
    by = 2 * (M * K + M * N * (1 + (epi in (1, 2, 3))))
    print(f'M={M} N={N} K={K} epi={epi}: {t:7.1f} us  {by / t / 1e3:6.0f} GB/s', flush=True)
for shape in ((860160, 48, 48, 2), (860160, 48, 48, 0), (860160, 144, 48, 0), (860160, 192, 48, 1), (860160, 48, 192, 2)):
    run(*shape)
