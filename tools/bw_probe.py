"""HBM bandwidth by access mix on this GPU (torch elementwise kernels): pure write, pure read (sum), copy, 1 read : 8 writes."""
import torch
n = 1 << 30   # 1 GiB
a = torch.empty(n, dtype=torch.uint8, device='cuda'); b = torch.empty(n, dtype=torch.uint8, device='cuda')
af, bf = a.view(torch.float32), b.view(torch.float32)
def t(fn, bytes_, it=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return bytes_ * it / (e0.elapsed_time(e1) * 1e-3) / 1e9
print('fill  (write only)   %.0f GB/s' % t(lambda: af.fill_(1.0), n))
print('sum   (read only)    %.0f GB/s' % t(lambda: af.sum(), n))
print('copy  (1R : 1W)      %.0f GB/s' % t(lambda: bf.copy_(af), 2 * n))
x = torch.empty(n // 4 // 8, dtype=torch.float32, device='cuda')
out = bf.view(8, -1)
print('bcast (1R : 8W)      %.0f GB/s' % t(lambda: out.copy_(x[None, :].expand(8, -1)), n + n // 8))
print('memset               %.0f GB/s' % t(lambda: a.zero_(), n))
