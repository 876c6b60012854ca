"""HBM bandwidth by access mix on this GPU (torch kernels, CUDA events, best of 10 over 2 GiB): copy (1 read : 1 write, what
MEASURED_PEAKS.json's hbm_gbs is), write-only (fill), read-only (sum), and 1 read : 4 writes / 4 reads : 1 write (the mixes
of the store-heavy / load-heavy conv kernels).  Context for roofline fractions quoted against the copy figure."""
import torch
dev = torch.device('cuda:0')
n = 1 << 29                                   # 2^29 float32 = 2 GiB
a = torch.rand(n, device=dev)
b = torch.empty(n, device=dev)
small = torch.rand(n // 4, device=dev)


def best(fn, bytes_):
  fn(); torch.cuda.synchronize()
  ts = []
  for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
  return bytes_ / (min(ts) * 1e-3) / 1e9


print('copy   (1r:1w)  %6.0f GB/s' % best(lambda: b.copy_(a), 8 * n))
print('fill   (0r:1w)  %6.0f GB/s' % best(lambda: b.fill_(1.5), 4 * n))
print('sum    (1r:0w)  %6.0f GB/s' % best(lambda: a.sum(), 4 * n))
print('expand (1r:4w)  %6.0f GB/s' % best(lambda: b.view(4, n // 4).copy_(small.unsqueeze(0).expand(4, n // 4)), 4 * n + n))
print('reduce (4r:1w)  %6.0f GB/s' % best(lambda: torch.sum(a.view(4, n // 4), dim=0, out=small), 4 * n + n))
