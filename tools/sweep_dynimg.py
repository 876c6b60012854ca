"""Rank-pooling sweep (BASELINE config 5) over cluster sizes: GB/s per (px, K, cluster).  cluster 0 = the library's choice,
-1 = two-pass kernels."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geeco_b200 import ops
dev = torch.device('cuda:0')
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
clusters = [int(c) for c in (sys.argv[1].split(',') if len(sys.argv) > 1 else '0,2,4,8,16,-1'.split(','))]
for px in (128, 256, 512):
  for K in (2, 3, 4, 6, 8, 12, 16):
    per = K * px * px * 3 * 4
    n = min(max(2, int(2 ** 30 // per) + 1), 60000)
    x = torch.rand((n, K, px, px, 3), device=dev)
    y = torch.empty((n, px, px, 3), device=dev)
    row = []
    for cl in clusters:
      try:
        ops.dynimg(x, out=y, cluster=cl); torch.cuda.synchronize()
        ts = []
        for _ in range(3):
          flush.fill_(1)
          a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
          a.record(); ops.dynimg(x, out=y, cluster=cl); b.record(); torch.cuda.synchronize()
          ts.append(a.elapsed_time(b))
        row.append('%5.0f' % ((K + 1) * px * px * 3 * 4 * n / (np.median(ts) * 1e-3) / 1e9))
      except Exception as e:
        row.append('  err')
    print('px %3d K %2d N %5d  GB/s by cluster %s: %s' % (px, K, n, clusters, ' '.join(row)), flush=True)
    del x, y
