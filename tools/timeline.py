"""In-situ kernel timeline of one train step (CUPTI through torch.profiler): start, duration and gap to the previous
kernel for every launch, without the serialisation / cold caches of an ncu pass."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geeco_b200 import create_e2evmc_config  # noqa: E402
from geeco_b200.data import synthetic_batch  # noqa: E402
from geeco_b200.engine import Engine  # noqa: E402


def main():
  N = 64
  dev = torch.device('cuda:0')
  cfg = create_e2evmc_config(dict(proc_obs='dynimg', proc_tgt='dyndiff', batch_size=N))
  eng = Engine(cfg, batch_size=N, precision='bf16', training=True, device=dev)
  eng.init_params(seed=0)
  f, l = synthetic_batch(N, seed=1, structured=False)
  b = {k: torch.from_numpy(v).to(dev) for k, v in f.items() if k != 'step'}
  b['cmd'] = torch.from_numpy(l['cmd']).to(dev)
  for _ in range(5):
    eng.train_step(b, b)
  torch.cuda.synchronize()
  with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
      eng.train_step(b, b)
    torch.cuda.synchronize()
  ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
  ev.sort(key=lambda e: e.time_range.start)
  starts = [i for i, e in enumerate(ev) if 'preprocess' in e.name]
  lo, hi = starts[-1], len(ev)
  t0 = ev[lo].time_range.start
  prev_end = None
  tot_k, tot_gap = 0.0, 0.0
  for e in ev[lo:hi]:
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    gap = 0.0 if prev_end is None else e.time_range.start - prev_end
    prev_end = e.time_range.end
    tot_k += d; tot_gap += max(gap, 0.0)
    print('%8.1f %8.1f gap %6.1f  %s' % (s, d, gap, e.name[:70]))
  print('kernels %.1f us, gaps %.1f us, span %.1f us' % (tot_k, tot_gap, prev_end - t0))


if __name__ == '__main__':
  main()
