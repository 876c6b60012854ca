"""Phase timing of the data-parallel step (run under torchrun): forward / backward buckets / all-reduce / update."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geeco_b200 import create_e2evmc_config, parallel  # noqa: E402
from geeco_b200.data import synthetic_batch  # noqa: E402
from geeco_b200.engine import Engine  # noqa: E402


def main():
  rank, world, lr = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
  torch.cuda.set_device(lr)
  dev = torch.device('cuda:%d' % lr)
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)
  N = 64
  cfg = create_e2evmc_config(dict(proc_obs='dynimg', proc_tgt='dyndiff', batch_size=N))
  eng = Engine(cfg, batch_size=N, precision='bf16', training=True, device=dev)
  eng.init_params(seed=0)
  f, l = synthetic_batch(N, seed=1, structured=False)
  b = {k: torch.from_numpy(v).to(dev) for k, v in f.items() if k != 'step'}
  b['cmd'] = torch.from_numpy(l['cmd']).to(dev)
  mode = os.environ.get('DP_MODE', 'full')
  for it in range(6):
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    t = [time.perf_counter()]
    eng.step_forward(b, b); torch.cuda.synchronize(); t.append(time.perf_counter())
    works = []
    for bi in range(len(eng.buckets)):
      eng.step_backward(bi)
      if mode == 'sync':
        torch.cuda.synchronize()
      t.append(time.perf_counter())
      if world > 1 and mode != 'nocomm':
        works.append(parallel.allreduce_bucket(eng.grad, eng.buckets[bi], async_op=True))
      t.append(time.perf_counter())
    for w in works:
      w.wait()
    torch.cuda.synchronize(); t.append(time.perf_counter())
    eng.step_update(1.0 / world); torch.cuda.synchronize(); t.append(time.perf_counter())
    if rank == 0:
      print('iter %d  total %.2f ms  ' % (it, (t[-1] - t[0]) * 1e3) + ' '.join('%.2f' % ((t[i + 1] - t[i]) * 1e3) for i in range(len(t) - 1)),
            flush=True)
  # free-running loop (no host synchronisation inside), as bench.py times it
  for rep in range(2):
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for it in range(20):
      parallel.data_parallel_step(eng, b, b)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    if rank == 0:
      print('free-run: %.2f ms/step on device, host enqueue %.2f ms/step' % (e0.elapsed_time(e1) / 20, (t1 - t0) * 1e3 / 20), flush=True)
  if world > 1:
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
