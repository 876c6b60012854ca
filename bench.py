#!/usr/bin/env python
"""Benchmark of the GEECO-F train step (BASELINE.json metric) -- see DESIGN.md "Measurement".

  python bench.py --gpus N --steps K --warmup W            our arm  (one process per GPU under torchrun for N>1)
  python bench.py --impl reference ...                      the CPU restatement of the reference graph

Prints ONE JSON line on rank 0.  A "step" is one full train step (rank pooling -> 3 conv encoders ->
LSTM cell -> heads -> losses -> backward -> Adam) on a synthetic batch of 64 windows per GPU
(BASELINE config 2; weak scaling for N>1 with an NCCL gradient all-reduce overlapped with backward).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GEECO-F train samples/sec"
UNIT = "samples/s"
FLOP_PER_SAMPLE_TRAIN = 9.895e9
WORKLOAD = ("GEECO-F train step (fwd+bwd+Adam), synthetic 256x256 RGB pick-pad2-cube2-shaped batches, "
            "K=4, --observation_format rgb --goal_condition target --proc_obs dynimg --proc_tgt dyndiff")


def parse_args():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=20)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', type=str, default='ours', choices=['ours', 'reference'])
  ap.add_argument('--precision', type=str, default=os.environ.get('GEECO_PRECISION', 'bf16'), choices=['bf16', 'fp32'])
  ap.add_argument('--batch', type=int, default=64, help='windows per GPU per step')
  ap.add_argument('--cpu-batch', type=int, default=4, help='windows per step of the CPU baseline sample')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-e2e', action='store_true')
  ap.add_argument('--no-kernels', action='store_true', help='skip the isolated kernel roofline timings')
  return ap.parse_args()


def load_peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    with open(path) as fp:
      d = json.load(fp)
    return {'hbm_gbs': d['hbm_gbs'], 'bf16_tflops': d['bf16_tflops'], 'bf16_tflops_sustained': d['bf16_tflops_sustained'],
            'source': 'measured'}
  return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}


class ClockSampler(object):
  """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
  Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
       'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
       'clocks_event_reasons.sw_power_cap')

  def __init__(self, gpu_index=0):
    self.rows, self.proc, self.gpu = [], None, gpu_index

  def start(self):
    try:
      self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                    '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                   stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append(line.strip())

  def stop(self):
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    time.sleep(0.15)
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:
      self.proc.kill()
    sm, mx, reasons, pw = [], [], set(), []
    for r in self.rows:
      f = [x.strip() for x in r.split(',')]
      if len(f) < 9:
        continue
      try:
        sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
      except ValueError:
        continue
      for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
        if val.lower().startswith('active'):
          reasons.add(name)
    if not sm:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
    return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'power_w_max': float(max(pw)),
            'samples': len(sm), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU restatement of the reference graph (the oracle port) -- reported baseline / reference arm
# ------------------------------------------------------------------------------------------------
def time_cpu_reference(batch, steps, warmup):
  import torch
  from oracle import geeco_oracle as O
  from geeco_b200.data import synthetic_batch
  torch.set_num_threads(os.cpu_count() or 1)
  cfg = O.make_config(batch_size=batch)
  P = O.init_params(cfg, seed=0, dtype=torch.float32)
  opt = O.adam_init(P)
  feats, labels = synthetic_batch(batch, seed=1)
  for _ in range(warmup):
    O.train_step(P, opt, feats, labels, cfg)
  times = []
  for _ in range(steps):
    t0 = time.perf_counter()
    O.train_step(P, opt, feats, labels, cfg)
    times.append(time.perf_counter() - t0)
  med = float(np.median(times))
  cpu_model = ''
  try:
    with open('/proc/cpuinfo') as fp:
      for line in fp:
        if line.startswith('model name'):
          cpu_model = line.split(':', 1)[1].strip()
          break
  except Exception:
    pass
  return {'value': batch / med, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
          'sample': 'oracle/geeco_oracle.py train_step (fp32 torch-CPU restatement of the TF-1.15 graph; TF not '
                    'installable offline), batch %d, median of %d steps after %d warm-up' % (batch, steps, warmup),
          'ms_per_step': med * 1e3, 'cpu_model': cpu_model, 'os_cpu_count': os.cpu_count()}


def run_reference(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  steps = max(1, min(args.steps, 10))
  warm = max(1, min(args.warmup, 2))
  cb = time_cpu_reference(args.cpu_batch, steps, warm)
  line = {
      'impl': 'reference', 'metric': METRIC, 'value': cb['value'], 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
      'warmup': warm, 'ms_per_step': cb['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': WORKLOAD, 'per_step_sample': 'batch %d on host cores' % args.cpu_batch},
      'cpu_baseline': cb,
      'e2e': {'value': cb['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
      'gpu_launches': 0,
  }
  print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
  import torch
  import torch.distributed as dist
  from geeco_b200 import create_e2evmc_config, _lib
  from geeco_b200.data import synthetic_batch
  from geeco_b200.engine import Engine

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  if not torch.cuda.is_available():
    raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback for the product path")
  torch.cuda.set_device(local_rank)
  dev = torch.device('cuda:%d' % local_rank)
  if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=dev)
  N = args.batch
  cfg = create_e2evmc_config(dict(proc_obs='dynimg', proc_tgt='dyndiff', batch_size=N))
  eng = Engine(cfg, batch_size=N, precision=args.precision, training=True, device=dev)
  eng.init_params(seed=0)
  lib = _lib.load()

  # >= 3 distinct device-resident batches (3 x 252 MB of frames >> 126 MB L2), rotated every step
  NB = 3
  host_batches, dev_batches = [], []
  for i in range(NB):
    f, l = synthetic_batch(N, seed=1234 + rank * 16 + i, structured=False)
    hb = {k: torch.from_numpy(v).pin_memory() for k, v in f.items() if k != 'step'}
    hb['cmd'] = torch.from_numpy(l['cmd']).pin_memory()
    host_batches.append(hb)
    dev_batches.append({k: v.to(dev) for k, v in hb.items()})

  def step_dev(i):
    b = dev_batches[i % NB]
    if world == 1:
      return eng.train_step(b, b)
    eng.step_forward(b, b)
    works = []
    for bk in range(len(eng.buckets)):
      g = eng.step_backward(bk)
      works.append(dist.all_reduce(g, op=dist.ReduceOp.SUM, async_op=True))
    for w in works:
      w.wait()
    eng.step_update(1.0 / world)
    return eng.out_losses

  def barrier():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
      torch.cuda.synchronize()

  for i in range(args.warmup):
    step_dev(i)
  barrier()
  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
  lib.geeco_launch_count(1)
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  ev0.record()
  for i in range(args.steps):
    step_dev(i)
  ev1.record()
  barrier()
  launches = int(lib.geeco_launch_count(1))
  clocks = sampler.stop() if rank == 0 else None
  ms = ev0.elapsed_time(ev1)
  t = torch.tensor([ms], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  ms = float(t.item())
  final_loss = float(eng.out_losses[5].item())
  value = world * N * args.steps / (ms * 1e-3)

  # ---- end-to-end through the public API with HOST (pinned) inputs, loss read back every step
  e2e = None
  if not args.no_e2e:
    def step_host(i):
      hb = host_batches[i % NB]
      if world == 1:
        losses = eng.train_step(hb, hb)
      else:
        eng.step_forward(hb, hb)
        works = []
        for bk in range(len(eng.buckets)):
          works.append(dist.all_reduce(eng.step_backward(bk), op=dist.ReduceOp.SUM, async_op=True))
        for w in works:
          w.wait()
        eng.step_update(1.0 / world)
        losses = eng.out_losses
      return losses.cpu()      # device -> host read of the step's result (32 bytes)
    for i in range(max(1, min(args.warmup, 3))):
      step_host(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_e2e = max(3, min(args.steps, 10))
    e0.record()
    for i in range(n_e2e):
      step_host(i)
    e1.record()
    barrier()
    t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
      dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e = {'value': world * N * n_e2e / (float(t2.item()) * 1e-3), 'unit': UNIT,
           'h2d_bytes_per_step': eng.h2d_bytes(True), 'd2h_bytes_per_step': 32, 'steps': n_e2e,
           'api': 'geeco_b200.engine.Engine.train_step(host pinned features, labels) -> losses.cpu()'}

  peaks = load_peaks()
  roofline, extra = kernel_rooflines(eng, dev, peaks, args) if (rank == 0 and not args.no_kernels) else (None, None)
  cpu_baseline = None
  if rank == 0 and world == 1 and not args.no_cpu_baseline:
    cpu_baseline = time_cpu_reference(args.cpu_batch, 5, 2)

  if rank == 0:
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'per_gpu_batch': N, 'global_batch': N * world, 'window_size': 4,
                   'parallelism': 'dp%d' % world,
                   'l2_policy': 'inputs larger than L2: %d rotating batches x %.0f MB of frames' % (
                       NB, eng.h2d_bytes(True) / 1e6),
                   'master_weights': 'fp32', 'optimizer': 'TF-Adam fused'},
        'step_tflops': value * FLOP_PER_SAMPLE_TRAIN / 1e12,
        'final_loss': final_loss,
        'clocks': clocks, 'gpu_launches': launches, 'e2e': e2e, 'roofline': roofline, 'kernels': extra,
        'cpu_baseline': cpu_baseline, 'peaks': peaks,
    }
    print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


def kernel_rooflines(eng, dev, peaks, args):
  """Times the dominant kernels in isolation with CUDA events (inputs > L2 or L2 flushed between reps)."""
  import torch
  from geeco_b200 import ops
  out = {}
  flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

  def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
      flush.fill_(1)
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record(); fn(); b.record()
      torch.cuda.synchronize()
      ts.append(a.elapsed_time(b))
    return float(np.mean(ts)) * 1e-3

  # rank pooling: dynimg over K=4 frames, 256 samples of 256x256x3 (805 MB in, 201 MB out)
  Nr, K, H, W, C = 256, 4, 256, 256, 3
  x = torch.rand((Nr, K, H, W, C), device=dev)
  y = torch.empty((Nr, H, W, C), device=dev)
  sec = timed(lambda: ops.dynimg(x, out=y))
  bytes_alg = (K + 1) * H * W * C * 4 * Nr
  out['dynimg_cluster_kernel<4>'] = {'bound': 'hbm', 'achieved': bytes_alg / sec / 1e9, 'peak': peaks['hbm_gbs'],
                                     'unit': 'GB/s', 'frac': bytes_alg / sec / 1e9 / peaks['hbm_gbs'], 'traffic': None,
                                     'peak_source': peaks['source'], 'ms': sec * 1e3,
                                     'algorithmic_bytes_per_launch': bytes_alg}
  del x, y
  dom = dominant_kernel_roofline(eng, dev, peaks, timed, args)
  out.update(dom[1])
  return dom[0], out


def dominant_kernel_roofline(eng, dev, peaks, timed, args):
  """conv2 forward of the three encoders (39.8 % of the step's FLOPs) at the bench batch."""
  import torch
  from geeco_b200 import ops
  N = 3 * args.batch
  if args.precision == 'bf16' and hasattr(ops, 'conv2d_same_bf16'):
    x = (torch.rand((N, 256, 256, 32), device=dev)).to(torch.bfloat16)
    w = ((torch.rand((3, 3, 32, 48), device=dev) - 0.5) * 0.2)
    b = torch.zeros(48, device=dev)
    sec = timed(lambda: ops.conv2d_same_bf16(x, w, b, stride=2))
    name, dt = 'conv_tc fwd conv2 (32->48, s2, 256px)', 2
  else:
    x = torch.rand((N, 256, 256, 32), device=dev)
    w = ((torch.rand((3, 3, 32, 48), device=dev) - 0.5) * 0.2)
    b = torch.zeros(48, device=dev)
    sec = timed(lambda: ops.conv2d_same(x, w, b, stride=2))
    name, dt = 'gemm_nn_f32_kernel<3> conv2 fwd (32->48, s2, 256px)', 4
  flops = 2.0 * N * 128 * 128 * 48 * 288
  bytes_alg = N * (256 * 256 * 32 + 128 * 128 * 48) * dt
  r_hbm = {'bound': 'hbm', 'achieved': bytes_alg / sec / 1e9, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
           'frac': bytes_alg / sec / 1e9 / peaks['hbm_gbs'], 'traffic': None, 'kernel': name, 'ms': sec * 1e3,
           'peak_source': peaks['source'], 'algorithmic_bytes_per_launch': bytes_alg,
           'tensor_tflops': flops / sec / 1e12, 'tensor_frac_of_burst': flops / sec / 1e12 / peaks['bf16_tflops']}
  return r_hbm, {name: r_hbm}


def main():
  args = parse_args()
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_ours(args)


if __name__ == '__main__':
  main()
