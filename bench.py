#!/usr/bin/env python
"""Benchmark of the GEECO-F train step (BASELINE.json metric) -- see DESIGN.md "Measurement".

  python bench.py --gpus N --steps K --warmup W            our arm  (one process per GPU under torchrun for N>1)
  python bench.py --impl reference ...                      the CPU restatement of the reference graph
  python bench.py --extras                                  BASELINE configs 4 (batched policy steps) and 5 (rank-
                                                             pooling sweep) in full, written to profiles/; every 1-GPU
                                                             line carries a one-chunk / six-point version of both
  python bench.py --scaling strong --gpus N                 BASELINE config 3: global batch 512 split over N GPUs

Prints ONE JSON line on rank 0.  A "step" is one full train step (rank pooling -> 3 conv encoders ->
LSTM cell -> heads -> losses -> backward -> Adam) on a synthetic batch of 64 windows per GPU
(BASELINE config 2; weak scaling for N>1 with an NCCL gradient all-reduce overlapped with backward).

  value     device-resident inputs (3 rotating batches, 252 MB of frames each: larger than the 126 MB L2)
  e2e       the same metric through the public training entry `Estimator.train(input_fn)` with HOST (pinned)
            batches: every step uploads its 252 MB batch (overlapped with the previous step on a copy stream)
            and reads the step's losses back to the host
  roofline  the kernel with the largest share of the step, timed alone with CUDA events (L2 flushed between
            repetitions) against the measured HBM copy bandwidth; `kernels` holds the other hot kernels
  input_pipeline  (N=1) host-side rate of the native episode decoder (libgeeco_io.so) on this box's cores, measured
            beside cpu_baseline on one synthetic recorded episode; never part of the timed GPU region
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
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
  ap.add_argument('--steps', type=int, default=50)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', type=str, default='ours', choices=['ours', 'reference'])
  ap.add_argument('--precision', type=str, default=os.environ.get('GEECO_PRECISION', 'bf16'), choices=['bf16', 'fp32'])
  ap.add_argument('--batch', type=int, default=64, help='windows per GPU per step')
  ap.add_argument('--cpu-batch', type=int, default=0,
                  help='windows per step of the CPU arm / baseline (0: the per-GPU batch of the workload, at most 64)')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-e2e', action='store_true')
  ap.add_argument('--no-kernels', action='store_true', help='skip the isolated kernel roofline timings')
  ap.add_argument('--extras', action='store_true',
                  help='BASELINE config 4 over all 4096 environments and the full config 5 sweep (default: one '
                       '1024-environment chunk and a 6-point sweep subset)')
  ap.add_argument('--scaling', type=str, default='weak', choices=['weak', 'strong'],
                  help='weak: --batch windows per GPU; strong: BASELINE config 3, global batch 512 split over the GPUs')
  ap.add_argument('--global-batch', type=int, default=512, help='global batch of --scaling strong')
  return ap.parse_args()


def load_peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    with open(path) as fp:
      d = json.load(fp)
    return {'hbm_gbs': d['hbm_gbs'], 'bf16_tflops': d['bf16_tflops'], 'bf16_tflops_sustained': d['bf16_tflops_sustained'],
            'source': 'measured'}
  return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}


def load_traffic():
  """DRAM bytes per launch from the committed ncu --set full captures (profiles/r02_traffic.json, else round 1's)."""
  for name in ('r02_traffic.json', 'r01_traffic.json'):
    path = os.path.join(ROOT, 'profiles', name)
    if os.path.exists(path):
      with open(path) as fp:
        return json.load(fp)
  return {}


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
                                    '--format=csv,noheader,nounits', '-lms', '20'], stdout=subprocess.PIPE,
                                   stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
      time.sleep(0.3)      # let the sampler come up before the timed region starts
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append((time.time(), line.strip()))

  def stop(self, t0=None, t1=None):
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    time.sleep(0.12)
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:
      self.proc.kill()
    sm, mx, reasons, pw = [], [], set(), []
    for ts, r in self.rows:
      if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.1):
        continue
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


def time_input_pipeline(frames=24):
  """Host-side leg next to cpu_baseline: the native episode decoder (libgeeco_io.so) on this box's host cores.  One
  synthetic 256x256 episode in the recorder's format (float-encoded pixels, one zlib stream) is written to a
  scratch directory and decoded to uint8 frames; the rate is per decode thread (episodes decode in parallel).  A
  failure here never takes the bench line down: it is reported under 'error'."""
  import shutil
  import tempfile
  d = tempfile.mkdtemp(prefix='geeco_bench_ds_')
  try:
    from geeco_b200.data import write_synthetic_dataset
    from geeco_b200.input_pipeline import decode_episode, get_meta_v4
    write_synthetic_dataset(d, episodes=1, episode_length=frames, eval_episodes=0)
    meta = get_meta_v4(d)
    path = os.path.join(d, 'data', '000000.tfrecord.zlib')
    decode_episode(path, meta, True, 'uint8', want_depth=False)
    ts = []
    for _ in range(3):
      t0 = time.perf_counter()
      ep = decode_episode(path, meta, True, 'uint8', want_depth=False)
      ts.append(time.perf_counter() - t0)
    t = sorted(ts)[1]
    return {'frames_per_s_per_thread': frames / t, 'ms_per_episode': 1e3 * t, 'episode_frames': frames,
            'file_mb': os.path.getsize(path) / 1e6, 'protobuf_mb': frames * 256 * 256 * 4 * 4 / 1e6,
            'decoded': {k: list(v.shape) for k, v in ep.items() if k in ('rgb', 'jnt_state', 'target_rgb')},
            'host_cores': os.cpu_count(),
            'sample': 'decode_episode (inflate + CRC-32C + SequenceExample -> uint8 frames, states, targets) of one '
                      'synthetic %d-frame 256x256 episode, median of 3; pixel noise makes it the worst case for zlib'
                      % frames}
  except Exception as e:    # noqa: BLE001
    return {'error': repr(e)}
  finally:
    shutil.rmtree(d, ignore_errors=True)


def bench_config(args, world, precision=None):
  """`config` of the JSON line: the workload both arms measure."""
  per_gpu = args.global_batch // world if args.scaling == 'strong' else args.batch
  return {'workload': WORKLOAD, 'per_gpu_batch': per_gpu, 'global_batch': per_gpu * world, 'window_size': 4,
          'parallelism': 'dp%d' % world}


def run_reference(args):
  """Reference arm: the reference's CPU implementation of the path (the oracle port -- TensorFlow 1.15 is not
  installable here) on all host cores, rank 0 only, EXACTLY --steps timed steps after --warmup warm-up steps.  One
  step = one full train step on one batch of the workload's per-GPU size (a bounded sample of the N-GPU job: the CPU
  has no second replica to add)."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  world = int(os.environ.get('WORLD_SIZE', str(args.gpus)))
  cfg = bench_config(args, max(world, 1))
  batch = args.cpu_batch if args.cpu_batch > 0 else min(cfg['per_gpu_batch'], 64)
  cb = time_cpu_reference(batch, args.steps, args.warmup)
  line = {
      'impl': 'reference', 'metric': METRIC, 'value': cb['value'], 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
      'warmup': args.warmup, 'ms_per_step': cb['ms_per_step'], 'higher_is_better': True, 'scaling': args.scaling,
      'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': cfg,
      'cpu_baseline': cb,
      'e2e': {'value': cb['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
      'gpu_launches': 0,
  }
  emit(line)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
  import torch
  import torch.distributed as dist
  from geeco_b200 import create_e2evmc_config, _lib, parallel
  from geeco_b200.data import synthetic_batch
  from geeco_b200.engine import Engine
  from geeco_b200.estimator import Estimator, RunConfig, goal_e2evmc_model_fn

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  parallel.tune_for_data_parallel(world)
  if not torch.cuda.is_available():
    raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback for the product path")
  torch.cuda.set_device(local_rank)
  dev = torch.device('cuda:%d' % local_rank)
  numa_cores = parallel.pin_to_gpu_numa_node(local_rank, world) if (world > 1 and not os.environ.get('GEECO_NO_NUMA_PIN')) else None
  if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=dev)
  bcfg = bench_config(args, world)
  N = bcfg['per_gpu_batch']
  if args.scaling == 'strong' and args.global_batch % world:
    raise ValueError("--global-batch %d is not divisible by %d GPUs" % (args.global_batch, world))
  cfg = create_e2evmc_config(dict(proc_obs='dynimg', proc_tgt='dyndiff', batch_size=N))
  eng = Engine(cfg, batch_size=N, precision=args.precision, training=True, device=dev)
  eng.init_params(seed=0)
  parallel.broadcast_parameters(eng)
  lib = _lib.load()

  # >= 3 distinct batches (3 x 252 MB of frames >> 126 MB L2), rotated every step
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
    return parallel.data_parallel_step(eng, b, b)

  def barrier():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
      torch.cuda.synchronize()

  for i in range(args.warmup):
    step_dev(i)
  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()          # before the barrier: its start-up delay must not skew rank 0 against the others
  barrier()
  lib.geeco_launch_count(1)
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  t_wall0 = time.time()
  ev0.record()
  for i in range(args.steps):
    step_dev(i)
  ev1.record()
  barrier()
  t_wall1 = time.time()
  launches = int(lib.geeco_launch_count(1))
  clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
  ms = ev0.elapsed_time(ev1)
  t = torch.tensor([ms], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  ms = float(t.item())
  final_loss = float(eng.out_losses[5].item())
  value = world * N * args.steps / (ms * 1e-3)

  # ---- end to end through the public training entry: Estimator.train(input_fn) over HOST batches
  e2e = None
  if not args.no_e2e:
    n_e2e = max(args.steps, 50)            # long enough that the one-off upload of the first batch (pipeline fill) amortises
    mdir = tempfile.mkdtemp(prefix='geeco_bench_')
    est = Estimator(goal_e2evmc_model_fn, mdir, RunConfig(save_checkpoints_steps=0, keep_checkpoint_max=1),
                    {'e2evmc_config': cfg, 'log_steps': 1, 'debug': False, 'save_final_checkpoint': False},
                    precision=args.precision, batch_size=N)
    est._engine = eng                      # same replica; Estimator drives it through model_fn

    def timed_train(batches, frames_u8):
      def host_input(n):
        return lambda: ((batches[i % NB], batches[i % NB]) for i in range(n))
      est.train(host_input(2), steps=2)      # warm-up (allocates staging buffers)
      barrier()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      est.train(host_input(n_e2e), steps=n_e2e)          # log_steps=1: every step's losses are copied to the host (async, read one step later)
      e1.record()
      barrier()
      t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
      if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
      return {'value': world * N * n_e2e / (float(t2.item()) * 1e-3), 'unit': UNIT,
              'h2d_bytes_per_step': eng.h2d_bytes(True, frames_u8=frames_u8), 'd2h_bytes_per_step': 4 * 12,
              'steps': n_e2e}

    # float32 frames in [0,1]: exactly what the reference's model_fn is handed (estimator.py:160-176); 252 MB of
    # frames per 64-sample step, so at >= 12k samples/s this leg measures the PCIe link, not the step
    e2e_f32 = timed_train(host_batches, False)
    e2e_f32['api'] = 'Estimator.train(input_fn over pinned host batches), float32 frames in [0,1] (model_fn boundary)'
    # headline: the same entry fed the frames as RECORDED (uint8); the `/= 255.0` of the reference's input pipeline
    # (geeco_gym.py:310) runs on the device and produces the bit-identical network input (tests/test_gpu_boundary.py)
    u8_batches = []
    for hb in host_batches:
      ub = dict(hb)
      for k in ('rgb', 'target_rgb'):
        ub[k] = torch.round(hb[k] * 255.0).to(torch.uint8).pin_memory()
      u8_batches.append(ub)
    e2e_u8 = timed_train(u8_batches, True)
    e2e_u8['api'] = ('Estimator.train(input_fn over pinned host batches, uint8 frames as recorded, dense [N,K,H,W,C] windows)')
    # headline: the batches as the input pipeline emits them for consecutive windows (layout='pool': every distinct
    # frame once + an int32 index; the 64 windows of a batch are consecutive windows of two episodes, as in the
    # reference's batches of 32 consecutive windows).  Same pixels reach the network (bit-identical step:
    # tests/test_gpu_switches.py::test_frame_pool_layout_trains_bit_identically), ~4x fewer bytes cross PCIe.
    from geeco_b200.data import synthetic_pool_batch
    pool_batches = []
    for i in range(NB):
      pf, pl = synthetic_pool_batch(N, pieces=max(1, N // 32), seed=4321 + rank * 16 + i, structured=False)
      hb = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in pf.items() if k != 'step'}
      hb['cmd'] = torch.from_numpy(pl['cmd']).pin_memory()
      pool_batches.append(hb)
    e2e = timed_train(pool_batches, True)
    e2e['h2d_bytes_per_step'] = eng.h2d_bytes(True, frames_u8=True, features=pool_batches[0])
    e2e['api'] = ('geeco_b200.estimator.Estimator.train(input_fn over pinned host batches in the frame-pool layout of '
                  "pickplace_input_fn(layout='pool'): uint8 frames as recorded, each distinct frame once + index), "
                  'log_steps=1: every step uploads its batch and copies its losses back')
    e2e['dense_uint8_windows'] = e2e_u8
    e2e['float32_frames'] = e2e_f32

  peaks = load_peaks()
  roofline, extra = kernel_rooflines(eng, dev, peaks, args) if (rank == 0 and not args.no_kernels) else (None, None)
  cpu_baseline = None
  if rank == 0 and world == 1 and not args.no_cpu_baseline:
    # bounded sample of the same workload on the host cores: one batch of the per-GPU size, 1 warm-up + 4 timed steps
    cpu_baseline = time_cpu_reference(args.cpu_batch if args.cpu_batch > 0 else min(N, 64), 4, 1)
  input_pipeline = time_input_pipeline() if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
  # the other half of BASELINE's metric (batched policy steps, config 4) and the rank-pooling sweep (config 5) ride in
  # every single-GPU line: one 1024-environment chunk / six sweep points by default, everything with --extras
  extras = run_extras(eng, dev, peaks, args, full=args.extras) if (rank == 0 and world == 1 and not args.no_kernels) else None

  if rank == 0:
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
        'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
        'config': dict(bcfg, l2_policy='inputs larger than L2: %d rotating batches x %.0f MB of frames' % (
            NB, eng.h2d_bytes(True) / 1e6), master_weights='fp32', optimizer='TF-Adam fused'),
        'step_tflops': value * FLOP_PER_SAMPLE_TRAIN / 1e12,
        'final_loss': final_loss,
        'clocks': clocks, 'gpu_launches': launches, 'e2e': e2e, 'roofline': roofline, 'kernels': extra,
        'cpu_baseline': cpu_baseline, 'input_pipeline': input_pipeline, 'peaks': peaks,
        'host': {'cpu_count': os.cpu_count(), 'rank0_numa_cores': len(numa_cores) if numa_cores else None},
        'policy_steps': (extras or {}).get('policy_steps'), 'rankpool_sweep': (extras or {}).get('rankpool_sweep'),
    }
    emit(line)
  if world > 1:
    dist.destroy_process_group()


def _timed(dev, fn, reps=5, flush=None):
  import torch
  fn(); torch.cuda.synchronize()
  ts = []
  for _ in range(reps):
    if flush is not None:
      flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
  return float(np.mean(ts)) * 1e-3


def kernel_rooflines(eng, dev, peaks, args):
  """Times the hot kernels in isolation with CUDA events (L2 flushed between repetitions).  Algorithmic bytes
  per launch are the figures of DESIGN.md "Kernels"; traffic comes from the committed ncu captures."""
  import torch
  from geeco_b200 import ops
  out = {}
  traffic = load_traffic()
  flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
  hbm = peaks['hbm_gbs']

  def entry(name, sec, bytes_alg, flops=None, launches=1):
    e = {'bound': 'hbm', 'achieved': bytes_alg / sec / 1e9, 'peak': hbm, 'unit': 'GB/s',
         'frac': bytes_alg / sec / 1e9 / hbm, 'traffic': traffic.get(name), 'kernel': name, 'ms': sec * 1e3,
         'peak_source': peaks['source'], 'algorithmic_bytes_per_launch': bytes_alg, 'launches_timed': launches}
    if flops:
      e['tflops'] = flops / sec / 1e12
      e['tensor_frac_of_burst'] = flops / sec / 1e12 / peaks['bf16_tflops']
    out[name] = e
    return e

  # rank pooling: dynimg over K=4 frames, 256 samples of 256x256x3 (805 MB in, 201 MB out)
  Nr, K, H, W, C = 256, 4, 256, 256, 3
  x = torch.rand((Nr, K, H, W, C), device=dev)
  y = torch.empty((Nr, H, W, C), device=dev)
  sec = _timed(dev, lambda: ops.dynimg(x, out=y), flush=flush)
  entry('dynimg_cluster_kernel<4>', sec, (K + 1) * H * W * C * 4 * Nr)
  del x, y
  if args.precision != 'bf16':
    N3 = 3 * args.batch
    x = torch.rand((N3, 256, 256, 32), device=dev)
    w = ((torch.rand((3, 3, 32, 48), device=dev) - 0.5) * 0.2)
    b = torch.zeros(48, device=dev)
    sec = _timed(dev, lambda: ops.conv2d_same(x, w, b, stride=2), flush=flush)
    dom = entry('gemm_nn_f32_kernel<3> conv2 fwd', sec, N3 * (256 * 256 * 32 + 128 * 128 * 48) * 4,
                2.0 * N3 * 128 * 128 * 48 * 288)
    return dom, out
  N3 = 3 * args.batch
  # conv1 (4-channel padded input 256x256 -> 32 ch) and conv2 (32 -> 48, stride 2): the HBM-bound bulk of the step.
  # Each entry times ONE stand-alone op = the named kernel plus its few-microsecond helpers (weight repack of that layer,
  # split reduction); algorithmic bytes = every tensor the op must read or write once (DESIGN.md "Kernels").
  x1 = torch.rand((N3, 256, 256, 4), device=dev).to(torch.bfloat16); x1[..., 3] = 0
  w1 = ((torch.rand((3, 3, 3, 32), device=dev) - 0.5) * 0.2)
  b1 = torch.zeros(32, device=dev)
  g1 = (torch.rand((N3, 256, 256, 32), device=dev) - 0.5).to(torch.bfloat16)
  P1, P2 = N3 * 65536, N3 * 16384                       # pixels of the 256x256 and 128x128 maps
  # the step's own conv1 + conv2 forward: ONE fused kernel (conv12_fused.cu), run again on the buffers of the last train
  # step.  Algorithmic bytes: x0 in, y1 out once (the backward needs it; it is never read back by the forward), both
  # 1-bit ReLU masks out, y2 out.  The separate kernels below are what it replaces (and what other image sizes still run).
  fused = None
  try:
    sec = _timed(dev, lambda: eng.profile_kernel('conv12'), flush=flush)
    fused = entry('conv12_fused_kernel conv1+conv2 fwd', sec, P1 * 4 * 2 + P1 * 32 * 2 + P1 * 4 + P2 * 48 * 2 + P2 * 6,
                  2.0 * P1 * 32 * 27 + 2.0 * P2 * 48 * 288, launches=1)
  except Exception as exc:                      # GEECO_NO_FUSE12=1 or another image size: the separate kernels ran
    sys.stderr.write('fused conv1->conv2 kernel not timed: %s\n' % exc)
  # the step's conv2 data gradient + conv1 weight gradient: ONE fused kernel (conv21_bwd_fused.cu) + its partial reduce.
  # Algorithmic bytes: G2 in, the 1-bit ReLU mask of y1 in, x0 in; dL/d(pre-activation of conv1) (0.8 GB) never exists.
  fused_bwd = None
  try:
    sec = _timed(dev, lambda: eng.profile_kernel('bwd21'), flush=flush)
    fused_bwd = entry('conv21_bwd_fused_kernel conv2 dgrad + conv1 wgrad', sec, P2 * 48 * 2 + P1 * 4 + P1 * 4 * 2,
                      2.0 * P2 * 48 * 288 + 2.0 * P1 * 32 * 27, launches=2)
    fused_bwd['limiter'] = 'shared-memory data pipe (MMA operand reads + staging stores), see profiles/r02_ncu_notes.md'
  except Exception as exc:                      # GEECO_NO_FUSE_BWD21=1 or another image size
    sys.stderr.write('fused conv2-dgrad -> conv1-wgrad kernel not timed: %s\n' % exc)
  # the step's conv2 weight gradient: whole y1 rows by TMA into a ring, windows stacked into M = 128 (conv2_wgrad_fused.cu)
  rows_wg = None
  try:
    sec = _timed(dev, lambda: eng.profile_kernel('wgrad2'), flush=flush)
    rows_wg = entry('conv2_wgrad_rows_kernel conv2 wgrad', sec, (P1 * 32 + P2 * 48) * 2, 2.0 * P2 * 48 * 288, launches=2)
  except Exception as exc:
    sys.stderr.write('row-wise conv2 weight gradient not timed: %s\n' % exc)
  sec = _timed(dev, lambda: ops.conv2d_same_bf16(x1, w1, b1, stride=1), flush=flush)
  entry('tc_nn_kernel<4,4> conv1 fwd', sec, P1 * (4 + 32) * 2, 2.0 * P1 * 32 * 27, launches=2)
  sec = _timed(dev, lambda: ops.conv2d_same_bwd_bf16(x1, w1, g1, stride=1, need_dx=False), flush=flush)
  entry('tc_wgrad_kernel<4,256> conv1 wgrad', sec, P1 * (4 + 32) * 2, 2.0 * P1 * 32 * 27, launches=2)
  x2 = g1.abs()                                          # post-ReLU activations of conv1 (the mask of the data gradient)
  w2 = ((torch.rand((3, 3, 32, 48), device=dev) - 0.5) * 0.2)
  b2 = torch.zeros(48, device=dev)
  sec = _timed(dev, lambda: ops.conv2d_same_bf16(x2, w2, b2, stride=2), flush=flush)
  entry('tc_rows_kernel<2,12> conv2 fwd', sec, (P1 * 32 + P2 * 48) * 2, 2.0 * P2 * 48 * 288, launches=2)
  g2 = (torch.rand((N3, 128, 128, 48), device=dev) - 0.5).to(torch.bfloat16)
  sec = _timed(dev, lambda: ops.conv2d_same_bwd_bf16(x2, w2, g2, stride=2, need_dx=False), flush=flush)
  wg2 = entry('tc_wgrad_kernel<8,512> conv2 wgrad', sec, (P1 * 32 + P2 * 48) * 2, 2.0 * P2 * 48 * 288, launches=2)
  # conv2 data gradient as a separate kernel (what the fused backward replaces; other image sizes still run it):
  # G2 in, 1-bit ReLU mask of y1 in, G1 out
  bits1 = ops.relu_mask_bits(x2)
  sec = _timed(dev, lambda: ops.conv2d_same_bwd_bf16(x2, w2, g2, stride=2, relu_mask_bits=bits1, need_dx=True, need_dw=False),
               flush=flush)
  dom = entry('tc_rows_kernel<1,8,MASKBITS> conv2 dgrad', sec, (P2 * 48 + P1 * 32) * 2 + P1 * 4, 2.0 * P2 * 48 * 288,
              launches=5)
  # the same with the bf16 activation as the mask (stand-alone C-ABI default): 0.8 GB more to read
  sec = _timed(dev, lambda: ops.conv2d_same_bwd_bf16(x2, w2, g2, stride=2, relu_mask_x=x2, need_dx=True, need_dw=False),
               flush=flush)
  entry('tc_rows_kernel<1,8,MASK> conv2 dgrad (bf16 mask)', sec, (P2 * 48 + 2 * P1 * 32) * 2, 2.0 * P2 * 48 * 288, launches=5)
  # a tensor-bound layer for the tensor-pipe view: conv5 (128 -> 192, 32x32 -> 16x16)
  x5 = torch.rand((N3, 32, 32, 128), device=dev).to(torch.bfloat16)
  w5 = ((torch.rand((3, 3, 128, 192), device=dev) - 0.5) * 0.1)
  b5 = torch.zeros(192, device=dev)
  sec = _timed(dev, lambda: ops.conv2d_same_bf16(x5, w5, b5, stride=2), flush=flush)
  e5 = entry('tc_nn_kernel<8,0> conv5 fwd', sec, N3 * (1024 * 128 + 256 * 192) * 2 + 192 * 1152 * 2, 2.0 * N3 * 256 * 192 * 1152,
             launches=2)
  e5['bound'] = 'tensor'
  # tensor-pipe view of every layer above the ridge (conv4-conv8), all three passes, as stand-alone ops on 3 x batch
  # images (each = the tcgen05 kernel + its few-microsecond weight repack / split reduce): FLOPs / time against the
  # measured bf16 burst.  conv7 / conv8 hold 2 % of the FLOPs in 2-30 us launches: latency, not tensor, bound.
  Hin, Cin = 64, 64
  for li, Cout in ((4, 128), (5, 192), (6, 256), (7, 256), (8, 256)):
    Ho = Hin // 2
    x = torch.rand((N3, Hin, Hin, Cin), device=dev).to(torch.bfloat16)
    w = ((torch.rand((3, 3, Cin, Cout), device=dev) - 0.5) * 0.1)
    b = torch.zeros(Cout, device=dev)
    g = (torch.rand((N3, Ho, Ho, Cout), device=dev) - 0.5).to(torch.bfloat16)
    bits = ops.relu_mask_bits(x)
    flops = 2.0 * N3 * Ho * Ho * Cout * 9 * Cin
    act = N3 * (Hin * Hin * Cin + Ho * Ho * Cout) * 2 + 9 * Cin * Cout * 2
    for tag, fn in (('fwd', lambda: ops.conv2d_same_bf16(x, w, b, stride=2)),
                    ('dgrad', lambda: ops.conv2d_same_bwd_bf16(x, w, g, stride=2, relu_mask_bits=bits, need_dx=True, need_dw=False)),
                    ('wgrad', lambda: ops.conv2d_same_bwd_bf16(x, w, g, stride=2, need_dx=False))):
      if li == 5 and tag == 'fwd':
        continue                                            # timed above
      sec = _timed(dev, fn, flush=flush)
      e = entry('conv%d %s' % (li, tag), sec, act, flops, launches=2)
      e['bound'] = 'tensor'
    del x, w, b, g, bits
    Hin, Cin = Ho, Cout
  # `roofline` = the longest kernel the step actually runs: the fused conv1 -> conv2 forward, conv2's (row-wise) weight
  # gradient or the fused backward (the separate kernels only where the fused ones do not apply)
  if fused is not None:
    dom = max([e for e in (fused, rows_wg if rows_wg is not None else wg2, fused_bwd) if e is not None], key=lambda e: e['ms'])
  return dom, out


def run_extras(eng, dev, peaks, args, full=False):
  """BASELINE config 4 (batched closed-loop policy steps) and config 5 (rank-pooling sweep).
  full=False: one chunk of 1024 environments and six sweep points (a second or two, part of every 1-GPU line);
  full=True: all 4096 environments (4 chunks) and the whole K x resolution sweep."""
  import torch
  from geeco_b200 import ops
  from geeco_b200.predictor import BatchedGoalPredictor
  res = {}
  # ---- config 5: dynimg K=2..16 x {128,256,512} px, inputs >= 1 GiB
  sweep = []
  points = [(px, K) for px in (128, 256, 512) for K in (2, 3, 4, 6, 8, 12, 16)] if full else [
      (128, 4), (256, 3), (256, 4), (256, 8), (256, 16), (512, 4)]
  for px, K in points:
    per = K * px * px * 3 * 4
    n = min(max(2, int(2 ** 30 // per) + 1), 60000)
    x = torch.rand((n, K, px, px, 3), device=dev)
    y = torch.empty((n, px, px, 3), device=dev)
    sec = _timed(dev, lambda: ops.dynimg(x, out=y), reps=3)
    gbs = (K + 1) * px * px * 3 * 4 * n / sec / 1e9
    sweep.append({'px': px, 'K': K, 'N': n, 'GBps': gbs, 'frac_of_hbm_peak': gbs / peaks['hbm_gbs']})
    del x, y
  res['rankpool_sweep'] = sweep
  # ---- config 4: 4096 environments as chunks of 1024; K-frame histories in a device ring of uint8 frames (what the
  # hook renders, gym_pickplace.py:869-872), LSTM state carried across control steps, per-environment resets on the
  # device.  Every control step pushes one NEW frame per environment and runs the forward.
  total_envs, chunk = (4096, 1024) if full else (1024, 1024)
  cfg = eng.cfg
  bp = BatchedGoalPredictor(cfg, chunk, precision=args.precision, carry_state=True, frame_dtype='uint8')
  bp.engine.theta.copy_(eng.theta); bp.engine.params_changed()
  bp.set_goal(torch.randint(0, 256, (chunk, 256, 256, 3), dtype=torch.uint8, device=dev))
  frames = [torch.randint(0, 256, (chunk, 256, 256, 3), dtype=torch.uint8, device=dev) for _ in range(3)]
  jn = torch.rand((chunk, 7), device=dev)
  lib = bp.engine.lib
  for i in range(4):
    bp.predict_batch(frames[i % 3], jn)
  torch.cuda.synchronize()
  lib.geeco_launch_count(1)
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  steps = 3
  e0.record()
  for s in range(steps):
    for c in range(total_envs // chunk):
      out = bp.predict_batch(frames[(s + c) % 3], jn)
  e1.record()
  torch.cuda.synchronize()
  sec = e0.elapsed_time(e1) * 1e-3 / steps
  res['policy_steps'] = {'metric': 'batched policy env-steps/s (BASELINE config 4)', 'envs': total_envs, 'chunk': chunk,
                         'ms_per_control_step': sec * 1e3, 'env_steps_per_s': total_envs / sec,
                         'tflops': total_envs * 3.415e9 / sec / 1e12,
                         'tensor_frac_of_sustained': total_envs * 3.415e9 / sec / 1e12 / peaks['bf16_tflops_sustained'],
                         'carry_state': True, 'frame_dtype': 'uint8', 'ring': 'device ring, index rotation',
                         'gpu_launches_per_control_step': int(lib.geeco_launch_count(1)) // steps,
                         'note': 'MuJoCo stepping excluded; frames synthetic, device resident; ' + (
                             '4 chunks of 1024 environments' if full else
                             'one chunk of 1024 environments (the 4096-environment job is 4 such chunks back to back: --extras)')}
  del bp
  if full:
    for d in ('profiles', 'gpurun_out'):                   # gpurun_out/ is what travels back from the GPU box
      os.makedirs(os.path.join(ROOT, d), exist_ok=True)
      with open(os.path.join(ROOT, d, 'r02_extras.json'), 'w') as fp:
        json.dump(res, fp, indent=1)
  return res


_REAL_STDOUT = None


def _quiet_stdout():
  """stdout carries exactly ONE line (the JSON result): everything libraries print there (NCCL's version banner,
  ...) is routed to stderr; emit() writes to the saved descriptor."""
  global _REAL_STDOUT
  sys.stdout.flush()
  _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
  os.dup2(2, 1)


def emit(obj):
  out = _REAL_STDOUT or sys.stdout
  out.write(json.dumps(obj) + '\n')
  out.flush()


def main():
  args = parse_args()
  _quiet_stdout()
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_ours(args)


if __name__ == '__main__':
  main()
