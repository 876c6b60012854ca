#!/usr/bin/env python
"""Throughput of training from RECORDED episodes, next to bench.py's synthetic-batch numbers.

  python tools/bench_recorded.py [--episodes 4] [--frames 100] [--batch 64] [--epochs 3] [--threads 8]

Writes a synthetic dataset in the recorder's format to a scratch directory (float-encoded pixels, one zlib stream per
episode: ~20 s of CPU per 100-frame episode to WRITE), then measures, each as samples/s over `--epochs` passes:

  cold          pickplace_input_fn alone, first pass (inflate + decode + cache write), no GPU
  warm_host     pipeline alone, warm episode cache, host windows (pinned when a GPU is present)
  warm_device_hostside  pipeline alone, device-resident mode with the gather stubbed out (what the host must sustain)
  train_host    Estimator.train over pinned uint8 host windows            (GPU only)
  train_device  Estimator.train over device-resident frames (`device='cuda'`) (GPU only)

GPU legs time whole epochs with a synchronize on both sides (the input pipeline is inside the timed region: that is the
point); the first epoch of each GPU leg is a warm-up.  One JSON line on stdout.
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--episodes', type=int, default=4)
  ap.add_argument('--frames', type=int, default=100)
  ap.add_argument('--batch', type=int, default=64)
  ap.add_argument('--epochs', type=int, default=3)
  ap.add_argument('--threads', type=int, default=min(8, os.cpu_count() or 1))
  ap.add_argument('--precision', default='bf16')
  args = ap.parse_args()
  import torch
  from geeco_b200 import create_e2evmc_config
  from geeco_b200.data import write_synthetic_dataset
  from geeco_b200.input_pipeline import pickplace_input_fn_v4
  real_stdout = sys.stdout
  sys.stdout = sys.stderr                                   # library banners must not pollute the JSON line
  d = tempfile.mkdtemp(prefix='geeco_recorded_')
  out = {'episodes': args.episodes, 'frames': args.frames, 'batch': args.batch, 'threads': args.threads,
         'host_cores': os.cpu_count(), 'unit': 'samples/s'}
  try:
    t0 = time.perf_counter()
    write_synthetic_dataset(d, episodes=args.episodes, episode_length=args.frames, eval_episodes=0)
    out['dataset_write_s'] = time.perf_counter() - t0
    cache = os.path.join(d, 'cache')
    cuda = torch.cuda.is_available()

    def pipe(**kw):
      return pickplace_input_fn_v4(d, 'default', 'train', 4, True, batch_size=args.batch, num_threads=args.threads,
                                   frame_format='uint8', drop_remainder=True, cache_dir=cache, want_depth=False, seed=1,
                                   **kw)

    def drain(it):
      t = time.perf_counter()
      n = sum(int(f['jnt_state'].shape[0]) for f, _ in it)
      return n / (time.perf_counter() - t)

    out['cold'] = drain(pipe())
    out['warm_host'] = drain(pipe(num_epochs=args.epochs, pin_memory=cuda))
    it = pipe(num_epochs=args.epochs, device='cpu')
    it._device_images = lambda feats, plan, episodes: feats
    out['warm_device_hostside'] = drain(it)
    if cuda:
      from geeco_b200.estimator import Estimator, RunConfig, goal_e2evmc_model_fn
      cfg = create_e2evmc_config(dict(proc_obs='dynimg', proc_tgt='dyndiff', batch_size=args.batch))
      for name, kw in (('train_host', dict(pin_memory=True)), ('train_device', dict(device='cuda'))):
        est = Estimator(goal_e2evmc_model_fn, os.path.join(d, name), RunConfig(save_checkpoints_steps=0),
                        {'e2evmc_config': cfg, 'log_steps': 10 ** 9, 'save_final_checkpoint': False},
                        precision=args.precision, batch_size=args.batch)
        est.train(lambda kw=kw: pipe(**kw))                 # warm-up epoch (kernels, allocator, page cache)
        torch.cuda.synchronize()
        step0, t = est.engine.global_step, time.perf_counter()
        est.train(lambda kw=kw: pipe(num_epochs=args.epochs, **kw))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        out[name] = (est.engine.global_step - step0) * args.batch / dt
        out[name + '_ms_per_step'] = 1e3 * dt / max(est.engine.global_step - step0, 1)
    else:
      out['gpu'] = 'no CUDA device: pipeline legs only'
  finally:
    shutil.rmtree(d, ignore_errors=True)
    sys.stdout = real_stdout
  print(json.dumps(out))


if __name__ == '__main__':
  main()
