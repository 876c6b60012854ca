#!/usr/bin/env python
"""Training script for end-to-end visuomotor controllers on the B200-native engine.

Same command line as the reference's scripts/train_e2evmc.py:22-124 (every flag is accepted with the same
name, type and default), same run-directory protocol (:213-252: `<ts>-runcmd.json`, `e2evmc_config.json`
which, when it already exists, overrides the model flags), same epoch loop (:288-291: train -> evaluate ->
best-k snapshot export, :143-205).

`--dataset_dir <dir>` reads a recorded dataset (meta/meta_info.json, data/*.tfrecord.zlib, splits/<split>/) through
the native input pipeline (geeco_b200/input_pipeline.py over libgeeco_io.so: same stages as the reference's
pickplace_input_fn, :266-278; frames travel as the recorded bytes and are divided by 255 on the device; only
full batches are fed because the engine, like the reference's lstm_memory variable, has a static batch size).
No dataset ships offline: `--dataset_dir synthetic[:<episodes>]` streams synthetic episodes that follow the
pipeline's layout and index contract (geeco_b200/data.py) without touching the disk.
Launch under torchrun for data-parallel training (one process per GPU, NCCL gradient all-reduce).
"""
import argparse
import json
import os
import pprint
import re
import shutil
import sys
from stat import ST_CTIME

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from geeco_b200.params import create_e2evmc_config, load_model_config, save_model_config  # noqa: E402
from geeco_b200.runscript import save_run_command  # noqa: E402

ARGPARSER = argparse.ArgumentParser(description='Train E2E VMC.')
_A = ARGPARSER.add_argument
# --- directory parameters
_A('--dataset_dir', type=str, default='../data/gym-pick-pad2-cube2-v4')
_A('--split_name', type=str, default='default')
_A('--model_dir', type=str, default='../tmp/models/geeco-f')
# --- model parameters
_A('--observation_format', type=str, default='rgb', help='rgb | rgbd')
_A('--control_mode', type=str, default='cartesian', help='cartesian | velocity')
_A('--goal_condition', type=str, default='none', help='none | target')
_A('--window_size', type=int, default=4)
_A('--dim_h_lstm', type=int, default=128)
_A('--dim_h_fc', type=int, default=128)
_A('--dim_s_obs', type=int, default=256)
_A('--dim_s_dyn', type=int, default=256)
_A('--dim_s_diff', type=int, default=256)
_A('--proc_obs', type=str, default='sequence', help='sequence | dynimg')
_A('--proc_tgt', type=str, default='constant', help='constant | residual | dyndiff')
_A('--l2_regularizer', type=float, default=0.0)
_A('--lambda_aux', type=float, default=1.0)
# --- data parameters
_A('--data_encoding', type=str, default='v4')
# --- training parameters
_A('--lr', type=float, default=1e-4)
_A('--train_epochs', type=int, default=10)
# --- snapshot management
_A('--ckpt_steps', type=int, default=10000)
_A('--num_last_ckpt', type=int, default=2)
_A('--num_best_ckpt', type=int, default=5)
# --- memory management
_A('--batch_size', type=int, default=32)
_A('--memcap', type=float, default=0.8)
_A('--num_threads', type=int, default=4)
_A('--prefetch_size', type=int, default=4)
_A('--shuffle_buffer', type=int, default=64)
# --- logging
_A('--log_steps', type=int, default=1000)
_A('--debug', default=False, action='store_true')
_A('--initial_eval', default=False, action='store_true')
# --- execution switches of this build (not in the reference)
_A('--precision', type=str, default='bf16', help='bf16 (tcgen05 tensor cores) | fp32')
_A('--cache_dir', type=str, default='', help='directory for decoded-episode caches (recorded datasets)')
_A('--device_frames', default=False, action='store_true',
   help='recorded datasets: upload every episode once and gather the K-frame windows on the device')
_A('--checkpoint_format', type=str, default='npz', help='npz | bundle (TF V2 .index/.data files)')
_A('--frame_layout', type=str, default='pool',
   help='recorded datasets, host batches: pool (distinct frames + index: ~window_size x fewer bytes to upload) | windows')

_OBSERVATION_FORMAT_TO_CHANNELS = {'rgb': 3, 'rgbd': 4}


def export_snapshot(model_dir, eval_results, num_best_ckpt):
  """Keeps the `num_best_ckpt` best checkpoints (by eval loss) under <model_dir>/snapshots/ with a JSON index
  {name: {step, loss, dir}}; evicts the worst when the slots are exceeded (train_e2evmc.py:143-205)."""
  from geeco_b200.estimator import latest_checkpoint
  snap_root = os.path.join(model_dir, 'snapshots')
  os.makedirs(snap_root, exist_ok=True)
  index_path = os.path.join(snap_root, 'snapshot_index.json')
  index = {}
  if os.path.exists(index_path):
    with open(index_path) as fp:
      index = json.load(fp)
  ckpt_name = os.path.basename(latest_checkpoint(model_dir))
  step = int(re.search(r'\d+', ckpt_name).group(0))
  loss = float(eval_results['loss'])

  def newest(suffix):
    files = [os.path.join(model_dir, f) for f in os.listdir(model_dir) if f.endswith(suffix)]
    return max(files, key=lambda f: os.stat(f)[ST_CTIME])

  dst = os.path.join(snap_root, ckpt_name)
  os.makedirs(dst, exist_ok=True)
  for path in (newest('runcmd.json'), newest('config.json')):
    shutil.copy(src=path, dst=dst)
  for f in os.listdir(model_dir):
    if f.startswith(ckpt_name):
      shutil.copy(src=os.path.join(model_dir, f), dst=dst)
  with open(os.path.join(dst, 'checkpoint'), 'w') as fp:
    fp.write('model_checkpoint_path: "%s"\n' % ckpt_name)
  index[ckpt_name] = {'step': step, 'loss': loss, 'dir': dst}
  if len(index) > num_best_ckpt:
    worst = max(index.items(), key=lambda kv: kv[1]['loss'])[0]
    shutil.rmtree(index[worst]['dir'], ignore_errors=True)
    index.pop(worst)
  with open(index_path, 'w') as fp:
    json.dump(index, fp, indent=2, sort_keys=True)
  return dst


def synthetic_input_fn(spec, config, batch_size, mode, rank=0, world=1):
  """Synthetic stand-in for pickplace_input_fn: `episodes` episodes of length 100 -> (L-1-K+1) windows each,
  batched in stream order (no window-level shuffle), each rank taking its contiguous share of a batch."""
  from geeco_b200.data import EPISODE_LENGTH, num_windows, synthetic_batch
  episodes = int(spec.split(':', 1)[1]) if ':' in spec else 2
  per_episode = num_windows(EPISODE_LENGTH, config.window_size)
  total = episodes * per_episode
  global_batch = batch_size * world

  def gen():
    seed0 = 0 if mode == 'train' else 10 ** 6
    for b in range(total // global_batch):
      lo = b * global_batch + rank * batch_size
      yield synthetic_batch(batch_size, window_size=config.window_size, height=config.img_height,
                            width=config.img_width, channels=config.img_channels, seed=seed0 + lo,
                            first_stream_pos=lo)
  return gen


def recorded_input_fn(args, config, mode, rank=0, world=1):
  """pickplace_input_fn over a recorded dataset with the arguments of the reference's call (:266-278)."""
  from geeco_b200.input_pipeline import pickplace_input_fn
  epoch = [0]

  def make():
    epoch[0] += 1
    # every rank must shuffle the episodes identically: a per-epoch seed replaces numpy's global generator
    return pickplace_input_fn(
        dataset_dir=args.dataset_dir, split_name=args.split_name, mode=mode, encoding=args.data_encoding,
        window_size=config.window_size, fetch_target=(args.goal_condition == 'target'),
        shuffle_buffer=args.shuffle_buffer, batch_size=config.batch_size, num_epochs=1,
        num_threads=args.num_threads, prefetch_size=args.prefetch_size, seed=epoch[0] if world > 1 else None,
        # RGB-D frames share one float tensor with the depth channel (estimator.py:166-172): no uint8 wire format
        frame_format='float32' if args.observation_format == 'rgbd' else 'uint8',
        drop_remainder=True, rank=rank, world=world, pin_memory=not args.device_frames,
        cache_dir=args.cache_dir or None, want_depth=(args.observation_format == 'rgbd'),
        device='cuda' if args.device_frames else None,
        layout='windows' if args.device_frames else args.frame_layout)
  return make


def main(args, argv=None):
  from geeco_b200 import parallel
  from geeco_b200.estimator import Estimator, RunConfig, e2evmc_model_fn, goal_e2evmc_model_fn
  # train_e2evmc.py:127-130, :258: the model_fn follows --goal_condition (KeyError for anything else)
  model_fn = {'none': e2evmc_model_fn, 'target': goal_e2evmc_model_fn}[args.goal_condition]
  # the model config: an existing run directory overrides the model flags (train_e2evmc.py:229-232)
  config_name = 'e2evmc_config'
  have_config = os.path.exists(os.path.join(args.model_dir, config_name + '.json'))
  if have_config:
    e2evmc_config = create_e2evmc_config(load_model_config(args.model_dir, config_name))
    print(">>> Loaded existing model config from %s" % (args.model_dir,))
  else:
    e2evmc_config = create_e2evmc_config({
        'img_channels': _OBSERVATION_FORMAT_TO_CHANNELS[args.observation_format],
        'control_mode': args.control_mode, 'window_size': args.window_size, 'dim_h_lstm': args.dim_h_lstm,
        'dim_h_fc': args.dim_h_fc, 'dim_s_obs': args.dim_s_obs, 'dim_s_dyn': args.dim_s_dyn,
        'dim_s_diff': args.dim_s_diff, 'proc_obs': args.proc_obs, 'proc_tgt': args.proc_tgt,
        'l2_regularizer': args.l2_regularizer, 'lambda_aux': args.lambda_aux, 'batch_size': args.batch_size,
        'lr': args.lr})
  # everything that can be refused is refused before the run directory is touched
  from geeco_b200.engine import _check_switches
  _check_switches(e2evmc_config, args.goal_condition)
  if not args.dataset_dir.startswith('synthetic') and not os.path.isdir(args.dataset_dir):
    raise FileNotFoundError("--dataset_dir %s does not exist (use synthetic[:<episodes>] for synthetic episodes)"
                            % (args.dataset_dir,))
  import torch
  import torch.distributed as dist
  if not torch.cuda.is_available():
    raise RuntimeError("geeco_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
  parallel.tune_for_data_parallel(int(os.environ.get('WORLD_SIZE', '1')))
  if int(os.environ.get('WORLD_SIZE', '1')) > 1 and not dist.is_initialized():
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    parallel.pin_to_gpu_numa_node(int(os.environ.get('LOCAL_RANK', '0')), int(os.environ.get('LOCAL_WORLD_SIZE', '1')))
    dist.init_process_group('nccl')
  rank, world = parallel.world_info()
  os.makedirs(args.model_dir, exist_ok=True)
  if rank == 0:
    save_run_command(ARGPARSER, args.model_dir, argv)
    if not have_config:
      save_model_config(e2evmc_config._asdict(), args.model_dir, config_name)
  run_config = RunConfig(save_checkpoints_steps=args.ckpt_steps, keep_checkpoint_max=args.num_last_ckpt)
  estimator = Estimator(model_fn=model_fn, model_dir=args.model_dir, config=run_config,
                        params={'e2evmc_config': e2evmc_config, 'log_steps': args.log_steps, 'debug': args.debug,
                                'checkpoint_format': args.checkpoint_format},
                        precision=args.precision, batch_size=e2evmc_config.batch_size)
  if args.dataset_dir.startswith('synthetic'):
    train_input = synthetic_input_fn(args.dataset_dir, e2evmc_config, e2evmc_config.batch_size, 'train', rank, world)
    eval_input = synthetic_input_fn(args.dataset_dir, e2evmc_config, e2evmc_config.batch_size, 'eval', rank, world)
  else:
    train_input = recorded_input_fn(args, e2evmc_config, 'train', rank, world)
    eval_input = recorded_input_fn(args, e2evmc_config, 'eval', rank, world)
  results = []
  if args.initial_eval:
    results.append(estimator.evaluate(input_fn=eval_input))
  for _ in range(args.train_epochs):
    estimator.train(input_fn=train_input)
    eval_results = estimator.evaluate(input_fn=eval_input)
    results.append(eval_results)
    if rank == 0:
      export_snapshot(args.model_dir, eval_results, args.num_best_ckpt)
      pprint.pprint(eval_results)
  return results


if __name__ == '__main__':
  print(">>> Training E2E VMC.")
  PARSED_ARGS, UNPARSED_ARGS = ARGPARSER.parse_known_args()
  pprint.pprint(PARSED_ARGS)
  main(PARSED_ARGS)
