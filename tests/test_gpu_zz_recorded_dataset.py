"""GPU tests of the steps either side of the train step: recorded episodes -> native input
pipeline -> Estimator, and TF V2 bundle checkpoints -> Estimator / predictor."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import geeco_oracle as O

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def recorded(tmp_path_factory):
  from geeco_b200.data import write_synthetic_dataset
  d = str(tmp_path_factory.mktemp('recorded'))
  write_synthetic_dataset(d, episodes=1, episode_length=10, height=256, width=256, seed=2, eval_episodes=1)
  return d


def test_recorded_dataset_trains_like_host_batches(cuda_device, recorded, tmp_path):
  from geeco_b200 import create_e2evmc_config
  from geeco_b200.estimator import Estimator, RunConfig, goal_e2evmc_model_fn, latest_checkpoint
  from geeco_b200.input_pipeline import pickplace_input_fn_v4
  cfg_d = O.make_config(batch_size=2, lr=1e-3)
  cfg = create_e2evmc_config(cfg_d)

  def make(md, fmt, **kw):
    est = Estimator(goal_e2evmc_model_fn, md, RunConfig(), {'e2evmc_config': cfg, 'log_steps': 1, 'checkpoint_format': fmt},
                    precision='fp32', batch_size=2)
    inp = lambda: pickplace_input_fn_v4(recorded, 'default', 'train', 4, True, batch_size=2, drop_remainder=True, **kw)
    return est, inp

  # A: recorded bytes in pinned batches (the production path); B: the reference's float32 tensors in plain numpy
  estA, inpA = make(str(tmp_path / 'a'), 'bundle', frame_format='uint8', pin_memory=True)
  estB, inpB = make(str(tmp_path / 'b'), 'npz', frame_format='float32')
  P0 = {k: torch.tensor(v) for k, v in estA.engine.get_params().items()}
  first = next(iter(inpB()))
  ref_losses, _ = O.forward_losses(P0, first[0], first[1], cfg_d)
  estA.train(inpA)
  estB.train(inpB)
  assert estA.engine.global_step == estB.engine.global_step == 3            # (10 - 1) - 4 + 1 = 6 windows / 2
  assert torch.equal(estA.engine.theta, estB.engine.theta)                  # bit-identical training
  step1 = estA.last_train_losses[0][1]['loss']
  assert abs(step1 - float(ref_losses['loss'])) <= 1e-4 * abs(float(ref_losses['loss']))
  # checkpoints: A wrote a TF V2 bundle, B an npz; each restores into a fresh Estimator, to the same state
  pa, pb = latest_checkpoint(str(tmp_path / 'a')), latest_checkpoint(str(tmp_path / 'b'))
  assert os.path.exists(pa + '.index') and os.path.exists(pa + '.data-00000-of-00001') and os.path.exists(pb + '.npz')
  est2, _ = make(str(tmp_path / 'a'), 'bundle')
  assert est2.engine.global_step == 3 and torch.equal(est2.engine.theta, estA.engine.theta)
  assert torch.equal(est2.engine.adam_m, estA.engine.adam_m) and torch.equal(est2.engine.adam_v, estA.engine.adam_v)
  evA = estA.evaluate(lambda: pickplace_input_fn_v4(recorded, 'default', 'eval', 4, True, batch_size=2, drop_remainder=True,
                                                    frame_format='uint8'))
  ev2 = est2.evaluate(lambda: pickplace_input_fn_v4(recorded, 'default', 'eval', 4, True, batch_size=2, drop_remainder=True))
  assert evA['loss'] == ev2['loss'] and evA['global_step'] == 3


def test_train_cli_on_a_recorded_dataset_and_predictor_from_bundle(cuda_device, recorded, tmp_path):
  sys.path.insert(0, os.path.join(ROOT, 'scripts'))
  import importlib
  m = importlib.import_module('train_e2evmc')
  md = os.path.join(str(tmp_path), 'run')
  argv = ['--dataset_dir', recorded, '--split_name', 'default', '--model_dir', md, '--goal_condition', 'target',
          '--proc_obs', 'dynimg', '--proc_tgt', 'dyndiff', '--batch_size', '2', '--train_epochs', '2', '--lr', '1e-3',
          '--precision', 'bf16', '--checkpoint_format', 'bundle', '--cache_dir', os.path.join(str(tmp_path), 'cache'),
          '--num_threads', '2']
  res = m.main(m.ARGPARSER.parse_args(argv), argv)
  assert len(res) == 2 and res[1]['global_step'] == 6
  assert len(os.listdir(os.path.join(str(tmp_path), 'cache'))) == 2            # train + eval episode, decoded once
  files = os.listdir(md)
  assert 'model.ckpt-6.index' in files and 'model.ckpt-6.data-00000-of-00001' in files
  snap = os.listdir(os.path.join(md, 'snapshots'))
  assert 'model.ckpt-6' in snap and 'model.ckpt-6.index' in os.listdir(os.path.join(md, 'snapshots', 'model.ckpt-6'))
  # the controller hook loads the bundle (predictor.py:87-93) and predicts
  from geeco_b200.checkpoint import read_bundle
  from geeco_b200.predictor import GoalE2EVMCPredictor
  pred = GoalE2EVMCPredictor(md)
  stored = read_bundle(os.path.join(md, 'model.ckpt-6'))
  got = pred.sess.get_params()
  for k, v in got.items():
    np.testing.assert_array_equal(v, stored[k])
  rng = np.random.default_rng(0)
  pred.set_goal(rng.integers(0, 256, size=(256, 256, 3)) / 255.0)
  out = pred.predict(rng.integers(0, 256, size=(256, 256, 3)) / 255.0, rng.uniform(-1, 1, size=7).astype(np.float32))
  assert out['cmd_ee'].shape == (3,) and out['cmd_grp'].shape == (1,) and np.all(np.isfinite(out['cmd_ee']))


def test_device_resident_frames_train_like_host_batches(cuda_device, recorded, tmp_path):
  """device='cuda' pipeline mode: every episode is uploaded once, the windows are gathered on the device; training is
  bit-identical to the host-window batches and the upload is one frame per frame, not K."""
  from geeco_b200 import create_e2evmc_config
  from geeco_b200.estimator import Estimator, RunConfig, goal_e2evmc_model_fn
  from geeco_b200.input_pipeline import pickplace_input_fn_v4
  cfg = create_e2evmc_config(O.make_config(batch_size=2, lr=1e-3))
  thetas, pipes = [], []
  for name, kw in (('host', dict(pin_memory=True)), ('dev', dict(device='cuda'))):
    est = Estimator(goal_e2evmc_model_fn, str(tmp_path / name), RunConfig(), {'e2evmc_config': cfg, 'log_steps': 1},
                    precision='fp32', batch_size=2)

    def inp(kw=kw):
      pipes.append(pickplace_input_fn_v4(recorded, 'default', 'train', 4, True, batch_size=2, drop_remainder=True,
                                         frame_format='uint8', want_depth=False, num_epochs=2, **kw))
      return pipes[-1]
    est.train(inp)
    assert est.engine.global_step == 6
    thetas.append(est.engine.theta.clone())
  assert torch.equal(thetas[0], thetas[1])
  first = next(iter(pickplace_input_fn_v4(recorded, 'default', 'train', 4, True, batch_size=2, frame_format='uint8',
                                          want_depth=False, device='cuda')))[0]
  assert first['rgb'].is_cuda and first['rgb'].dtype == torch.uint8 and tuple(first['rgb'].shape) == (2, 4, 256, 256, 3)
  assert first['target_rgb'].is_cuda and not torch.is_tensor(first['jnt_state'])
  frame = 256 * 256 * 3
  assert pipes[1].uploaded_bytes == 2 * (9 + 1) * frame          # 2 epochs x (9 frames + the target), once each
