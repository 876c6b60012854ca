"""GPU tests of the drop-in boundary: predictor hook, Estimator train/evaluate/checkpoint cycle, train CLI."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from oracle import geeco_oracle as O
from tests.util import rel_max

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model_dir(tmp_path, P, cfg_d):
  """A run directory as train_e2evmc.py leaves it: e2evmc_config.json + a checkpoint."""
  from geeco_b200 import create_e2evmc_config, save_model_config
  md = str(tmp_path)
  save_model_config(create_e2evmc_config(cfg_d)._asdict(), md, 'e2evmc_config')
  arrays = {k: v.numpy() for k, v in P.items()}
  arrays['global_step'] = np.array(123, dtype=np.int64)
  np.savez(os.path.join(md, 'model.ckpt-123.npz'), **arrays)
  with open(os.path.join(md, 'checkpoint'), 'w') as fp:
    fp.write('model_checkpoint_path: "model.ckpt-123"\n')
  return md


def test_predictor_hook_matches_oracle(cuda_device, tmp_path):
  from geeco_b200.predictor import GoalE2EVMCPredictor
  cfg_d = O.make_config(batch_size=32)
  P = O.init_params(cfg_d, seed=7, dtype=torch.float32, bias_scale=0.05)
  pred = GoalE2EVMCPredictor(_model_dir(tmp_path, P, cfg_d))
  assert pred.cfg.batch_size == 1 and pred.cfg.proc_obs == 'dynimg' and pred.cfg.proc_tgt == 'dyndiff'
  rng = np.random.default_rng(0)
  goal = rng.integers(0, 256, size=(256, 256, 4)) / 255.0          # RGB-D goal: channels are cropped (predictor.py:208)
  pred.reset()
  pred.set_goal(goal)
  frames = [rng.integers(0, 256, size=(256, 256, 3)) / 255.0 for _ in range(6)]     # float64 like gym_pickplace.py:869-872
  jnts = [rng.uniform(-np.pi, np.pi, size=7).astype(np.float32) for _ in range(6)]
  fifo = []
  cfg1 = dict(cfg_d, batch_size=1)
  for t in range(1, 6):              # t = 0 would be the degenerate all-equal buffer
    if t == 1:
      out0 = pred.predict(frames[0], jnts[0])        # first step: buffer padded with copies
      assert out0['cmd_grp'].shape == (1,) and out0['cmd_grp'].dtype == np.float32
      assert np.all(out0['dyndiff'] >= 0) and out0['dynbuff'].shape == (256, 256, 3)
      fifo = [(frames[0], jnts[0])] * 4
    out = pred.predict(frames[t], jnts[t])
    fifo = (fifo + [(frames[t], jnts[t])])[-4:]
    rgb = torch.tensor(np.stack([f for f, _ in fifo])[None].astype(np.float32))
    jn = torch.tensor(np.stack([j for _, j in fifo])[None])
    _, ep = O.goal_e2evmc(rgb, jn, torch.tensor(goal[None, :, :, :3].astype(np.float32)), P, cfg1)
    ref = O.predictor_postprocess(ep)
    assert set(out.keys()) == {'cmd_ee', 'cmd_grp', 'pos_ee', 'pos_obj', 'dynbuff', 'dyndiff'}
    for k in ('cmd_ee', 'pos_ee', 'pos_obj'):
      assert out[k].shape == (3,) and rel_max(out[k], ref[k]) <= 1e-4, (t, k)
    assert out['cmd_grp'][0] == ref['cmd_grp'][0] and out['cmd_grp'][0] in (-1.0, 0.0, 1.0)
    assert np.abs(out['dynbuff'] - ref['dynbuff']).max() <= 1e-5 and np.abs(out['dyndiff'] - ref['dyndiff']).max() <= 1e-5
  with pytest.raises(AssertionError):
    pred.predict(frames[0][:128], jnts[0])
  with pytest.raises(AssertionError):
    pred.predict(frames[0] * 2.0, jnts[0])


def test_estimator_cycle_and_checkpoints(cuda_device, tmp_path):
  from geeco_b200 import create_e2evmc_config
  from geeco_b200.data import synthetic_batch
  from geeco_b200.estimator import Estimator, RunConfig, goal_e2evmc_model_fn, latest_checkpoint
  cfg = create_e2evmc_config(O.make_config(batch_size=4, lr=1e-3))
  batches = [synthetic_batch(4, seed=s) for s in range(3)]
  inp = lambda: iter(batches)
  md = str(tmp_path)
  est = Estimator(goal_e2evmc_model_fn, md, RunConfig(save_checkpoints_steps=2, keep_checkpoint_max=2),
                  {'e2evmc_config': cfg, 'log_steps': 1, 'debug': False}, precision='fp32', batch_size=4)
  ev0 = est.evaluate(inp)
  assert set(ev0.keys()) == {'loss', 'cmd_ee', 'pos_ee', 'pos_obj', 'cmd_grp', 'global_step'} and ev0['global_step'] == 0
  est.train(inp)
  assert est.engine.global_step == 3
  assert os.path.basename(latest_checkpoint(md)) == 'model.ckpt-3'
  assert sorted(f for f in os.listdir(md) if f.endswith('.npz')) == ['model.ckpt-2.npz', 'model.ckpt-3.npz']
  with np.load(os.path.join(md, 'model.ckpt-3.npz')) as d:
    assert 'GoalVMC/ConvEncoder/conv1/kernel' in d and 'GoalVMC/ConvEncoder/conv1/kernel/Adam_1' in d
    assert np.all(d['GoalVMC/LSTMDecoder/lstm_memory'] == 0)
  ev1 = est.evaluate(inp)
  # the oracle agrees on the streaming metrics for the restored parameters
  P = {k: torch.tensor(v) for k, v in est.engine.get_params().items()}
  se, cnt, loss = 0.0, 0, 0.0
  for f, l in batches:
    losses, ep = O.forward_losses(P, f, l, O.make_config(batch_size=4))
    m = O.eval_metrics_batch(ep, f, l, O.make_config(batch_size=4))
    se += m['cmd_ee'][0]; cnt += m['cmd_ee'][1]; loss += float(losses['loss'])
  assert abs(ev1['cmd_ee'] - se / cnt) <= 1e-4 * (se / cnt) and abs(ev1['loss'] - loss / 3) <= 1e-4 * (loss / 3)
  # a second Estimator on the same directory resumes from the checkpoint (TF Estimator semantics)
  est2 = Estimator(goal_e2evmc_model_fn, md, RunConfig(), {'e2evmc_config': cfg}, precision='fp32', batch_size=4)
  assert est2.engine.global_step == 3
  assert torch.equal(est2.engine.theta, est.engine.theta) and torch.equal(est2.engine.adam_v, est.engine.adam_v)
  ev2 = est2.evaluate(inp)
  assert ev2['loss'] == ev1['loss']
  with pytest.raises(ValueError):
    est.train(lambda: iter([synthetic_batch(3, seed=0)]))


def test_train_cli_end_to_end(cuda_device, tmp_path):
  sys.path.insert(0, os.path.join(ROOT, 'scripts'))
  import importlib
  m = importlib.import_module('train_e2evmc')
  md = os.path.join(str(tmp_path), 'run')
  argv = ['--dataset_dir', 'synthetic:1', '--model_dir', md, '--observation_format', 'rgb', '--goal_condition', 'target',
          '--proc_obs', 'dynimg', '--proc_tgt', 'dyndiff', '--batch_size', '32', '--train_epochs', '2', '--lr', '1e-3',
          '--precision', 'bf16']
  res = m.main(m.ARGPARSER.parse_args(argv), argv)
  assert len(res) == 2 and res[1]['global_step'] == 6           # 96 windows / 32 = 3 steps per epoch
  assert res[1]['loss'] < res[0]['loss']                          # it learns the (fixed) synthetic epoch
  files = os.listdir(md)
  assert 'e2evmc_config.json' in files and 'checkpoint' in files and any(f.endswith('-runcmd.json') for f in files)
  with open(os.path.join(md, 'snapshots', 'snapshot_index.json')) as fp:
    assert len(json.load(fp)) == 2
  # a later run on the same directory ignores the model flags and reloads the stored config (train_e2evmc.py:229-232)
  argv2 = [a if a != '256' else a for a in argv] + ['--dim_h_lstm', '64', '--train_epochs', '1']
  res2 = m.main(m.ARGPARSER.parse_args(argv2), argv2)
  assert res2[0]['global_step'] == 9
  with open(os.path.join(md, 'e2evmc_config.json')) as fp:
    assert json.load(fp)['dim_h_lstm'] == 128


def test_batched_predictor_equals_single_predictor(cuda_device, tmp_path):
  from geeco_b200 import create_e2evmc_config
  from geeco_b200.engine import Engine
  from geeco_b200.predictor import BatchedGoalPredictor
  cfg_d = O.make_config(batch_size=3)
  cfg = create_e2evmc_config(cfg_d)
  P = O.init_params(cfg_d, seed=9, dtype=torch.float32, bias_scale=0.05)
  bp = BatchedGoalPredictor(cfg, 3, precision='fp32')
  bp.engine.set_params(P)
  rng = np.random.default_rng(1)
  goals = (rng.integers(0, 256, size=(3, 256, 256, 3)) / 255.0).astype(np.float32)
  bp.set_goal(goals)
  hist = []
  for t in range(3):
    fr = (rng.integers(0, 256, size=(3, 256, 256, 3)) / 255.0).astype(np.float32)
    jn = rng.uniform(-1, 1, size=(3, 7)).astype(np.float32)
    hist.append((fr, jn))
    out = bp.predict_batch(fr, jn)
  fifo = ([hist[0]] * 4 + hist[1:])[-4:]
  rgb = torch.tensor(np.stack([f for f, _ in fifo], axis=1))
  jn = torch.tensor(np.stack([j for _, j in fifo], axis=1))
  _, ep = O.goal_e2evmc(rgb, jn, torch.tensor(goals), P, cfg_d)
  assert rel_max(out['cmd_ee'].cpu().numpy(), ep['pred_cmd_ee'].numpy()) <= 1e-4
  assert np.array_equal(out['cmd_grp'].cpu().numpy(), (ep['logits_cmd_grp'].argmax(dim=1) - 1).float().numpy())


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_uint8_frames_equal_host_divided_frames(cuda_device, precision):
  """Recorded uint8 frames divided by 255 on the device == the input pipeline's host division (geeco_gym.py:310):
  same network input bit for bit, hence the same step; covers pinned, pageable and device-resident callers."""
  from geeco_b200 import create_e2evmc_config
  from geeco_b200.data import synthetic_batch
  from geeco_b200.engine import Engine
  N = 2
  cfg_d = O.make_config(batch_size=N)
  P = O.init_params(cfg_d, seed=3, dtype=torch.float32, bias_scale=0.05)
  f8, labels = synthetic_batch(N, seed=5, frame_format='uint8')
  f32 = dict(f8)
  for k in ('rgb', 'target_rgb'):
    assert f8[k].dtype == np.uint8
    f32[k] = f8[k].astype(np.float32) / np.float32(255.0)
  res = {}
  for tag, feats in (('f32', f32), ('u8', f8), ('u8_cuda', {k: (torch.from_numpy(v).cuda() if k in ('rgb', 'target_rgb')
                                                                else v) for k, v in f8.items()})):
    eng = Engine(create_e2evmc_config(cfg_d), batch_size=N, precision=precision, training=True)
    eng.set_params(P)
    out = eng.forward(feats, labels, want_dyn=True)
    torch.cuda.synchronize()
    x0 = eng.debug_buffer('x0')
    x0 = (x0.view(torch.int16) if x0.dtype == torch.bfloat16 else x0.view(torch.int32)).cpu().numpy().copy()
    fwd = {k: out[k].cpu().numpy().copy() for k in ('pred_cmd_ee', 'logits_cmd_grp', 'dynbuff', 'dyndiff', 'losses')}
    eng.train_step(feats, labels)
    torch.cuda.synchronize()
    res[tag] = (x0, fwd, eng.get_grads(), eng.get_params())
    eng.close()
  for tag in ('u8', 'u8_cuda'):
    assert np.array_equal(res[tag][0], res['f32'][0]), tag
    for k, v in res['f32'][1].items():
      assert np.array_equal(res[tag][1][k], v), (tag, k)
    for k, v in res['f32'][2].items():
      assert np.array_equal(res[tag][2][k], v), (tag, k)
    for k, v in res['f32'][3].items():
      assert np.array_equal(res[tag][3][k], v), (tag, k)
  # oracle on the host-divided frames (the division itself: float32(u) / 255.0f)
  ref_losses, _, ep = O.train_step({k: v.clone() for k, v in P.items()}, O.adam_init(P), f32, labels, cfg_d)
  tol = 1e-4 if precision == 'fp32' else 2e-2
  assert abs(float(res['u8'][1]['losses'][5]) - ref_losses['loss']) <= tol * abs(ref_losses['loss'])
  assert np.abs(res['u8'][1]['dynbuff'] - ep['dynbuff'].detach().numpy()).max() <= 1e-5


def test_mixed_frame_formats_are_rejected(cuda_device):
  from geeco_b200 import create_e2evmc_config
  from geeco_b200.data import synthetic_batch
  from geeco_b200.engine import Engine
  cfg_d = O.make_config(batch_size=1)
  f8, labels = synthetic_batch(1, seed=5, frame_format='uint8')
  f8['target_rgb'] = f8['target_rgb'].astype(np.float32) / np.float32(255.0)
  eng = Engine(create_e2evmc_config(cfg_d), batch_size=1, precision='fp32', training=False)
  eng.init_params(seed=0)
  with pytest.raises(ValueError):
    eng.forward(f8, None)
  eng.close()


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_rgbd_observation_format(cuda_device, precision):
  """--observation_format rgbd: model_fn concatenates depth behind RGB (estimator.py:164-172), conv1 sees 4 channels."""
  from geeco_b200 import create_e2evmc_config
  from geeco_b200.data import synthetic_batch
  from geeco_b200.engine import Engine
  from geeco_b200.estimator import ModeKeys, goal_e2evmc_model_fn
  N = 2
  cfg_d = O.make_config(batch_size=N, img_channels=4)
  cfg = create_e2evmc_config(cfg_d)
  P = O.init_params(cfg_d, seed=4, dtype=torch.float32, bias_scale=0.05)
  assert tuple(P['GoalVMC/ConvEncoder/conv1/kernel'].shape) == (3, 3, 4, 32)
  feats, labels = synthetic_batch(N, seed=6)
  rng = np.random.default_rng(7)
  feats['depth'] = rng.uniform(0, 1, size=feats['rgb'].shape[:-1] + (1,)).astype(np.float32)
  feats['target_depth'] = rng.uniform(0, 1, size=feats['target_rgb'].shape[:-1] + (1,)).astype(np.float32)
  eng = Engine(cfg, batch_size=N, precision=precision, training=True)
  eng.set_params(P)
  spec = goal_e2evmc_model_fn(feats, labels, ModeKeys.EVAL, {'e2evmc_config': cfg, 'engine': eng})
  torch.cuda.synchronize()
  f4 = dict(feats)
  f4['rgb'] = np.concatenate([feats['rgb'], feats['depth']], axis=-1)
  f4['target_rgb'] = np.concatenate([feats['target_rgb'], feats['target_depth']], axis=-1)
  ref_losses, ref_grads, ep = O.train_step({k: v.clone() for k, v in P.items()}, O.adam_init(P), f4, labels, cfg_d,
                                           emulate_bf16=(precision == 'bf16'))
  tol = 1e-4 if precision == 'fp32' else 2e-2
  got = eng.losses_dict(spec.loss)
  assert abs(got['loss'] - ref_losses['loss']) <= tol * abs(ref_losses['loss'])
  for k in ('pred_cmd_ee', 'pred_aux_ee', 'pred_aux_obj'):
    assert rel_max(spec.endpoints[k].cpu().numpy(), ep[k].detach().numpy()) <= tol, k
  spec = goal_e2evmc_model_fn(feats, labels, ModeKeys.TRAIN, {'e2evmc_config': cfg, 'engine': eng})
  torch.cuda.synchronize()
  g = eng.get_grads()
  gtol = 1e-4 if precision == 'fp32' else 1e-1
  for k in ('GoalVMC/ConvEncoder/conv1/kernel', 'GoalVMC/DynDiffEncoder/conv1/kernel', 'GoalVMC/LSTMDecoder/lstm_cell/kernel'):
    assert rel_max(g[k], ref_grads[k].numpy()) <= gtol, k
  eng.close()
