"""CPU-only tests of the host side: config record, input index contract, C-ABI surface, gloo path."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config_record_matches_reference_defaults(tmp_path):
  from geeco_b200 import (E2E_VMC_DEFAULT_CONFIG, E2E_VMC_DEFAULT_PARAM_DICT, create_e2evmc_config,
                          load_model_config, save_model_config)
  # params.py:7-28 of the reference
  expect = dict(img_height=256, img_width=256, img_channels=3, dim_jnt_state=7, dim_grp_command=2,
                control_mode='cartesian', num_grp_states=3, dim_action=4, proc_obs='sequence', proc_tgt='constant',
                dim_s_obs=256, dim_s_dyn=256, dim_s_diff=256, dim_h_lstm=128, dim_h_fc=128, window_size=4,
                l2_regularizer=0.0, lambda_aux=1.0, batch_size=32, lr=1e-4)
  assert E2E_VMC_DEFAULT_PARAM_DICT == expect
  assert list(E2E_VMC_DEFAULT_CONFIG._fields) == list(expect.keys())
  cfg = create_e2evmc_config({'proc_obs': 'dynimg', 'unknown_key': 1, 'lr': 3e-4})
  assert cfg.proc_obs == 'dynimg' and cfg.lr == 3e-4 and not hasattr(cfg, 'unknown_key')
  with pytest.raises(AttributeError):
    cfg.lr = 1.0
  save_model_config(cfg._asdict(), str(tmp_path), 'e2evmc_config')
  with open(os.path.join(str(tmp_path), 'e2evmc_config.json')) as fp:
    txt = fp.read()
  assert txt.startswith('{\n  "batch_size": 32')          # indent=2, sort_keys=True
  assert create_e2evmc_config(load_model_config(str(tmp_path), 'e2evmc_config')) == cfg


def _literal_windows(L, K):
  """Literal transcription of _preprocess_targets_v3 + _window_v3 on frame indices."""
  seq = list(range(L))[:-1]                 # drop last frame
  S = L - 1
  return [seq[i:i + K] for i in range(S - K + 1)]


def test_window_index_contract_property():
  from hypothesis import given, settings, strategies as st
  from geeco_b200 import data as D
  from oracle import geeco_oracle as O

  @settings(max_examples=200, deadline=None)
  @given(st.integers(min_value=3, max_value=300), st.integers(min_value=1, max_value=16), st.integers(0, 10 ** 6))
  def prop(L, K, g):
    if K > L - 1:
      return
    lit = _literal_windows(L, K)
    idx = D.window_frame_indices(L, K)
    assert idx.dtype == np.int64 and idx.tolist() == lit
    assert np.array_equal(idx, O.window_indices(L, K))
    nw = D.num_windows(L, K)
    assert nw == len(lit) == L - K
    e, w, cur, tgt = D.locate(g, L, K)
    assert e * nw + w == g and 0 <= w < nw
    assert cur == lit[w][-1] and tgt == L - 1
    assert O.stream_index(g, L, K)[:4] == (e, w, cur, tgt)
  prop()
  assert D.num_windows(100, 4) == 96           # 3 batches of 32 per episode


def test_rank_slices_partition_the_global_batch():
  from geeco_b200.data import rank_slice
  for world in (1, 2, 4, 8):
    seen = []
    for r in range(world):
      lo, hi = rank_slice(3, 512, r, world)
      seen.extend(range(lo, hi))
    assert seen == list(range(3 * 512, 4 * 512))
  with pytest.raises(ValueError):
    rank_slice(0, 510, 0, 4)


def test_synthetic_batch_layout():
  from geeco_b200.data import synthetic_batch
  f, l = synthetic_batch(3, seed=0, first_stream_pos=94)
  assert f['rgb'].shape == (3, 4, 256, 256, 3) and f['rgb'].dtype == np.float32
  assert f['target_rgb'].shape == (3, 256, 256, 3)
  assert f['jnt_state'].shape == (3, 4, 7) and f['ee_state'].shape == (3, 4, 7) and f['obj_state'].shape == (3, 4, 7)
  assert f['step'].dtype == np.int64 and f['step'].tolist() == [[94, 95, 96, 97], [95, 96, 97, 98], [0, 1, 2, 3]]
  assert l['cmd'].shape == (3, 4) and set(np.unique(l['cmd'][:, 3])).issubset({-1.0, 0.0, 1.0})
  q = f['rgb'] * 255.0
  assert np.abs(q - np.rint(q)).max() < 1e-4 and f['rgb'].min() >= 0 and f['rgb'].max() <= 1
  assert not np.array_equal(f['rgb'][:, 0], f['rgb'][:, 1])      # never a static buffer
  f2, _ = synthetic_batch(3, seed=0, first_stream_pos=94)
  assert np.array_equal(f['rgb'], f2['rgb'])


def _header_functions():
  with open(os.path.join(ROOT, 'include', 'geeco_b200.h')) as fp:
    src = fp.read()
  src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
  return sorted(set(re.findall(r'\b(geeco_[a-z0-9_]+)\s*\(', src)))


def test_c_abi_library_loads_and_exports_every_declared_symbol():
  import ctypes
  from geeco_b200 import _lib
  from geeco_b200.build import build_library
  build_library()
  names = _header_functions()
  assert len(names) >= 20
  assert sorted(_lib.SYMBOLS.keys()) == names, set(names) ^ set(_lib.SYMBOLS.keys())
  lib = _lib.load()
  raw = ctypes.CDLL(_lib.LIB_PATH)
  for n in names:
    assert getattr(raw, n) is not None
  assert lib.geeco_version() >= 100
  # host-only entry points work without a GPU
  a = _lib.alpha_table(4)
  assert np.allclose(a, [-2.416668, 0.58333254, 1.0833325, 0.7499994], atol=2e-7)
  from oracle import geeco_oracle as O
  for K in range(2, 17):
    assert np.array_equal(_lib.alpha_table(K), O.alpha_table_f32(K)), K
  c = _lib.GeecoConfig()
  for k, v in dict(img_height=256, img_width=256, img_channels=3, dim_jnt_state=7, window_size=4, dim_s_obs=256,
                   dim_s_dyn=256, dim_s_diff=256, dim_h_lstm=128, dim_h_fc=128, num_grp_states=3, batch_size=64,
                   precision=1, training=1).items():
    setattr(c, k, v)
  s = _lib.GeecoSizes()
  _lib.check(lib.geeco_query_sizes(ctypes.byref(c), ctypes.byref(s)))
  assert s.num_params == 60 and s.num_buckets == 4 and 7552796 <= s.arena_floats < 7552796 + 4 * 60
  assert s.workspace_bytes > 2 ** 30
  c.img_channels = 5
  with pytest.raises(ValueError):
    _lib.check(lib.geeco_query_sizes(ctypes.byref(c), ctypes.byref(s)))
  # K-step LSTM op: scratch planning is host arithmetic; unsupported shapes answer -1 with a message
  need = lib.geeco_lstm_seq_scratch_floats(64, 4, 2076, 128)
  assert need >= 4 * 64 * (2076 + 128) + 4 * 64 * 512 and need % 64 == 0
  assert lib.geeco_lstm_seq_scratch_floats(64, 4, 2075, 128) == -1 and 'multiple of 4' in _lib.last_error()
  assert lib.geeco_lstm_seq_scratch_floats(64, 0, 2076, 128) == -1


def test_product_fails_loudly_without_gpu():
  import torch
  if torch.cuda.is_available():
    pytest.skip("GPU present")
  from geeco_b200 import create_e2evmc_config
  from geeco_b200.engine import Engine
  with pytest.raises(RuntimeError):
    Engine(create_e2evmc_config(dict(proc_obs='dynimg', proc_tgt='dyndiff')), batch_size=1)


def test_product_never_imports_the_oracle():
  for dirpath, _, files in os.walk(os.path.join(ROOT, 'geeco_b200')):
    for fn in files:
      if fn.endswith(('.py', '.cu', '.cuh', '.h')):
        with open(os.path.join(dirpath, fn)) as fp:
          txt = fp.read()
        assert 'import oracle' not in txt and 'from oracle' not in txt, fn


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
from geeco_b200 import parallel
from geeco_b200.data import rank_slice
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo', rank=rank, world_size=world)
class FakeEngine:
  def __init__(self):
    self.grad = torch.zeros(10); self.buckets = [(0, 4), (4, 4), (8, 2)]; self.theta = torch.zeros(10)
    self.adam_m = None; self.adam_v = None; self.log = []; self.out_losses = torch.zeros(8)
  def step_forward(self, f, l): self.log.append('fwd')
  def step_backward(self, b):
    off, cnt = self.buckets[b]; self.grad[off:off + cnt] = float(rank + 1) * (b + 1); self.log.append('bwd%%d' %% b)
  def step_update(self, scale): self.theta -= self.grad * scale; self.log.append('upd')
  def params_changed(self): pass
e = FakeEngine()
parallel.data_parallel_step(e, None, None)
exp = torch.tensor([1.5] * 4 + [3.0] * 4 + [4.5] * 2)      # mean over ranks of (rank+1)*(b+1)
assert torch.allclose(-e.theta, exp), e.theta
assert e.log == ['fwd', 'bwd0', 'bwd1', 'bwd2', 'upd']
# engines with a per-bucket optimizer step (geeco_step_update_buckets): buckets 0 .. n-2 are updated while the last
# bucket's all-reduce is in flight, then the last one; same result, every bucket exactly once
class SplitEngine(FakeEngine):
  def step_update_buckets(self, scale, first, last):
    lo, hi = self.buckets[first][0], self.buckets[last][0] + self.buckets[last][1]
    self.theta[lo:hi] -= self.grad[lo:hi] * scale; self.log.append('upd%%d-%%d' %% (first, last))
  def step_update(self, scale): raise AssertionError('the split engine must be updated per bucket')
e2 = SplitEngine()
parallel.data_parallel_step(e2, None, None)
assert torch.allclose(-e2.theta, exp), e2.theta
assert e2.log == ['fwd', 'bwd0', 'bwd1', 'bwd2', 'upd0-1', 'upd2-2']
e.theta = torch.full((10,), float(rank))
parallel.broadcast_parameters(e, src=0)
assert float(e.theta.abs().max()) == 0.0
lo, hi = rank_slice(0, 8, rank, world)
assert (lo, hi) == (rank * 4, rank * 4 + 4)
# evaluation metrics: accumulators add across ranks, every rank gets the whole-set result
assert parallel.allreduce_sums([rank + 1.0, 2, 0.5]) == [3.0, 4.0, 1.0]
from geeco_b200.estimator import Estimator, ModeKeys, EstimatorSpec
class EvalEngine:
  global_step = 7
def model_fn(features, labels, mode, params):
  v = torch.tensor([features['mse'], 0., features['mse'] * 2, features['mse'] * 3, 0., features['loss'], features['hits'], 4.])
  return EstimatorSpec(mode, v, None, None, None, None)
est = Estimator.__new__(Estimator)
from geeco_b200 import create_e2evmc_config
est._model_fn, est.params, est._batch, est._engine = model_fn, {'summaries': False}, 4, EvalEngine()
est._cfg = create_e2evmc_config({})
est._check_batch = lambda f: None
shard = [{'mse': 1.0 + rank, 'loss': 2.0 + rank, 'hits': 1.0 + 2 * rank}, {'mse': 3.0, 'loss': 4.0 - rank, 'hits': 2.0}]
res = est.evaluate(lambda: iter([(f, None) for f in shard]))
# whole set: 4 rank-batches of 4 rows; mse accumulators are sums of (mse * n * 3) over 16 * 3 values
assert abs(res['cmd_ee'] - (1.0 + 3.0 + 2.0 + 3.0) / 4) < 1e-12, res
assert abs(res['loss'] - (2.0 + 4.0 + 3.0 + 3.0) / 4) < 1e-12 and abs(res['cmd_grp'] - (1 + 2 + 3 + 2) / 16) < 1e-12, res
assert res['global_step'] == 7
dist.destroy_process_group()
sys.stdout.write('ok %%d\n' %% rank); sys.stdout.flush()      # one write: the two ranks' lines never interleave
'''


def _free_port():
  import socket
  with socket.socket() as sk:
    sk.bind(('127.0.0.1', 0))
    return sk.getsockname()[1]


@pytest.mark.timeout(180)
def test_data_parallel_step_world_size_2_gloo(tmp_path):
  script = os.path.join(str(tmp_path), 'worker.py')
  with open(script, 'w') as fp:
    fp.write(_GLOO_WORKER % {'root': ROOT})
  cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
         '127.0.0.1', '--master-port', str(_free_port()), script]       # a fixed port can still be in TIME_WAIT
  r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=170)
  assert r.returncode == 0, r.stdout[-3000:]
  assert 'ok 0' in r.stdout and 'ok 1' in r.stdout


# reference scripts/train_e2evmc.py:22-124 -- (flag, type, default)
_REFERENCE_FLAGS = [
    ('dataset_dir', str, '../data/gym-pick-pad2-cube2-v4'), ('split_name', str, 'default'),
    ('model_dir', str, '../tmp/models/geeco-f'), ('observation_format', str, 'rgb'), ('control_mode', str, 'cartesian'),
    ('goal_condition', str, 'none'), ('window_size', int, 4), ('dim_h_lstm', int, 128), ('dim_h_fc', int, 128),
    ('dim_s_obs', int, 256), ('dim_s_dyn', int, 256), ('dim_s_diff', int, 256), ('proc_obs', str, 'sequence'),
    ('proc_tgt', str, 'constant'), ('l2_regularizer', float, 0.0), ('lambda_aux', float, 1.0),
    ('data_encoding', str, 'v4'), ('lr', float, 1e-4), ('train_epochs', int, 10), ('ckpt_steps', int, 10000),
    ('num_last_ckpt', int, 2), ('num_best_ckpt', int, 5), ('batch_size', int, 32), ('memcap', float, 0.8),
    ('num_threads', int, 4), ('prefetch_size', int, 4), ('shuffle_buffer', int, 64), ('log_steps', int, 1000),
    ('debug', bool, False), ('initial_eval', bool, False)]


def _train_module():
  import importlib
  sys.path.insert(0, os.path.join(ROOT, 'scripts'))
  return importlib.import_module('train_e2evmc')


def test_train_cli_accepts_the_reference_flags():
  m = _train_module()
  d = vars(m.ARGPARSER.parse_args([]))
  for name, typ, default in _REFERENCE_FLAGS:
    assert name in d, name
    assert d[name] == default and isinstance(d[name], typ), (name, d[name])
  a = m.ARGPARSER.parse_args(['--observation_format', 'rgb', '--goal_condition', 'target', '--proc_obs', 'dynimg',
                              '--proc_tgt', 'dyndiff', '--batch_size', '64', '--debug'])
  assert (a.observation_format, a.goal_condition, a.proc_obs, a.proc_tgt, a.batch_size, a.debug) == \
      ('rgb', 'target', 'dynimg', 'dyndiff', 64, True)


def test_train_cli_refuses_before_touching_the_run_directory(tmp_path):
  m = _train_module()
  md = str(tmp_path / 'run')
  geecof = ['--goal_condition', 'target', '--proc_obs', 'dynimg', '--proc_tgt', 'dyndiff', '--model_dir', md]
  import torch
  if not torch.cuda.is_available():
    # every switch value of the reference is on the CUDA path: its default flags (sequence / constant, unconditional
    # model) get as far as the device check -- there is no CPU fallback -- and still leave no run directory behind
    with pytest.raises(RuntimeError, match='no CPU fallback'):
      m.main(m.ARGPARSER.parse_args(['--goal_condition', 'target', '--dataset_dir', 'synthetic:1', '--model_dir', md]))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
      m.main(m.ARGPARSER.parse_args(['--goal_condition', 'none', '--dataset_dir', 'synthetic:1', '--model_dir', md]))
  with pytest.raises(FileNotFoundError):
    m.main(m.ARGPARSER.parse_args(geecof + ['--dataset_dir', str(tmp_path / 'nonexistent')]))
  with pytest.raises(ValueError):
    m.main(m.ARGPARSER.parse_args(geecof + ['--control_mode', 'torque', '--dataset_dir', 'synthetic:1']))
  with pytest.raises(KeyError):
    m.main(m.ARGPARSER.parse_args(['--goal_condition', 'sometimes', '--model_dir', md]))
  assert not os.path.exists(md)


def test_snapshot_export_protocol(tmp_path):
  m = _train_module()
  from geeco_b200.estimator import latest_checkpoint
  md = str(tmp_path)
  with open(os.path.join(md, '20260101_000000000-runcmd.json'), 'w') as fp:
    fp.write('{}')
  with open(os.path.join(md, 'e2evmc_config.json'), 'w') as fp:
    fp.write('{}')
  losses = [0.5, 0.3, 0.9, 0.1]
  for i, loss in enumerate(losses):
    step = 10 * (i + 1)
    np.savez(os.path.join(md, 'model.ckpt-%d.npz' % step), global_step=np.array(step))
    with open(os.path.join(md, 'checkpoint'), 'w') as fp:
      fp.write('model_checkpoint_path: "model.ckpt-%d"\n' % step)
    assert latest_checkpoint(md).endswith('model.ckpt-%d' % step)
    m.export_snapshot(md, {'loss': loss}, num_best_ckpt=3)
  with open(os.path.join(md, 'snapshots', 'snapshot_index.json')) as fp:
    idx = json.load(fp)
  assert sorted(v['loss'] for v in idx.values()) == [0.1, 0.3, 0.5]          # the worst (0.9) was evicted
  assert not os.path.exists(os.path.join(md, 'snapshots', 'model.ckpt-30'))
  snap = os.path.join(md, 'snapshots', 'model.ckpt-40')
  assert sorted(os.listdir(snap)) == ['20260101_000000000-runcmd.json', 'checkpoint', 'e2evmc_config.json',
                                      'model.ckpt-40.npz']
  with open(os.path.join(snap, 'checkpoint')) as fp:
    assert fp.read() == 'model_checkpoint_path: "model.ckpt-40"\n'


def test_run_command_file(tmp_path):
  m = _train_module()
  from geeco_b200.runscript import save_run_command
  p = save_run_command(m.ARGPARSER, str(tmp_path), ['--lr', '0.5', '--bogus', '1'])
  assert re.match(r'\d{8}_\d{9}-runcmd\.json$', os.path.basename(p))
  with open(p) as fp:
    d = json.load(fp)
  assert d['parsed_args']['lr'] == 0.5 and d['unparsed_args'] == ['--bogus', '1']


def test_compat_import_paths():
  # the import lines of the reference's scripts, verbatim (gym_pickplace.py:41-44, train_e2evmc.py:13-16)
  code = ("import sys; sys.path.insert(0, %r); "
          "from data.geeco_gym import load_target_frame; "
          "from data.geeco_gym import pickplace_input_fn; "
          "from models.e2evmc.predictor import E2EVMCPredictor, GoalE2EVMCPredictor; "
          "from models.e2evmc.estimator import e2evmc_model_fn, goal_e2evmc_model_fn; "
          "from models.e2evmc.utils import save_model_config, load_model_config; "
          "from models.e2evmc.params import create_e2evmc_config; "
          "from models.e2evmc.utils import load_model_config; "
          "from utils.runscript import save_run_command; print('ok')") % os.path.join(ROOT, 'compat')
  r = subprocess.run([sys.executable, '-c', code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
  assert r.returncode == 0 and 'ok' in r.stdout, r.stdout


def test_unconditional_twins_need_the_cuda_path(tmp_path):
  """`E2EVMCPredictor` / `e2evmc_model_fn` (--goal_condition none) have bodies now; without a CUDA device they raise
  like every other entry point (no CPU fallback), after the reference's own argument errors."""
  import torch
  from geeco_b200 import create_e2evmc_config, save_model_config
  from geeco_b200.estimator import ModeKeys, e2evmc_model_fn
  from geeco_b200.predictor import E2EVMCPredictor
  with pytest.raises(FileNotFoundError):                    # load_model_config(model_dir, 'e2evmc_config'), predictor.py:224
    E2EVMCPredictor(str(tmp_path))
  cfg = create_e2evmc_config({})
  feats = {'rgb': np.zeros((1, 4, 256, 256, 3), np.float32), 'jnt_state': np.zeros((1, 4, 7), np.float32)}
  with pytest.raises(RuntimeError, match='Unknown estimator mode'):                     # estimator.py:138-140
    e2evmc_model_fn(feats, None, 'serve', {'e2evmc_config': cfg})
  with pytest.raises(ValueError, match='number of channels'):                           # estimator.py:27-29
    e2evmc_model_fn({}, None, ModeKeys.PREDICT, {'e2evmc_config': cfg._replace(img_channels=5)})
  if not torch.cuda.is_available():
    save_model_config(cfg._asdict(), str(tmp_path), 'e2evmc_config')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
      E2EVMCPredictor(str(tmp_path))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
      e2evmc_model_fn(feats, None, ModeKeys.PREDICT, {'e2evmc_config': cfg})


def test_target_frame_loaders(tmp_path):
  """load_target_frame / load_target_frames / load_keyframes of geeco_gym.py:168-217 on a dataset directory."""
  from PIL import Image
  from geeco_b200.input_pipeline import load_keyframes, load_target_frame, load_target_frames
  rng = np.random.default_rng(0)
  d = str(tmp_path)
  for sub in ('images/targets/rgb', 'images/targets/depth', 'images/keyframes/rgb', 'images/keyframes/depth', 'data'):
    os.makedirs(os.path.join(d, sub))
  rgb = rng.integers(0, 256, size=(8, 6, 3), dtype=np.uint8)
  depth = rng.uniform(0.5, 2.0, size=(8, 6)).astype(np.float32)
  Image.fromarray(rgb).save(os.path.join(d, 'images/targets/rgb/000007.png'))
  np.save(os.path.join(d, 'images/targets/depth/000007.npy'), depth)
  f = load_target_frame(d, 'some/dir/000007.tfrecord.zlib')
  assert f.shape == (8, 6, 4) and f.dtype == np.float32
  assert np.array_equal(f[..., :3], rgb.astype(np.float32) / 255.0) and np.array_equal(f[..., 3], depth)
  assert load_target_frame(d, '000007.tfrecord.zlib', load_depth=False).shape == (8, 6, 3)
  assert len(load_target_frames(d, '000007.tfrecord.zlib')) == 1
  for i in (1, 0):
    Image.fromarray(np.full((8, 6, 3), 10 * i, np.uint8)).save(os.path.join(d, 'images/keyframes/rgb/000007_%d.png' % i))
    np.save(os.path.join(d, 'images/keyframes/depth/000007_%d.npy' % i), depth + i)
  keys = load_keyframes(d, '000007.tfrecord.zlib')
  assert len(keys) == 2 and keys[0][0, 0, 0] == 0.0 and keys[1][0, 0, 3] == depth[0, 0] + 1
  open(os.path.join(d, 'data', 'key_frames_000007.json'), 'w').write('{}')
  assert len(load_target_frames(d, '000007.tfrecord.zlib')) == 2


def test_decode_observation_concatenates_depth_behind_rgb():
  """estimator.py:160-172: rgbd observations = depth frames concatenated behind the RGB channels."""
  import numpy as np
  import pytest
  from geeco_b200 import create_e2evmc_config
  from geeco_b200.estimator import decode_observation
  cfg3 = create_e2evmc_config({})
  cfg4 = create_e2evmc_config({'img_channels': 4})
  rng = np.random.default_rng(0)
  f = {'rgb': rng.uniform(size=(2, 4, 8, 8, 3)).astype(np.float32), 'target_rgb': rng.uniform(size=(2, 8, 8, 3)).astype(np.float32),
       'depth': rng.uniform(size=(2, 4, 8, 8, 1)).astype(np.float32), 'target_depth': rng.uniform(size=(2, 8, 8, 1)).astype(np.float32)}
  assert decode_observation(f, cfg3) is f                       # RGB: untouched
  # recorded uint8 pixels are divided by 255 before they are concatenated with the float depth
  u8 = dict(f, rgb=(f['rgb'] * 255).astype(np.uint8), target_rgb=(f['target_rgb'] * 255).astype(np.uint8))
  g8 = decode_observation(u8, cfg4)
  assert g8['rgb'].dtype == np.float32 and g8['rgb'].shape == (2, 4, 8, 8, 4)
  assert np.array_equal(g8['rgb'][..., :3], u8['rgb'].astype(np.float32) / np.float32(255.0))
  assert np.array_equal(g8['rgb'][..., 3:], f['depth']) and float(g8['target_rgb'][..., :3].max()) <= 1.0
  import torch
  t8 = decode_observation({k: torch.from_numpy(v) for k, v in u8.items()}, cfg4)
  assert torch.equal(t8['rgb'], torch.from_numpy(g8['rgb']))
  out = decode_observation(f, cfg4)
  assert out['rgb'].shape == (2, 4, 8, 8, 4) and out['target_rgb'].shape == (2, 8, 8, 4)
  assert np.array_equal(out['rgb'][..., :3], f['rgb']) and np.array_equal(out['rgb'][..., 3:], f['depth'])
  assert np.array_equal(out['target_rgb'][..., 3:], f['target_depth'])
  assert decode_observation(out, cfg4) is out                   # already 4 channels: untouched
  with pytest.raises(ValueError):
    decode_observation({'rgb': f['rgb'], 'target_rgb': f['target_rgb']}, cfg4)


def test_bench_input_pipeline_leg_runs_on_the_host():
  """bench.py's `input_pipeline` object (host decode rate next to cpu_baseline) needs no GPU and never raises."""
  import importlib.util
  spec = importlib.util.spec_from_file_location('geeco_bench', os.path.join(ROOT, 'bench.py'))
  bench = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(bench)
  r = bench.time_input_pipeline(frames=6)
  assert 'error' not in r, r
  assert r['frames_per_s_per_thread'] > 0 and r['episode_frames'] == 6 and r['decoded']['rgb'] == [5, 256, 256, 3]
  assert r['decoded']['target_rgb'] == [256, 256, 3] and r['host_cores'] == os.cpu_count()
