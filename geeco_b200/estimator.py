"""Training / evaluation entry with the shape of the reference's tf.estimator usage.

Reference: `goal_e2evmc_model_fn(features, labels, mode, params)` (src/models/e2evmc/estimator.py:144-279)
wrapped in `tf.estimator.Estimator(model_fn, model_dir, config, params)` and driven by
`estimator.train(input_fn)` / `estimator.evaluate(input_fn)` (scripts/train_e2evmc.py:254-291).

Here an `Estimator` owns one `Engine` (the CUDA context); `input_fn()` returns an iterable of
`(features, labels)` dicts in the layout of `_prepare_v4` (src/data/geeco_gym.py:373-399).  The TF
mechanics that the caller can observe are kept:
  * `train()` restores the latest checkpoint of `model_dir`, runs ONE pass over the input, writes
    `model.ckpt-<global_step>` files plus the `checkpoint` header, keeps `keep_checkpoint_max` of them;
  * `evaluate()` returns {'loss' (mean of per-batch losses), 'cmd_ee','pos_ee','pos_obj' (streaming MSE),
    'cmd_grp' (accuracy), 'global_step'} (estimator.py:246-258);
  * checkpoints store variables under their TF names (GoalVMC/.../kernel, .../kernel/Adam, ...), either as
    one `.npz` per step (default) or as TF V2 bundles `model.ckpt-<step>.index` + `.data-00000-of-00001`
    (`params['checkpoint_format'] = 'bundle'`, geeco_b200/checkpoint.py); both are restored, so a run directory
    written by the reference's tf.estimator resumes here.
"""
from __future__ import annotations

import collections
import os
import re

import numpy as np

from . import parallel
from .params import E2EVMCConfig


class ModeKeys(object):
  TRAIN, EVAL, PREDICT = 'train', 'eval', 'infer'


RunConfig = collections.namedtuple('RunConfig', ['save_checkpoints_steps', 'keep_checkpoint_max'])
RunConfig.__new__.__defaults__ = (10000, 2)

EstimatorSpec = collections.namedtuple('EstimatorSpec', ['mode', 'loss', 'train_op', 'eval_metric_ops', 'predictions',
                                                         'endpoints'])


# ------------------------------------------------------------------------------------------------
# checkpoints (TF-named variables in .npz; the `checkpoint` header file of tf.train.Saver)
# ------------------------------------------------------------------------------------------------
def latest_checkpoint(model_dir):
  """Counterpart of tf.train.latest_checkpoint: path prefix `<model_dir>/model.ckpt-<step>` or None."""
  header = os.path.join(model_dir, 'checkpoint')
  if not os.path.exists(header):
    return None
  with open(header) as fp:
    m = re.search(r'model_checkpoint_path:\s*"([^"]+)"', fp.read())
  if not m:
    return None
  prefix = os.path.join(model_dir, os.path.basename(m.group(1)))
  return prefix if (os.path.exists(prefix + '.npz') or os.path.exists(prefix + '.index')) else None


_CKPT_FILE = re.compile(r'model\.ckpt-(\d+)(\.npz|\.index|\.meta|\.data-\d{5}-of-\d{5})$')


def save_checkpoint(engine, model_dir, keep_max=2, fmt='npz'):
  if fmt not in ('npz', 'bundle'):
    raise ValueError("checkpoint_format must be 'npz' or 'bundle', got %r" % (fmt,))
  step = int(engine.global_step)
  name = 'model.ckpt-%d' % step
  arrays = {'global_step': np.array(step, dtype=np.int64)}
  for n in engine.param_names():
    arrays[n] = engine.view(n).detach().cpu().numpy()
    if engine.training:
      arrays[n + '/Adam'] = engine.view(n, engine.adam_m).detach().cpu().numpy()
      arrays[n + '/Adam_1'] = engine.view(n, engine.adam_v).detach().cpu().numpy()
  # tf.train.AdamOptimizer initialises its power accumulators to beta and multiplies AFTER every step: a checkpoint at
  # global_step = t holds beta ** (t + 1)
  arrays['beta1_power'] = np.array(0.9 ** (step + 1), dtype=np.float32)
  arrays['beta2_power'] = np.array(0.999 ** (step + 1), dtype=np.float32)
  # the reference saves the (never written, all-zero) lstm_memory variable as well (graph.py:219,226)
  arrays['%s/LSTMDecoder/lstm_memory' % getattr(engine, 'scope', 'GoalVMC')] = np.zeros(
      (engine.N, 2 * engine.cfg.dim_h_lstm), dtype=np.float32)
  os.makedirs(model_dir, exist_ok=True)
  if fmt == 'bundle':
    from .checkpoint import write_bundle
    write_bundle(os.path.join(model_dir, name), arrays)
  else:
    np.savez(os.path.join(model_dir, name + '.npz'), **arrays)
  files = [(int(m.group(1)), f) for f in os.listdir(model_dir) for m in [_CKPT_FILE.match(f)] if m]
  existing = sorted({s for s, _ in files})
  if keep_max > 0:
    for old, f in files:
      if old in existing[:-keep_max]:
        os.remove(os.path.join(model_dir, f))
    existing = existing[-keep_max:]
  with open(os.path.join(model_dir, 'checkpoint'), 'w') as fp:
    fp.write('model_checkpoint_path: "%s"\n' % name)
    for s in existing:
      fp.write('all_model_checkpoint_paths: "model.ckpt-%d"\n' % s)
  return os.path.join(model_dir, name)


class _BundleView(object):
  """The `name in data` / `data[name]` face of np.load over a TF V2 bundle."""

  def __init__(self, prefix):
    from .checkpoint import BundleReader
    self._r = BundleReader(prefix)

  def __contains__(self, name):
    return self._r.has_tensor(name)

  def __getitem__(self, name):
    return self._r.get_tensor(name)

  def __enter__(self):
    return self

  def __exit__(self, *exc):
    self._r.close()


def verify_checkpoint(prefix, config, goal_condition='target'):
  """Checks, before anything is loaded, that the checkpoint holds the variables of `config`'s graph variant with
  the right shapes (geeco_b200.graph.variable_table); a checkpoint trained with other switches raises ValueError."""
  from .graph import check_checkpoint_variables, variable_table
  if os.path.exists(prefix + '.npz'):
    with np.load(prefix + '.npz') as data:
      have = {n: data[n].shape for n, _ in variable_table(config, goal_condition) if n in data}
  else:
    from .checkpoint import BundleReader
    with BundleReader(prefix) as r:
      have = {n: r.shape(n) for n in r.names()}
  check_checkpoint_variables(have, config, goal_condition)


def restore_checkpoint(engine, prefix):
  """Restores every variable except lstm_memory (predictor.py:87) and, when present, the Adam slots; `prefix`
  names either `<prefix>.npz` or a TF V2 bundle `<prefix>.index` / `.data-*`."""
  with (np.load(prefix + '.npz') if os.path.exists(prefix + '.npz') else _BundleView(prefix)) as data:
    engine.set_params({n: data[n] for n in engine.param_names()})
    if engine.training and all((n + '/Adam') in data for n in engine.param_names()):
      import torch
      for n in engine.param_names():
        engine.view(n, engine.adam_m).copy_(torch.from_numpy(data[n + '/Adam']))
        engine.view(n, engine.adam_v).copy_(torch.from_numpy(data[n + '/Adam_1']))
    engine.set_global_step(int(data['global_step']))
  return engine.global_step


# ------------------------------------------------------------------------------------------------
# model_fn
# ------------------------------------------------------------------------------------------------
_ENGINES = {}


def _engine_for(config: E2EVMCConfig, batch, precision, training, goal_condition='target'):
  from .engine import Engine
  key = (tuple(config), int(batch), precision, bool(training), goal_condition)
  if key not in _ENGINES:
    _ENGINES[key] = Engine(config, batch_size=batch, precision=precision, training=training,
                           goal_condition=goal_condition)
    _ENGINES[key].init_params(seed=0)
  return _ENGINES[key]


def predictions_from_endpoints(ep, control_mode='cartesian'):
  """estimator.py:32-47 / :183-197."""
  if control_mode == 'cartesian':
    return {'cmd_ee': ep['pred_cmd_ee'], 'logits_cmd_grp': ep['logits_cmd_grp'], 'pos_ee': ep['pred_aux_ee'],
            'pos_obj': ep['pred_aux_obj']}
  return {'cmd_vel': ep['pred_cmd_vel'], 'cmd_ee': ep['pred_cmd_ee'], 'cmd_grp': ep['pred_cmd_grp'],
          'pos_ee': ep['pred_aux_ee'], 'pos_obj': ep['pred_aux_obj']}


def decode_observation(features, config, goal=True):
  """estimator.py:17-29 / :160-176: RGB-D observations are the depth frames concatenated behind the RGB channels
  (`in_rgbd_frames`, `tgt_rgbd_frame`).  Features that already carry img_channels channels pass through."""
  if config.img_channels != 4 or np.shape(features['rgb'])[-1] == 4:
    return features
  if 'depth' not in features or (goal and 'target_depth' not in features):
    raise ValueError("observation_format rgbd needs features['depth']" + (" and features['target_depth']" if goal else ""))
  import torch

  def cat(a, b):
    # recorded uint8 pixels become the float32 values of geeco_gym.py:310 before they share a tensor with the depth
    if torch.is_tensor(a) or torch.is_tensor(b):
      a, b = torch.as_tensor(a), torch.as_tensor(b)
      a = a.float() / 255.0 if a.dtype == torch.uint8 else a.float()
      return torch.cat([a, b.to(a.device).float()], dim=-1)
    a = np.asarray(a)
    a = a.astype(np.float32) / np.float32(255.0) if a.dtype == np.uint8 else a.astype(np.float32)
    return np.concatenate([a, np.asarray(b, dtype=np.float32)], axis=-1)

  out = dict(features)
  out['rgb'] = cat(features['rgb'], features['depth'])
  if goal:
    out['target_rgb'] = cat(features['target_rgb'], features['target_depth'])
  return out


def _model_fn(features, labels, mode, params, goal_condition):
  config = params['e2evmc_config']
  if config.img_channels not in (3, 4):
    raise ValueError("Unsupported number of channels for input frame: %d!" % config.img_channels)
  if mode not in (ModeKeys.TRAIN, ModeKeys.EVAL, ModeKeys.PREDICT):
    raise RuntimeError("Unknown estimator mode: %s" % (mode,))
  features = decode_observation(features, config, goal=goal_condition == 'target')
  batch = int(np.shape(features['jnt_state'])[0])
  eng = params.get('engine') or _engine_for(config, batch, params.get('precision', 'bf16'), mode == ModeKeys.TRAIN,
                                            goal_condition)
  if getattr(eng, 'goal_condition', goal_condition) != goal_condition:
    raise ValueError("engine was built for goal_condition=%s, model_fn is the one of %s" % (eng.goal_condition, goal_condition))
  if mode == ModeKeys.PREDICT:
    ep = eng.forward(features, None)
    return EstimatorSpec(mode, None, None, None, {k: v.detach().cpu().numpy() for k, v in
                                                   predictions_from_endpoints(ep, config.control_mode).items()}, ep)
  if mode == ModeKeys.TRAIN:
    losses = parallel.data_parallel_step(eng, features, labels)
    return EstimatorSpec(mode, losses, 'adam', None, None, None)
  ep = eng.forward(features, labels)
  return EstimatorSpec(mode, ep['losses'], None, ep['losses'], None, ep)


def e2evmc_model_fn(features, labels, mode, params):
  """Eager counterpart of the unconditional controller's model_fn (estimator.py:14-141): `e2e_vmc` (graph.py:268-319)
  over features['rgb'] / ['jnt_state'], cartesian or velocity losses (:63-98), Adam, the eval metrics of :104-117.
  `params` as for goal_e2evmc_model_fn."""
  return _model_fn(features, labels, mode, params, 'none')


def goal_e2evmc_model_fn(features, labels, mode, params):
  """Eager counterpart of estimator.py:144-279.  `params`: {'e2evmc_config', 'log_steps', 'debug'} plus the
  optional execution keys 'precision' ('bf16' | 'fp32') and 'engine' (reuse an existing Engine)."""
  return _model_fn(features, labels, mode, params, 'target')


class Estimator(object):
  """`tf.estimator.Estimator`-shaped driver of one Engine."""

  def __init__(self, model_fn, model_dir, config=None, params=None, precision='bf16', batch_size=None):
    self._model_fn = model_fn
    self.model_dir = model_dir
    self.config = config or RunConfig()
    self.params = dict(params or {})
    self.precision = precision
    self._cfg = self.params['e2evmc_config']
    self._batch = int(batch_size or self._cfg.batch_size)
    self._engine = None
    self._ckpt_format = self.params.get('checkpoint_format', 'npz')
    # train_e2evmc.py:258-259 picks the model_fn from --goal_condition; the graph variant follows from it
    self._goal = 'none' if model_fn is e2evmc_model_fn else self.params.get('goal_condition', 'target')
    self._writers = {}
    self.last_train_losses = []
    os.makedirs(model_dir, exist_ok=True)

  @property
  def engine(self):
    if self._engine is None:
      from .engine import Engine
      self._engine = Engine(self._cfg, batch_size=self._batch, precision=self.precision, training=True,
                            goal_condition=self._goal)
      self._engine.init_params(seed=int(self.params.get('seed', 0)))
      parallel.broadcast_parameters(self._engine)
      ckpt = latest_checkpoint(self.model_dir)
      if ckpt:
        verify_checkpoint(ckpt, self._cfg, self._goal)
        restore_checkpoint(self._engine, ckpt)
    return self._engine

  def _check_batch(self, features):
    n = int(np.shape(features['jnt_state'])[0])          # (rgb may be a frame pool: its first axis counts frames)
    if n != self._batch:
      # lstm_memory is created with the static shape [batch_size, 2*dim_h_lstm] (graph.py:212,218): every
      # batch the reference sees has exactly batch_size rows
      raise ValueError("batch of %d rows but the model was built for batch_size=%d" % (n, self._batch))

  def train(self, input_fn, steps=None):
    eng = self.engine
    rank, _ = parallel.world_info()
    p = dict(self.params, engine=eng)
    done, logged, self.last_train_losses = 0, 0, []
    log_steps = int(self.params.get('log_steps', 1000) or 1000)
    pending = None            # (global_step, pinned losses, copy-done event) of the last logged step

    def resolve(entry):
      step, pin, ev = entry
      ev.synchronize()
      self.last_train_losses.append((step, eng.losses_dict(pin)))
      if rank == 0 and self.params.get('summaries', True):
        # estimator.py:262-265, :305-313: one scalar per member of GraphKeys.LOSSES every log_steps, into model_dir
        from .summaries import loss_scalars
        self._summary_writer('').scalars(step, loss_scalars(self.last_train_losses[-1][1], self._cfg.control_mode))
      if rank == 0 and self.params.get('debug'):
        print('step %d: %s' % self.last_train_losses[-1])

    for features, labels in self._prefetched(eng, input_fn()):
      self._check_batch(features)
      spec = self._model_fn(features, labels, ModeKeys.TRAIN, p)
      done += 1
      if eng.global_step % log_steps == 0 or done == 1:
        # the losses of a logged step travel to the host asynchronously and are read one step later, after the
        # next step has been enqueued, so the device never idles on the host round trip
        if pending is not None:
          resolve(pending)              # before its pinned slot can be reused
        entry = eng.read_losses_async(spec.loss, logged & 1)
        logged += 1
        pending = (eng.global_step,) + entry
      if rank == 0 and self.config.save_checkpoints_steps and eng.global_step % self.config.save_checkpoints_steps == 0:
        save_checkpoint(eng, self.model_dir, self.config.keep_checkpoint_max, self._ckpt_format)
      if steps is not None and done >= steps:
        break
    if pending is not None:
      resolve(pending)
    if rank == 0 and self.params.get('save_final_checkpoint', True):
      save_checkpoint(eng, self.model_dir, self.config.keep_checkpoint_max, self._ckpt_format)
    for w in self._writers.values():
      w.flush()
    return self

  def _summary_writer(self, sub):
    if sub not in self._writers:
      from .summaries import SummaryWriter
      self._writers[sub] = SummaryWriter(os.path.join(self.model_dir, sub) if sub else self.model_dir)
    return self._writers[sub]

  @staticmethod
  def _prefetched(eng, batches):
    """Yields device-resident batches; the upload of batch i+1 (pinned host -> device, copy stream) overlaps
    the step of batch i.  Device tensors yielded by input_fn pass through untouched."""
    it = iter(batches)
    slot = 0
    try:
      f, l = next(it)
    except StopIteration:
      return
    cur = eng.stage(f, l, slot)
    while cur is not None:
      try:
        f, l = next(it)
        nxt = eng.stage(f, l, slot ^ 1)
      except StopIteration:
        nxt = None
      feats, labels, ev = cur
      eng.wait_staged(ev, slot)
      yield feats, labels
      eng.release_staged(slot)
      cur, slot = nxt, slot ^ 1

  def evaluate(self, input_fn, steps=None):
    """estimator.py:100-117 / :246-258: streaming MSE per prediction key (sum of squared errors / element count over
    the whole evaluation set), accuracy of the gripper class (cartesian), 'loss' = mean of the per-batch losses."""
    eng = self.engine
    p = dict(self.params, engine=eng)
    cfg = self._cfg
    if cfg.control_mode == 'cartesian':
      mse_keys = [('cmd_ee', 0, 3), ('pos_ee', 2, 3), ('pos_obj', 3, 3)]            # (metric, loss slot, width)
    else:
      from .engine import LOSS_SLOT_CMD_VEL
      mse_keys = [('cmd_vel', LOSS_SLOT_CMD_VEL, cfg.dim_jnt_state), ('cmd_ee', 0, 3),
                  ('cmd_grp', 1, cfg.dim_grp_command), ('pos_ee', 2, 3), ('pos_obj', 3, 3)]
    loss_sum, nb = 0.0, 0
    se = {k: 0.0 for k, _, _ in mse_keys}
    correct, rows = 0.0, 0
    for features, labels in input_fn():
      self._check_batch(features)
      spec = self._model_fn(features, labels, ModeKeys.EVAL, p)
      v = spec.loss.detach().cpu().numpy().astype(np.float64)
      n = int(v[7])
      loss_sum += v[5]; nb += 1
      # the loss slots are per-batch MSEs (mean over n*width): recover the sums for the streaming metric
      for k, slot, width in mse_keys:
        se[k] += v[slot] * n * width
      correct += v[6]; rows += n
      if steps is not None and nb >= steps:
        break
    # data parallel: every rank evaluated its share of each global batch; the accumulators add across ranks, so all
    # ranks return the metrics of the whole evaluation set (what the single-process reference computes)
    red = parallel.allreduce_sums([loss_sum, nb, correct, rows] + [se[k] for k, _, _ in mse_keys])
    loss_sum, nb, correct, rows = red[:4]
    if nb == 0:
      raise ValueError("evaluate(): input_fn yielded no batches")
    out = {k: red[4 + i] / (rows * width) for i, (k, _, width) in enumerate(mse_keys)}
    if cfg.control_mode == 'cartesian':
      out['cmd_grp'] = correct / rows
    out.update({'loss': loss_sum / nb, 'global_step': eng.global_step})
    rank, _ = parallel.world_info()
    if rank == 0 and self.params.get('summaries', True):
      w = self._summary_writer('eval')                       # tf.estimator writes evaluation summaries to model_dir/eval
      w.scalars(eng.global_step, {k: v for k, v in out.items() if k != 'global_step'})
      w.flush()
    return out

  def predict(self, input_fn):
    eng = self.engine
    p = dict(self.params, engine=eng)
    for features in input_fn():
      if isinstance(features, tuple):
        features = features[0]
      spec = self._model_fn(features, None, ModeKeys.PREDICT, p)
      n = spec.predictions['cmd_ee'].shape[0]
      for i in range(n):
        yield {k: v[i] for k, v in spec.predictions.items()}

  def latest_checkpoint(self):
    return latest_checkpoint(self.model_dir)
