"""High-level predictor API: the controller hook of gym_pickplace.py / gym_pushing.py.

Reference: `GoalE2EVMCPredictor` (src/models/e2evmc/predictor.py:43-208) as used at
scripts/gym_pickplace.py:670-683, 850-853, 898-907.  Same constructor, `reset()`, `set_goal()`,
`predict()`, `.cfg`, `.sess`; same assertions on the fed frame; same K-frame FIFO padded by repetition;
same post-processing (np.squeeze, argmax(logits) - 1 as float32 shape (1,)).

`BatchedGoalPredictor` is the extension for many environments at once (BASELINE config 4): the K-frame
history lives in a device-resident ring, one forward per control step for all environments.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .estimator import latest_checkpoint, restore_checkpoint, verify_checkpoint
from .params import create_e2evmc_config, load_model_config

TOL_FRAME_RANGE = 1e-6   # predictor.py:17


class GoalE2EVMCPredictor(object):
  """High-level API to run goal-conditioned E2EVMC (batch size 1)."""

  def __init__(self, model_dir, checkpoint_name=None, memcap=0.8, precision='fp32'):
    from .engine import Engine
    self._model_dir = model_dir
    cfg = load_model_config(model_dir, 'e2evmc_config')
    cfg['batch_size'] = 1          # predictor.py:56
    self._cfg = create_e2evmc_config(cfg)
    self._engine = Engine(self._cfg, batch_size=1, precision=precision, training=False)
    prefix = os.path.join(model_dir, checkpoint_name) if checkpoint_name else latest_checkpoint(model_dir)
    if prefix is None:
      raise FileNotFoundError("no checkpoint found in %s" % (model_dir,))
    if prefix.endswith('.npz'):
      prefix = prefix[:-4]
    verify_checkpoint(prefix, self._cfg)
    restore_checkpoint(self._engine, prefix)
    self._frame_buffer = []
    self._buffer_size = self._cfg.window_size
    self._target_frame = None

  @property
  def sess(self):
    """The reference exposes its tf.Session here; this build exposes the Engine that plays that role."""
    return self._engine

  @property
  def cfg(self):
    return self._cfg

  def _feed_frame(self, rgb_frame, jnt_state):
    expected_shape = (self._cfg.img_height, self._cfg.img_width, self._cfg.img_channels)
    received_shape = rgb_frame.shape
    assert received_shape == expected_shape, \
        "Fed frame has wrong dimensions! Expected %s, got %s!" % (expected_shape, received_shape)
    received_range = (np.amin(rgb_frame), np.amax(rgb_frame))
    expected_range = (0 - TOL_FRAME_RANGE, 1 + TOL_FRAME_RANGE)
    assert expected_range[0] <= received_range[0] <= expected_range[1] \
        and expected_range[0] <= received_range[1] <= expected_range[1], \
        "Fed frame exceeds range! Expected %s, got %s!" % (expected_range, received_range)
    self._frame_buffer.append({'rgb': rgb_frame, 'jnt_state': jnt_state})
    if len(self._frame_buffer) > self._buffer_size:
      self._frame_buffer.pop(0)

  def _predict_command(self, reset):
    if self._target_frame is None:
      raise RuntimeError("set_goal() must be called before predict()")
    feats = {
        'rgb': np.expand_dims(np.array([f['rgb'] for f in self._frame_buffer], dtype=np.float32), axis=0),
        'jnt_state': np.expand_dims(np.array([f['jnt_state'] for f in self._frame_buffer], dtype=np.float32), axis=0),
        'target_rgb': np.expand_dims(np.asarray(self._target_frame, dtype=np.float32), axis=0),
    }
    # `reset` only selects between two all-zero LSTM states in the reference (graph.py:218-220,226)
    ep = self._engine.forward(feats, None, want_dyn=True)
    predictions = {
        'cmd_ee': ep['pred_cmd_ee'], 'cmd_grp': ep['logits_cmd_grp'], 'pos_ee': ep['pred_aux_ee'],
        'pos_obj': ep['pred_aux_obj'], 'dynbuff': ep['dynbuff'], 'dyndiff': ep['dyndiff'],
    }
    predictions = {k: np.squeeze(v.detach().cpu().numpy()) for k, v in predictions.items()}
    cmd_grp = np.argmax(predictions['cmd_grp']).reshape((1,))
    cmd_grp -= 1
    predictions['cmd_grp'] = cmd_grp.astype(np.float32)
    return predictions

  def predict(self, rgb_frame, jnt_state):
    """Feeds the frame (padding the buffer on the first step of an episode) and returns the predictions."""
    reset = (len(self._frame_buffer) == 0)
    self._feed_frame(rgb_frame, jnt_state)
    while len(self._frame_buffer) < self._buffer_size:
      self._feed_frame(rgb_frame, jnt_state)
    return self._predict_command(reset)

  def reset(self):
    self._frame_buffer.clear()

  def set_goal(self, tgt_frame):
    self._target_frame = np.copy(tgt_frame[:, :, :self._cfg.img_channels])


class E2EVMCPredictor(object):
  """Unconditional twin of GoalE2EVMCPredictor (src/models/e2evmc/predictor.py:212-379): same constructor, so the
  controller scripts' `from models.e2evmc.predictor import E2EVMCPredictor, GoalE2EVMCPredictor`
  (scripts/gym_pickplace.py:43, gym_pushing.py:40) resolves.  The unconditional graph (`e2e_vmc`, graph.py:268-319)
  is not on the CUDA path yet: constructing one raises NotImplementedError, there is no fallback."""

  def __init__(self, model_dir, checkpoint_name=None, memcap=0.8):
    raise NotImplementedError("E2EVMCPredictor (--goal_condition none, the unconditional e2e_vmc graph) is not on "
                              "the CUDA path yet; GoalE2EVMCPredictor (--goal_condition target) is")


class BatchedGoalPredictor(object):
  """N environments at once; frame history [N,K,H,W,C] is a device ring (extension, not in the reference).

  carry_state=False reproduces what the reference executes (zero LSTM state every step); True carries
  [c | m] per environment across steps and clears it for environments whose reset flag is set."""

  def __init__(self, cfg, num_envs, precision='bf16', carry_state=False, engine=None):
    from .engine import Engine
    self.cfg = cfg
    self.N = int(num_envs)
    self.K = cfg.window_size
    self.engine = engine or Engine(cfg, batch_size=self.N, precision=precision, training=False,
                                   carry_state=carry_state)
    dev = self.engine.device
    shape = (self.N, self.K, cfg.img_height, cfg.img_width, cfg.img_channels)
    self.frames = torch.zeros(shape, dtype=torch.float32, device=dev)
    self.jnt = torch.zeros((self.N, self.K, cfg.dim_jnt_state), dtype=torch.float32, device=dev)
    self.goal = torch.zeros((self.N,) + shape[2:], dtype=torch.float32, device=dev)
    self.filled = torch.zeros(self.N, dtype=torch.bool, device=dev)
    self.carry_state = carry_state

  def set_goal(self, goals):
    self.goal.copy_(torch.as_tensor(goals)[..., :self.cfg.img_channels].to(self.goal.device))

  def reset(self, mask=None):
    if mask is None:
      self.filled.zero_()
    else:
      self.filled &= ~torch.as_tensor(mask, device=self.filled.device, dtype=torch.bool)

  def predict_batch(self, frames, jnt_states):
    """frames [N,H,W,C] in [0,1], jnt_states [N,J] (device or host).  Returns a dict of DEVICE tensors:
    cmd_ee [N,3], cmd_grp [N] in {-1,0,1}, pos_ee, pos_obj."""
    dev = self.frames.device
    f = torch.as_tensor(frames).to(dev, dtype=torch.float32)
    j = torch.as_tensor(jnt_states).to(dev, dtype=torch.float32)
    fresh = ~self.filled
    # shift the window by one slot (oldest first, as the FIFO of predictor.py:140-146), then pad fresh envs
    self.frames[:, :-1] = self.frames[:, 1:].clone()
    self.jnt[:, :-1] = self.jnt[:, 1:].clone()
    self.frames[:, -1] = f
    self.jnt[:, -1] = j
    if bool(fresh.any()):
      self.frames[fresh] = f[fresh].unsqueeze(1).expand(-1, self.K, -1, -1, -1)
      self.jnt[fresh] = j[fresh].unsqueeze(1).expand(-1, self.K, -1)
      if self.carry_state:
        st = self.engine.out_state.clone()
        st[fresh] = 0
        self.engine.set_lstm_state(st)
    self.filled |= True
    ep = self.engine.forward({'rgb': self.frames, 'jnt_state': self.jnt, 'target_rgb': self.goal}, None)
    return {'cmd_ee': ep['pred_cmd_ee'], 'cmd_grp': (ep['logits_cmd_grp'].argmax(dim=1) - 1).float(),
            'pos_ee': ep['pred_aux_ee'], 'pos_obj': ep['pred_aux_obj']}
