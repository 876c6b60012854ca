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


class _PredictorBase(object):
  """What the two predictor classes of the reference share (predictor.py:43-208 and :212-379): the batch-1 engine in
  place of the session, checkpoint restore without `lstm_memory`, the K-frame FIFO with its assertions and padding.
  `_feed_frame`, `predict`, `reset` and `set_goal` are transcribed from the reference method by method: their
  assertion messages, FIFO order and padding rule ARE the drop-in contract of the controller hook."""

  _goal_condition = 'target'

  def __init__(self, model_dir, checkpoint_name=None, memcap=0.8, precision='fp32'):
    from .engine import Engine
    self._model_dir = model_dir
    cfg = load_model_config(model_dir, 'e2evmc_config')
    cfg['batch_size'] = 1          # predictor.py:56, :225
    self._cfg = create_e2evmc_config(cfg)
    self._engine = Engine(self._cfg, batch_size=1, precision=precision, training=False,
                          goal_condition=self._goal_condition)
    prefix = os.path.join(model_dir, checkpoint_name) if checkpoint_name else latest_checkpoint(model_dir)
    if prefix is None:
      raise FileNotFoundError("no checkpoint found in %s" % (model_dir,))
    if prefix.endswith('.npz'):
      prefix = prefix[:-4]
    verify_checkpoint(prefix, self._cfg, self._goal_condition)
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

  def _features(self):
    return {
        'rgb': np.expand_dims(np.array([f['rgb'] for f in self._frame_buffer], dtype=np.float32), axis=0),
        'jnt_state': np.expand_dims(np.array([f['jnt_state'] for f in self._frame_buffer], dtype=np.float32), axis=0),
    }

  def _postprocess(self, ep, extra=()):
    """predictor.py:149-190 / :326-366: the fetch dict per control mode, np.squeeze, argmax(logits) - 1 as float32 (1,)."""
    if self._cfg.control_mode == 'cartesian':
      predictions = {'cmd_ee': ep['pred_cmd_ee'], 'cmd_grp': ep['logits_cmd_grp'], 'pos_ee': ep['pred_aux_ee'],
                     'pos_obj': ep['pred_aux_obj']}
    else:
      predictions = {'cmd_vel': ep['pred_cmd_vel'], 'cmd_ee': ep['pred_cmd_ee'], 'cmd_grp': ep['pred_cmd_grp'],
                     'pos_ee': ep['pred_aux_ee'], 'pos_obj': ep['pred_aux_obj']}
    for k in extra:
      predictions[k] = ep[k]
    predictions = {k: np.squeeze(v.detach().cpu().numpy()) for k, v in predictions.items()}
    if self._cfg.control_mode == 'cartesian':
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


class GoalE2EVMCPredictor(_PredictorBase):
  """High-level API to run goal-conditioned E2EVMC (batch size 1), every --proc_obs / --proc_tgt / --control_mode."""

  _goal_condition = 'target'

  def _predict_command(self, reset):
    if self._target_frame is None:
      raise RuntimeError("set_goal() must be called before predict()")
    feats = self._features()
    feats['target_rgb'] = np.expand_dims(np.asarray(self._target_frame, dtype=np.float32), axis=0)
    # `reset` only selects between two all-zero LSTM states in the reference (graph.py:218-220,226)
    ep = self._engine.forward(feats, None, want_dyn=True)
    extra = (['dynbuff'] if self._cfg.proc_obs == 'dynimg' else []) + (
        ['dyndiff'] if self._cfg.proc_tgt == 'dyndiff' else [])            # predictor.py:166-169
    return self._postprocess(ep, [k for k in extra if k in ep])

  def set_goal(self, tgt_frame):
    self._target_frame = np.copy(tgt_frame[:, :, :self._cfg.img_channels])


class E2EVMCPredictor(_PredictorBase):
  """Unconditional twin (src/models/e2evmc/predictor.py:212-379): `e2e_vmc` over the K-frame FIFO, no goal frame.
  The controller scripts import both (scripts/gym_pickplace.py:43, gym_pushing.py:40)."""

  _goal_condition = 'none'

  def __init__(self, model_dir, checkpoint_name=None, memcap=0.8, precision='fp32'):
    super().__init__(model_dir, checkpoint_name, memcap, precision)

  def _predict_command(self, reset):
    return self._postprocess(self._engine.forward(self._features(), None))


class BatchedGoalPredictor(object):
  """N environments at once (extension, not in the reference; BASELINE config 4).

  The K-frame histories live in a device RING [N,K,H,W,C]: a control step writes the new frame of every environment
  into one slot (`geeco_ring_push`; an environment whose history is empty gets it in all K slots, the padding rule
  of predictor.py:196-198) and the forward reads the ring through `ring_start` -- nothing is shifted, cloned or
  synchronised with the host.  frame_dtype='uint8' keeps the recorded bytes (the real hook renders uint8 and
  divides by 255, gym_pickplace.py:869-872; the division then runs inside the pre-process kernel, bit-identically).

  carry_state=False reproduces what the reference executes (zero LSTM state every step); True carries [c | m] per
  environment across steps and starts environments that were reset from the zero state (reset mask on the device)."""

  def __init__(self, cfg, num_envs, precision='bf16', carry_state=False, engine=None, frame_dtype='float32'):
    from .engine import Engine
    self.cfg = cfg
    self.N = int(num_envs)
    self.K = cfg.window_size
    self.engine = engine or Engine(cfg, batch_size=self.N, precision=precision, training=False,
                                   carry_state=carry_state)
    dev = self.engine.device
    self.frame_dtype = {'float32': torch.float32, 'uint8': torch.uint8}[frame_dtype]
    shape = (self.N, self.K, cfg.img_height, cfg.img_width, cfg.img_channels)
    self.frames = torch.zeros(shape, dtype=self.frame_dtype, device=dev)
    self.jnt = torch.zeros((self.N, self.K, cfg.dim_jnt_state), dtype=torch.float32, device=dev)
    self.goal = torch.zeros((self.N,) + shape[2:], dtype=self.frame_dtype, device=dev)
    self.fresh = torch.ones(self.N, dtype=torch.uint8, device=dev)       # 1: history empty (before the first frame)
    self.slot = 0                                                        # ring slot the NEXT frame is written to
    self.carry_state = carry_state

  def _as_frames(self, x):
    t = torch.as_tensor(x)
    if t.dtype != self.frame_dtype:
      if self.frame_dtype == torch.uint8:
        raise ValueError("this predictor keeps uint8 frames; got %s" % (t.dtype,))
      t = t.to(torch.float32)
    return t.to(self.frames.device, non_blocking=True).contiguous()

  def set_goal(self, goals):
    self.goal.copy_(self._as_frames(goals)[..., :self.cfg.img_channels])

  def reset(self, mask=None):
    """Clears the history (and, with carry_state, the LSTM state at the next step) of all / the masked environments."""
    if mask is None:
      self.fresh.fill_(1)
    else:
      self.fresh |= torch.as_tensor(mask).to(self.fresh.device).to(torch.uint8)

  def predict_batch(self, frames, jnt_states):
    """frames [N,H,W,C] (float in [0,1], or uint8 with frame_dtype='uint8'), jnt_states [N,J] (device or host).
    Returns a dict of DEVICE tensors: cmd_ee [N,3], cmd_grp [N] in {-1,0,1}, pos_ee, pos_obj."""
    import ctypes as C
    from . import _lib
    eng = self.engine
    f = self._as_frames(frames)
    j = torch.as_tensor(jnt_states).to(self.jnt.device, dtype=torch.float32).contiguous()
    if tuple(f.shape) != tuple(self.goal.shape) or tuple(j.shape) != (self.N, self.cfg.dim_jnt_state):
      raise ValueError("frames %s / jnt_states %s do not match %d environments" % (tuple(f.shape), tuple(j.shape), self.N))
    st = eng._stream()
    row = f[0].numel() * f.element_size()
    p = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(eng.lib.geeco_ring_push(p(self.frames), p(f), p(self.fresh), self.N, self.K, row, self.slot, st))
    _lib.check(eng.lib.geeco_ring_push(p(self.jnt), p(j), p(self.fresh), self.N, self.K, 4 * self.cfg.dim_jnt_state,
                                       self.slot, st))
    self.slot = (self.slot + 1) % self.K          # the slot after the newest frame holds the oldest one
    ep = eng.forward({'rgb': self.frames, 'jnt_state': self.jnt, 'target_rgb': self.goal}, None,
                     ring_start=self.slot, reset_mask=self.fresh if self.carry_state else None)
    self.fresh.zero_()                            # stream-ordered after the forward that read it
    return {'cmd_ee': ep['pred_cmd_ee'], 'cmd_grp': (ep['logits_cmd_grp'].argmax(dim=1) - 1).float(),
            'pos_ee': ep['pred_aux_ee'], 'pos_obj': ep['pred_aux_obj']}
