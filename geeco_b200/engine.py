"""Engine: owns the device memory of one model replica and drives the C-ABI.

PyTorch is used for plumbing only (device allocations, streams, pinned host buffers,
torch.distributed for the gradient all-reduce); every arithmetic op of the hot path runs in
libgeeco_b200.so.  One Engine == one `tf.Session` worth of state in the reference: parameters
(TF variable names / layouts), Adam slots, global step, LSTM memory.
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from .params import E2EVMCConfig

LOSS_KEYS = ('loss_cmd_ee', 'loss_cmd_grp', 'loss_pos_ee', 'loss_pos_obj', 'loss_reg', 'loss')
LOSS_SLOT_CMD_VEL = 8          # velocity control: loss_cmd_vel (include/geeco_b200.h, geeco_outputs.losses)
_FEATURE_KEYS = ('rgb', 'target_rgb', 'jnt_state', 'ee_state', 'obj_state')
_FRAME_KEYS = ('rgb', 'target_rgb')
_INDEX_KEYS = ('rgb_index', 'target_index')
_LABEL_KEYS = {'cartesian': ('cmd',), 'velocity': ('vel_target', 'ee_target', 'grp_target')}   # estimator.py:206-236


def _check_switches(cfg: E2EVMCConfig, goal_condition='target'):
  """Same ValueErrors as graph.py:250-252,357-359,382-384,408-410 and estimator.py:173-175 for unknown values; the
  KeyError of `_GOAL_CONDITION_TO_MODEL[goal_condition]` (train_e2evmc.py:258) for an unknown goal condition."""
  if goal_condition not in _lib.GOAL_CONDITION:
    raise KeyError(goal_condition)
  if cfg.control_mode not in ('cartesian', 'velocity'):
    raise ValueError("Unknown control mode '%s'" % (cfg.control_mode,))
  if goal_condition == 'target':
    if cfg.proc_tgt not in ('constant', 'residual', 'dyndiff'):
      raise ValueError("Unknown processing mode for target image: %s!" % (cfg.proc_tgt,))
    if cfg.proc_obs not in ('sequence', 'dynimg'):
      raise ValueError("Unknown processing mode for frame buffer: %s!" % (cfg.proc_obs,))
  if cfg.img_channels not in (3, 4):
    raise ValueError("Unsupported number of channels for input frame: %d!" % cfg.img_channels)


class Engine(object):
  """One replica of the controller on one GPU: `goal_e2evmc` (goal_condition='target', every --proc_obs / --proc_tgt
  value) or the unconditional `e2e_vmc` (goal_condition='none'), cartesian or velocity control."""

  def __init__(self, cfg: E2EVMCConfig, batch_size=None, precision='fp32', training=True, carry_state=False,
               device=None, goal_condition='target'):
    _check_switches(cfg, goal_condition)
    self.goal_condition = goal_condition
    self.scope = 'GoalVMC' if goal_condition == 'target' else 'VMC'
    if not torch.cuda.is_available():
      raise RuntimeError("geeco_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    self.lib = _lib.load()
    self.cfg = cfg
    self.N = int(batch_size if batch_size is not None else cfg.batch_size)
    self.precision = precision
    self.training = bool(training)
    self.device = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
    torch.cuda.set_device(self.device)
    c = _lib.GeecoConfig()
    for k in ('img_height', 'img_width', 'img_channels', 'dim_jnt_state', 'window_size', 'dim_s_obs', 'dim_s_dyn',
              'dim_s_diff', 'dim_h_lstm', 'dim_h_fc', 'num_grp_states'):
      setattr(c, k, int(getattr(cfg, k)))
    c.batch_size = self.N
    c.goal_condition = _lib.GOAL_CONDITION[goal_condition]
    c.proc_obs = _lib.PROC_OBS.get(cfg.proc_obs, 0)
    c.proc_tgt = _lib.PROC_TGT.get(cfg.proc_tgt, 0)
    c.control_mode = _lib.CONTROL_MODE[cfg.control_mode]
    c.dim_grp_command = int(cfg.dim_grp_command)
    c.precision = {'fp32': _lib.GEECO_FP32, 'bf16': _lib.GEECO_BF16}[precision]
    c.carry_state = 1 if carry_state else 0
    c.training = 1 if training else 0
    c.lr, c.lambda_aux, c.l2_regularizer = float(cfg.lr), float(cfg.lambda_aux), float(cfg.l2_regularizer)
    c.adam_beta1, c.adam_beta2, c.adam_eps = 0.9, 0.999, 1e-8
    self._c = c
    sizes = _lib.GeecoSizes()
    _lib.check(self.lib.geeco_query_sizes(C.byref(c), C.byref(sizes)))
    self.sizes = sizes
    n = int(sizes.arena_floats)
    self.theta = torch.zeros(n, dtype=torch.float32, device=self.device)
    self.grad = torch.zeros(n, dtype=torch.float32, device=self.device) if training else None
    self.adam_m = torch.zeros(n, dtype=torch.float32, device=self.device) if training else None
    self.adam_v = torch.zeros(n, dtype=torch.float32, device=self.device) if training else None
    self.workspace = torch.zeros(int(sizes.workspace_bytes) + 256, dtype=torch.uint8, device=self.device)
    ws_ptr = (self.workspace.data_ptr() + 255) // 256 * 256
    ctx = C.c_void_p()
    _lib.check(self.lib.geeco_create(C.byref(c), C.byref(ctx)))
    self._ctx = ctx
    ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    _lib.check(self.lib.geeco_bind(ctx, ptr(self.theta), ptr(self.grad), ptr(self.adam_m), ptr(self.adam_v),
                                   C.c_void_p(ws_ptr), int(sizes.workspace_bytes)))
    # named views (TF variable names, TF layouts)
    self.param_table = OrderedDict()
    for i in range(sizes.num_params):
      d = _lib.GeecoParamDesc()
      _lib.check(self.lib.geeco_param_info(ctx, i, C.byref(d)))
      shape = tuple(int(d.shape[j]) for j in range(d.ndim))
      self.param_table[d.name.decode()] = (int(d.offset), int(d.numel), shape)
    self.buckets = []
    for b in range(sizes.num_buckets):
      off, cnt = C.c_int64(), C.c_int64()
      _lib.check(self.lib.geeco_grad_bucket(ctx, b, C.byref(off), C.byref(cnt)))
      self.buckets.append((int(off.value), int(cnt.value)))
    self.global_step = 0
    self.NH = int(self.lib.geeco_head_columns(C.byref(c)))
    # persistent device inputs/outputs + pinned host staging (allocated lazily)
    self._dev_in, self._pin_in, self._pin_busy = {}, {}, {}
    self.out_heads = torch.zeros(self.N, self.NH, dtype=torch.float32, device=self.device)
    self.out_fc1 = torch.zeros(self.N, cfg.dim_h_fc, dtype=torch.float32, device=self.device)
    self.out_state = torch.zeros(self.N, 2 * cfg.dim_h_lstm, dtype=torch.float32, device=self.device)
    self.out_losses = torch.zeros(_lib.NUM_LOSS_SLOTS, dtype=torch.float32, device=self.device)
    self._out_dyn = None

  # ------------------------------------------------------------------ lifecycle
  def close(self):
    if getattr(self, '_ctx', None) is not None:
      self.lib.geeco_destroy(self._ctx)
      self._ctx = None

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass

  def _stream(self):
    return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

  # ------------------------------------------------------------------ parameters
  def view(self, name, arena=None):
    off, cnt, shape = self.param_table[name]
    arena = self.theta if arena is None else arena
    return arena[off:off + cnt].view(shape)

  def param_names(self):
    return list(self.param_table.keys())

  def count_parameters(self):
    """What utils.count_parameters (src/models/e2evmc/utils.py:10-14) prints."""
    return int(sum(cnt for _, cnt, _ in self.param_table.values()))

  def init_params(self, seed=0):
    """glorot_uniform kernels, zero biases: the TF defaults of tf.layers.conv2d / dense / LSTMCell."""
    gen = torch.Generator(device='cpu')
    gen.manual_seed(int(seed))
    self.theta.zero_()
    for name, (off, cnt, shape) in self.param_table.items():
      if not name.endswith('/kernel'):
        continue
      if len(shape) == 4:
        rf = shape[0] * shape[1]
        fan_in, fan_out = rf * shape[2], rf * shape[3]
      else:
        fan_in, fan_out = shape[0], shape[1]
      lim = math.sqrt(6.0 / (fan_in + fan_out))
      w = (torch.rand(cnt, generator=gen, dtype=torch.float32) * 2.0 - 1.0) * lim
      self.theta[off:off + cnt].copy_(w)
    self.params_changed()

  def set_params(self, named: dict):
    for name, arr in named.items():
      if name.endswith('lstm_memory'):
        continue
      t = torch.as_tensor(np.asarray(arr, dtype=np.float32)) if not torch.is_tensor(arr) else arr.detach().float()
      v = self.view(name)
      if tuple(t.shape) != tuple(v.shape):
        raise ValueError("shape mismatch for %s: %s vs %s" % (name, tuple(t.shape), tuple(v.shape)))
      v.copy_(t.to(self.device))
    self.params_changed()

  def get_params(self, arena=None):
    return OrderedDict((n, self.view(n, arena).detach().cpu().numpy().copy()) for n in self.param_table)

  def get_grads(self):
    return self.get_params(self.grad)

  def params_changed(self):
    _lib.check(self.lib.geeco_params_changed(self._ctx, self._stream()))

  def set_global_step(self, t):
    self.global_step = int(t)
    _lib.check(self.lib.geeco_set_step(self._ctx, int(t), self._stream()))

  def state_dict(self):
    sd = {'global_step': self.global_step, 'theta': self.theta.detach().cpu().numpy()}
    if self.training:
      sd['adam_m'] = self.adam_m.detach().cpu().numpy()
      sd['adam_v'] = self.adam_v.detach().cpu().numpy()
    return sd

  # ------------------------------------------------------------------ inputs
  def _shapes(self):
    cfg, N, K = self.cfg, self.N, self.cfg.window_size
    H, W, Cc = cfg.img_height, cfg.img_width, cfg.img_channels
    return {'rgb': (N, K, H, W, Cc), 'target_rgb': (N, H, W, Cc), 'jnt_state': (N, K, cfg.dim_jnt_state),
            'ee_state': (N, K, 7), 'obj_state': (N, K, 7), 'cmd': (N, 4), 'vel_target': (N, cfg.dim_jnt_state),
            'ee_target': (N, 7), 'grp_target': (N, cfg.dim_grp_command)}

  def _feature_keys(self, with_labels, features=None):
    goal = self.goal_condition == 'target'
    keys = ['rgb'] + (['target_rgb'] if goal else []) + ['jnt_state']
    if features is not None and 'rgb_index' in features:          # frame-pool layout (include/geeco_b200.h: frame_index)
      keys += ['rgb_index'] + (['target_index'] if goal else [])
    return keys + (['ee_state', 'obj_state'] if with_labels else [])

  def _label_keys(self):
    return _LABEL_KEYS[self.cfg.control_mode]

  def _expect(self, key, tensor, pooled):
    """(capacity shape, wire dtype) of an input; raises ValueError on a shape the graph cannot take.  In the
    frame-pool layout `rgb` / `target_rgb` are [F,H,W,C] with any F up to the dense frame count."""
    cfg, N, K = self.cfg, self.N, self.cfg.window_size
    if key in _INDEX_KEYS:
      shape = (N, K) if key == 'rgb_index' else (N,)
      if tuple(tensor.shape) != shape:
        raise ValueError("%s: expected shape %s, got %s" % (key, shape, tuple(tensor.shape)))
      return shape, torch.int32
    shape = self._shapes()[key]
    wd = self._wire_dtype(key, tensor.dtype)
    if pooled and key in _FRAME_KEYS:
      cap = (N * K if key == 'rgb' else N,) + shape[-3:]
      if tensor.dim() != 4 or tuple(tensor.shape[1:]) != cap[1:] or not 1 <= tensor.shape[0] <= cap[0]:
        raise ValueError("%s pool: expected [F<=%d,%d,%d,%d], got %s" % ((key, cap[0]) + cap[1:] + (tuple(tensor.shape),)))
      return cap, wd
    if tuple(tensor.shape) != shape:
      raise ValueError("%s: expected shape %s, got %s" % (key, shape, tuple(tensor.shape)))
    return shape, wd

  @staticmethod
  def _wire_dtype(key, dtype):
    """Frames recorded as uint8 [0..255] stay uint8 on the wire and are divided by 255.0f on the device
    (geeco_gym.py:310); everything else travels as float32."""
    return torch.uint8 if key in _FRAME_KEYS and dtype == torch.uint8 else torch.float32

  def _to_device(self, key, value, pooled=False):
    """Accepts a CUDA tensor (used in place) or a host array (staged through pinned memory)."""
    if torch.is_tensor(value) and value.is_cuda:
      _, wd = self._expect(key, value, pooled)
      return value.contiguous().to(wd)
    arr = value if torch.is_tensor(value) else torch.from_numpy(np.ascontiguousarray(value))
    cap, wd = self._expect(key, arr, pooled)
    dk = (key, wd, pooled)
    if dk not in self._dev_in:
      self._dev_in[dk] = torch.empty(cap, dtype=wd, device=self.device)
    dst = self._dev_in[dk][:arr.shape[0]] if len(cap) else self._dev_in[dk]
    if arr.dtype == wd and arr.is_pinned():
      dst.copy_(arr, non_blocking=True)                     # caller-pinned: straight H2D
      return dst
    if dk not in self._pin_in:
      self._pin_in[dk] = torch.empty(cap, dtype=wd).pin_memory()
    self._pin_wait(dk)                                     # the upload that last read this pinned buffer is done
    pin = self._pin_in[dk][:arr.shape[0]]
    pin.copy_(arr)
    dst.copy_(pin, non_blocking=True)
    self._pin_mark(dk, torch.cuda.current_stream(self.device))
    return dst

  # A pinned staging buffer is written by the HOST (`pin.copy_`) and read by an asynchronous H2D copy.  Nothing in
  # a train step synchronises the host, so the host may run several steps ahead of the device: before it rewrites
  # a pinned buffer it waits for the event recorded right after the copy that last read that buffer.
  def _pin_wait(self, key):
    ev = self._pin_busy.get(key)
    if ev is not None:
      ev.synchronize()

  def _pin_mark(self, key, stream):
    ev = self._pin_busy.get(key)
    if ev is None:
      ev = self._pin_busy[key] = torch.cuda.Event()
    ev.record(stream)

  # ---- double-buffered staging: upload batch i+1 on a copy stream while batch i computes -------------
  def stage(self, features, labels, slot):
    """Starts the host->device copy of one (features, labels) batch into staging slot 0/1 on the copy stream.
    Returns (device_features, device_labels, ready_event); pass the event to `wait_staged` before the step."""
    if not hasattr(self, '_stage_bufs'):
      self._stage_bufs = [{}, {}]
      self._stage_free = [None, None]
      self._copy_stream = torch.cuda.Stream(device=self.device)
    bufs = self._stage_bufs[slot]
    out = {}
    pooled = 'rgb_index' in features
    with torch.cuda.stream(self._copy_stream):
      if self._stage_free[slot] is not None:
        self._copy_stream.wait_event(self._stage_free[slot])      # the step that last read this slot is done
      items = [(k, features[k]) for k in self._feature_keys(True, features) if k in features]
      if labels is not None:
        items += [(k, labels[k]) for k in self._label_keys()]
      for k, v in items:
        if torch.is_tensor(v) and v.is_cuda:
          out[k] = v
          continue
        t = v if torch.is_tensor(v) else torch.from_numpy(np.ascontiguousarray(v))
        cap, wd = self._expect(k, t, pooled)
        bk = (k, wd, pooled)
        if bk not in bufs:
          bufs[bk] = torch.empty(cap, dtype=wd, device=self.device)
        dst = bufs[bk][:t.shape[0]]
        if not (t.dtype == wd and t.is_pinned()):
          pk = bk + (slot,)
          if pk not in self._pin_in:
            self._pin_in[pk] = torch.empty(cap, dtype=wd).pin_memory()
          pin = self._pin_in[pk][:t.shape[0]]
          self._pin_wait(pk)
          pin.copy_(t)
          dst.copy_(pin, non_blocking=True)
          self._pin_mark(pk, self._copy_stream)
        else:
          dst.copy_(t, non_blocking=True)
        out[k] = dst
      ev = torch.cuda.Event()
      ev.record(self._copy_stream)
    return out, ({k: out[k] for k in self._label_keys()} if labels is not None else None), ev

  def wait_staged(self, event, slot):
    torch.cuda.current_stream(self.device).wait_event(event)

  def release_staged(self, slot):
    """Call after enqueueing the step that consumed `slot`: the next upload into it waits for that step."""
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(self.device))
    self._stage_free[slot] = ev

  def h2d_bytes(self, with_labels=True, frames_u8=False, features=None):
    """Bytes one step uploads.  With `features` (a host batch) the actual tensor sizes are counted, which is what
    matters for the frame-pool layout; else the dense layout of the graph's input shapes."""
    keys = self._feature_keys(with_labels, features) + (list(self._label_keys()) if with_labels else [])
    if features is not None:
      return int(sum(int(np.prod(np.shape(features[k]))) * (1 if (k in _FRAME_KEYS and frames_u8) else 4)
                     for k in keys if k in features))
    return int(sum((1 if frames_u8 and k in _FRAME_KEYS else 4) * int(np.prod(self._shapes()[k])) for k in keys))

  def _batch(self, features, labels, ring_start=0, reset_mask=None):
    b = _lib.GeecoBatch()
    keep = []
    pooled = 'rgb_index' in features
    fkeys = self._feature_keys(False, features)
    for k in fkeys:
      t = self._to_device(k, features[k], pooled)
      keep.append(t)
      setattr(b, {'rgb_index': 'frame_index'}.get(k, k), t.data_ptr())
    if 'target_rgb' in fkeys and keep[0].dtype != keep[1].dtype:
      raise ValueError("rgb is %s but target_rgb is %s: both must be float32 in [0,1] or both uint8 [0..255]"
                       % (keep[0].dtype, keep[1].dtype))
    b.frame_format = _lib.FRAMES_U8 if keep[0].dtype == torch.uint8 else _lib.FRAMES_F32
    b.ring_start = int(ring_start)
    if reset_mask is not None:
      if not (torch.is_tensor(reset_mask) and reset_mask.is_cuda and reset_mask.dtype in (torch.uint8, torch.bool)
              and reset_mask.numel() == self.N and reset_mask.is_contiguous()):
        raise ValueError("reset_mask must be a contiguous CUDA uint8 / bool tensor of batch_size entries")
      keep.append(reset_mask)
      b.reset_mask = reset_mask.data_ptr()
    if labels is not None:
      for k in ('ee_state', 'obj_state'):
        t = self._to_device(k, features[k])
        keep.append(t)
        setattr(b, k, t.data_ptr())
      for k in self._label_keys():
        if k not in labels:
          raise KeyError("labels['%s'] is needed for control_mode=%s" % (k, self.cfg.control_mode))
        t = self._to_device(k, labels[k])
        keep.append(t)
        setattr(b, k, t.data_ptr())
    return b, keep

  def _outputs(self, want_dyn=False, want_losses=True):
    o = _lib.GeecoOutputs()
    o.heads, o.fc1, o.lstm_state = self.out_heads.data_ptr(), self.out_fc1.data_ptr(), self.out_state.data_ptr()
    if want_losses:
      o.losses = self.out_losses.data_ptr()
    if want_dyn:
      if self._out_dyn is None:
        cfg = self.cfg
        shp = (self.N, cfg.img_height, cfg.img_width, cfg.img_channels)
        self._out_dyn = (torch.zeros(shp, dtype=torch.float32, device=self.device),
                         torch.zeros(shp, dtype=torch.float32, device=self.device))
      o.dynbuff, o.dyndiff = self._out_dyn[0].data_ptr(), self._out_dyn[1].data_ptr()
    return o

  # ------------------------------------------------------------------ steps
  def forward(self, features, labels=None, want_dyn=False, ring_start=0, reset_mask=None):
    """Forward of the graph (+ losses when labels are given).  Returns a dict of DEVICE tensors (views of
    persistent output buffers, overwritten by the next call).  ring_start / reset_mask: geeco_batch in
    include/geeco_b200.h (frame history as a ring; rows whose carried LSTM state is cleared)."""
    b, keep = self._batch(features, labels, ring_start, reset_mask)
    o = self._outputs(want_dyn=want_dyn, want_losses=labels is not None)
    _lib.check(self.lib.geeco_forward(self._ctx, C.byref(b), C.byref(o), self._stream()))
    return self._endpoints(want_dyn, labels is not None)

  def head_slices(self):
    """Endpoint name -> column slice of the heads matrix, in the order the graph creates the heads (graph.py:233-259)."""
    cfg = self.cfg
    if cfg.control_mode == 'cartesian':
      heads = [('pred_cmd_ee', 3), ('logits_cmd_grp', cfg.num_grp_states)]
    else:
      heads = [('pred_cmd_vel', cfg.dim_jnt_state), ('pred_cmd_ee', 3), ('pred_cmd_grp', cfg.dim_grp_command)]
    out, col = OrderedDict(), 0
    for name, width in heads + [('pred_aux_ee', 3), ('pred_aux_obj', 3)]:
      out[name] = slice(col, col + width)
      col += width
    return out

  def _endpoints(self, want_dyn, with_losses):
    h = self.out_heads
    ep = OrderedDict((name, h[:, sl]) for name, sl in self.head_slices().items())
    ep['fc1'] = self.out_fc1
    ep['lstm_state'] = self.out_state
    if want_dyn:
      if self.goal_condition == 'target' and self.cfg.proc_obs == 'dynimg':
        ep['dynbuff'] = self._out_dyn[0]
      if self.goal_condition == 'target' and (self.cfg.proc_obs == 'dynimg' or self.cfg.proc_tgt == 'dyndiff'):
        ep['dyndiff'] = self._out_dyn[1]
    if with_losses:
      ep['losses'] = self.out_losses
    return ep

  def train_step(self, features, labels, grad_scale=1.0):
    """One model_fn(TRAIN) step (estimator.py:144-244).  Returns the device tensor of 8 floats:
    loss_cmd_ee, loss_cmd_grp, loss_pos_ee, loss_pos_obj, loss_reg, loss, #correct, N."""
    b, keep = self._batch(features, labels)
    o = self._outputs()
    _lib.check(self.lib.geeco_train_step(self._ctx, C.byref(b), C.byref(o), float(grad_scale), self._stream()))
    self.global_step += 1
    return self.out_losses

  def step_forward(self, features, labels):
    b, keep = self._batch(features, labels)
    o = self._outputs()
    _lib.check(self.lib.geeco_step_forward(self._ctx, C.byref(b), C.byref(o), self._stream()))
    return self.out_losses

  def step_backward(self, bucket):
    _lib.check(self.lib.geeco_step_backward(self._ctx, int(bucket), self._stream()))
    off, cnt = self.buckets[bucket]
    return self.grad[off:off + cnt]

  def step_update(self, grad_scale=1.0):
    _lib.check(self.lib.geeco_step_update(self._ctx, float(grad_scale), self._stream()))
    self.global_step += 1

  def step_update_buckets(self, grad_scale, first, last):
    """Adam over gradient buckets [first, last] only; the call that includes the last bucket closes the step."""
    _lib.check(self.lib.geeco_step_update_buckets(self._ctx, float(grad_scale), int(first), int(last), self._stream()))
    if last == len(self.buckets) - 1:
      self.global_step += 1

  def set_lstm_state(self, state_cm=None):
    p = C.c_void_p(state_cm.data_ptr()) if state_cm is not None else None
    _lib.check(self.lib.geeco_set_lstm_state(self._ctx, p, self._stream()))

  def read_losses_async(self, losses, slot):
    """Starts the device->host copy of one step's loss vector into pinned slot 0/1; returns (pinned, event)."""
    if not hasattr(self, '_loss_pins'):
      self._loss_pins = [torch.empty(_lib.NUM_LOSS_SLOTS, dtype=torch.float32).pin_memory() for _ in range(2)]
    pin = self._loss_pins[slot]
    pin.copy_(losses, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(self.device))
    return pin, ev

  def losses_dict(self, losses=None):
    v = (self.out_losses if losses is None else losses).detach().cpu().numpy()
    d = OrderedDict(zip(LOSS_KEYS, [float(x) for x in v[:6]]))
    if self.cfg.control_mode == 'velocity':
      d['loss_cmd_vel'] = float(v[LOSS_SLOT_CMD_VEL])
    return d

  # ------------------------------------------------------------------ debugging
  def profile_kernel(self, name):
    """Runs one kernel of the bf16 step again on the buffers of the last step ('conv12': fused conv1 -> conv2 forward; 'bwd21': fused conv2 data gradient -> conv1 weight gradient)."""
    _lib.check(self.lib.geeco_profile_kernel(self._ctx, name.encode(), self._stream()))

  def debug_buffer(self, name):
    ptr, cnt, dt = C.c_void_p(), C.c_int64(), C.c_int32()
    _lib.check(self.lib.geeco_debug_buffer(self._ctx, name.encode(), C.byref(ptr), C.byref(cnt), C.byref(dt)))
    return wrap_device_pointer(ptr.value, int(cnt.value), torch.bfloat16 if dt.value == 1 else torch.float32,
                               self.device)


class _CudaArray(object):
  def __init__(self, ptr, nbytes):
    self.__cuda_array_interface__ = {'shape': (nbytes,), 'typestr': '|u1', 'data': (ptr, False), 'version': 3}


def wrap_device_pointer(ptr, count, dtype, device):
  """torch view (no copy) of `count` elements of `dtype` at device address `ptr`."""
  esz = torch.empty(0, dtype=dtype).element_size()
  raw = torch.as_tensor(_CudaArray(ptr, count * esz), device=device)
  return raw.view(dtype)
