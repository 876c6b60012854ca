"""CPU oracle for the GEECO e2evmc hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

PARITY PINNED ON THE REFERENCE'S GRAPH CODE, NOT ON TENSORFLOW ITSELF.  The reference
(ogroth/geeco) ships no tests, golden vectors or fixtures for this path, and its arithmetic
lives in TensorFlow 1.15.0 (environment.yml:189, not vendored, not installable offline), so
TensorFlow's kernels were never executed here.  What pins this file:
  * tests/golden/geeco_graph_golden.npz -- produced by tests/golden/make_golden.py, which imports
    the reference's src/models/e2evmc/graph.py UNMODIFIED and executes it over a small
    torch-backed stand-in for the TF symbols it uses (tests/golden/tf_shim; the per-op semantics
    stated there -- SAME padding, LSTMCell gate order / forget bias, loss reductions, L2 term --
    are ours, from TF-1.15's documented behaviour).  The wiring, variable names and creation
    order, loss functions and (through autograd) every gradient are the reference's own code;
    tests/test_golden_reference_graph.py holds this oracle to them for eight configurations, plus two
    `--control_mode velocity` configurations in geeco_graph_golden_velocity.npz (make_golden.py --velocity);
  * the closed-form known answers of tests/test_oracle_pins.py;
  * the independent loop-level restatement in oracle/np_restatement.py.
Residual risk = a TF-1.15 op semantic misread identically in the shim and here.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  Nothing under geeco_b200/ does.

Everything is written with torch CPU ops so that the same code runs in float32
(the reference's dtype) and float64 (for tolerance budgeting), and so autograd
provides the backward pass that `optimizer.minimize` derives in the reference
(src/models/e2evmc/estimator.py:243-244).

Layouts follow TensorFlow: activations NHWC, conv kernels HWIO [3,3,Cin,Cout],
dense kernels [in,out], LSTM kernel [in+h, 4h] with gate order i,j,f,o.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------
# configuration (mirrors src/models/e2evmc/params.py:7-28)
# --------------------------------------------------------------------------
DEFAULTS = dict(
    img_height=256, img_width=256, img_channels=3, dim_jnt_state=7, dim_grp_command=2,
    control_mode='cartesian', num_grp_states=3, dim_action=4, proc_obs='sequence',
    proc_tgt='constant', dim_s_obs=256, dim_s_dyn=256, dim_s_diff=256, dim_h_lstm=128,
    dim_h_fc=128, window_size=4, l2_regularizer=0.0, lambda_aux=1.0, batch_size=32, lr=1e-4)

GEECO_F = dict(DEFAULTS, proc_obs='dynimg', proc_tgt='dyndiff')

ENCODER_CHANNELS = (32, 48, 64, 128, 192, 256, 256)   # conv1..conv7, graph.py:76-110
ENCODER_STRIDES = (1, 2, 2, 2, 2, 2, 2, 2)            # graph.py:78-113
FORGET_BIAS = 1.0                                      # tf.nn.rnn_cell.LSTMCell default
ADAM_BETA1, ADAM_BETA2, ADAM_EPS = 0.9, 0.999, 1e-8    # tf.train.AdamOptimizer defaults


def make_config(**overrides):
  cfg = dict(GEECO_F)
  for k, v in overrides.items():
    if k in cfg:
      cfg[k] = v
  return cfg


# --------------------------------------------------------------------------
# rank pooling  (graph.py:17-55)
# --------------------------------------------------------------------------
def harmonic_f32(t: int) -> np.float32:
  """_H(t), graph.py:17-23: reduce_sum(1.0 / range(1, t+1)) in float32."""
  acc = np.float32(0.0)
  for i in range(1, t + 1):
    acc = np.float32(acc + np.float32(1.0) / np.float32(i))
  return acc


def alpha_table_f32(K: int) -> np.ndarray:
  """_alpha(t, T=K) for t = 1..K in float32 arithmetic (graph.py:25-28, :41-42).

  a_t = 2*(T - t + 1) - (T + 1) * (H_T - H_{t-1}), every operand a float32
  tensor in the reference graph.
  """
  T = np.float32(K)
  HT = harmonic_f32(K)
  out = np.empty(K, dtype=np.float32)
  for t in range(1, K + 1):
    tf_ = np.float32(t)
    lhs = np.float32(np.float32(2.0) * np.float32(np.float32(T - tf_) + np.float32(1.0)))
    rhs = np.float32(np.float32(T + np.float32(1.0)) * np.float32(HT - harmonic_f32(t - 1)))
    out[t - 1] = np.float32(lhs - rhs)
  return out


def alpha_table_exact(K: int):
  """Same coefficients as exact rationals (used by the known-answer tests)."""
  from fractions import Fraction
  H = [Fraction(0)]
  for i in range(1, K + 1):
    H.append(H[-1] + Fraction(1, i))
  return [2 * (K - t + 1) - (K + 1) * (H[K] - H[t - 1]) for t in range(1, K + 1)]


def dynimg(frames: torch.Tensor, alpha=None) -> torch.Tensor:
  """graph.py:30-55.  frames [N,K,H,W,C] -> [N,H,W,C].

  d = sum_k alpha_k x_k ; out = (d - min) / (max - min + 1e-6), min/max per
  sample over (H,W,C).
  """
  N, K = frames.shape[0], frames.shape[1]
  if alpha is None:
    alpha = alpha_table_f32(K)
  a = torch.as_tensor(np.asarray(alpha, dtype=np.float64), dtype=frames.dtype).reshape(1, K, 1, 1, 1)
  d = (a * frames).sum(dim=1)
  flat = d.reshape(N, -1)
  mn = flat.min(dim=1).values.reshape(N, 1, 1, 1)
  mx = flat.max(dim=1).values.reshape(N, 1, 1, 1)
  rng = mx - mn + 1e-6
  return (d - mn) / rng


def dyndiff(cur: torch.Tensor, tgt: torch.Tensor) -> torch.Tensor:
  """graph.py:397-400: dynimg(concat([cur, tgt], axis=1)), i.e. K = 2."""
  return dynimg(torch.stack([cur, tgt], dim=1))


# --------------------------------------------------------------------------
# TF 'SAME' padding and the conv encoder (graph.py:61-117)
# --------------------------------------------------------------------------
def same_pad(in_size: int, k: int, s: int):
  """TensorFlow SAME rule: out = ceil(in/s); total = max((out-1)*s + k - in, 0);
  before = total // 2; after = total - before."""
  out = -(-in_size // s)
  total = max((out - 1) * s + k - in_size, 0)
  before = total // 2
  return out, before, total - before


def conv2d_same(x_nhwc: torch.Tensor, w_hwio: torch.Tensor, b: torch.Tensor, stride: int, relu=True):
  """tf.layers.conv2d(kernel_size=3, strides=stride, padding='SAME', activation=relu)."""
  H, W = x_nhwc.shape[1], x_nhwc.shape[2]
  kh, kw = w_hwio.shape[0], w_hwio.shape[1]
  _, pt, pb = same_pad(H, kh, stride)
  _, pl, pr = same_pad(W, kw, stride)
  x = x_nhwc.permute(0, 3, 1, 2)
  x = F.pad(x, (pl, pr, pt, pb))
  y = F.conv2d(x, w_hwio.permute(3, 2, 0, 1), b, stride=stride)
  if relu:
    y = torch.relu(y)
  return y.permute(0, 2, 3, 1)


class _RoundBF16(torch.autograd.Function):
  """forward: round to bfloat16 (value kept in the working dtype); backward: identity."""

  @staticmethod
  def forward(ctx, x):
    return x.to(torch.bfloat16).to(x.dtype)

  @staticmethod
  def backward(ctx, g):
    return g


class _GradRoundBF16(torch.autograd.Function):
  """forward: identity; backward: round the incoming gradient to bfloat16."""

  @staticmethod
  def forward(ctx, x):
    return x.view_as(x)

  @staticmethod
  def backward(ctx, g):
    return g.to(torch.bfloat16).to(g.dtype)


def round_bf16(x):
  return _RoundBF16.apply(x)


def conv_encoder(x_nhwc, params, scope, return_all=False, emulate_bf16=False, relu_masks=None):
  """graph.py:61-117; `scope` e.g. 'GoalVMC/ConvEncoder'.

  relu_masks (test aid): eight 0/1 tensors shaped like the layer outputs.  Layer l then computes `z * mask_l`
  instead of `relu(z)`, i.e. the ReLU decisions are GIVEN (taken from the path under test) instead of re-derived from
  this run's own roundings.  A unit whose pre-activation is within rounding error of zero has a continuous output but
  a discontinuous gradient mask, so two correct implementations in different precisions (fp32 / bf16 kernels vs this
  float64 graph) disagree on a few such units; with the masks given the gradient comparison is free of those flips.

  emulate_bf16=True restates the SAME graph with the storage roundings of the library's bf16 mode
  (DESIGN.md "bf16 policy"): encoder input, conv kernels, every post-ReLU activation except conv8's
  and every pre-activation gradient are rounded to bfloat16; accumulation, biases, master weights and
  everything after conv8 stay in the working dtype.  One exception: conv1's bias is rounded to bfloat16 as well,
  because the library folds it into the conv1 GEMM as one more column of the packed bf16 weights.  It exists because ReLU-mask flips make the bf16
  gradients differ from the fp32 graph's by O(sqrt(eps)), which says nothing about kernel correctness."""
  acts = []
  net = round_bf16(x_nhwc) if emulate_bf16 else x_nhwc
  for li in range(8):
    w = params['%s/conv%d/kernel' % (scope, li + 1)]
    b = params['%s/conv%d/bias' % (scope, li + 1)]
    if emulate_bf16:
      z = conv2d_same(net, round_bf16(w), round_bf16(b) if li == 0 else b, ENCODER_STRIDES[li], relu=False)
      z = _GradRoundBF16.apply(z)
      net = torch.relu(z) if relu_masks is None else z * relu_masks[li].to(z.dtype)
      if li < 7:
        net = round_bf16(net)
    elif relu_masks is not None:
      net = conv2d_same(net, w, b, ENCODER_STRIDES[li], relu=False) * relu_masks[li].to(net.dtype)
    else:
      net = conv2d_same(net, w, b, ENCODER_STRIDES[li], relu=True)
    acts.append(net)
  return (net, acts) if return_all else net


# --------------------------------------------------------------------------
# state concatenation (graph.py:123-192)
# --------------------------------------------------------------------------
def _tile_jnt(jnt, like):
  N = jnt.shape[0]
  return jnt.reshape(N, 1, 1, -1).expand(N, like.shape[1], like.shape[2], jnt.shape[1])


def state_concatenation(feat, jnt):                      # graph.py:123-144
  return torch.cat([feat, _tile_jnt(jnt, feat)], dim=-1).reshape(feat.shape[0], -1)


def representation_concatenation(obs, tgt, jnt):         # graph.py:146-167
  return torch.cat([obs, _tile_jnt(jnt, obs), tgt], dim=-1).reshape(obs.shape[0], -1)


def representation_concatenation_v2(obs, dyn, jnt, tgt):  # graph.py:169-192
  return torch.cat([obs, dyn, _tile_jnt(jnt, obs), tgt], dim=-1).reshape(obs.shape[0], -1)


# --------------------------------------------------------------------------
# LSTM decoder (graph.py:198-260)
# --------------------------------------------------------------------------
def lstm_cell(x, state_cm, kernel, bias):
  """tf.nn.rnn_cell.LSTMCell(num_units, state_is_tuple=False).__call__.

  state = [c | m];  gates = [x, m] @ kernel + bias ; i, j, f, o = split(gates, 4)
  c' = sigmoid(f + forget_bias) * c + sigmoid(i) * tanh(j) ;  m' = sigmoid(o) * tanh(c')
  """
  h = kernel.shape[1] // 4
  c_prev, m_prev = state_cm[:, :h], state_cm[:, h:]
  gates = torch.cat([x, m_prev], dim=1) @ kernel + bias
  i, j, f, o = gates.split(h, dim=1)
  c = torch.sigmoid(f + FORGET_BIAS) * c_prev + torch.sigmoid(i) * torch.tanh(j)
  m = torch.sigmoid(o) * torch.tanh(c)
  return m, torch.cat([c, m], dim=1)


def lstm_decoder(feat_list, params, cfg, scope='GoalVMC/LSTMDecoder', init_state=None):
  """graph.py:198-260.  In the reference the initial state is zeros at EVERY call
  (the tf.assign at graph.py:226 is never executed), so init_state=None is the
  reference-faithful mode; passing a state emulates the intended carry mode."""
  N = feat_list[0].shape[0]
  h = cfg['dim_h_lstm']
  dt = feat_list[0].dtype
  state = torch.zeros(N, 2 * h, dtype=dt) if init_state is None else init_state
  out = None
  for feat in feat_list:
    out, state = lstm_cell(feat, state, params[scope + '/lstm_cell/kernel'], params[scope + '/lstm_cell/bias'])
  ep = OrderedDict()
  ep['lstm_out'] = out
  ep['lstm_state'] = state
  fc1 = torch.relu(out @ params[scope + '/fc1/kernel'] + params[scope + '/fc1/bias'])
  ep['fc1'] = fc1
  if cfg['control_mode'] == 'cartesian':
    heads = ['pred_cmd_ee', 'logits_cmd_grp']
  elif cfg['control_mode'] == 'velocity':
    heads = ['pred_cmd_vel', 'pred_cmd_ee', 'pred_cmd_grp']
  else:
    raise ValueError("Unknown control mode '%s'" % (cfg['control_mode'],))
  for name in heads + ['pred_aux_ee', 'pred_aux_obj']:
    ep[name] = fc1 @ params['%s/%s/kernel' % (scope, name)] + params['%s/%s/bias' % (scope, name)]
  return fc1, ep


# --------------------------------------------------------------------------
# model fns (graph.py:268-416)
# --------------------------------------------------------------------------
def goal_e2evmc(rgb_frames, jnt_states, tgt_frame, params, cfg, scope='GoalVMC', init_state=None,
                alpha=None, emulate_bf16=False, relu_masks=None):
  """graph.py:321-416.  rgb_frames [N,K,H,W,C], jnt_states [N,K,7], tgt_frame [N,H,W,C].
  relu_masks (see conv_encoder): dynimg branch {'obs' | 'dyn' | 'diff': eight masks}; sequence branch
  {'frames': [eight masks per frame], 'tgt': eight masks, 'diff': [eight masks per frame]}."""
  ep = OrderedDict()
  K = cfg['window_size']
  frames = [rgb_frames[:, k] for k in range(K)]
  jnts = [jnt_states[:, k] for k in range(K)]
  proc_obs, proc_tgt = cfg['proc_obs'], cfg['proc_tgt']
  rm = relu_masks or {}
  if proc_tgt in ('constant', 'residual'):
    tgt_feat = conv_encoder(tgt_frame, params, scope + '/ConvEncoder', emulate_bf16=emulate_bf16,
                            relu_masks=rm.get('tgt') if proc_obs == 'sequence' else None)
  elif proc_tgt != 'dyndiff':
    raise ValueError("Unknown processing mode for target image: %s!" % (proc_tgt,))
  feat_list = []
  if proc_obs == 'sequence':
    for t, (frame, jnt) in enumerate(zip(frames, jnts)):
      feat = conv_encoder(frame, params, scope + '/ConvEncoder', emulate_bf16=emulate_bf16,
                          relu_masks=rm['frames'][t] if 'frames' in rm else None)
      if proc_tgt == 'constant':
        state = representation_concatenation(feat, tgt_feat, jnt)
      elif proc_tgt == 'residual':
        state = state_concatenation(tgt_feat - feat, jnt)
      else:
        dd = dyndiff(frame, tgt_frame)
        ep['dyndiff'] = dd
        dfeat = conv_encoder(dd, params, scope + '/DynDiffEncoder', emulate_bf16=emulate_bf16,
                             relu_masks=rm['diff'][t] if 'diff' in rm else None)
        state = representation_concatenation(feat, dfeat, jnt)
      feat_list.append(state)
  elif proc_obs == 'dynimg':
    cur, jnt = frames[-1], jnts[-1]
    eb = emulate_bf16
    feat, acts = conv_encoder(cur, params, scope + '/ConvEncoder', return_all=True, emulate_bf16=eb,
                              relu_masks=rm.get('obs'))
    ep['obs_acts'] = acts
    dyn_buff = dynimg(rgb_frames, alpha)
    ep['dynbuff'] = dyn_buff
    dyn_feat = conv_encoder(dyn_buff, params, scope + '/DynBuffEncoder', emulate_bf16=eb, relu_masks=rm.get('dyn'))
    dd = dyndiff(cur, tgt_frame)
    ep['dyndiff'] = dd
    tgt_feat = conv_encoder(dd, params, scope + '/DynDiffEncoder', emulate_bf16=eb, relu_masks=rm.get('diff'))
    ep['conv8_obs'], ep['conv8_dyn'], ep['conv8_diff'] = feat, dyn_feat, tgt_feat
    feat_list.append(representation_concatenation_v2(feat, dyn_feat, jnt, tgt_feat))
  else:
    raise ValueError("Unknown processing mode for frame buffer: %s!" % (proc_obs,))
  ep['flat_representation'] = feat_list[-1]
  net, ep_dec = lstm_decoder(feat_list, params, cfg, scope + '/LSTMDecoder', init_state)
  ep.update(ep_dec)
  return net, ep


def e2e_vmc(rgb_frames, jnt_states, params, cfg, scope='VMC', init_state=None, emulate_bf16=False, relu_masks=None):
  """graph.py:268-319 (unconditional baseline).  relu_masks: {'frames': [eight masks per frame]}, see conv_encoder."""
  ep = OrderedDict()
  feat_list = []
  for k in range(cfg['window_size']):
    feat = conv_encoder(rgb_frames[:, k], params, scope + '/ConvEncoder', emulate_bf16=emulate_bf16,
                        relu_masks=relu_masks['frames'][k] if relu_masks else None)
    feat_list.append(state_concatenation(feat, jnt_states[:, k]))
  ep['flat_state'] = feat_list[-1]
  net, ep_dec = lstm_decoder(feat_list, params, cfg, scope + '/LSTMDecoder', init_state)
  ep.update(ep_dec)
  return net, ep


# --------------------------------------------------------------------------
# losses (graph.py:430-500, estimator.py:206-239)
# --------------------------------------------------------------------------
def mean_squared_error(pred, target):
  """tf.losses.mean_squared_error, default SUM_BY_NONZERO_WEIGHTS -> mean over all elements."""
  return ((pred - target) ** 2).mean()


def gripper_classes(cmd_grp_col: torch.Tensor) -> torch.Tensor:
  """estimator.py:213-216: int32(rint(x)) + 1 ; rint is round-half-to-even."""
  return torch.round(cmd_grp_col).to(torch.int64) + 1


def softmax_cross_entropy(logits, classes, depth):
  """tf.one_hot + tf.losses.softmax_cross_entropy -> mean over the batch."""
  onehot = F.one_hot(classes, depth).to(logits.dtype)
  return -(onehot * torch.log_softmax(logits, dim=1)).sum(dim=1).mean()


def l2_reg_loss(params, scale):
  """tf.contrib.layers.l2_regularizer(scale): scale * sum(w**2)/2 on every variable created
  under the encoder / LSTM scopes (graph.py:13-15,72-73,213-214).  scale == 0 -> None -> 0."""
  if not scale:
    return None
  return scale * sum(0.5 * (p ** 2).sum() for name, p in params.items() if not name.endswith('lstm_memory'))


def losses_cartesian(ep, features, labels, params, cfg):
  """estimator.py:205-228,239 for control_mode == 'cartesian'."""
  cmd = labels['cmd']
  tgt_cmd_ee = cmd[:, :3]
  tgt_grp = gripper_classes(cmd[:, 3])
  tgt_pos_ee = features['ee_state'][:, -1, :3]
  tgt_pos_obj = features['obj_state'][:, -1, :3]
  out = OrderedDict()
  out['loss_cmd_ee'] = mean_squared_error(ep['pred_cmd_ee'], tgt_cmd_ee)
  out['loss_cmd_grp'] = softmax_cross_entropy(ep['logits_cmd_grp'], tgt_grp, cfg['num_grp_states'])
  out['loss_pos_ee'] = mean_squared_error(ep['pred_aux_ee'], tgt_pos_ee)
  out['loss_pos_obj'] = mean_squared_error(ep['pred_aux_obj'], tgt_pos_obj)
  reg = l2_reg_loss(params, cfg['l2_regularizer'])
  out['loss_reg'] = reg if reg is not None else torch.zeros((), dtype=cmd.dtype)
  out['loss'] = (out['loss_cmd_ee'] + out['loss_cmd_grp']) + cfg['lambda_aux'] * (
      out['loss_pos_ee'] + out['loss_pos_obj']) + out['loss_reg']
  return out


VELOCITY_LOSS_KEYS = ('cmd_vel', 'cmd_ee', 'cmd_grp', 'pos_ee', 'pos_obj')     # _PREDICTION_KEYS, graph.py:421-428


def losses_velocity(ep, features, labels, params, cfg):
  """estimator.py:229-239 for control_mode == 'velocity': targets :230-236, `mse_loss` (graph.py:430-450) = the
  plain sum of five `tf.losses.mean_squared_error` terms (no lambda_aux), plus the regularisation term."""
  pred = {'cmd_vel': ep['pred_cmd_vel'], 'cmd_ee': ep['pred_cmd_ee'], 'cmd_grp': ep['pred_cmd_grp'],
          'pos_ee': ep['pred_aux_ee'], 'pos_obj': ep['pred_aux_obj']}                      # estimator.py:191-197
  tgt = {'cmd_vel': labels['vel_target'], 'cmd_ee': labels['ee_target'][:, :3], 'cmd_grp': labels['grp_target'],
         'pos_ee': features['ee_state'][:, -1, :3], 'pos_obj': features['obj_state'][:, -1, :3]}
  out = OrderedDict()
  for k in VELOCITY_LOSS_KEYS:
    out['loss_' + k] = mean_squared_error(pred[k], tgt[k])
  reg = l2_reg_loss(params, cfg['l2_regularizer'])
  out['loss_reg'] = reg if reg is not None else torch.zeros((), dtype=pred['cmd_vel'].dtype)
  total = out['loss_' + VELOCITY_LOSS_KEYS[0]]
  for k in VELOCITY_LOSS_KEYS[1:]:
    total = total + out['loss_' + k]
  out['loss'] = total + out['loss_reg']
  return out


# --------------------------------------------------------------------------
# parameters: names, shapes, init
# --------------------------------------------------------------------------
def encoder_param_shapes(scope, cin, dim_out):
  shapes = OrderedDict()
  chans = (cin,) + ENCODER_CHANNELS + (dim_out,)
  for li in range(8):
    shapes['%s/conv%d/kernel' % (scope, li + 1)] = (3, 3, chans[li], chans[li + 1])
    shapes['%s/conv%d/bias' % (scope, li + 1)] = (chans[li + 1],)
  return shapes


def lstm_input_dim(cfg):
  cells = 4  # 2x2 spatial cells, hard-coded tiling graph.py:139,163,188
  j = cfg['dim_jnt_state']
  if cfg['proc_obs'] == 'dynimg':
    return cells * (cfg['dim_s_obs'] + cfg['dim_s_dyn'] + j + cfg['dim_s_diff'])
  if cfg['proc_tgt'] == 'residual':
    return cells * (cfg['dim_s_obs'] + j)
  if cfg['proc_tgt'] == 'constant':
    return cells * (2 * cfg['dim_s_obs'] + j)
  return cells * (cfg['dim_s_obs'] + j + cfg['dim_s_diff'])


def param_shapes(cfg, goal=True):
  """Trainable variables in creation order with TF names."""
  scope = 'GoalVMC' if goal else 'VMC'
  C = cfg['img_channels']
  shapes = OrderedDict()
  if not goal:
    shapes.update(encoder_param_shapes(scope + '/ConvEncoder', C, 256))
    xdim = 4 * (256 + cfg['dim_jnt_state'])
  else:
    shapes.update(encoder_param_shapes(scope + '/ConvEncoder', C, cfg['dim_s_obs']))
    if cfg['proc_obs'] == 'dynimg':
      shapes.update(encoder_param_shapes(scope + '/DynBuffEncoder', C, cfg['dim_s_dyn']))
    if cfg['proc_tgt'] == 'dyndiff':
      shapes.update(encoder_param_shapes(scope + '/DynDiffEncoder', C, cfg['dim_s_diff']))
    xdim = lstm_input_dim(cfg)
  h, fc = cfg['dim_h_lstm'], cfg['dim_h_fc']
  d = scope + '/LSTMDecoder'
  shapes[d + '/lstm_cell/kernel'] = (xdim + h, 4 * h)
  shapes[d + '/lstm_cell/bias'] = (4 * h,)
  shapes[d + '/fc1/kernel'] = (h, fc)
  shapes[d + '/fc1/bias'] = (fc,)
  if cfg['control_mode'] == 'cartesian':
    heads = [('pred_cmd_ee', 3), ('logits_cmd_grp', cfg['num_grp_states'])]
  else:
    heads = [('pred_cmd_vel', cfg['dim_jnt_state']), ('pred_cmd_ee', 3), ('pred_cmd_grp', cfg['dim_grp_command'])]
  for name, n in heads + [('pred_aux_ee', 3), ('pred_aux_obj', 3)]:
    shapes['%s/%s/kernel' % (d, name)] = (fc, n)
    shapes['%s/%s/bias' % (d, name)] = (n,)
  return shapes


def glorot_limit(shape):
  """glorot_uniform (TF default initializer for get_variable / tf.layers kernels)."""
  if len(shape) == 4:
    rf = shape[0] * shape[1]
    fan_in, fan_out = rf * shape[2], rf * shape[3]
  else:
    fan_in, fan_out = shape[0], shape[1]
  return math.sqrt(6.0 / (fan_in + fan_out))


def init_params(cfg, seed=0, goal=True, dtype=torch.float32, bias_scale=0.0):
  """Glorot-uniform kernels, zero biases (TF defaults).  `bias_scale` > 0 draws small
  random biases instead, used by tests so bias gradients/ReLU masks are exercised."""
  rng = np.random.default_rng(seed)
  params = OrderedDict()
  for name, shape in param_shapes(cfg, goal).items():
    if name.endswith('/kernel'):
      lim = glorot_limit(shape)
      arr = rng.uniform(-lim, lim, size=shape)
    else:
      arr = rng.uniform(-bias_scale, bias_scale, size=shape) if bias_scale else np.zeros(shape)
    params[name] = torch.tensor(arr.astype(np.float32)).to(dtype)
  return params


def count_parameters(params):
  return int(sum(int(np.prod(p.shape)) for p in params.values()))


# --------------------------------------------------------------------------
# optimizer: tf.train.AdamOptimizer (training_ops.apply_adam semantics)
# --------------------------------------------------------------------------
def adam_init(params):
  return {'t': 0,
          'm': OrderedDict((k, torch.zeros_like(v)) for k, v in params.items()),
          'v': OrderedDict((k, torch.zeros_like(v)) for k, v in params.items())}


def adam_update(params, grads, state, lr, beta1=ADAM_BETA1, beta2=ADAM_BETA2, eps=ADAM_EPS):
  """lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t); m += (g - m)(1 - b1); v += (g^2 - v)(1 - b2);
  theta -= lr_t * m / (sqrt(v) + eps)."""
  state['t'] += 1
  t = state['t']
  lr_t = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
  with torch.no_grad():
    for k in params:
      g = grads[k]
      m, v = state['m'][k], state['v'][k]
      m += (g - m) * (1.0 - beta1)
      v += (g * g - v) * (1.0 - beta2)
      params[k] -= lr_t * m / (v.sqrt() + eps)
  return lr_t


# --------------------------------------------------------------------------
# one training / eval step (estimator.py:144-279)
# --------------------------------------------------------------------------
def _to(x, dtype):
  return torch.as_tensor(np.asarray(x)).to(dtype) if not torch.is_tensor(x) else x.to(dtype)


def forward_losses(params, features, labels, cfg, init_state=None, emulate_bf16=False, relu_masks=None, goal=True):
  """goal=True: goal_e2evmc_model_fn (estimator.py:144-279); goal=False: e2evmc_model_fn (:14-141, the
  unconditional `e2e_vmc` graph, no target frame)."""
  dt = next(iter(params.values())).dtype
  rgb = _to(features['rgb'], dt)
  jnt = _to(features['jnt_state'], dt)
  f2 = {'ee_state': _to(features['ee_state'], dt), 'obj_state': _to(features['obj_state'], dt)}
  if goal:
    tgt = _to(features['target_rgb'], dt)
    net, ep = goal_e2evmc(rgb, jnt, tgt, params, cfg, init_state=init_state, emulate_bf16=emulate_bf16,
                          relu_masks=relu_masks)
  else:
    net, ep = e2e_vmc(rgb, jnt, params, cfg, init_state=init_state, emulate_bf16=emulate_bf16, relu_masks=relu_masks)
  if cfg['control_mode'] == 'velocity':
    l2 = {k: _to(labels[k], dt) for k in ('vel_target', 'ee_target', 'grp_target')}
    return losses_velocity(ep, f2, l2, params, cfg), ep
  l2 = {'cmd': _to(labels['cmd'], dt)}
  losses = losses_cartesian(ep, f2, l2, params, cfg)
  return losses, ep


def train_step(params, opt_state, features, labels, cfg, retain=(), emulate_bf16=False, relu_masks=None, goal=True):
  """model_fn in TRAIN mode: forward, losses, gradients, one Adam update (in place).
  Returns (losses, grads, endpoints)."""
  leaves = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in params.items())
  losses, ep = forward_losses(leaves, features, labels, cfg, emulate_bf16=emulate_bf16, relu_masks=relu_masks,
                              goal=goal)
  for k in retain:
    ep[k].retain_grad()
  losses['loss'].backward()
  grads = OrderedDict((k, (v.grad if v.grad is not None else torch.zeros_like(v))) for k, v in leaves.items())
  adam_update(params, grads, opt_state, cfg['lr'])
  losses = OrderedDict((k, float(v.detach())) for k, v in losses.items())
  return losses, grads, ep


def eval_metrics_batch(ep, features, labels, cfg):
  """Per-batch pieces of estimator.py:246-254 (streaming MSE = sum sq err / count; accuracy)."""
  dt = ep['pred_cmd_ee'].dtype
  cmd = _to(labels['cmd'], dt)
  tgt = {'cmd_ee': cmd[:, :3], 'pos_ee': _to(features['ee_state'], dt)[:, -1, :3],
         'pos_obj': _to(features['obj_state'], dt)[:, -1, :3]}
  pred = {'cmd_ee': ep['pred_cmd_ee'], 'pos_ee': ep['pred_aux_ee'], 'pos_obj': ep['pred_aux_obj']}
  out = {}
  for k in tgt:
    out[k] = (float(((pred[k] - tgt[k]) ** 2).sum()), pred[k].numel())
  cls = gripper_classes(cmd[:, 3])
  out['cmd_grp'] = (float((ep['logits_cmd_grp'].argmax(dim=1) == cls).sum()), cls.numel())
  return out


# --------------------------------------------------------------------------
# predictor post-processing (predictor.py:148-200)
# --------------------------------------------------------------------------
def predictor_postprocess(ep):
  out = {'cmd_ee': np.squeeze(ep['pred_cmd_ee'].detach().numpy()),
         'pos_ee': np.squeeze(ep['pred_aux_ee'].detach().numpy()),
         'pos_obj': np.squeeze(ep['pred_aux_obj'].detach().numpy())}
  logits = np.squeeze(ep['logits_cmd_grp'].detach().numpy())
  out['cmd_grp'] = (np.argmax(logits).reshape((1,)) - 1).astype(np.float32)
  if 'dynbuff' in ep:
    out['dynbuff'] = np.squeeze(ep['dynbuff'].detach().numpy())
  if 'dyndiff' in ep:
    out['dyndiff'] = np.squeeze(ep['dyndiff'].detach().numpy())
  return out


# --------------------------------------------------------------------------
# input index contract (geeco_gym.py:598-631, 373-399, 401-474)
# --------------------------------------------------------------------------
def window_indices(episode_length: int, window_size: int):
  """_preprocess_targets_v3 drops the last frame (S = L-1); _window_v3 emits
  num_windows = S-K+1 windows, window i = frames i..i+K-1.  Returns int64 [num_windows, K]."""
  S = episode_length - 1
  nw = S - window_size + 1
  return np.arange(nw, dtype=np.int64)[:, None] + np.arange(window_size, dtype=np.int64)[None, :]


def stream_index(g: int, episode_length: int, window_size: int):
  """Stream position g (no window-level shuffle, geeco_gym.py:447-448,471-473) ->
  (episode e, window w, current frame index, target frame index, label frame index)."""
  nw = episode_length - 1 - window_size + 1
  e, w = divmod(g, nw)
  cur = w + window_size - 1
  return e, w, cur, episode_length - 1, cur


def batch_of(g: int, batch_size: int):
  return g // batch_size, g % batch_size
