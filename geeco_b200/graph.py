"""Functional graph API of the reference (src/models/e2evmc/graph.py) on CUDA tensors.

`dynimg` (:30-55), `conv_encoder` (:61-117, one scope's eight conv layers, fp32 kernels) and
`goal_e2evmc` (:321-416, the whole GEECO-F wiring through an Engine).  Endpoint keys follow the
reference: conv8, flat_representation is internal, fc1, pred_cmd_ee, logits_cmd_grp, pred_aux_ee,
pred_aux_obj, dynbuff, dyndiff.
"""
from __future__ import annotations

from . import ops
from .ops import dynimg  # noqa: F401

ENCODER_STRIDES = (1, 2, 2, 2, 2, 2, 2, 2)


def conv_encoder(rgb_frame, params, scope='GoalVMC/ConvEncoder'):
  """rgb_frame [N,256,256,C(4-padded)] float32 CUDA; params: dict name -> CUDA tensor (TF names/layouts).
  Returns (net [N,2,2,dim_out], endpoints)."""
  net = rgb_frame
  for li in range(8):
    w = params['%s/conv%d/kernel' % (scope, li + 1)]
    b = params['%s/conv%d/bias' % (scope, li + 1)]
    if li == 0 and net.shape[-1] != w.shape[2]:
      # channel-padded input: pad the kernel's input-channel axis with zeros
      import torch
      wp = torch.zeros((3, 3, net.shape[-1], w.shape[3]), dtype=w.dtype, device=w.device)
      wp[:, :, :w.shape[2]] = w
      w = wp.contiguous()
    net = ops.conv2d_same(net.contiguous(), w.contiguous(), b.contiguous(), stride=ENCODER_STRIDES[li], relu=True)
  return net, {'conv8': net}


def goal_e2evmc(rgb_frames, jnt_states, tgt_frame, reset, engine):
  """graph.py:321-416 through an Engine built for params (E2EVMCConfig), every --proc_obs / --proc_tgt value.
  `reset` is accepted for signature compatibility; in the reference it only selects between two all-zero LSTM states
  (graph.py:218-220,226).  Returns (net = fc1 [N,dim_h_fc], endpoints dict) of device tensors."""
  ep = engine.forward({'rgb': rgb_frames, 'jnt_state': jnt_states, 'target_rgb': tgt_frame}, None, want_dyn=True)
  return ep['fc1'], ep


def e2e_vmc(rgb_frames, jnt_states, reset, engine):
  """graph.py:268-319 (the unconditional baseline) through an Engine built with goal_condition='none'."""
  ep = engine.forward({'rgb': rgb_frames, 'jnt_state': jnt_states}, None)
  return ep['fc1'], ep


# ------------------------------------------------------------------------------------------------
# variable tables of every graph variant (names, shapes, creation order of the TF variables)
# ------------------------------------------------------------------------------------------------
ENCODER_CHANNELS = (32, 48, 64, 128, 192, 256, 256)     # conv1..conv7 (graph.py:76-110); conv8 = dim_out


def _encoder_variables(scope, in_channels, dim_out):
  out, cin = [], in_channels
  for li, cout in enumerate(ENCODER_CHANNELS + (dim_out,)):
    out.append(('%s/conv%d/kernel' % (scope, li + 1), (3, 3, cin, cout)))
    out.append(('%s/conv%d/bias' % (scope, li + 1), (cout,)))
    cin = cout
  return out


def lstm_input_width(config, goal_condition='target'):
  """Width of one LSTM input vector: the flattened 2x2 map of the channel concatenation that the graph's
  `*_concatenation` function builds (graph.py:123-192)."""
  J = config.dim_jnt_state
  if goal_condition == 'none':                              # state_concatenation: [feat, jnt]; conv_encoder's
    return 4 * (256 + J)                                    # default dim_out = 256 (graph.py:61, :310)
  if config.proc_obs == 'dynimg':                           # representation_concatenation_v2: [obs, dyn, jnt, tgt]
    return 4 * (config.dim_s_obs + config.dim_s_dyn + J + config.dim_s_diff)
  if config.proc_tgt == 'residual':                         # state_concatenation: [tgt - feat, jnt]
    return 4 * (config.dim_s_obs + J)
  tgt = config.dim_s_diff if config.proc_tgt == 'dyndiff' else config.dim_s_obs
  return 4 * (config.dim_s_obs + J + tgt)                   # representation_concatenation: [obs, jnt, tgt]


def variable_table(config, goal_condition='target'):
  """[(TF variable name, shape)] of the trainable variables, in creation order, for every switch combination of
  `goal_e2evmc` (graph.py:321-416) and for `e2e_vmc` (`goal_condition='none'`, :268-319): what a checkpoint of
  that model holds (plus `<name>/Adam`, `<name>/Adam_1`, `beta1_power`, `beta2_power`, `global_step` and the
  never-assigned `lstm_memory`).  Raises the reference's ValueErrors for unknown switch values."""
  if goal_condition not in ('none', 'target'):
    raise KeyError(goal_condition)                          # _GOAL_CONDITION_TO_MODEL lookup, train_e2evmc.py:258
  if config.control_mode not in ('cartesian', 'velocity'):
    raise ValueError("Unknown control mode '%s'" % (config.control_mode,))
  C = config.img_channels
  if goal_condition == 'none':
    scope = 'VMC'
    table = _encoder_variables(scope + '/ConvEncoder', C, 256)
  else:
    scope = 'GoalVMC'
    if config.proc_tgt not in ('constant', 'residual', 'dyndiff'):
      raise ValueError("Unknown processing mode for target image: %s!" % (config.proc_tgt,))
    if config.proc_obs not in ('sequence', 'dynimg'):
      raise ValueError("Unknown processing mode for frame buffer: %s!" % (config.proc_obs,))
    # the target encoder of the constant / residual modes shares ConvEncoder's variables (AUTO_REUSE, :353-355)
    table = _encoder_variables(scope + '/ConvEncoder', C, config.dim_s_obs)
    if config.proc_obs == 'dynimg':
      table += _encoder_variables(scope + '/DynBuffEncoder', C, config.dim_s_dyn)
      table += _encoder_variables(scope + '/DynDiffEncoder', C, config.dim_s_diff)
    elif config.proc_tgt == 'dyndiff':
      table += _encoder_variables(scope + '/DynDiffEncoder', C, config.dim_s_diff)
  d = scope + '/LSTMDecoder/'
  Hl, Fc = config.dim_h_lstm, config.dim_h_fc
  table += [(d + 'lstm_cell/kernel', (lstm_input_width(config, goal_condition) + Hl, 4 * Hl)),
            (d + 'lstm_cell/bias', (4 * Hl,)), (d + 'fc1/kernel', (Hl, Fc)), (d + 'fc1/bias', (Fc,))]
  if config.control_mode == 'cartesian':
    heads = [('pred_cmd_ee', 3), ('logits_cmd_grp', config.num_grp_states)]
  else:                                                     # graph.py:240-249
    heads = [('pred_cmd_vel', config.dim_jnt_state), ('pred_cmd_ee', 3), ('pred_cmd_grp', config.dim_grp_command)]
  for name, units in heads + [('pred_aux_ee', 3), ('pred_aux_obj', 3)]:
    table += [(d + name + '/kernel', (Fc, units)), (d + name + '/bias', (units,))]
  return table


def check_checkpoint_variables(names_and_shapes, config, goal_condition='target'):
  """Raises ValueError naming the first difference between a checkpoint's variables ({name: shape}) and the
  variable table of `config`: a checkpoint trained with other switches fails here, before any tensor is loaded."""
  for name, shape in variable_table(config, goal_condition):
    if name not in names_and_shapes:
      raise ValueError("checkpoint has no variable '%s': it was not trained with proc_obs=%s proc_tgt=%s "
                       "control_mode=%s goal_condition=%s" % (name, config.proc_obs, config.proc_tgt,
                                                              config.control_mode, goal_condition))
    if tuple(names_and_shapes[name]) != tuple(shape):
      raise ValueError("checkpoint variable '%s' has shape %s, the model config needs %s"
                       % (name, tuple(names_and_shapes[name]), tuple(shape)))
