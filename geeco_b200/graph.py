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
  """graph.py:321-416 through an Engine built for params (E2EVMCConfig).  `reset` is accepted for signature
  compatibility; in the reference it only selects between two all-zero LSTM states (graph.py:218-220,226).
  Returns (net = fc1 [N,dim_h_fc], endpoints dict) of device tensors."""
  ep = engine.forward({'rgb': rgb_frames, 'jnt_state': jnt_states, 'target_rgb': tgt_frame}, None, want_dyn=True)
  return ep['fc1'], ep
