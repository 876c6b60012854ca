"""Generates tests/golden/geeco_graph_golden.npz by EXECUTING the reference's own graph code.

Run in the development container only (needs /root/reference):

    python tests/golden/make_golden.py

`/root/reference/src/models/e2evmc/graph.py` is imported unmodified; its `import tensorflow` resolves
to tests/golden/tf_shim (a torch-backed eager stand-in, see its docstring for the op semantics it
states).  For every case of cases.py the script runs the reference's graph_fn + loss functions in
float64 with the oracle's seeded parameter set, assembles the total loss exactly as
estimator.py:205-239 does, and back-propagates with torch autograd.  Stored: the variable
names/shapes in the order the reference creates them, the head outputs, fc1, conv8, state vector,
sub-sampled dynamic images, every loss term, and per-variable gradient (L2 norm, sum, sampled entries).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference/src')
sys.path.insert(0, os.path.join(HERE, 'tf_shim'))

import tensorflow as tf                                    # noqa: E402  (the shim)
from models.e2evmc import graph as G                       # noqa: E402  (the reference)
from models.e2evmc import params as RP                     # noqa: E402  (the reference)

from oracle import geeco_oracle as O                       # noqa: E402
from tests.golden import cases as C                        # noqa: E402


def run_case(over, N, seed, goal):
  cfg_d = O.make_config(batch_size=N, **over)
  ref_cfg = RP.create_e2evmc_config(cfg_d)
  K = ref_cfg.window_size
  feats, labels = C.make_inputs(N, K, seed, C=cfg_d['img_channels'])
  P = O.init_params(cfg_d, seed=seed, goal=goal, dtype=torch.float64, bias_scale=C.BIAS_SCALE)
  tf.reset_state()
  for k, v in P.items():
    tf.VARIABLES[k] = v.clone().requires_grad_(True)
  f = {k: torch.tensor(v) for k, v in feats.items()}
  cmd = torch.tensor(labels['cmd'])
  # estimator.py:176-182: reset = (prod(step) == 0); graph entry
  reset = tf.equal(tf.reduce_prod(f['step']), tf.constant(0, dtype=tf.int64))
  if goal:
    net, ep = G.goal_e2evmc(f['rgb'], f['jnt_state'], f['target_rgb'], reset, params=ref_cfg)
  else:
    net, ep = G.e2e_vmc(f['rgb'], f['jnt_state'], reset, params=ref_cfg)
  if ref_cfg.control_mode == 'velocity':
    return run_velocity_tail(ep, f, labels, ref_cfg, seed)
  predictions = {'cmd_ee': ep['pred_cmd_ee'], 'logits_cmd_grp': ep['logits_cmd_grp'],
                 'pos_ee': ep['pred_aux_ee'], 'pos_obj': ep['pred_aux_obj']}
  # estimator.py:205-239 (cartesian): regularisation term, targets, class shift, loss assembly
  loss_reg = tf.reduce_sum(tf.get_collection(tf.GraphKeys.REGULARIZATION_LOSSES))
  targets = {'cmd_ee': cmd[:, :3], 'cmd_grp': cmd[:, 3], 'pos_ee': f['ee_state'][:, -1, :3],
             'pos_obj': f['obj_state'][:, -1, :3]}
  targets['cmd_grp'] = tf.dtypes.cast(tf.math.rint(targets['cmd_grp']), tf.int32) + 1
  l_ee, _ = G.cartesian_cmd_loss(predictions, targets, ref_cfg)
  l_grp, _ = G.gripper_cmd_loss(predictions, targets, ref_cfg)
  l_pe, _ = G.ee_pose_loss(predictions, targets, ref_cfg)
  l_po, _ = G.obj_pose_loss(predictions, targets, ref_cfg)
  loss = tf.add_n([l_ee, l_grp]) + ref_cfg.lambda_aux * tf.add_n([l_pe, l_po])
  loss = tf.add(loss, loss_reg)
  names = [n for n, _ in tf.CREATED]
  grads = torch.autograd.grad(loss, [tf.VARIABLES[n] for n in names], allow_unused=True)
  out = {
      'var_names': np.array(names), 'var_shapes': np.array([str(s) for _, s in tf.CREATED]),
      'classes': targets['cmd_grp'].numpy(),
      'losses': np.array([float(l_ee), float(l_grp), float(l_pe), float(l_po), float(loss_reg), float(loss)]),
  }
  for k in ('pred_cmd_ee', 'logits_cmd_grp', 'pred_aux_ee', 'pred_aux_obj', 'fc1', 'conv8'):
    out['ep_' + k] = ep[k].detach().numpy()
  for k in ('flat_state', 'flat_representation'):
    if k in ep:
      out['ep_' + k] = ep[k].detach().numpy()
  for k in ('dynbuff', 'dyndiff'):
    if k in ep:
      img = ep[k].detach().numpy()
      out['ep_' + k + '_sub'] = img[:, ::8, ::8, :].copy()
      out['ep_' + k + '_stats'] = np.array([img.mean(), img.min(), img.max(), (img ** 2).sum()])
  for n, g in zip(names, grads):
    g = torch.zeros_like(tf.VARIABLES[n]) if g is None else g
    g = g.detach().numpy().ravel()
    idx = C.sample_indices(n, g.size, seed)
    out['grad_stats/' + n] = np.array([np.linalg.norm(g), g.sum()])
    out['grad_samples/' + n] = g[idx]
  return out


def run_velocity_tail(ep, f, labels, ref_cfg, seed):
  """estimator.py:190-197 (predictions), :201-204 (regularisation), :229-239 (targets, mse_loss, total)."""
  predictions = {'cmd_vel': ep['pred_cmd_vel'], 'cmd_ee': ep['pred_cmd_ee'], 'cmd_grp': ep['pred_cmd_grp'],
                 'pos_ee': ep['pred_aux_ee'], 'pos_obj': ep['pred_aux_obj']}
  loss_reg = tf.reduce_sum(tf.get_collection(tf.GraphKeys.REGULARIZATION_LOSSES))
  targets = {'cmd_vel': torch.tensor(labels['vel_target']), 'cmd_ee': torch.tensor(labels['ee_target'])[:, :3],
             'cmd_grp': torch.tensor(labels['grp_target']), 'pos_ee': f['ee_state'][:, -1, :3],
             'pos_obj': f['obj_state'][:, -1, :3]}
  loss, ep_losses = G.mse_loss(predictions, targets)
  loss = tf.add(loss, loss_reg)
  names = [n for n, _ in tf.CREATED]
  grads = torch.autograd.grad(loss, [tf.VARIABLES[n] for n in names], allow_unused=True)
  keys = ('cmd_vel', 'cmd_ee', 'cmd_grp', 'pos_ee', 'pos_obj')
  out = {
      'var_names': np.array(names), 'var_shapes': np.array([str(s) for _, s in tf.CREATED]),
      'losses': np.array([float(ep_losses['loss_' + k]) for k in keys] + [float(loss_reg), float(loss)]),
  }
  for k in ('pred_cmd_vel', 'pred_cmd_ee', 'pred_cmd_grp', 'pred_aux_ee', 'pred_aux_obj', 'fc1'):
    out['ep_' + k] = ep[k].detach().numpy()
  for n, g in zip(names, grads):
    g = torch.zeros_like(tf.VARIABLES[n]) if g is None else g
    g = g.detach().numpy().ravel()
    idx = C.sample_indices(n, g.size, seed)
    out['grad_stats/' + n] = np.array([np.linalg.norm(g), g.sum()])
    out['grad_samples/' + n] = g[idx]
  return out


def main_velocity():
  blob = {}
  for name, (over, N, seed, goal) in C.VELOCITY_CASES.items():
    res = run_case(over, N, seed, goal)
    for k, v in res.items():
      blob[name + '::' + k] = v
    print('%-22s loss=%.9f  vars=%d' % (name, res['losses'][-1], len(res['var_names'])))
  path = os.path.join(HERE, 'geeco_graph_golden_velocity.npz')
  np.savez_compressed(path, **blob)
  print('wrote %s (%.1f KiB)' % (path, os.path.getsize(path) / 1024.0))


def main():
  if '--velocity' in sys.argv:       # the velocity cases live in their own fixture file
    return main_velocity()
  blob = {}
  for name, (over, N, seed, goal) in C.CASES.items():
    res = run_case(over, N, seed, goal)
    for k, v in res.items():
      blob[name + '::' + k] = v
    print('%-14s loss=%.9f  vars=%d' % (name, res['losses'][5], len(res['var_names'])))
  path = os.path.join(HERE, 'geeco_graph_golden.npz')
  np.savez_compressed(path, **blob)
  print('wrote %s (%.1f KiB)' % (path, os.path.getsize(path) / 1024.0))


if __name__ == '__main__':
  main()
