"""Inputs of the golden cases (shared by the generator and the tests; touches nothing under
/root/reference).  Every input is re-derived from the seed, so only OUTPUTS are stored in the fixture."""
from collections import OrderedDict

import numpy as np

# name -> (config overrides, batch rows, seed, goal-conditioned graph?)
CASES = OrderedDict([
    ('geecof_n2', (dict(proc_obs='dynimg', proc_tgt='dyndiff'), 2, 11, True)),
    ('geecof_l2_aux', (dict(proc_obs='dynimg', proc_tgt='dyndiff', l2_regularizer=1e-3, lambda_aux=0.5), 1, 12, True)),
    ('geecof_k8', (dict(proc_obs='dynimg', proc_tgt='dyndiff', window_size=8), 1, 13, True)),
    ('seq_constant', (dict(proc_obs='sequence', proc_tgt='constant', window_size=2), 1, 14, True)),
    ('seq_residual', (dict(proc_obs='sequence', proc_tgt='residual', window_size=2), 1, 15, True)),
    ('seq_dyndiff', (dict(proc_obs='sequence', proc_tgt='dyndiff', window_size=2), 1, 16, True)),
    ('vmc_baseline', (dict(window_size=2), 1, 17, False)),
    ('geecof_rgbd', (dict(proc_obs='dynimg', proc_tgt='dyndiff', img_channels=4), 1, 18, True)),
])

# --control_mode velocity (graph.py:240-249, estimator.py:190-197, :229-237): its own fixture file
# (geeco_graph_golden_velocity.npz) so that the cases above keep the bytes they were generated with
VELOCITY_CASES = OrderedDict([
    ('geecof_velocity', (dict(proc_obs='dynimg', proc_tgt='dyndiff', control_mode='velocity', l2_regularizer=1e-3), 2, 19, True)),
    ('seq_constant_velocity', (dict(proc_obs='sequence', proc_tgt='constant', control_mode='velocity', window_size=2), 1, 20, True)),
])

GRAD_SAMPLES = 24
BIAS_SCALE = 0.05


def make_inputs(N, K, seed, H=256, W=256, C=3, dj=7):
  """Smooth-ish frames in [0,1] (so the rank-pooling min/max are not degenerate), float64."""
  rng = np.random.default_rng(seed)
  base = rng.uniform(0.2, 0.8, size=(N, 1, H // 16, W // 16, C)).repeat(16, axis=2).repeat(16, axis=3)
  rgb = np.clip(base + rng.normal(0.0, 0.08, size=(N, K, H, W, C)), 0.0, 1.0)
  tgt = np.clip(base[:, 0] + rng.normal(0.0, 0.15, size=(N, H, W, C)), 0.0, 1.0)
  feats = {
      'rgb': rgb, 'target_rgb': tgt,
      'jnt_state': rng.normal(0.0, 0.5, size=(N, K, dj)),
      'ee_state': rng.normal(0.0, 0.5, size=(N, K, 7)),
      'obj_state': rng.normal(0.0, 0.5, size=(N, K, 7)),
      'step': np.zeros((N, K), dtype=np.int64),
  }
  cmd = rng.normal(0.0, 0.05, size=(N, 4))
  cmd[:, 3] = rng.integers(-1, 2, size=N).astype(np.float64) + rng.uniform(-0.3, 0.3, size=N)
  labels = {'cmd': cmd}
  # velocity-mode labels (geeco_gym.py:392-398), drawn after everything else so the inputs above never change
  labels['vel_target'] = rng.normal(0.0, 0.3, size=(N, dj))
  labels['ee_target'] = rng.normal(0.0, 0.5, size=(N, 7))
  labels['grp_target'] = rng.uniform(0.0, 0.05, size=(N, 2))
  return feats, labels


def sample_indices(name, size, seed):
  """Flat indices at which a gradient is stored (the fixture also holds its L2 norm and sum)."""
  h = sum(ord(c) * (i + 1) for i, c in enumerate(name)) % 100003
  rng = np.random.default_rng(seed * 1000003 + h)
  return rng.integers(0, size, size=min(GRAD_SAMPLES, size))
