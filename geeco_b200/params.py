"""Hyper-parameter record of the e2evmc controller.

Same field names, defaults and merge rule as the reference's `E2EVMCConfig`
(src/models/e2evmc/params.py:7-47): an immutable namedtuple; `create_e2evmc_config`
overlays only the keys it knows and silently drops the rest.  The JSON helpers mirror
src/models/e2evmc/utils.py:16-27 (file `<name>.json`, indent 2, sorted keys).
"""
from __future__ import annotations

import json
import os
from collections import OrderedDict, namedtuple

_FIELDS = OrderedDict([
    # observation geometry
    ('img_height', 256), ('img_width', 256), ('img_channels', 3),
    # proprioception / action spaces
    ('dim_jnt_state', 7), ('dim_grp_command', 2), ('control_mode', 'cartesian'),
    ('num_grp_states', 3), ('dim_action', 4),
    # buffer / goal processing switches
    ('proc_obs', 'sequence'), ('proc_tgt', 'constant'),
    # embedding widths
    ('dim_s_obs', 256), ('dim_s_dyn', 256), ('dim_s_diff', 256), ('dim_h_lstm', 128), ('dim_h_fc', 128),
    ('window_size', 4),
    # optimisation
    ('l2_regularizer', 0.0), ('lambda_aux', 1.0), ('batch_size', 32), ('lr', 1e-4),
])

E2E_VMC_DEFAULT_PARAM_DICT = dict(_FIELDS)
E2EVMCConfig = namedtuple('E2EVMCConfig', list(_FIELDS.keys()))
E2E_VMC_DEFAULT_CONFIG = E2EVMCConfig(**E2E_VMC_DEFAULT_PARAM_DICT)


def create_e2evmc_config(custom_params: dict) -> E2EVMCConfig:
  merged = dict(E2E_VMC_DEFAULT_PARAM_DICT)
  merged.update({k: v for k, v in custom_params.items() if k in merged})
  return E2EVMCConfig(**merged)


def save_model_config(config: dict, run_dir, name):
  with open(os.path.join(run_dir, '%s.json' % (name,)), 'w') as fp:
    json.dump(config, fp, indent=2, sort_keys=True)


def load_model_config(run_dir, name):
  with open(os.path.join(run_dir, '%s.json' % (name,)), 'r') as fp:
    return json.load(fp)
