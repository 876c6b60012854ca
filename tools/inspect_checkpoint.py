#!/usr/bin/env python
"""Lists the tensors of a checkpoint (TF V2 bundle `<prefix>.index/.data-*` or this build's `<prefix>.npz`) and, given a
run directory's e2evmc_config.json, says whether it fits that model -- the counterpart of TensorFlow's
`inspect_checkpoint` for the files around this path.

  python tools/inspect_checkpoint.py <prefix | model_dir> [--config <dir with e2evmc_config.json>] [--values NAME]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('prefix')
  ap.add_argument('--config', default=None)
  ap.add_argument('--values', default=None, help='print this tensor')
  args = ap.parse_args()
  from geeco_b200.checkpoint import BundleReader
  from geeco_b200.estimator import latest_checkpoint, verify_checkpoint
  prefix = args.prefix
  if os.path.isdir(prefix):
    config_dir = args.config or prefix
    prefix = latest_checkpoint(prefix)
    if prefix is None:
      sys.exit("no checkpoint header in %s" % args.prefix)
  else:
    config_dir = args.config
  if prefix.endswith('.npz') or prefix.endswith('.index'):
    prefix = prefix.rsplit('.', 1)[0]
  if os.path.exists(prefix + '.npz'):
    with np.load(prefix + '.npz') as z:
      rows = [(n, z[n].dtype.name, z[n].shape) for n in sorted(z.files)]
      value = z[args.values] if args.values else None
  else:
    with BundleReader(prefix) as r:
      rows = []
      for n in r.names():
        t = r.get_tensor(n)                                   # also checks every tensor's CRC-32C
        rows.append((n, t.dtype.name, t.shape))
      value = r.get_tensor(args.values) if args.values else None
  total = 0
  for n, dt, shape in rows:
    print('%-64s %-8s %s' % (n, dt, tuple(shape)))
    if not (n.endswith('/Adam') or n.endswith('/Adam_1') or n.endswith('lstm_memory') or '/' not in n):
      total += int(np.prod(shape, dtype=np.int64))
  print('%d tensors, %d model parameters' % (len(rows), total))
  if config_dir and os.path.exists(os.path.join(config_dir, 'e2evmc_config.json')):
    from geeco_b200 import create_e2evmc_config, load_model_config
    cfg = create_e2evmc_config(load_model_config(config_dir, 'e2evmc_config'))
    goal = 'none' if any(n.startswith('VMC/') for n, _, _ in rows) else 'target'
    try:
      verify_checkpoint(prefix, cfg, goal)
      print('fits e2evmc_config.json (proc_obs=%s proc_tgt=%s control_mode=%s, goal_condition=%s)'
            % (cfg.proc_obs, cfg.proc_tgt, cfg.control_mode, goal))
    except ValueError as e:
      print('DOES NOT fit e2evmc_config.json: %s' % e)
  if value is not None:
    print(value)


if __name__ == '__main__':
  main()
