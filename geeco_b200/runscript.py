"""Persists the CLI invocation of a run (reference: src/utils/runscript.py:13-30): `<ts>-runcmd.json` with
{'parsed_args', 'unparsed_args'}, timestamp formatted %Y%m%d_%H%M%S + milliseconds."""
import datetime
import json
import os
import time


def save_run_command(argparser, run_dir, argv=None):
  stamp = datetime.datetime.fromtimestamp(time.time()).strftime('%Y%m%d_%H%M%S%f')[:-3]
  parsed, unparsed = argparser.parse_known_args(argv)
  path = os.path.join(run_dir, '%s-runcmd.json' % (stamp,))
  with open(path, 'w') as fp:
    json.dump({'parsed_args': vars(parsed), 'unparsed_args': unparsed}, fp, indent=2, sort_keys=True)
  return path
