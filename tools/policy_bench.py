"""BASELINE config 4 in isolation: one 1024-environment chunk of batched policy steps on the device ring (what bench.py's
`policy_steps` times), plus the kernel timeline of one control step."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geeco_b200 import create_e2evmc_config  # noqa: E402
from geeco_b200.predictor import BatchedGoalPredictor  # noqa: E402


def main():
  chunk = int(os.environ.get('GEECO_ENVS', '1024'))
  dev = torch.device('cuda:0')
  cfg = create_e2evmc_config(dict(proc_obs='dynimg', proc_tgt='dyndiff', batch_size=chunk))
  bp = BatchedGoalPredictor(cfg, chunk, precision='bf16', carry_state=True, frame_dtype='uint8')
  bp.engine.init_params(seed=0)
  bp.set_goal(torch.randint(0, 256, (chunk, 256, 256, 3), dtype=torch.uint8, device=dev))
  frames = [torch.randint(0, 256, (chunk, 256, 256, 3), dtype=torch.uint8, device=dev) for _ in range(3)]
  jn = torch.rand((chunk, 7), device=dev)
  for i in range(4):
    out = bp.predict_batch(frames[i % 3], jn)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  steps = 5
  e0.record()
  for s in range(steps):
    out = bp.predict_batch(frames[s % 3], jn)
  e1.record()
  torch.cuda.synchronize()
  sec = e0.elapsed_time(e1) * 1e-3 / steps
  print('envs %d  ms/control step %.3f  env-steps/s %.0f  checksum %.6f' % (chunk, sec * 1e3, chunk / sec,
                                                                             float(out['cmd_ee'].float().abs().sum()) if isinstance(out, dict) and 'cmd_ee' in out else 0.0))
  with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    bp.predict_batch(frames[0], jn)
    torch.cuda.synchronize()
  ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
  ev.sort(key=lambda e: e.time_range.start)
  t0 = ev[0].time_range.start
  for e in ev:
    print('%9.1f %9.1f  %s' % (e.time_range.start - t0, e.time_range.end - e.time_range.start, e.name[:90]))


if __name__ == '__main__':
  main()
