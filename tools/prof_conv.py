"""Runs one conv layer of the encoder (forward, data-gradient, weight-gradient) in isolation through the
C-ABI ops, for ncu captures and event timing.  python tools/prof_conv.py --layer 1 --batch 192 --reps 3"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geeco_b200 import ops  # noqa: E402

LAYERS = {1: (256, 4, 3, 32, 1), 2: (256, 32, 32, 48, 2), 3: (128, 48, 48, 64, 2), 4: (64, 64, 64, 128, 2),
          5: (32, 128, 128, 192, 2), 6: (16, 192, 192, 256, 2), 7: (8, 256, 256, 256, 2), 8: (4, 256, 256, 256, 2)}


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--layer', type=int, default=1)
  ap.add_argument('--batch', type=int, default=192)
  ap.add_argument('--reps', type=int, default=3)
  ap.add_argument('--what', type=str, default='fwd,bwd')     # fwd, bwd (wgrad + dgrad), wgrad, dgrad
  ap.add_argument('--nomask', action='store_true')
  a = ap.parse_args()
  H, Cin, Cw, Cout, s = LAYERS[a.layer]
  dev = torch.device('cuda:0')
  x = torch.rand((a.batch, H, H, Cin), device=dev).to(torch.bfloat16)
  if Cw < Cin:
    x[..., Cw:] = 0
  w = (torch.rand((3, 3, Cw, Cout), device=dev) - 0.5) * 0.2
  b = torch.zeros(Cout, device=dev)
  Ho = H // s
  dy = (torch.rand((a.batch, Ho, Ho, Cout), device=dev) - 0.5).to(torch.bfloat16)
  flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

  def timed(fn, name, flops, bytes_):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(a.reps):
      flush.fill_(1)
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record(); fn(); e1.record(); torch.cuda.synchronize()
      ts.append(e0.elapsed_time(e1))
    t = float(np.median(ts)) * 1e-3
    print("%-22s %8.1f us  %7.1f TFLOP/s  %7.1f GB/s (algorithmic)" % (name, t * 1e6, flops / t / 1e12, bytes_ / t / 1e9))

  M = a.batch * Ho * Ho
  flops = 2.0 * M * Cout * 9 * Cw
  in_b, out_b = a.batch * H * H * Cin * 2, M * Cout * 2
  if 'fwd' in a.what:
    timed(lambda: ops.conv2d_same_bf16(x, w, b, stride=s), 'conv%d fwd' % a.layer, flops, in_b + out_b)
  extra(a, ops, x, w, dy, s, timed, flops, in_b, out_b)
  if 'bwd' in a.what:
    need_dx = a.layer > 1
    timed(lambda: ops.conv2d_same_bwd_bf16(x, w, dy, stride=s, relu_mask_x=x if need_dx else None, need_dx=need_dx),
          'conv%d wgrad%s' % (a.layer, '+dgrad' if need_dx else ''), flops * (2 if need_dx else 1),
          in_b + out_b + (2 * in_b + out_b if need_dx else 0))


def extra(a, ops, x, w, dy, s, timed, flops, in_b, out_b):
  if 'wgrad' in a.what.split(','):
    timed(lambda: ops.conv2d_same_bwd_bf16(x, w, dy, stride=s, need_dx=False), 'conv%d wgrad' % a.layer, flops, in_b + out_b)
  if 'dgrad' in a.what.split(',') and a.layer > 1:
    timed(lambda: ops.conv2d_same_bwd_bf16(x, w, dy, stride=s, relu_mask_x=None if a.nomask else x, need_dx=True, need_dw=False),
          'conv%d dgrad%s' % (a.layer, ' (no mask)' if a.nomask else ''), flops, out_b + in_b * (1 if a.nomask else 2))


if __name__ == '__main__':
  main()
