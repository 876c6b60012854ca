"""Per-tap check of the row-resident conv kernels against the gather path (GEECO_TC_NO_ROWS=1 in a second process is
not needed: the reference here is torch's conv on the same bf16 operands)."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geeco_b200 import ops  # noqa: E402


def ref_fwd(x, w, stride):
  xx = x.float().permute(0, 3, 1, 2)
  H = x.shape[1]
  if stride == 2:
    xx = F.pad(xx, (0, 1, 0, 1))
  else:
    xx = F.pad(xx, (1, 1, 1, 1))
  return F.conv2d(xx, w.to(torch.bfloat16).float().permute(3, 2, 0, 1), stride=stride).permute(0, 2, 3, 1)


def main():
  dev = torch.device('cuda:0')
  N, H, Cin, Cout, s = 1, 256, 32, 48, 2
  torch.manual_seed(0)
  x = torch.rand((N, H, H, Cin), device=dev).to(torch.bfloat16)
  for tap in list(range(9)) + [-1]:
    w = torch.zeros((3, 3, Cin, Cout), device=dev)
    if tap >= 0:
      w[tap // 3, tap % 3] = (torch.rand((Cin, Cout), device=dev) - 0.5) * 0.2
    else:
      w = (torch.rand((3, 3, Cin, Cout), device=dev) - 0.5) * 0.2
    y, y32 = ops.conv2d_same_bf16(x, w, None, stride=s, relu=False, want_f32=True)
    r = ref_fwd(x, w, s)
    err = (y32 - r).abs()
    print('fwd tap %2d: max err %.4g (ref max %.3g)  bad rows: %s  bad cols(x): %s  bad ch: %s' % (
        tap, err.max().item(), r.abs().max().item(),
        torch.nonzero(err.amax(dim=(0, 2, 3)) > 1e-2).flatten()[:6].tolist(),
        torch.nonzero(err.amax(dim=(0, 1, 3)) > 1e-2).flatten()[:6].tolist(),
        torch.nonzero(err.amax(dim=(0, 1, 2)) > 1e-2).flatten()[:6].tolist()))
  # data gradient: dx = conv_transpose(dy, w); reference via autograd
  dy = (torch.rand((N, H // 2, H // 2, Cout), device=dev) - 0.5).to(torch.bfloat16)
  for tap in list(range(9)) + [-1]:
    w = torch.zeros((3, 3, Cin, Cout), device=dev)
    if tap >= 0:
      w[tap // 3, tap % 3] = (torch.rand((Cin, Cout), device=dev) - 0.5) * 0.2
    else:
      w = (torch.rand((3, 3, Cin, Cout), device=dev) - 0.5) * 0.2
    dw, db, dx = ops.conv2d_same_bwd_bf16(x, w, dy, stride=s, relu_mask_x=None, need_dx=True)
    xr = x.float().requires_grad_(True)
    ref_fwd(xr, w, s).backward(dy.float())
    err = (dx.float() - xr.grad).abs()
    print('dgrad tap %2d: max err %.4g (ref max %.3g)  bad rows: %s  bad cols: %s' % (
        tap, err.max().item(), xr.grad.abs().max().item(),
        torch.nonzero(err.amax(dim=(0, 2, 3)) > 2e-2).flatten()[:6].tolist(),
        torch.nonzero(err.amax(dim=(0, 1, 3)) > 2e-2).flatten()[:6].tolist()))


if __name__ == '__main__':
  main()
