"""GPU parity of the tcgen05 (bf16) conv kernels against the CPU oracle evaluated on the same
bf16-rounded operands: the fp32 copy of the output must agree to fp32-accumulation accuracy, the bf16
output to bf16 rounding (2^-9 relative)."""
import numpy as np
import pytest
import torch

from oracle import geeco_oracle as O
from tests.util import rel_l2, rel_max

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

CASES = [
    # N, H, Cin(storage), Cw(real), Cout, stride
    (2, 16, 4, 3, 32, 1),        # conv1: RGB padded to 4 channels, 8-byte gather pieces, K 36 -> 64
    (1, 32, 4, 4, 32, 1),        # conv1 with rgbd input
    (2, 16, 8, 3, 32, 1),        # 8-channel storage of a 3-channel input, K 72 -> 128
    (2, 16, 32, 32, 48, 2),      # conv2-like: K 288 -> 320
    (3, 8, 48, 48, 64, 2),       # conv3-like: K 432 -> 448
    (2, 8, 64, 64, 128, 2),
    (5, 4, 128, 128, 192, 2),    # M = 20 rows: one partial tile
    (2, 4, 256, 256, 256, 2),    # N = 256: full TMEM double buffer
    (1, 64, 8, 4, 32, 1),        # 4096 rows: 32 tiles, rgbd-like
    (3, 32, 32, 32, 48, 2),      # 768 rows: multi-tile persistent loop
    (40, 2, 256, 256, 256, 2),   # conv8-like at batch 40
    (2, 256, 32, 32, 48, 2),     # conv2 itself: row-resident kernels (pixel-pair forward, 4-class data gradient)
    (1, 512, 32, 32, 48, 2),     # 256-pixel output rows: two row tiles per row
]


def _bf(x):
  return torch.from_numpy(x).to(torch.bfloat16)


@pytest.mark.parametrize("case", CASES)
def test_conv2d_bf16_fwd_bwd_matches_oracle(cuda_device, case):
  from geeco_b200 import ops
  N, H, Cin, Cw, Cout, stride = case
  rng = np.random.default_rng(sum(case))
  x = rng.uniform(0, 1, size=(N, H, H, Cin)).astype(np.float32)
  x[..., Cw:] = 0.0
  w = rng.uniform(-0.2, 0.2, size=(3, 3, Cw, Cout)).astype(np.float32)
  b = rng.uniform(-0.1, 0.1, size=(Cout,)).astype(np.float32)
  xb, wb = _bf(x), _bf(w)
  xt = xb[..., :Cw].double().requires_grad_(True)
  wt = wb.double().requires_grad_(True)
  bt = torch.from_numpy(b).double().requires_grad_(True)
  y_ref = O.conv2d_same(xt, wt, bt, stride, relu=True)
  up = _bf(rng.uniform(-1, 1, size=tuple(y_ref.shape)).astype(np.float32))
  dy_pre_ref = up.double() * (y_ref.detach() > 0)
  y_ref.backward(dy_pre_ref)      # d/dy of relu output: equivalent to feeding dy_pre into the pre-activation
  xd = xb.to(cuda_device)
  wd = wb.float().to(cuda_device)
  bd = torch.from_numpy(b).to(cuda_device)
  y, y32 = ops.conv2d_same_bf16(xd, wd, bd, stride=stride, relu=True, want_f32=True)
  torch.cuda.synchronize()
  assert rel_max(y32.cpu().numpy(), y_ref.detach().numpy()) <= 2e-5, 'fwd f32'
  assert rel_max(y.float().cpu().numpy(), y_ref.detach().numpy()) <= 5e-3, 'fwd bf16'
  dyb = _bf(dy_pre_ref.float().numpy()).to(cuda_device).contiguous()
  want_dx = Cw == Cin and Cin >= 16          # the network input (conv1) never needs a data gradient
  mask = xd if want_dx else None
  dw, db, dx = ops.conv2d_same_bwd_bf16(xd, wd, dyb, stride=stride, relu_mask_x=mask, need_dx=want_dx)
  torch.cuda.synchronize()
  assert rel_max(dw.cpu().numpy(), wt.grad.numpy()) <= 2e-5, 'dw'
  assert rel_max(db.cpu().numpy(), bt.grad.numpy()) <= 2e-5, 'db'
  if want_dx:
    ref_dx = xt.grad.numpy() * (xb.double().numpy() > 0)
    assert rel_max(dx.float().cpu().numpy(), ref_dx) <= 5e-3, 'dx'


@pytest.mark.parametrize("case", [(3, 8, 48, 48, 64, 2), (2, 256, 32, 32, 48, 2)])
def test_bit_mask_data_gradient_equals_bf16_mask(cuda_device, case):
  """The model step's 1-bit ReLU mask (uint16 per pixel and 16 channels) gives the same data gradient, bit for bit,
  as masking with the bf16 activation itself."""
  from geeco_b200 import ops
  N, H, Cin, Cw, Cout, stride = case
  rng = np.random.default_rng(sum(case) + 1)
  x = torch.from_numpy(np.maximum(rng.uniform(-1, 1, size=(N, H, H, Cin)), 0).astype(np.float32)).to(torch.bfloat16).cuda()
  w = torch.from_numpy(rng.uniform(-0.2, 0.2, size=(3, 3, Cw, Cout)).astype(np.float32)).cuda()
  dy = torch.from_numpy(rng.uniform(-1, 1, size=(N, H // stride, H // stride, Cout)).astype(np.float32)).to(torch.bfloat16).cuda()
  bits = ops.relu_mask_bits(x)
  xb = x.view(torch.int16).cpu().numpy().reshape(-1, Cin // 16, 8, 2)          # [pixel][chunk][j][lo/hi]
  want = ((xb[..., 0] != 0) << np.arange(8)).sum(-1) + (((xb[..., 1] != 0) << (8 + np.arange(8))).sum(-1))
  assert np.array_equal(bits.cpu().numpy().astype(np.uint16), want.astype(np.uint16))
  _, _, dx_mask = ops.conv2d_same_bwd_bf16(x, w, dy, stride=stride, relu_mask_x=x, need_dw=False)
  _, _, dx_bits = ops.conv2d_same_bwd_bf16(x, w, dy, stride=stride, relu_mask_bits=bits, need_dw=False)
  torch.cuda.synchronize()
  assert torch.equal(dx_mask.view(torch.int16), dx_bits.view(torch.int16))
  assert float((dx_bits.float().abs() * (x == 0)).max()) == 0.0                    # masked positions are exactly zero
