"""GPU parity of the functional ops (fp32 kernels) against the CPU oracle.  Calls go through the
C-ABI (geeco_b200.ops -> libgeeco_b200.so)."""
import numpy as np
import pytest
import torch

from oracle import geeco_oracle as O
from tests.util import rel_l2, rel_max

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def _frames(N, K, H, W, C, seed):
  rng = np.random.default_rng(seed)
  return (rng.integers(0, 256, size=(N, K, H, W, C)) / 255.0).astype(np.float32)


@pytest.mark.parametrize("K", [2, 3, 4, 8, 16])
@pytest.mark.parametrize("shape", [(3, 32, 32, 3), (2, 128, 128, 3), (2, 256, 256, 3), (1, 64, 64, 4)])
def test_dynimg_matches_oracle(cuda_device, K, shape):
  from geeco_b200 import ops
  N, H, W, C = shape
  x = _frames(N, K, H, W, C, seed=K * 7 + H)
  ref = O.dynimg(torch.from_numpy(x)).numpy()
  ref64 = O.dynimg(torch.from_numpy(x).double(), alpha=O.alpha_table_f32(K)).numpy()
  xd = torch.from_numpy(x).to(cuda_device)
  for cluster in (0, 1, 2, 4, 8, -1):
    got = ops.dynimg(xd, cluster=cluster).cpu().numpy()
    # tolerance: SURVEY 7.2 -- <= 1e-5 abs vs the fp64 oracle on non-degenerate inputs
    assert np.abs(got - ref64).max() <= 1e-5, (cluster, np.abs(got - ref64).max())
    assert np.abs(got - ref).max() <= 1e-5
    assert got.min() >= 0.0 and got.max() <= 1.0


def test_dynimg_cluster16_and_512px(cuda_device):
  from geeco_b200 import ops
  x = _frames(2, 4, 512, 512, 3, seed=5)
  ref64 = O.dynimg(torch.from_numpy(x).double(), alpha=O.alpha_table_f32(4)).numpy()
  xd = torch.from_numpy(x).to(cuda_device)
  for cluster in (0, 16, -1):
    got = ops.dynimg(xd, cluster=cluster).cpu().numpy()
    assert np.abs(got - ref64).max() <= 1e-5, cluster


def test_dyndiff_identical_frames_is_exactly_zero(cuda_device):
  """graph.py:397-400 with tgt == cur: d = 0.5*(tgt-cur) = 0 -> (0-0)/(0+1e-6) = 0 exactly."""
  from geeco_b200 import ops
  f = _frames(2, 1, 64, 64, 3, seed=1)
  x = np.concatenate([f, f], axis=1)
  got = ops.dynimg(torch.from_numpy(x).to(cuda_device)).cpu().numpy()
  assert np.all(got == 0.0)


def test_dynimg_custom_alpha_and_empty_batch(cuda_device):
  from geeco_b200 import ops
  x = _frames(2, 4, 32, 32, 3, seed=2)
  a = np.array([0.25, -1.0, 0.5, 2.0], dtype=np.float32)
  ref = O.dynimg(torch.from_numpy(x).double(), alpha=a).numpy()
  got = ops.dynimg(torch.from_numpy(x).to(cuda_device), alpha=a).cpu().numpy()
  assert np.abs(got - ref).max() <= 1e-5
  empty = torch.zeros((0, 4, 32, 32, 3), device=cuda_device)
  assert ops.dynimg(empty).shape == (0, 32, 32, 3)
  with pytest.raises(ValueError):
    ops.dynimg(torch.zeros((1, 17, 8, 8, 4), device=cuda_device))


CONV_CASES = [
    # N, H, Cin, Cout, stride
    (2, 16, 4, 32, 1), (2, 16, 32, 48, 2), (3, 8, 48, 64, 2), (2, 8, 64, 128, 2), (5, 4, 128, 192, 2),
    (2, 4, 256, 256, 2), (1, 32, 8, 16, 1), (2, 6, 16, 24, 2),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_fwd_bwd_matches_oracle(cuda_device, case):
  from geeco_b200 import ops
  N, H, Cin, Cout, stride = case
  rng = np.random.default_rng(sum(case))
  x = rng.uniform(0, 1, size=(N, H, H, Cin)).astype(np.float32)
  w = rng.uniform(-0.2, 0.2, size=(3, 3, Cin, Cout)).astype(np.float32)
  b = rng.uniform(-0.1, 0.1, size=(Cout,)).astype(np.float32)
  xt = torch.from_numpy(x).double().requires_grad_(True)
  wt = torch.from_numpy(w).double().requires_grad_(True)
  bt = torch.from_numpy(b).double().requires_grad_(True)
  y_ref = O.conv2d_same(xt, wt, bt, stride, relu=True)
  up = torch.from_numpy(rng.uniform(-1, 1, size=tuple(y_ref.shape))).double()
  (y_ref * up).sum().backward()
  xd, wd, bd = (torch.from_numpy(a).to(cuda_device) for a in (x, w, b))
  y = ops.conv2d_same(xd, wd, bd, stride=stride, relu=True)
  assert tuple(y.shape) == tuple(y_ref.shape)
  assert rel_max(y.cpu().numpy(), y_ref.detach().numpy()) <= 1e-5
  dy_pre = (up * (y_ref.detach() > 0)).float().to(cuda_device).contiguous()
  dw, db, dx = ops.conv2d_same_bwd(xd, wd, dy_pre, stride=stride)
  assert rel_max(dw.cpu().numpy(), wt.grad.numpy()) <= 1e-5
  assert rel_max(db.cpu().numpy(), bt.grad.numpy()) <= 1e-5
  assert rel_max(dx.cpu().numpy(), xt.grad.numpy()) <= 1e-5


def test_conv2d_same_padding_index_map(cuda_device):
  """SURVEY 8c pin (3): stride-2 SAME on a 4x4 delta image pads (0 before, 1 after)."""
  from geeco_b200 import ops
  Cin, Cout = 4, 4
  w = np.zeros((3, 3, Cin, Cout), dtype=np.float32)
  for ky in range(3):
    for kx in range(3):
      w[ky, kx, 0, 0] = 1.0 + ky * 3 + kx       # tap id + 1 in channel 0
  for (iy, ix) in [(0, 0), (2, 2), (3, 3), (1, 2)]:
    x = np.zeros((1, 4, 4, Cin), dtype=np.float32)
    x[0, iy, ix, 0] = 1.0
    y = ops.conv2d_same(torch.from_numpy(x).to(cuda_device), torch.from_numpy(w).to(cuda_device), None, stride=2,
                        relu=False).cpu().numpy()[0, :, :, 0]
    exp = np.zeros((2, 2), dtype=np.float32)
    for oy in range(2):
      for ox in range(2):
        ky, kx = iy - 2 * oy, ix - 2 * ox       # pad_before = 0
        if 0 <= ky < 3 and 0 <= kx < 3:
          exp[oy, ox] = 1.0 + ky * 3 + kx
    assert np.array_equal(y, exp), (iy, ix, y, exp)


@pytest.mark.parametrize('N,K,xdim,Hl', [(3, 1, 1052, 128), (5, 4, 2076, 128), (2, 3, 8, 8), (70, 2, 1052, 64)])
def test_lstm_sequence_forward_and_bptt(cuda_device, N, K, xdim, Hl):
  """K-step LSTM recurrence (graph.py:212-225, `--proc_obs sequence` / e2e_vmc) against the oracle's lstm_cell chain
  and its autograd gradients in float64."""
  from geeco_b200 import ops
  g = torch.Generator().manual_seed(100 * N + K)
  x = torch.randn(K, N, xdim, generator=g, dtype=torch.float64) * 0.5
  lim = (6.0 / (xdim + Hl + 4 * Hl)) ** 0.5
  kernel = ((torch.rand(xdim + Hl, 4 * Hl, generator=g, dtype=torch.float64) * 2 - 1) * lim * 3).requires_grad_(True)
  bias = (torch.randn(4 * Hl, generator=g, dtype=torch.float64) * 0.1).requires_grad_(True)
  r = torch.randn(N, Hl, generator=g, dtype=torch.float64)
  xr = x.clone().requires_grad_(True)
  state = torch.zeros(N, 2 * Hl, dtype=torch.float64)
  for t in range(K):
    out, state = O.lstm_cell(xr[t], state, kernel, bias)
  (out * r).sum().backward()
  dev = lambda t: t.detach().float().to(cuda_device).contiguous()
  m_last, st, saved = ops.lstm_sequence([dev(x[t]) for t in range(K)], dev(kernel), dev(bias))
  assert rel_max(m_last.cpu().double(), out.detach()) <= 1e-5
  assert rel_max(st.cpu().double(), state.detach()) <= 1e-5
  dk, db, dx = ops.lstm_sequence_bwd(saved, dev(r))
  assert rel_max(dk.cpu().double(), kernel.grad) <= 1e-4
  assert rel_max(db.cpu().double(), bias.grad) <= 1e-4
  assert rel_max(dx.cpu().double(), xr.grad) <= 1e-4
  dk2, db2, none = ops.lstm_sequence_bwd(saved, dev(r), need_dx=False)      # deterministic, dx optional
  assert none is None and torch.equal(dk2, dk) and torch.equal(db2, db)
  with pytest.raises(ValueError):
    ops.lstm_sequence(dev(x)[:, :, :xdim - 1].contiguous(), dev(kernel), dev(bias))
