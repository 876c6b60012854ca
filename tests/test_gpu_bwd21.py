"""The fused conv2-data-gradient -> conv1-weight-gradient kernel (csrc/conv21_bwd_fused.cu) against the two separate
tensor-core kernels it replaces.  Both paths round dL/d(pre-activation of conv1) to bf16 at the same point and apply the
same 1-bit ReLU mask; what differs is the order in which the fp32 accumulators sum over pixels (per-CTA pixel ranges
here, split-K ranges there), so conv1's gradients agree to fp32 summation error and every other gradient, which the
fused kernel does not touch, agrees bit for bit.  The separate kernels are themselves held to the oracle in
test_gpu_ops_bf16.py / test_gpu_step_bf16.py / test_gpu_parity_sizes.py (which now run the fused kernel, too)."""
import numpy as np
import pytest
import torch

from oracle import geeco_oracle as O
from tests.util import rel_l2, rel_max

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

CONV1 = ('ConvEncoder/conv1/', 'DynBuffEncoder/conv1/', 'DynDiffEncoder/conv1/')


def _run(N, monkeypatch, fused, steps=1, wg2=True, want_y1=False, wg2_mode='load'):
  from geeco_b200 import create_e2evmc_config
  from geeco_b200.data import synthetic_batch
  from geeco_b200.engine import Engine
  if fused:
    monkeypatch.delenv('GEECO_NO_FUSE_BWD21', raising=False)
  else:
    monkeypatch.setenv('GEECO_NO_FUSE_BWD21', '1')
  # conv2 weight gradient on whole y1 rows: 'load' (default: rows by TMA, windows stacked into M = 128), 'nostack' (six
  # M = 64 accumulators), 'recompute' (rows rebuilt from x0, the forward does not store y1)
  monkeypatch.delenv('GEECO_WG2_RECOMPUTE', raising=False)
  monkeypatch.delenv('GEECO_WG2_NO_STACK', raising=False)
  if wg2_mode == 'recompute':
    monkeypatch.setenv('GEECO_WG2_RECOMPUTE', '1')
  elif wg2_mode == 'nostack':
    monkeypatch.setenv('GEECO_WG2_NO_STACK', '1')
  if wg2:
    monkeypatch.delenv('GEECO_NO_FUSE_WG2', raising=False)
  else:
    monkeypatch.setenv('GEECO_NO_FUSE_WG2', '1')
  cfg_d = O.make_config(batch_size=N)
  P = O.init_params(cfg_d, seed=3, dtype=torch.float32, bias_scale=0.05)
  feats, labels = synthetic_batch(N, seed=4)
  eng = Engine(create_e2evmc_config(cfg_d), batch_size=N, precision='bf16', training=True)
  eng.set_params(P)
  losses = []
  for _ in range(steps):
    losses.append(eng.train_step(feats, labels).detach().cpu().numpy().copy())
  torch.cuda.synchronize()
  grads = {k: v.copy() for k, v in eng.get_grads().items()}
  theta = {k: v.copy() for k, v in eng.get_params().items()}
  y1 = eng.debug_buffer('y1').view(torch.int16).cpu().numpy().copy() if want_y1 else None
  eng.close()
  return np.stack(losses), grads, theta, y1


# N = 1: CTA ranges of 2-3 units inside one image; 5: ranges that start mid-image and cross image boundaries;
# 50: 150 images per encoder group on 49 CTAs
@pytest.mark.parametrize('N', [1, 5, 50])
def test_fused_backward_matches_separate_kernels(cuda_device, monkeypatch, N):
  la, ga, ta, _ = _run(N, monkeypatch, fused=True)
  lb, gb, tb, _ = _run(N, monkeypatch, fused=False)
  assert np.array_equal(la, lb)
  seen = 0
  for k in gb:
    if any(c in k for c in CONV1):
      seen += 1
      # fp32 sums of ~65 k * N bf16 products in a different order
      assert rel_l2(ga[k], gb[k]) <= 2e-5, (k, rel_l2(ga[k], gb[k]))
      assert rel_max(ga[k], gb[k]) <= 1e-4, (k, rel_max(ga[k], gb[k]))
      assert np.abs(gb[k]).max() > 0
    else:
      assert np.array_equal(ga[k], gb[k]), 'the fused backward changed %s' % k
  assert seen == 6          # kernel + bias of the three encoders


def test_fused_backward_is_reproducible_and_trains(cuda_device, monkeypatch):
  la, ga, ta, _ = _run(3, monkeypatch, fused=True, steps=4)
  lb, gb, tb, _ = _run(3, monkeypatch, fused=True, steps=4)
  for k in ga:
    assert np.array_equal(ga[k], gb[k]), 'run-to-run difference in %s' % k
  for k in ta:
    assert np.array_equal(ta[k], tb[k])
  lc, gc, tc, _ = _run(3, monkeypatch, fused=False, steps=4)
  assert np.isfinite(la).all()
  # four Adam steps later the two paths are still the same model to fp32 rounding of conv1's update
  np.testing.assert_allclose(la, lc, rtol=2e-4, atol=1e-6)


def test_tile_repack_is_bit_identical(cuda_device, monkeypatch):
  """From a context's second step on, the conv4-conv8 forward operands and all data-gradient operands are repacked by
  pack_tiles_kernel (one 32 x 32 tile per block) instead of the element-wise kernel: same values, same rounding."""
  la, ga, ta, _ = _run(2, monkeypatch, fused=True, steps=4)
  monkeypatch.setenv('GEECO_PACK_NO_TILES', '1')
  lb, gb, tb, _ = _run(2, monkeypatch, fused=True, steps=4)
  assert np.array_equal(la, lb)
  for k in ta:
    assert np.array_equal(ta[k], tb[k]), 'tile repack changed %s' % k
  for k in ga:
    assert np.array_equal(ga[k], gb[k]), 'tile repack changed the gradient of %s' % k


CONV2 = ('ConvEncoder/conv2/', 'DynBuffEncoder/conv2/', 'DynDiffEncoder/conv2/')


# conv2's weight gradient on whole y1 rows (csrc/conv2_wgrad_fused.cu; rows loaded by TMA or recomputed on chip) against
# the generic split-K kernel: the recomputed rows are bit-identical to the stored ones, so only the fp32 summation order
# over pixels differs
@pytest.mark.parametrize('mode', ['load', 'nostack', 'recompute'])
@pytest.mark.parametrize('N', [1, 5, 50])
def test_row_weight_gradient_matches_generic_kernel(cuda_device, monkeypatch, N, mode):
  la, ga, ta, ya = _run(N, monkeypatch, fused=True, wg2=True, want_y1=True, wg2_mode=mode)
  lb, gb, tb, yb = _run(N, monkeypatch, fused=True, wg2=False, want_y1=True)
  assert np.array_equal(la, lb)
  # 'recompute': y1 was not stored by the first run; geeco_debug_buffer rebuilds it, bit-identical to the stored one
  assert np.array_equal(ya, yb)
  seen = 0
  for k in gb:
    if any(c in k for c in CONV2):
      seen += 1
      assert rel_l2(ga[k], gb[k]) <= 2e-5, (k, rel_l2(ga[k], gb[k]))
      assert rel_max(ga[k], gb[k]) <= 1e-4, (k, rel_max(ga[k], gb[k]))
      assert np.abs(gb[k]).max() > 0
    else:
      assert np.array_equal(ga[k], gb[k]), 'the fused conv2 weight gradient changed %s' % k
  assert seen == 6
