"""The fused conv1 -> conv2 forward kernel (csrc/conv12_fused.cu) against the two separate tensor-core kernels it
replaces.  Its default (per-pixel) form has the same bf16 roundings and the same accumulation order, so every byte must
agree -- y1, y2, both 1-bit ReLU masks (seen through the gradients they gate), head outputs, losses, and the parameters
after several Adam steps.  The opt-in form GEECO_CONV12_PAIR=1 runs conv1 on pixel pairs: the same products summed in a
different order inside the tensor core, so y1 agrees to fp32 rounding (a bf16 value may land on the neighbouring
representable number) and everything downstream to bf16 noise.  The separate kernels are themselves held to the oracle in
test_gpu_ops_bf16.py / test_gpu_step_bf16.py."""
import numpy as np
import pytest
import torch

from oracle import geeco_oracle as O

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def _engine(N, training, seed=0):
  from geeco_b200 import create_e2evmc_config
  from geeco_b200.data import synthetic_batch
  from geeco_b200.engine import Engine
  cfg_d = O.make_config(batch_size=N)
  P = O.init_params(cfg_d, seed=seed, dtype=torch.float32, bias_scale=0.05)
  feats, labels = synthetic_batch(N, seed=seed + 1)
  eng = Engine(create_e2evmc_config(cfg_d), batch_size=N, precision='bf16', training=training)
  eng.set_params(P)
  return eng, feats, labels


def _run(N, training, steps, monkeypatch, fused, pair=False):
  if fused:
    monkeypatch.delenv('GEECO_NO_FUSE12', raising=False)
  else:
    monkeypatch.setenv('GEECO_NO_FUSE12', '1')
  # the row-wise conv2 weight gradient only runs next to the fused forward: keep the generic kernel on both sides so that
  # the gradients compare bit for bit (tests/test_gpu_bwd21.py holds the row-wise kernel to the generic one)
  monkeypatch.setenv('GEECO_NO_FUSE_WG2', '1')
  if pair:
    monkeypatch.setenv('GEECO_CONV12_PAIR', '1')
  else:
    monkeypatch.delenv('GEECO_CONV12_PAIR', raising=False)
  eng, feats, labels = _engine(N, training)
  out = eng.forward(feats, labels)
  torch.cuda.synchronize()
  res = {k: out[k].detach().cpu().numpy().copy() for k in ('pred_cmd_ee', 'logits_cmd_grp', 'pred_aux_ee', 'pred_aux_obj')}
  res['y2'] = eng.debug_buffer('y2').view(torch.int16).cpu().numpy().copy()
  res['y8'] = eng.debug_buffer('y8').view(torch.int16).cpu().numpy().copy()
  if training:
    res['y1'] = eng.debug_buffer('y1').view(torch.int16).cpu().numpy().copy()
    losses = []
    for _ in range(steps):
      losses.append(eng.train_step(feats, labels).detach().cpu().numpy().copy())
    torch.cuda.synchronize()
    res['losses'] = np.stack(losses)
    res['grads'] = np.concatenate([g.ravel() for g in eng.get_grads().values()])
    res['theta'] = np.concatenate([p.ravel() for p in eng.get_params().values()])
  return res


@pytest.mark.parametrize('N', [1, 5])
def test_fused_training_step_is_bit_identical(cuda_device, monkeypatch, N):
  a = _run(N, True, 3, monkeypatch, fused=True)
  b = _run(N, True, 3, monkeypatch, fused=False)
  for k in b:
    assert np.array_equal(a[k], b[k]), 'fused conv1->conv2 differs from the separate kernels in %s' % k
  assert np.isfinite(a['losses']).all()


@pytest.mark.parametrize('N', [2, 50])
def test_fused_inference_is_bit_identical(cuda_device, monkeypatch, N):
  """Inference contexts do not write y1 at all; N = 50 gives 150 images = CTA ranges that start inside an image and
  ranges that cross image boundaries."""
  a = _run(N, False, 0, monkeypatch, fused=True)
  b = _run(N, False, 0, monkeypatch, fused=False)
  for k in b:
    assert np.array_equal(a[k], b[k]), 'fused conv1->conv2 differs from the separate kernels in %s' % k


def _bf16(a):
  return torch.from_numpy(a.copy()).view(torch.bfloat16).float().numpy()


@pytest.mark.parametrize('N,training', [(1, True), (5, True), (50, False)])
def test_pixel_pair_conv1_matches_separate_kernels(cuda_device, monkeypatch, N, training):
  a = _run(N, training, 1 if training else 0, monkeypatch, fused=True, pair=True)
  b = _run(N, training, 1 if training else 0, monkeypatch, fused=False)
  for k in (('y1', 'y2', 'y8') if training else ('y2', 'y8')):
    x, y = _bf16(a[k]), _bf16(b[k])
    diff = np.abs(x - y)
    # bf16 has 8 bits of mantissa: one representable step is 2^-7 relative
    assert (diff <= np.maximum(np.abs(y), 1e-6) * 2.0 ** -6).all() if k == 'y1' else np.abs(x - y).max() <= 2e-2 * max(np.abs(y).max(), 1e-6), k
    if k == 'y1':
      assert (x != y).mean() < 2e-3, 'more than 0.2 %% of y1 differs: %g' % (x != y).mean()
      assert ((x > 0) != (y > 0)).mean() < 1e-4
  for k in ('pred_cmd_ee', 'logits_cmd_grp', 'pred_aux_ee', 'pred_aux_obj'):
    np.testing.assert_allclose(a[k], b[k], rtol=5e-3, atol=5e-4)
  if training:
    np.testing.assert_allclose(a['losses'], b['losses'], rtol=2e-3, atol=1e-5)
    assert np.isfinite(a['grads']).all()
    den = np.linalg.norm(b['grads'])
    assert np.linalg.norm(a['grads'] - b['grads']) <= 2e-2 * den
