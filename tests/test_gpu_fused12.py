"""The fused conv1 -> conv2 forward kernel (csrc/conv12_fused.cu) against the two separate tensor-core kernels it
replaces: same bf16 roundings, same accumulation order, so every byte must agree -- y1, y2, both 1-bit ReLU masks (seen
through the gradients they gate), head outputs, losses, and the parameters after several Adam steps.  The separate
kernels are themselves held to the oracle in test_gpu_ops_bf16.py / test_gpu_step_bf16.py."""
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


def _run(N, training, steps, monkeypatch, fused):
  if fused:
    monkeypatch.delenv('GEECO_NO_FUSE12', raising=False)
  else:
    monkeypatch.setenv('GEECO_NO_FUSE12', '1')
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
