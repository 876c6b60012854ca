"""GPU parity of the full GEECO-F train step in bf16 mode (tcgen05 convs, fp32 master weights) against
the fp32 CPU oracle.  north_star tolerance: <= 2e-2 relative in bf16 after one step."""
import numpy as np
import pytest
import torch

from oracle import geeco_oracle as O
from tests.util import rel_l2, rel_max

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

BF16_REL = 2e-2


def _setup(N, seed=0, precision='bf16', **over):
  from geeco_b200 import create_e2evmc_config
  from geeco_b200.data import synthetic_batch
  from geeco_b200.engine import Engine
  cfg_d = O.make_config(batch_size=N, **over)
  P = O.init_params(cfg_d, seed=seed, dtype=torch.float32, bias_scale=0.05)
  feats, labels = synthetic_batch(N, seed=seed + 1)
  eng = Engine(create_e2evmc_config(cfg_d), batch_size=N, precision=precision, training=True)
  eng.set_params(P)
  return cfg_d, P, feats, labels, eng


def test_bf16_train_step_matches_oracle(cuda_device):
  N = 4
  cfg_d, P, feats, labels, eng = _setup(N)
  P64 = {k: v.double() for k, v in P.items()}
  ref_losses, ref_grads, ep = O.train_step(P64, O.adam_init(P64), feats, labels, cfg_d)
  out = eng.forward(feats, labels, want_dyn=True)
  torch.cuda.synchronize()
  report = {}
  for k in ('pred_cmd_ee', 'logits_cmd_grp', 'pred_aux_ee', 'pred_aux_obj', 'fc1'):
    report[k] = rel_max(out[k].cpu().numpy(), ep[k].detach().numpy())
  acts = ep['obs_acts']
  for li in range(8):
    y = eng.debug_buffer('y%d' % (li + 1)).float().cpu().numpy()
    ref = acts[li].detach().numpy()
    report['y%d' % (li + 1)] = rel_l2(y[:ref.size].reshape(ref.shape), ref)
  got_l = eng.losses_dict(out['losses'])
  for k in ('loss_cmd_ee', 'loss_cmd_grp', 'loss_pos_ee', 'loss_pos_obj', 'loss'):
    report[k] = abs(got_l[k] - ref_losses[k]) / abs(ref_losses[k])
  eng.train_step(feats, labels)
  torch.cuda.synchronize()
  grads = eng.get_grads()
  for k, g in ref_grads.items():
    report['grad:' + k] = rel_l2(grads[k], g.numpy())
  print("\n".join("%-60s %.3e" % kv for kv in report.items()))
  bad = {k: v for k, v in report.items() if not v <= BF16_REL}
  assert not bad, bad
  gk = grads['GoalVMC/LSTMDecoder/lstm_cell/kernel']
  assert np.all(gk[3100:, :] == 0.0) and np.all(gk[:, 256:384] == 0.0)


def test_bf16_step_is_deterministic_and_tracks_fp32_mode(cuda_device):
  N = 2
  cfg_d, P, feats, labels, eng = _setup(N, seed=2, lr=1e-3)
  l1 = eng.train_step(feats, labels).cpu().numpy().copy()
  g1 = eng.grad.clone()
  eng.set_params(P); eng.adam_m.zero_(); eng.adam_v.zero_(); eng.set_global_step(0)
  l2 = eng.train_step(feats, labels).cpu().numpy().copy()
  assert np.array_equal(l1, l2) and torch.equal(g1, eng.grad)
  _, _, _, _, ref = _setup(N, seed=2, precision='fp32', lr=1e-3)
  eng.set_params(P); eng.adam_m.zero_(); eng.adam_v.zero_(); eng.set_global_step(0)
  for step in range(5):
    a = eng.losses_dict(eng.train_step(feats, labels))['loss']
    b = ref.losses_dict(ref.train_step(feats, labels))['loss']
    assert abs(a - b) <= BF16_REL * abs(b), (step, a, b)
