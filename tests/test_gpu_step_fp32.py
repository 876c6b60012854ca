"""GPU parity of the full GEECO-F train / eval step (fp32 mode) against the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import geeco_oracle as O
from tests.util import rel_l2, rel_max

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

FP32_REL = 1e-4      # north_star: <= 1e-4 relative in fp32


def _setup(N, seed=0, **over):
  from geeco_b200 import create_e2evmc_config
  from geeco_b200.data import synthetic_batch
  from geeco_b200.engine import Engine
  cfg_d = O.make_config(batch_size=N, **over)
  cfg = create_e2evmc_config(cfg_d)
  P = O.init_params(cfg_d, seed=seed, dtype=torch.float64, bias_scale=0.05)
  feats, labels = synthetic_batch(N, seed=seed + 1)
  eng = Engine(cfg, batch_size=N, precision='fp32', training=True)
  eng.set_params({k: v.float() for k, v in P.items()})
  return cfg_d, P, feats, labels, eng


def test_parameter_table(cuda_device):
  cfg_d, P, feats, labels, eng = _setup(1)
  assert eng.count_parameters() == 7552796 == O.count_parameters(P)
  assert set(eng.param_names()) == set(P.keys())
  for k, v in P.items():
    assert tuple(eng.view(k).shape) == tuple(v.shape)


def test_train_step_matches_oracle(cuda_device):
  N = 2
  cfg_d, P, feats, labels, eng = _setup(N)
  opt = O.adam_init(P)
  ref_losses, ref_grads, ep = O.train_step(P, opt, feats, labels, cfg_d)
  # forward endpoints first (no update)
  out = eng.forward(feats, labels, want_dyn=True)
  torch.cuda.synchronize()
  assert np.abs(out['dynbuff'].cpu().numpy() - ep['dynbuff'].detach().numpy()).max() <= 1e-5
  assert np.abs(out['dyndiff'].cpu().numpy() - ep['dyndiff'].detach().numpy()).max() <= 1e-5
  for k in ('pred_cmd_ee', 'logits_cmd_grp', 'pred_aux_ee', 'pred_aux_obj', 'fc1'):
    assert rel_max(out[k].cpu().numpy(), ep[k].detach().numpy()) <= FP32_REL, k
  acts = ep['obs_acts']
  for li in range(8):
    y = eng.debug_buffer('y%d' % (li + 1)).cpu().numpy()
    ref = acts[li].detach().numpy()
    got = y[:ref.size].reshape(ref.shape)       # encoder 0 = current-frame encoder comes first
    assert rel_max(got, ref) <= FP32_REL, 'y%d' % (li + 1)
  got_l = eng.losses_dict(out['losses'])
  for k in ('loss_cmd_ee', 'loss_cmd_grp', 'loss_pos_ee', 'loss_pos_obj', 'loss'):
    assert abs(got_l[k] - ref_losses[k]) <= FP32_REL * abs(ref_losses[k]) + 1e-7, (k, got_l[k], ref_losses[k])
  # full step
  theta0 = eng.get_params()
  losses = eng.train_step(feats, labels)
  torch.cuda.synchronize()
  grads = eng.get_grads()
  worst = 0.0
  for k, g in ref_grads.items():
    e = rel_max(grads[k], g.numpy())
    worst = max(worst, e)
    assert e <= FP32_REL, (k, e)
  # Adam: theta moved by lr_1 * g/(|g| + eps/sqrt(1-b2)) (SURVEY 8c pin 7); compare against the oracle update
  theta1 = eng.get_params()
  lr = cfg_d['lr']
  for k, p_ref in P.items():
    d_ref = p_ref.numpy() - theta0[k].astype(np.float64)
    d_got = theta1[k].astype(np.float64) - theta0[k].astype(np.float64)
    g = ref_grads[k].numpy()
    sure = np.abs(g) > 1e-6 * (np.abs(g).max() + 1e-30)     # away from g ~ 0 the step is -lr*sign(g)
    assert np.abs(d_got - d_ref)[sure].max(initial=0.0) <= 0.02 * lr, k
    assert np.abs(d_got).max() <= lr * 1.0001
  # forget-gate columns and h-rows of the LSTM kernel get exactly zero gradient
  gk = grads['GoalVMC/LSTMDecoder/lstm_cell/kernel']
  assert np.all(gk[3100:, :] == 0.0) and np.all(gk[:, 256:384] == 0.0)
  assert eng.global_step == 1


def test_eval_forward_has_no_side_effects_and_is_deterministic(cuda_device):
  cfg_d, P, feats, labels, eng = _setup(2, seed=3)
  a = eng.forward(feats, labels)['losses'].cpu().numpy().copy()
  th = eng.theta.clone()
  b = eng.forward(feats, labels)['losses'].cpu().numpy().copy()
  assert np.array_equal(a, b)
  assert torch.equal(th, eng.theta)
  l1 = eng.train_step(feats, labels).cpu().numpy().copy()
  g1 = eng.grad.clone()
  eng.set_params({k: v.float() for k, v in P.items()})
  eng.adam_m.zero_(); eng.adam_v.zero_(); eng.set_global_step(0)
  l2 = eng.train_step(feats, labels).cpu().numpy().copy()
  assert np.array_equal(l1, l2) and torch.equal(g1, eng.grad)      # bitwise reproducible


def test_phased_step_equals_fused_step(cuda_device):
  cfg_d, P, feats, labels, eng = _setup(2, seed=4)
  eng.train_step(feats, labels)
  th_fused = eng.theta.clone()
  eng.set_params({k: v.float() for k, v in P.items()})
  eng.adam_m.zero_(); eng.adam_v.zero_(); eng.set_global_step(0)
  eng.step_forward(feats, labels)
  for b in range(len(eng.buckets)):
    eng.step_backward(b)
  eng.step_update(1.0)
  assert torch.equal(th_fused, eng.theta)
  # ... and with the optimizer step cut at a bucket boundary (what the data-parallel step does while the last
  # bucket's all-reduce is in flight); two steps, so that the step counter / bias correction is covered too
  eng.train_step(feats, labels)
  th2 = eng.theta.clone()
  eng.set_params({k: v.float() for k, v in P.items()})
  eng.adam_m.zero_(); eng.adam_v.zero_(); eng.set_global_step(0)
  nb = len(eng.buckets)
  for _ in range(2):
    eng.step_forward(feats, labels)
    for b in range(nb):
      eng.step_backward(b)
    eng.step_update_buckets(1.0, 0, nb - 2)
    eng.step_update_buckets(1.0, nb - 1, nb - 1)
  assert torch.equal(th2, eng.theta)
  assert eng.global_step == 2


def test_three_steps_track_oracle(cuda_device):
  N = 2
  cfg_d, P, feats, labels, eng = _setup(N, seed=5, lr=1e-3)
  opt = O.adam_init(P)
  for step in range(3):
    ref_losses, _, _ = O.train_step(P, opt, feats, labels, cfg_d)
    got = eng.losses_dict(eng.train_step(feats, labels))
    assert abs(got['loss'] - ref_losses['loss']) <= 2e-3 * abs(ref_losses['loss']), (step, got['loss'], ref_losses['loss'])


def test_error_conventions(cuda_device):
  from geeco_b200 import create_e2evmc_config
  from geeco_b200.engine import Engine
  with pytest.raises(ValueError):
    Engine(create_e2evmc_config(O.make_config(proc_obs='bogus')), batch_size=1)
  with pytest.raises(ValueError):
    Engine(create_e2evmc_config(O.make_config(img_channels=5)), batch_size=1)
  with pytest.raises(ValueError):
    Engine(create_e2evmc_config(O.make_config(img_height=128, img_width=128)), batch_size=1)
  cfg_d, P, feats, labels, eng = _setup(1, seed=6)
  bad = dict(feats)
  bad['rgb'] = feats['rgb'][:, :, :128]
  with pytest.raises(ValueError):
    eng.forward(bad)
