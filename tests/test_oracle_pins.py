"""Known-answer pins of the CPU oracle.  The reference ships no tests ("parity
unpinned"), so these closed-form answers are what the restatement is anchored on."""
from fractions import Fraction

import numpy as np
import pytest
import torch

from oracle import geeco_oracle as O
from oracle import np_restatement as R


def test_alpha_tables_exact_rationals():
  assert O.alpha_table_exact(2) == [Fraction(-1, 2), Fraction(1, 2)]
  assert O.alpha_table_exact(3) == [Fraction(-4, 3), Fraction(2, 3), Fraction(2, 3)]
  assert O.alpha_table_exact(4) == [Fraction(-29, 12), Fraction(7, 12), Fraction(13, 12), Fraction(3, 4)]
  for K in range(2, 17):
    assert sum(O.alpha_table_exact(K)) == 0
  k8 = [-8.460714, -1.460714, 1.039286, 2.039286, 2.289286, 2.089286, 1.589286, 0.875]
  assert np.allclose([float(a) for a in O.alpha_table_exact(8)], k8, atol=1e-6)
  k16 = [-25.472393, -10.472393, -3.972393, -0.305726, 1.944274, 3.344274, 4.177607, 4.606179, 4.731179, 4.620067,
         4.320067, 3.865522, 3.282189, 2.589881, 1.804167, 0.9375]
  assert np.allclose([float(a) for a in O.alpha_table_exact(16)], k16, atol=1e-6)


def test_alpha_table_fp32_emulation():
  a4 = O.alpha_table_f32(4)
  assert a4.dtype == np.float32
  # fp32 in-graph evaluation of graph.py:17-28
  assert np.allclose(a4, [-2.416668, 0.58333254, 1.0833325, 0.7499994], atol=2e-7)
  assert abs(float(a4.astype(np.float64).sum())) < 5e-6
  a2 = O.alpha_table_f32(2)
  assert a2[0] == np.float32(-0.5) and a2[1] == np.float32(0.5)
  for K in range(2, 17):
    assert np.allclose(O.alpha_table_f32(K), [float(a) for a in O.alpha_table_exact(K)], atol=2e-5)
    assert np.allclose(R.alpha_exact(K), [float(a) for a in O.alpha_table_exact(K)], atol=1e-12)


def test_dyndiff_is_normalised_half_difference():
  rng = np.random.default_rng(0)
  cur = torch.tensor(rng.uniform(size=(2, 8, 8, 3)))
  tgt = torch.tensor(rng.uniform(size=(2, 8, 8, 3)))
  d = 0.5 * (tgt - cur)
  mn = d.reshape(2, -1).min(dim=1).values.reshape(2, 1, 1, 1)
  mx = d.reshape(2, -1).max(dim=1).values.reshape(2, 1, 1, 1)
  assert torch.allclose(O.dyndiff(cur, tgt), (d - mn) / (mx - mn + 1e-6), atol=1e-12)
  assert torch.all(O.dyndiff(cur, cur) == 0)
  out = O.dynimg(torch.tensor(rng.uniform(size=(2, 4, 8, 8, 3))))
  assert float(out.min()) == 0.0 and float(out.max()) <= 1.0


def test_same_padding_rule_and_index_map():
  assert O.same_pad(256, 3, 1) == (256, 1, 1)
  assert O.same_pad(256, 3, 2) == (128, 0, 1)       # stride 2, even size: (0 before, 1 after)
  assert O.same_pad(4, 3, 2) == (2, 0, 1)
  assert O.same_pad(5, 3, 2) == (3, 1, 1)
  w = torch.zeros(3, 3, 1, 1, dtype=torch.float64)
  for ky in range(3):
    for kx in range(3):
      w[ky, kx, 0, 0] = 1 + ky * 3 + kx
  seen = {}
  for iy in range(4):
    for ix in range(4):
      x = torch.zeros(1, 4, 4, 1, dtype=torch.float64)
      x[0, iy, ix, 0] = 1
      y = O.conv2d_same(x, w, None, 2, relu=False)[0, :, :, 0]
      for oy in range(2):
        for ox in range(2):
          if y[oy, ox] != 0:
            seen.setdefault((oy, ox), set()).add((iy, ix))
            assert int(y[oy, ox]) - 1 == (iy - 2 * oy) * 3 + (ix - 2 * ox)
  assert seen[(0, 0)] == {(r, c) for r in range(3) for c in range(3)}     # rows/cols {0,1,2}
  assert seen[(1, 1)] == {(r, c) for r in (2, 3) for c in (2, 3)}          # {2,3,pad}
  cols, _ = R.im2col(np.arange(16, dtype=np.float64).reshape(1, 4, 4, 1), 2)
  assert cols.reshape(2, 2, 9)[1, 1].tolist() == [10, 11, 0, 14, 15, 0, 0, 0, 0]


def test_state_flatten_order():
  N, D, J = 2, 256, 7
  obs = torch.arange(N * 4 * D, dtype=torch.float64).reshape(N, 2, 2, D)
  dyn, tgt = obs + 10000, obs + 20000
  jnt = torch.arange(N * J, dtype=torch.float64).reshape(N, J) - 50
  st = O.representation_concatenation_v2(obs, dyn, jnt, tgt)
  assert st.shape == (N, 3100)
  for h in range(2):
    for w in range(2):
      cell = h * 2 + w
      assert st[1, cell * 775 + 5] == obs[1, h, w, 5]
      assert st[1, cell * 775 + 256 + 7] == dyn[1, h, w, 7]
      assert st[1, cell * 775 + 512 + 3] == jnt[1, 3]
      assert st[1, cell * 775 + 519 + 255] == tgt[1, h, w, 255]
      assert R.state_index(cell, 'tgt', 255, (256, 256, 256), 7) == cell * 775 + 519 + 255


def test_lstm_cell_by_hand_and_reset_invariance():
  x = torch.tensor([[1.0, -2.0]], dtype=torch.float64)
  h = 1
  W = torch.tensor([[0.5, -0.25, 0.1, 0.3], [0.2, 0.4, -0.6, 0.7], [9.0, 9.0, 9.0, 9.0]], dtype=torch.float64)
  b = torch.tensor([0.1, -0.1, 0.2, 0.0], dtype=torch.float64)
  state = torch.zeros(1, 2, dtype=torch.float64)
  m, new = O.lstm_cell(x, state, W, b)
  g = x @ W[:2] + b          # m_prev = 0: the h-row of the kernel contributes nothing
  sig = lambda v: 1 / (1 + np.exp(-v))
  i, j, f, o = [float(v) for v in g[0]]
  c = sig(f + 1.0) * 0.0 + sig(i) * np.tanh(j)
  assert abs(float(new[0, 0]) - c) < 1e-12 and abs(float(m[0, 0]) - sig(o) * np.tanh(c)) < 1e-12
  assert float(new[0, 1]) == float(m[0, 0])          # state = [c | m]
  st2 = torch.tensor([[0.3, -0.4]], dtype=torch.float64)
  m2, new2 = O.lstm_cell(x, st2, W, b)
  g2 = g + (-0.4) * W[2]
  i, j, f, o = [float(v) for v in g2[0]]
  c2 = sig(f + 1.0) * 0.3 + sig(i) * np.tanh(j)
  assert abs(float(new2[0, 0]) - c2) < 1e-12


def test_loss_reductions_and_rint():
  pred = torch.tensor([[1.0, 2.0, 3.0], [0.0, 0.0, 0.0]], dtype=torch.float64)
  tgt = torch.tensor([[0.0, 0.0, 0.0], [1.0, 1.0, 1.0]], dtype=torch.float64)
  assert float(O.mean_squared_error(pred, tgt)) == pytest.approx((1 + 4 + 9 + 3) / 6.0)
  cls = O.gripper_classes(torch.tensor([-0.5, 0.5, 1.5, -1.0, 0.49, 2.5]))
  assert cls.tolist() == [1, 1, 3, 0, 1, 3]          # rint is round-half-to-even: +-0.5 -> 0, 1.5 -> 2, 2.5 -> 2
  logits = torch.tensor([[1.0, 2.0, 3.0], [0.0, 0.0, 0.0]], dtype=torch.float64)
  ce = O.softmax_cross_entropy(logits, torch.tensor([2, 0]), 3)
  exp = 0.5 * (-(3.0 - np.log(np.exp(1) + np.exp(2) + np.exp(3))) + np.log(3.0))
  assert float(ce) == pytest.approx(exp)


def test_adam_step_one_closed_form():
  g = torch.tensor([0.3, -2.0, 1e-3, 0.0], dtype=torch.float64)
  p = {'w': torch.zeros(4, dtype=torch.float64)}
  st = O.adam_init(p)
  lr = 1e-4
  O.adam_update(p, {'w': g}, st, lr)
  exp = -lr * g / (g.abs() + 1e-8 / np.sqrt(0.001))       # TF: eps outside the bias correction (3.162e-7)
  assert torch.allclose(p['w'], exp, rtol=1e-9, atol=0)
  torch_style = -lr * g / (g.abs() + 1e-8)
  assert not torch.allclose(p['w'][2:3], torch_style[2:3], rtol=1e-5)
  assert st['t'] == 1


def test_parameter_count_and_names():
  cfg = O.make_config()
  P = O.init_params(cfg, seed=0)
  assert O.count_parameters(P) == 7552796
  assert P['GoalVMC/LSTMDecoder/lstm_cell/kernel'].shape == (3228, 512)
  assert P['GoalVMC/DynDiffEncoder/conv1/kernel'].shape == (3, 3, 3, 32)
  assert all(float(v.abs().max()) == 0.0 for k, v in P.items() if k.endswith('/bias'))
  lim = (6.0 / (27 + 288)) ** 0.5
  assert float(P['GoalVMC/ConvEncoder/conv1/kernel'].abs().max()) <= lim


@pytest.mark.timeout(300)
def test_torch_oracle_vs_numpy_restatement_and_finite_differences():
  from geeco_b200.data import synthetic_batch
  cfg = O.make_config(batch_size=1)
  P = O.init_params(cfg, seed=3, dtype=torch.float64, bias_scale=0.05)
  f, l = synthetic_batch(1, seed=4)
  losses, grads, ep = O.train_step({k: v.clone() for k, v in P.items()}, O.adam_init(P), f, l, cfg)
  Pn = {k: v.numpy() for k, v in P.items()}
  L, epn, G = R.geeco_f_forward_backward(
      Pn, f['rgb'].astype(np.float64), f['jnt_state'].astype(np.float64), f['target_rgb'].astype(np.float64),
      l['cmd'].astype(np.float64), f['ee_state'].astype(np.float64), f['obj_state'].astype(np.float64),
      alpha=O.alpha_table_f32(4).astype(np.float64))
  assert abs(L['loss'] - losses['loss']) < 1e-12
  for k in G:
    assert np.abs(G[k] - grads[k].numpy()).max() <= 1e-12 * (1 + np.abs(G[k]).max()), k
  # central finite differences on a few scalars of the fp64 oracle
  probes = [('GoalVMC/LSTMDecoder/fc1/bias', (5,)), ('GoalVMC/ConvEncoder/conv8/bias', (17,)),
            ('GoalVMC/DynBuffEncoder/conv3/kernel', (1, 2, 5, 7)), ('GoalVMC/DynDiffEncoder/conv1/kernel', (0, 1, 2, 3))]
  for name, idx in probes:
    vals = []
    for sgn in (+1, -1):
      Q = {k: v.clone() for k, v in P.items()}
      Q[name][idx] += sgn * 1e-5
      vals.append(float(O.forward_losses(Q, f, l, cfg)[0]['loss']))
    fd = (vals[0] - vals[1]) / 2e-5
    assert abs(fd - float(grads[name][idx])) <= 1e-5 * max(1.0, abs(fd)) + 1e-8, (name, fd, float(grads[name][idx]))


def test_bf16_emulation_only_changes_storage_roundings():
  from geeco_b200.data import synthetic_batch
  cfg = O.make_config(batch_size=1)
  P = O.init_params(cfg, seed=1, dtype=torch.float32, bias_scale=0.05)
  f, l = synthetic_batch(1, seed=2)
  a, _ = O.forward_losses(P, f, l, cfg)
  b, _ = O.forward_losses(P, f, l, cfg, emulate_bf16=True)
  assert abs(float(a['loss']) - float(b['loss'])) <= 2e-2 * abs(float(a['loss']))
  assert float(a['loss']) != float(b['loss'])
