"""Oracle vs. fixtures produced by EXECUTING the reference's graph.py (tests/golden/make_golden.py).

The fixture pins what the restatement could get wrong by misreading the reference: layer order and
strides, variable names / shapes / creation order, concat order of the state vector, which encoder
feeds which slot, zero LSTM state, loss assembly (lambda_aux, L2 term over every variable), class
shift of the gripper label, and the full backward pass (per-variable gradient norm, sum, samples).
"""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import geeco_oracle as O
from tests.golden import cases as C
from tests.util import rel_max

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'geeco_graph_golden.npz')
TOL = 2e-6        # float64 oracle vs float64 reference graph; the fp32 alpha table is the only fp32 piece


@pytest.fixture(scope='module')
def golden():
  with np.load(GOLDEN) as z:
    return {k: z[k] for k in z.files}


def _run_oracle(name):
  over, N, seed, goal = C.CASES[name]
  cfg_d = O.make_config(batch_size=N, **over)
  feats, labels = C.make_inputs(N, cfg_d['window_size'], seed, C=cfg_d['img_channels'])
  P = O.init_params(cfg_d, seed=seed, goal=goal, dtype=torch.float64, bias_scale=C.BIAS_SCALE)
  leaves = {k: v.clone().requires_grad_(True) for k, v in P.items()}
  f = {k: torch.tensor(v) for k, v in feats.items()}
  if goal:
    net, ep = O.goal_e2evmc(f['rgb'], f['jnt_state'], f['target_rgb'], leaves, cfg_d)
  else:
    net, ep = O.e2e_vmc(f['rgb'], f['jnt_state'], leaves, cfg_d)
  losses = O.losses_cartesian(ep, f, {'cmd': torch.tensor(labels['cmd'])}, leaves, cfg_d)
  losses['loss'].backward()
  grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)).numpy() for k, v in leaves.items()}
  return cfg_d, P, ep, losses, grads, seed, goal


@pytest.mark.parametrize('name', list(C.CASES))
def test_oracle_matches_reference_graph(golden, name):
  cfg_d, P, ep, losses, grads, seed, goal = _run_oracle(name)
  g = lambda k: golden[name + '::' + k]
  # variables: names, shapes and creation order are the reference's
  shapes = O.param_shapes(cfg_d, goal)
  assert list(g('var_names')) == list(shapes.keys())
  assert [str(tuple(s)) for s in shapes.values()] == list(g('var_shapes'))
  # endpoints
  for k in ('pred_cmd_ee', 'logits_cmd_grp', 'pred_aux_ee', 'pred_aux_obj', 'fc1'):
    assert rel_max(ep[k].detach().numpy(), g('ep_' + k)) <= TOL, k
  for k in ('flat_state', 'flat_representation'):
    if name + '::ep_' + k in golden:       # the LSTM input of the last frame, whatever the reference calls it
      mine = ep[k] if k in ep else ep['flat_representation' if k == 'flat_state' else 'flat_state']
      assert rel_max(mine.detach().numpy(), g('ep_' + k)) <= TOL, k
  if 'conv8_obs' in ep:
    assert rel_max(ep['conv8_obs'].detach().numpy(), g('ep_conv8')) <= TOL
  for k in ('dynbuff', 'dyndiff'):
    if name + '::ep_' + k + '_sub' in golden:
      img = ep[k].detach().numpy()
      assert np.abs(img[:, ::8, ::8, :] - g('ep_' + k + '_sub')).max() <= TOL, k
      st = np.array([img.mean(), img.min(), img.max(), (img ** 2).sum()])
      assert np.allclose(st, g('ep_' + k + '_stats'), rtol=TOL, atol=TOL), k
  # losses: cmd_ee, cmd_grp, pos_ee, pos_obj, reg, total
  got = [float(losses[k].detach()) for k in ('loss_cmd_ee', 'loss_cmd_grp', 'loss_pos_ee', 'loss_pos_obj', 'loss_reg', 'loss')]
  assert np.allclose(got, g('losses'), rtol=TOL, atol=1e-12)
  cmd = C.make_inputs(C.CASES[name][1], cfg_d['window_size'], seed, C=cfg_d['img_channels'])[1]['cmd']
  assert list(O.gripper_classes(torch.tensor(cmd)[:, 3]).numpy()) == list(g('classes'))
  # gradients
  for n in shapes:
    gr = grads[n].ravel()
    ref_norm, ref_sum = g('grad_stats/' + n)
    assert abs(np.linalg.norm(gr) - ref_norm) <= TOL * ref_norm + 1e-15, n
    scale = np.abs(gr).max() + 1e-30
    assert abs(gr.sum() - ref_sum) <= 1e-4 * scale * np.sqrt(gr.size) * TOL * 1e4 + 1e-15, n
    idx = C.sample_indices(n, gr.size, seed)
    assert np.abs(gr[idx] - g('grad_samples/' + n)).max() <= TOL * scale, n


def test_fixture_covers_every_variable_of_geecof(golden):
  names = list(golden['geecof_n2::var_names'])
  assert len(names) == 60
  total = sum(int(np.prod(eval(s))) for s in golden['geecof_n2::var_shapes'])
  assert total == 7552796                      # SURVEY 8c pin: GEECO-F parameter count


def test_variable_tables_match_the_reference_graph():
  """geeco_b200.graph.variable_table reproduces names, shapes and creation order of the variables that the
  reference's graph.py created in every golden configuration (all switch values, both graph functions)."""
  from geeco_b200 import create_e2evmc_config
  from geeco_b200.graph import check_checkpoint_variables, lstm_input_width, variable_table
  from tests.golden import cases as C
  with np.load(os.path.join(os.path.dirname(__file__), 'golden', 'geeco_graph_golden.npz'), allow_pickle=True) as z:
    for name, (over, _, _, goal) in C.CASES.items():
      cfg = create_e2evmc_config(dict(over))
      table = variable_table(cfg, 'target' if goal else 'none')
      assert [n for n, _ in table] == [str(n) for n in z[name + '::var_names']], name
      assert [tuple(s) for _, s in table] == [ast.literal_eval(str(s)) for s in z[name + '::var_shapes']], name
      width = dict(table)[('GoalVMC' if goal else 'VMC') + '/LSTMDecoder/lstm_cell/kernel'][0] - cfg.dim_h_lstm
      assert width == lstm_input_width(cfg, 'target' if goal else 'none')
  geecof = create_e2evmc_config(dict(proc_obs='dynimg', proc_tgt='dyndiff'))
  assert sum(int(np.prod(s)) for _, s in variable_table(geecof)) == 7552796          # SURVEY 8c pin (10)
  stored = {n: s for n, s in variable_table(create_e2evmc_config(dict(proc_obs='sequence', proc_tgt='constant')))}
  with pytest.raises(ValueError, match='DynBuffEncoder'):                           # another model's checkpoint
    check_checkpoint_variables(stored, geecof)
  check_checkpoint_variables(dict(variable_table(geecof)), geecof)
  wide = create_e2evmc_config(dict(proc_obs='dynimg', proc_tgt='dyndiff', dim_h_lstm=64))
  with pytest.raises(ValueError, match='shape'):
    check_checkpoint_variables(dict(variable_table(geecof)), wide)
  vel = create_e2evmc_config(dict(proc_obs='dynimg', proc_tgt='dyndiff', control_mode='velocity'))
  names = [n for n, _ in variable_table(vel)]
  assert names[-10:-4] == ['GoalVMC/LSTMDecoder/pred_cmd_vel/kernel', 'GoalVMC/LSTMDecoder/pred_cmd_vel/bias',
                           'GoalVMC/LSTMDecoder/pred_cmd_ee/kernel', 'GoalVMC/LSTMDecoder/pred_cmd_ee/bias',
                           'GoalVMC/LSTMDecoder/pred_cmd_grp/kernel', 'GoalVMC/LSTMDecoder/pred_cmd_grp/bias']
  for bad in (dict(control_mode='torque'), dict(proc_tgt='foo'), dict(proc_obs='bar')):
    with pytest.raises(ValueError):
      variable_table(create_e2evmc_config(bad))


@pytest.mark.parametrize('name', list(C.VELOCITY_CASES))
def test_oracle_velocity_mode_matches_reference_graph(name):
  """--control_mode velocity (heads graph.py:240-249, mse_loss :430-450, targets estimator.py:229-237) against the
  reference graph executed by tests/golden/make_golden.py --velocity."""
  from geeco_b200 import create_e2evmc_config
  from geeco_b200.graph import variable_table
  over, N, seed, goal = C.VELOCITY_CASES[name]
  with np.load(os.path.join(os.path.dirname(GOLDEN), 'geeco_graph_golden_velocity.npz')) as z:
    g = lambda k: z[name + '::' + k]
    cfg_d = O.make_config(batch_size=N, **over)
    feats, labels = C.make_inputs(N, cfg_d['window_size'], seed, C=cfg_d['img_channels'])
    P = O.init_params(cfg_d, seed=seed, goal=goal, dtype=torch.float64, bias_scale=C.BIAS_SCALE)
    leaves = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    losses, ep = O.forward_losses(leaves, feats, labels, cfg_d)
    losses['loss'].backward()
    shapes = O.param_shapes(cfg_d, goal)
    assert list(g('var_names')) == list(shapes.keys())
    assert [str(tuple(s)) for s in shapes.values()] == list(g('var_shapes'))
    table = variable_table(create_e2evmc_config(dict(over)))                 # the product's table agrees too
    assert [n for n, _ in table] == list(g('var_names')) and [str(tuple(s)) for _, s in table] == list(g('var_shapes'))
    for k in ('pred_cmd_vel', 'pred_cmd_ee', 'pred_cmd_grp', 'pred_aux_ee', 'pred_aux_obj', 'fc1'):
      assert rel_max(ep[k].detach().numpy(), g('ep_' + k)) <= TOL, k
    got = [float(losses['loss_' + k].detach()) for k in O.VELOCITY_LOSS_KEYS] + [float(losses['loss_reg'].detach()),
                                                                                float(losses['loss'].detach())]
    assert np.allclose(got, g('losses'), rtol=TOL, atol=1e-12)
    for n in shapes:
      gr = (leaves[n].grad if leaves[n].grad is not None else torch.zeros_like(leaves[n])).numpy().ravel()
      ref_norm, _ = g('grad_stats/' + n)
      assert abs(np.linalg.norm(gr) - ref_norm) <= TOL * ref_norm + 1e-15, n
      idx = C.sample_indices(n, gr.size, seed)
      assert np.abs(gr[idx] - g('grad_samples/' + n)).max() <= TOL * (np.abs(gr).max() + 1e-30), n
