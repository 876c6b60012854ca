"""GPU tests of the switch values beyond GEECO-F at the drop-in boundary: the reference CLI's default flags, the
unconditional model_fn / predictor, sequence-mode predictor, velocity control through the Estimator, the carried LSTM
state of the batched predictor (ring buffer, per-environment resets), and the pinned-staging hazard."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import geeco_oracle as O
from tests.util import rel_max

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model_dir(tmp_path, P, cfg_d):
  from geeco_b200 import create_e2evmc_config, save_model_config
  md = str(tmp_path)
  save_model_config(create_e2evmc_config(cfg_d)._asdict(), md, 'e2evmc_config')
  arrays = {k: v.numpy() for k, v in P.items()}
  arrays['global_step'] = np.array(5, dtype=np.int64)
  np.savez(os.path.join(md, 'model.ckpt-5.npz'), **arrays)
  with open(os.path.join(md, 'checkpoint'), 'w') as fp:
    fp.write('model_checkpoint_path: "model.ckpt-5"\n')
  return md


def test_train_cli_with_the_reference_default_flags(cuda_device, tmp_path):
  """scripts/train_e2evmc.py:34-75 defaults: --goal_condition none --proc_obs sequence --proc_tgt constant (the
  unconditional e2e_vmc graph), then --goal_condition target with the default sequence / constant processing."""
  sys.path.insert(0, os.path.join(ROOT, 'scripts'))
  import importlib
  m = importlib.import_module('train_e2evmc')
  for sub, extra, scope in (('vmc', [], 'VMC'), ('goal', ['--goal_condition', 'target'], 'GoalVMC')):
    md = os.path.join(str(tmp_path), sub)
    argv = ['--dataset_dir', 'synthetic:1', '--model_dir', md, '--train_epochs', '2', '--lr', '1e-3', '--log_steps', '1'] + extra
    res = m.main(m.ARGPARSER.parse_args(argv), argv)
    assert len(res) == 2 and res[1]['global_step'] == 6           # 96 windows / 32 = 3 steps per epoch
    assert np.isfinite(res[1]['loss']) and res[1]['loss'] < res[0]['loss']
    with np.load(os.path.join(md, 'model.ckpt-6.npz')) as d:
      assert scope + '/ConvEncoder/conv1/kernel' in d and scope + '/LSTMDecoder/lstm_memory' in d
      assert not any('DynBuff' in k or 'DynDiff' in k for k in d.files)
    # per-loss TensorBoard scalars under the reference's tags (estimator.py:262-265, :305-313)
    from geeco_b200.summaries import read_scalars
    tags = read_scalars(md)
    for t in ('loss', 'CartesianCmdLoss/mean_squared_error/value_0', 'GripperCmdLoss/softmax_cross_entropy_loss/value_0',
              'EEPoseAuxLoss/mean_squared_error/value_0', 'ObjPoseAuxLoss/mean_squared_error/value_0',
              'RegularizationLoss/l2_reg_loss_0'):
      assert t in tags and len(tags[t]) >= 2, (t, sorted(tags))
    assert 'loss' in read_scalars(os.path.join(md, 'eval'))


@pytest.mark.parametrize('variant', ['vmc', 'seq_residual', 'seq_dyndiff_velocity'])
def test_predictors_of_the_other_graphs_match_oracle(cuda_device, tmp_path, variant):
  """E2EVMCPredictor (predictor.py:212-379) and GoalE2EVMCPredictor on sequence graphs, closed loop over the FIFO."""
  from geeco_b200.predictor import E2EVMCPredictor, GoalE2EVMCPredictor
  over = {'vmc': dict(proc_obs='sequence', proc_tgt='constant'), 'seq_residual': dict(proc_obs='sequence', proc_tgt='residual'),
          'seq_dyndiff_velocity': dict(proc_obs='sequence', proc_tgt='dyndiff', control_mode='velocity')}[variant]
  goal = variant != 'vmc'
  cfg_d = O.make_config(batch_size=32, window_size=3, **over)
  P = O.init_params(cfg_d, seed=11, goal=goal, dtype=torch.float32, bias_scale=0.05)
  md = _model_dir(tmp_path, P, cfg_d)
  pred = (GoalE2EVMCPredictor if goal else E2EVMCPredictor)(md)
  rng = np.random.default_rng(3)
  tgt = rng.integers(0, 256, size=(256, 256, 3)) / 255.0
  pred.reset()
  if goal:
    pred.set_goal(tgt)
  cfg1 = dict(cfg_d, batch_size=1)
  fifo = []
  for t in range(4):
    fr = rng.integers(0, 256, size=(256, 256, 3)) / 255.0
    jn = rng.uniform(-np.pi, np.pi, size=7).astype(np.float32)
    out = pred.predict(fr, jn)
    fifo = ([(fr, jn)] * 3 if t == 0 else fifo + [(fr, jn)])[-3:]
    rgb = torch.tensor(np.stack([f for f, _ in fifo])[None].astype(np.float32))
    jnt = torch.tensor(np.stack([j for _, j in fifo])[None])
    if goal:
      _, ep = O.goal_e2evmc(rgb, jnt, torch.tensor(tgt[None].astype(np.float32)), P, cfg1)
    else:
      _, ep = O.e2e_vmc(rgb, jnt, P, cfg1)
    if cfg_d['control_mode'] == 'velocity':
      assert set(out) == {'cmd_vel', 'cmd_ee', 'cmd_grp', 'pos_ee', 'pos_obj', 'dyndiff'}
      pairs = [('cmd_vel', 'pred_cmd_vel'), ('cmd_ee', 'pred_cmd_ee'), ('cmd_grp', 'pred_cmd_grp'), ('pos_ee', 'pred_aux_ee')]
      assert out['cmd_vel'].shape == (7,) and out['cmd_grp'].shape == (2,)
      assert np.abs(out['dyndiff'] - ep['dyndiff'][0].numpy()).max() <= 1e-5
    else:
      assert set(out) == {'cmd_ee', 'cmd_grp', 'pos_ee', 'pos_obj'}
      pairs = [('cmd_ee', 'pred_cmd_ee'), ('pos_ee', 'pred_aux_ee'), ('pos_obj', 'pred_aux_obj')]
      assert out['cmd_grp'][0] == float(ep['logits_cmd_grp'][0].argmax()) - 1.0 and out['cmd_grp'].dtype == np.float32
    for k, ek in pairs:
      assert rel_max(out[k], ep[ek][0].numpy()) <= 1e-4, (t, k)


def test_velocity_control_through_the_estimator(cuda_device, tmp_path):
  """estimator.py:190-197 predictions, :229-237 targets + mse_loss, :255-258 eval metrics, for the GEECO-F wiring."""
  from geeco_b200 import create_e2evmc_config
  from geeco_b200.data import synthetic_batch
  from geeco_b200.estimator import Estimator, RunConfig, goal_e2evmc_model_fn
  cfg_d = O.make_config(batch_size=2, control_mode='velocity', lr=1e-3)
  cfg = create_e2evmc_config(cfg_d)
  batches = [synthetic_batch(2, seed=s) for s in range(2)]
  inp = lambda: iter(batches)
  est = Estimator(goal_e2evmc_model_fn, str(tmp_path), RunConfig(save_checkpoints_steps=0),
                  {'e2evmc_config': cfg, 'log_steps': 1, 'seed': 3}, precision='fp32', batch_size=2)
  ev0 = est.evaluate(inp)
  assert set(ev0) == {'loss', 'cmd_vel', 'cmd_ee', 'cmd_grp', 'pos_ee', 'pos_obj', 'global_step'}
  P = {k: torch.tensor(v) for k, v in est.engine.get_params().items()}
  se, cnt, loss = {k: 0.0 for k in O.VELOCITY_LOSS_KEYS}, {k: 0 for k in O.VELOCITY_LOSS_KEYS}, 0.0
  for f, l in batches:
    losses, ep = O.forward_losses(P, f, l, cfg_d)
    loss += float(losses['loss'])
    for k in O.VELOCITY_LOSS_KEYS:
      width = {'cmd_vel': 7, 'cmd_grp': 2}.get(k, 3)
      se[k] += float(losses['loss_' + k]) * 2 * width; cnt[k] += 2 * width
  assert abs(ev0['loss'] - loss / 2) <= 1e-4 * loss / 2
  for k in O.VELOCITY_LOSS_KEYS:
    assert abs(ev0[k] - se[k] / cnt[k]) <= 1e-4 * se[k] / cnt[k], k
  est.train(inp)
  assert est.evaluate(inp)['loss'] < ev0['loss']
  preds = list(est.predict(lambda: iter([batches[0][0]])))
  assert len(preds) == 2 and set(preds[0]) == {'cmd_vel', 'cmd_ee', 'cmd_grp', 'pos_ee', 'pos_obj'} and preds[0]['cmd_vel'].shape == (7,)
  from geeco_b200.summaries import read_scalars
  assert 'MSELoss/mean_squared_error_4/value_0' in read_scalars(str(tmp_path))


@pytest.mark.parametrize('frame_dtype', ['float32', 'uint8'])
def test_batched_predictor_ring_and_carried_state(cuda_device, frame_dtype):
  """BASELINE config 4 semantics: the K-frame history as a rotated device ring must feed the network exactly what the
  FIFO of predictor.py:140-146 would, and with carry_state = 1 the LSTM state [c | m] must flow from step to step per
  environment and restart from zero for environments that were reset.

  The oracle is `O.goal_e2evmc(..., init_state=...)` (the restatement with the dead assign of graph.py:226 made
  live).  A freshly reset environment sees K copies of one frame: that buffer's dynamic image is ill-conditioned
  (SURVEY 7.3 item 5), so at such a step the environment's outputs are compared with a non-carrying CUDA engine
  (which starts every step from the zero state: equality proves the reset) and the oracle adopts the CUDA state."""
  from geeco_b200 import create_e2evmc_config
  from geeco_b200.predictor import BatchedGoalPredictor
  E, K, steps = 3, 4, 7
  cfg_d = O.make_config(batch_size=E)
  cfg = create_e2evmc_config(cfg_d)
  P = O.init_params(cfg_d, seed=21, dtype=torch.float32, bias_scale=0.05)
  carry = BatchedGoalPredictor(cfg, E, precision='fp32', carry_state=True, frame_dtype=frame_dtype)
  plain = BatchedGoalPredictor(cfg, E, precision='fp32', carry_state=False, frame_dtype=frame_dtype)
  carry.engine.set_params(P); plain.engine.set_params(P)
  rng = np.random.default_rng(5)
  as_in = (lambda u: u.astype(np.uint8)) if frame_dtype == 'uint8' else (lambda u: (u / 255.0).astype(np.float32))
  as_f32 = lambda u: (u.astype(np.float32) / np.float32(255.0))
  goals_u = rng.integers(0, 256, size=(E, 256, 256, 3))
  carry.set_goal(as_in(goals_u)); plain.set_goal(as_in(goals_u))
  resets = {3: [1], 5: [0, 2]}                          # step -> environments reset BEFORE that step
  fifo = [[] for _ in range(E)]
  state = torch.zeros(E, 2 * cfg_d['dim_h_lstm'])
  for t in range(steps):
    if t in resets:
      mask = np.zeros(E, dtype=bool); mask[resets[t]] = True
      carry.reset(mask); plain.reset(torch.as_tensor(mask))
      for e in resets[t]:
        fifo[e] = []
    fr_u = rng.integers(0, 256, size=(E, 256, 256, 3))
    jn = rng.uniform(-1, 1, size=(E, 7)).astype(np.float32)
    out_c = {k: v.clone() for k, v in carry.predict_batch(as_in(fr_u), jn).items()}
    st_c = carry.engine.out_state.cpu().clone()
    out_p = {k: v.clone() for k, v in plain.predict_batch(as_in(fr_u), jn).items()}
    fresh = [len(fifo[e]) == 0 for e in range(E)]
    for e in range(E):
      fifo[e] = ([(fr_u[e], jn[e])] * K if fresh[e] else fifo[e] + [(fr_u[e], jn[e])])[-K:]
    rgb = torch.tensor(np.stack([np.stack([as_f32(f) for f, _ in fifo[e]]) for e in range(E)]))
    jnt = torch.tensor(np.stack([np.stack([j for _, j in fifo[e]]) for e in range(E)]))
    init = state.clone()
    for e in range(E):
      if fresh[e]:
        init[e] = 0
    _, ep = O.goal_e2evmc(rgb, jnt, torch.tensor(as_f32(goals_u)), P, cfg_d, init_state=init)
    state = ep['lstm_state'].detach().clone()
    for e in range(E):
      if fresh[e]:
        # zero initial state: identical to the engine that never carries (bit for bit: the extra h-rows multiply zeros)
        for k in ('cmd_ee', 'pos_ee', 'pos_obj', 'cmd_grp'):
          assert torch.equal(out_c[k][e], out_p[k][e]), (t, e, k)
        state[e] = st_c[e]
      else:
        for k, ek in (('cmd_ee', 'pred_cmd_ee'), ('pos_ee', 'pred_aux_ee'), ('pos_obj', 'pred_aux_obj')):
          assert rel_max(out_c[k][e].cpu().numpy(), ep[ek][e].numpy()) <= 1e-4, (t, e, k)
        assert rel_max(st_c[e].numpy(), ep['lstm_state'][e].numpy()) <= 1e-4, (t, e)
        assert float(out_c['cmd_grp'][e]) == float(ep['logits_cmd_grp'][e].argmax()) - 1.0
        if t >= 1:
          # the carried state matters: the non-carrying engine gives a different answer for the same frames
          assert not torch.equal(out_c['cmd_ee'][e], out_p['cmd_ee'][e]), (t, e)


def test_staged_batches_survive_a_host_that_runs_ahead(cuda_device):
  """`Engine.stage` writes a pinned buffer on the host and reads it with an asynchronous H2D copy.
  With the device held up (a long sleep kernel in front of the event the upload waits for) the host must not
  overwrite the pinned buffer of slot 0 before its upload has run: the first staged batch has to arrive intact."""
  from geeco_b200 import create_e2evmc_config
  from geeco_b200.data import synthetic_batch
  from geeco_b200.engine import Engine
  eng = Engine(create_e2evmc_config(O.make_config(batch_size=2)), batch_size=2, precision='fp32', training=True)
  fa, la = synthetic_batch(2, seed=1)
  fb, lb = synthetic_batch(2, seed=2)
  eng.stage(fa, la, 0)                                  # creates the staging state
  torch.cuda.synchronize()
  torch.cuda._sleep(int(2e9))                           # ~1 s of device time on the compute stream
  eng.release_staged(0)                                 # slot 0 is "free" only after the sleep
  out_a, lab_a, ev_a = eng.stage(fa, la, 0)             # its H2D copies queue up behind the sleep
  with torch.cuda.stream(eng._copy_stream):
    snap_cmd = lab_a['cmd'].clone()                     # stream-ordered after A's upload, before B's
    snap_jnt = out_a['jnt_state'].clone()
  eng.stage(fb, lb, 0)                                  # the host is a full batch ahead of the device now
  torch.cuda.synchronize()
  assert np.array_equal(snap_cmd.cpu().numpy(), la['cmd']) and np.array_equal(snap_jnt.cpu().numpy(), fa['jnt_state'])
  assert np.array_equal(eng._stage_bufs[0][('cmd', torch.float32, False)].cpu().numpy(), lb['cmd'])
  eng.close()


@pytest.mark.parametrize('variant', ['seq_constant', 'seq_residual', 'seq_dyndiff', 'vmc'])
def test_sequence_graphs_train_step_k4(cuda_device, variant):
  """K = 4 frames through the encoder, four LSTM steps (the persistent recurrent kernels of lstm_persistent.cu with W_h in
  shared memory), back-propagation through time; batch of 3 (not a multiple of any tile).  fp32: losses, head outputs and
  every gradient entry <= 1e-4 against the float64 oracle (ReLU decisions given, tests/util.py); bf16: forward <= 2e-2."""
  from geeco_b200 import create_e2evmc_config
  from geeco_b200.data import synthetic_batch
  from geeco_b200.engine import Engine
  from tests.util import engine_relu_masks
  over = {'seq_constant': dict(proc_obs='sequence', proc_tgt='constant'), 'seq_residual': dict(proc_obs='sequence', proc_tgt='residual'),
          'seq_dyndiff': dict(proc_obs='sequence', proc_tgt='dyndiff'), 'vmc': dict(proc_obs='sequence', proc_tgt='constant')}[variant]
  goal = variant != 'vmc'
  N = 3
  cfg_d = O.make_config(batch_size=N, window_size=4, **over)
  P = O.init_params(cfg_d, seed=31, goal=goal, dtype=torch.float32, bias_scale=0.05)
  feats, labels = synthetic_batch(N, seed=32)
  P64 = {k: v.double() for k, v in P.items()}
  ref_l, _, ep = O.train_step({k: v.clone() for k, v in P64.items()}, O.adam_init(P64), feats, labels, cfg_d, goal=goal)
  for precision, tol in (('fp32', 1e-4), ('bf16', 2e-2)):
    eng = Engine(create_e2evmc_config(cfg_d), batch_size=N, precision=precision, training=True,
                 goal_condition='target' if goal else 'none')
    eng.set_params(P)
    out = eng.forward(feats, labels)
    torch.cuda.synchronize()
    for k in ('pred_cmd_ee', 'logits_cmd_grp', 'pred_aux_ee', 'pred_aux_obj', 'fc1'):
      assert rel_max(out[k].cpu().numpy(), ep[k].detach().numpy()) <= tol, (precision, k)
    assert rel_max(out['lstm_state'].cpu().numpy(), ep['lstm_state'].detach().numpy()) <= tol
    got = eng.losses_dict(out['losses'])
    for k in ('loss_cmd_ee', 'loss_cmd_grp', 'loss_pos_ee', 'loss_pos_obj', 'loss'):
      assert abs(got[k] - ref_l[k]) <= tol * abs(ref_l[k]) + 1e-7, (precision, k, got[k], ref_l[k])
    eng.train_step(feats, labels)
    torch.cuda.synchronize()
    grads = eng.get_grads()
    if precision == 'fp32':
      _, mask_g, _ = O.train_step({k: v.clone() for k, v in P64.items()}, O.adam_init(P64), feats, labels, cfg_d,
                                  relu_masks=engine_relu_masks(eng), goal=goal)
      worst = max((rel_max(grads[k], g.numpy()), k) for k, g in mask_g.items())
      print("%s K=4 fp32 worst gradient entry %.3e (%s)" % (variant, worst[0], worst[1]))
      assert worst[0] <= 1e-4, worst
      # the recurrent rows of the LSTM kernel receive gradient now (steps 1..3 see a non-zero m_{t-1})
      gk = grads[('GoalVMC' if goal else 'VMC') + '/LSTMDecoder/lstm_cell/kernel']
      assert np.abs(gk[-cfg_d['dim_h_lstm']:]).max() > 0
    else:
      _, emu_g, _ = O.train_step({k: v.clone() for k, v in P64.items()}, O.adam_init(P64), feats, labels, cfg_d,
                                 emulate_bf16=True, relu_masks=engine_relu_masks(eng), goal=goal)
      from tests.util import rel_l2
      worst = max((rel_l2(grads[k], g.numpy()), k) for k, g in emu_g.items() if float(g.abs().max()) > 0)
      print("%s K=4 bf16 worst gradient rel-L2 (given ReLU masks, bf16 storage emulated) %.3e (%s)" % (variant, worst[0], worst[1]))
      assert worst[0] <= 2e-2, worst
    eng.close()


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_frame_pool_layout_trains_bit_identically(cuda_device, precision, tmp_path):
  """geeco_batch.frame_index / target_index (include/geeco_b200.h): a batch given as the distinct frames its windows
  touch plus an index reads the same pixels as the dense [N,K,H,W,C] layout -- same network input, same step, bit for
  bit -- while uploading ~K times fewer frame bytes.  Checked through Engine.train_step, through
  Estimator.train(input_fn) over host batches (the e2e entry of bench.py), and for a sequence graph."""
  from geeco_b200 import create_e2evmc_config
  from geeco_b200.data import expand_pool_layout, synthetic_pool_batch
  from geeco_b200.engine import Engine
  from geeco_b200.estimator import Estimator, RunConfig, goal_e2evmc_model_fn
  N = 6
  cfg_d = O.make_config(batch_size=N, lr=1e-3)
  cfg = create_e2evmc_config(cfg_d)
  P = O.init_params(cfg_d, seed=41, dtype=torch.float32, bias_scale=0.05)
  pooled, labels = synthetic_pool_batch(N, pieces=2, seed=42)
  dense = expand_pool_layout(pooled)
  assert pooled['rgb'].shape[0] == N + 2 * 3 and pooled['rgb'].dtype == np.uint8 and dense['rgb'].shape[:2] == (N, 4)
  res = {}
  for tag, feats in (('dense', dense), ('pool', pooled)):
    eng = Engine(cfg, batch_size=N, precision=precision, training=True)
    eng.set_params(P)
    out = eng.forward(feats, labels, want_dyn=True)
    torch.cuda.synchronize()
    x0 = eng.debug_buffer('x0')
    x0 = (x0.view(torch.int16) if x0.dtype == torch.bfloat16 else x0.view(torch.int32)).cpu().numpy().copy()
    fwd = {k: out[k].cpu().numpy().copy() for k in ('pred_cmd_ee', 'dynbuff', 'dyndiff', 'losses')}
    eng.train_step(feats, labels)
    torch.cuda.synchronize()
    res[tag] = (x0, fwd, eng.theta.cpu().numpy().copy())
    assert eng.h2d_bytes(True, frames_u8=True, features=dict(feats, **labels)) == sum(
        np.asarray(feats[k]).nbytes if k in ('rgb', 'target_rgb') else np.asarray(feats[k]).size * 4
        for k in eng._feature_keys(True, feats)) + labels['cmd'].size * 4
    eng.close()
  assert np.array_equal(res['pool'][0], res['dense'][0])
  for k, v in res['dense'][1].items():
    assert np.array_equal(res['pool'][1][k], v), k
  assert np.array_equal(res['pool'][2], res['dense'][2])
  # Estimator.train over host batches in both layouts: same parameters after two steps
  thetas = []
  for tag, feats in (('dense', dense), ('pool', pooled)):
    est = Estimator(goal_e2evmc_model_fn, str(tmp_path / tag), RunConfig(save_checkpoints_steps=0),
                    {'e2evmc_config': cfg, 'log_steps': 1, 'save_final_checkpoint': False, 'seed': 5},
                    precision=precision, batch_size=N)
    est.train(lambda: iter([(feats, labels), (feats, labels)]))
    torch.cuda.synchronize()
    thetas.append(est.engine.theta.cpu().numpy().copy())
  assert np.array_equal(thetas[0], thetas[1])
  # a sequence graph (every frame through the encoder) reads the pool the same way
  cfg_s = create_e2evmc_config(O.make_config(batch_size=N, proc_obs='sequence', proc_tgt='dyndiff'))
  outs = []
  for feats in (dense, pooled):
    eng = Engine(cfg_s, batch_size=N, precision=precision, training=False)
    eng.init_params(seed=6)
    outs.append(eng.forward(feats, None)['pred_cmd_ee'].cpu().numpy().copy())
    eng.close()
  assert np.array_equal(outs[0], outs[1])
  # malformed pools are refused
  eng = Engine(cfg, batch_size=N, precision=precision, training=False)
  bad = dict(pooled); bad['rgb'] = pooled['rgb'][:, :128]
  with pytest.raises(ValueError, match='pool'):
    eng.forward(bad, None)
  bad = dict(pooled); bad['rgb_index'] = pooled['rgb_index'][:, :3]
  with pytest.raises(ValueError, match='rgb_index'):
    eng.forward(bad, None)
  eng.close()
