"""Shared helpers of the parity tests."""
import numpy as np


def rel_l2(a, b):
  a = np.asarray(a, dtype=np.float64).ravel()
  b = np.asarray(b, dtype=np.float64).ravel()
  return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def rel_max(a, b):
  a = np.asarray(a, dtype=np.float64).ravel()
  b = np.asarray(b, dtype=np.float64).ravel()
  return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def oracle_step_chunked(O, params, feats, labels, cfg_d, chunk=4, emulate_bf16=False, relu_masks=None, keep=(),
                        goal=True):
  """Forward + backward of the oracle over a batch that is evaluated `chunk` rows at a time (float64 at batch 64
  would otherwise hold ~15 GB of autograd state).  Exact, not an approximation: every loss term of the graph is a
  MEAN over the batch rows (tf.losses reductions) and the L2 term does not depend on the batch,
  so with equal chunks  loss = mean_c loss_c  and  grad = mean_c grad_c.
  Returns (losses dict of floats, grads dict of tensors, {key: concatenated endpoint} for `keep`)."""
  import torch
  N = int(np.shape(feats['rgb'])[0])
  assert N % chunk == 0
  nchunks = N // chunk
  cfg_c = dict(cfg_d, batch_size=chunk)
  tot_l, tot_g, kept = None, None, {k: [] for k in keep}
  for c in range(nchunks):
    sl = slice(c * chunk, (c + 1) * chunk)
    f = {k: v[sl] for k, v in feats.items()}
    l = {k: v[sl] for k, v in labels.items()}
    rm = _slice_masks(relu_masks, sl)
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    losses, ep = O.forward_losses(leaves, f, l, cfg_c, emulate_bf16=emulate_bf16, relu_masks=rm, goal=goal)
    losses['loss'].backward()
    g = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    lv = {k: float(v.detach()) for k, v in losses.items()}
    tot_l = lv if tot_l is None else {k: tot_l[k] + lv[k] for k in lv}
    tot_g = g if tot_g is None else {k: tot_g[k] + g[k] for k in g}
    for k in keep:
      kept[k].append(ep[k].detach())
  losses = {k: v / nchunks for k, v in tot_l.items()}
  grads = {k: v / nchunks for k, v in tot_g.items()}
  return losses, grads, {k: torch.cat(v, dim=0) for k, v in kept.items()}


def _slice_masks(rm, sl):
  """Batch rows `sl` of a relu_masks structure (lists of eight tensors, possibly nested per frame)."""
  if rm is None:
    return None
  def cut(v):
    return [cut(x) for x in v] if isinstance(v, (list, tuple)) else v[sl]
  return {k: cut(v) for k, v in rm.items()}


def engine_relu_masks(eng, N=None):
  """The ReLU decisions the CUDA path took in its last forward, in the structure oracle.goal_e2evmc / e2e_vmc take as
  `relu_masks`: torch bool tensors [N,H,W,C] on the CPU, y_l > 0 read back through geeco_debug_buffer (y8 from its fp32
  copy in bf16 mode).  Layout of a layer buffer: [group][image][H][W][C], images of a group step-major (t*N + n), the
  target frame of the constant / residual graphs last (geeco_b200/csrc/tail.cuh)."""
  cfg = eng.cfg
  N, K = eng.N, cfg.window_size
  seq = eng.goal_condition == 'none' or cfg.proc_obs == 'sequence'
  if not seq:
    G, M = 3, N
  elif eng.goal_condition == 'none':
    G, M = 1, K * N
  elif cfg.proc_tgt == 'dyndiff':
    G, M = 2, K * N
  else:
    G, M = 1, (K + 1) * N
  chans = (32, 48, 64, 128, 192, 256, 256)
  layers = []                                   # per layer: list over groups of bool [M,H,H,C]
  H = 256
  for li in range(8):
    H = H if li == 0 else H // 2
    name = 'y%d' % (li + 1)
    if li == 7 and eng.precision == 'bf16':
      name = 'y8_f32'
    y = eng.debug_buffer(name).float().cpu()
    if li < 7:
      layers.append([g > 0 for g in y.view(G, M, H, H, chans[li])])
    else:                                       # conv8 widths may differ per group: consecutive blocks
      widths = {3: (cfg.dim_s_obs, cfg.dim_s_dyn, cfg.dim_s_diff), 2: (cfg.dim_s_obs, cfg.dim_s_diff),
                1: ((256,) if eng.goal_condition == 'none' else (cfg.dim_s_obs,))}[G]
      off, per = 0, []
      for w in widths:
        n = M * H * H * w
        per.append(y[off:off + n].view(M, H, H, w) > 0)
        off += n
      layers.append(per)
  group = lambda gi, lo, hi: [layers[li][gi][lo:hi] for li in range(8)]
  if not seq:
    return {'obs': group(0, 0, N), 'dyn': group(1, 0, N), 'diff': group(2, 0, N)}
  out = {'frames': [group(0, t * N, (t + 1) * N) for t in range(K)]}
  if eng.goal_condition == 'target' and cfg.proc_tgt in ('constant', 'residual'):
    out['tgt'] = group(0, K * N, (K + 1) * N)
  if eng.goal_condition == 'target' and cfg.proc_tgt == 'dyndiff':
    out['diff'] = [group(1, t * N, (t + 1) * N) for t in range(K)]
  return out
