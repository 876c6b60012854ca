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


def oracle_step_chunked(O, params, feats, labels, cfg_d, chunk=4, emulate_bf16=False, relu_masks=None, keep=()):
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
    rm = None if relu_masks is None else {e: [m[sl] for m in ms] for e, ms in relu_masks.items()}
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    losses, ep = O.forward_losses(leaves, f, l, cfg_c, emulate_bf16=emulate_bf16, relu_masks=rm)
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


def engine_relu_masks(eng, N):
  """The ReLU decisions the CUDA path took in its last forward, per encoder and layer, as torch bool tensors
  [N,H,W,C] on the CPU: y_l > 0 read back through geeco_debug_buffer (y8 from its fp32 copy in bf16 mode)."""
  import torch
  names = ('obs', 'dyn', 'diff')
  chans = (32, 48, 64, 128, 192, 256, 256)
  out = {e: [] for e in names}
  H = 256
  for li in range(8):
    H = H if li == 0 else H // 2
    name = 'y%d' % (li + 1)
    if li == 7 and eng.precision == 'bf16':
      name = 'y8_f32'
    y = eng.debug_buffer(name).float().cpu()
    C = y.numel() // (3 * N * H * H) if li == 7 else chans[li]
    y = y.view(3, N, H, H, C)
    for ei, e in enumerate(names):
      out[e].append(y[ei] > 0)
  return out
