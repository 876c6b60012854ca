"""Second, independent CPU restatement (NumPy, explicit index math, hand-written backward).

TEST INFRASTRUCTURE, NOT PRODUCT -- see oracle/geeco_oracle.py for the rules.
Pinned like geeco_oracle.py (see its header): the reference has no fixtures; TF itself never ran here.

Purpose: cross-check oracle/geeco_oracle.py (torch ops + autograd) with code that
shares nothing with it: the SAME-padding index map is spelled out per output pixel
(no F.pad / F.conv2d), the flatten/concat order is an explicit index formula, and
the backward pass is derived by hand (the same derivation the CUDA kernels follow).
Only the GEECO-F wiring (graph.py:386-407) is restated here.
"""
from __future__ import annotations

import numpy as np

STRIDES = (1, 2, 2, 2, 2, 2, 2, 2)   # graph.py:78-113


# ---- rank pooling (graph.py:17-55) ----------------------------------------
def alpha_exact(K):
  H = np.concatenate([[0.0], np.cumsum(1.0 / np.arange(1, K + 1))])
  t = np.arange(1, K + 1)
  return 2.0 * (K - t + 1) - (K + 1) * (H[K] - H[t - 1])


def dynimg(frames, alpha):
  N, K = frames.shape[:2]
  out = np.empty(frames.shape[:1] + frames.shape[2:], dtype=frames.dtype)
  for n in range(N):
    d = np.zeros(frames.shape[2:], dtype=frames.dtype)
    for k in range(K):
      d = d + frames.dtype.type(alpha[k]) * frames[n, k]
    lo, hi = d.min(), d.max()
    out[n] = (d - lo) / (hi - lo + frames.dtype.type(1e-6))
  return out


# ---- conv 3x3, TF SAME ------------------------------------------------------
def same_geometry(size, stride):
  out = (size + stride - 1) // stride
  total = max((out - 1) * stride + 3 - size, 0)
  return out, total // 2


def im2col(x, stride):
  """x [N,H,W,C] -> cols [N*Ho*Wo, 9*C] with k = (ky*3+kx)*C + c ; out-of-image taps are 0."""
  N, H, W, C = x.shape
  Ho, pt = same_geometry(H, stride)
  Wo, pl = same_geometry(W, stride)
  cols = np.zeros((N, Ho, Wo, 9, C), dtype=x.dtype)
  for ky in range(3):
    for kx in range(3):
      for oy in range(Ho):
        iy = oy * stride + ky - pt
        if iy < 0 or iy >= H:
          continue
        for ox in range(Wo):
          ix = ox * stride + kx - pl
          if 0 <= ix < W:
            cols[:, oy, ox, ky * 3 + kx, :] = x[:, iy, ix, :]
  return cols.reshape(N * Ho * Wo, 9 * C), (N, Ho, Wo)


def col2im(dcols, xshape, stride):
  N, H, W, C = xshape
  Ho, pt = same_geometry(H, stride)
  Wo, pl = same_geometry(W, stride)
  d = dcols.reshape(N, Ho, Wo, 9, C)
  dx = np.zeros(xshape, dtype=dcols.dtype)
  for ky in range(3):
    for kx in range(3):
      for oy in range(Ho):
        iy = oy * stride + ky - pt
        if iy < 0 or iy >= H:
          continue
        for ox in range(Wo):
          ix = ox * stride + kx - pl
          if 0 <= ix < W:
            dx[:, iy, ix, :] += d[:, oy, ox, ky * 3 + kx, :]
  return dx


def conv_fwd(x, w_hwio, b, stride):
  cols, (N, Ho, Wo) = im2col(x, stride)
  y = cols @ w_hwio.reshape(-1, w_hwio.shape[3]) + b
  y = np.maximum(y, 0)
  return y.reshape(N, Ho, Wo, -1), cols


def conv_bwd(dy, y, cols, xshape, w_hwio, stride, need_dx=True):
  """dy wrt post-ReLU output y."""
  co = w_hwio.shape[3]
  g = (dy * (y > 0)).reshape(-1, co)
  dw = (cols.T @ g).reshape(w_hwio.shape)
  db = g.sum(axis=0)
  dx = col2im(g @ w_hwio.reshape(-1, co).T, xshape, stride) if need_dx else None
  return dx, dw, db


def encoder_fwd(x, P, scope):
  cache = []
  net = x
  for li in range(8):
    w, b = P['%s/conv%d/kernel' % (scope, li + 1)], P['%s/conv%d/bias' % (scope, li + 1)]
    y, cols = conv_fwd(net, w, b, STRIDES[li])
    cache.append((net.shape, cols, y))
    net = y
  return net, cache


def encoder_bwd(dy, cache, P, scope, G):
  for li in reversed(range(8)):
    xshape, cols, y = cache[li]
    w = P['%s/conv%d/kernel' % (scope, li + 1)]
    dy, dw, db = conv_bwd(dy, y, cols, xshape, w, STRIDES[li], need_dx=(li > 0))
    G['%s/conv%d/kernel' % (scope, li + 1)] = dw
    G['%s/conv%d/bias' % (scope, li + 1)] = db


# ---- state layout (graph.py:187-190) ---------------------------------------
def state_index(cell, block, c, dims, J):
  """Flat LSTM-input index of channel c of `block` in spatial cell (h*2+w)."""
  d_obs, d_dyn, d_diff = dims
  per = d_obs + d_dyn + J + d_diff
  off = {'obs': 0, 'dyn': d_obs, 'jnt': d_obs + d_dyn, 'tgt': d_obs + d_dyn + J}[block]
  return cell * per + off + c


def build_state(f_obs, f_dyn, jnt, f_tgt):
  N = f_obs.shape[0]
  dims = (f_obs.shape[3], f_dyn.shape[3], f_tgt.shape[3])
  J = jnt.shape[1]
  per = sum(dims) + J
  st = np.zeros((N, 4 * per), dtype=f_obs.dtype)
  for h in range(2):
    for w in range(2):
      cell = h * 2 + w
      for c in range(dims[0]):
        st[:, state_index(cell, 'obs', c, dims, J)] = f_obs[:, h, w, c]
      for c in range(dims[1]):
        st[:, state_index(cell, 'dyn', c, dims, J)] = f_dyn[:, h, w, c]
      for c in range(J):
        st[:, state_index(cell, 'jnt', c, dims, J)] = jnt[:, c]
      for c in range(dims[2]):
        st[:, state_index(cell, 'tgt', c, dims, J)] = f_tgt[:, h, w, c]
  return st


def sigmoid(x):
  return 1.0 / (1.0 + np.exp(-x))


# ---- whole GEECO-F step -----------------------------------------------------
def geeco_f_forward_backward(P, rgb, jnt_states, tgt, cmd, ee_state, obj_state, lambda_aux=1.0,
                             alpha=None, scope='GoalVMC'):
  """Returns (losses dict, endpoints dict, grads dict).  All float64 NumPy."""
  N, K = rgb.shape[:2]
  alpha = alpha_exact(K) if alpha is None else alpha
  cur, jnt = rgb[:, -1], jnt_states[:, -1]
  dynb = dynimg(rgb, alpha)
  dynd = dynimg(np.stack([cur, tgt], axis=1), alpha_exact(2))
  f_obs, c_obs = encoder_fwd(cur, P, scope + '/ConvEncoder')
  f_dyn, c_dyn = encoder_fwd(dynb, P, scope + '/DynBuffEncoder')
  f_tgt, c_tgt = encoder_fwd(dynd, P, scope + '/DynDiffEncoder')
  x = build_state(f_obs, f_dyn, jnt, f_tgt)
  d = scope + '/LSTMDecoder'
  Wl, bl = P[d + '/lstm_cell/kernel'], P[d + '/lstm_cell/bias']
  h = Wl.shape[1] // 4
  xin = np.concatenate([x, np.zeros((N, h))], axis=1)   # m_prev = 0 (graph.py:218-220,226 dead assign)
  gates = xin @ Wl + bl
  gi, gj, gf, go = gates[:, :h], gates[:, h:2 * h], gates[:, 2 * h:3 * h], gates[:, 3 * h:]
  c_prev = np.zeros((N, h))
  si, sf, so, tj = sigmoid(gi), sigmoid(gf + 1.0), sigmoid(go), np.tanh(gj)
  c = sf * c_prev + si * tj
  tc = np.tanh(c)
  m = so * tc
  W1, b1 = P[d + '/fc1/kernel'], P[d + '/fc1/bias']
  fc1 = np.maximum(m @ W1 + b1, 0)
  heads = ['pred_cmd_ee', 'logits_cmd_grp', 'pred_aux_ee', 'pred_aux_obj']
  out = {k: fc1 @ P['%s/%s/kernel' % (d, k)] + P['%s/%s/bias' % (d, k)] for k in heads}
  # losses
  t_ee, t_pe, t_po = cmd[:, :3], ee_state[:, -1, :3], obj_state[:, -1, :3]
  cls = np.rint(cmd[:, 3]).astype(np.int64) + 1
  L = {}
  L['loss_cmd_ee'] = np.mean((out['pred_cmd_ee'] - t_ee) ** 2)
  L['loss_pos_ee'] = np.mean((out['pred_aux_ee'] - t_pe) ** 2)
  L['loss_pos_obj'] = np.mean((out['pred_aux_obj'] - t_po) ** 2)
  z = out['logits_cmd_grp']
  zs = z - z.max(axis=1, keepdims=True)
  logp = zs - np.log(np.exp(zs).sum(axis=1, keepdims=True))
  L['loss_cmd_grp'] = -np.mean(logp[np.arange(N), cls])
  L['loss'] = L['loss_cmd_ee'] + L['loss_cmd_grp'] + lambda_aux * (L['loss_pos_ee'] + L['loss_pos_obj'])
  # ---- backward
  G = {}
  dout = {
      'pred_cmd_ee': 2.0 * (out['pred_cmd_ee'] - t_ee) / (N * 3),
      'pred_aux_ee': lambda_aux * 2.0 * (out['pred_aux_ee'] - t_pe) / (N * 3),
      'pred_aux_obj': lambda_aux * 2.0 * (out['pred_aux_obj'] - t_po) / (N * 3),
  }
  p = np.exp(logp)
  p[np.arange(N), cls] -= 1.0
  dout['logits_cmd_grp'] = p / N
  dfc1 = np.zeros_like(fc1)
  for k in heads:
    G['%s/%s/kernel' % (d, k)] = fc1.T @ dout[k]
    G['%s/%s/bias' % (d, k)] = dout[k].sum(axis=0)
    dfc1 += dout[k] @ P['%s/%s/kernel' % (d, k)].T
  dfc1 = dfc1 * (fc1 > 0)
  G[d + '/fc1/kernel'] = m.T @ dfc1
  G[d + '/fc1/bias'] = dfc1.sum(axis=0)
  dm = dfc1 @ W1.T
  dso = dm * tc
  dc = dm * so * (1 - tc ** 2)
  dsi, dtj, dsf = dc * tj, dc * si, dc * c_prev
  dgates = np.concatenate([dsi * si * (1 - si), dtj * (1 - tj ** 2), dsf * sf * (1 - sf), dso * so * (1 - so)], axis=1)
  G[d + '/lstm_cell/kernel'] = xin.T @ dgates
  G[d + '/lstm_cell/bias'] = dgates.sum(axis=0)
  dx = (dgates @ Wl.T)[:, :x.shape[1]]
  dims = (f_obs.shape[3], f_dyn.shape[3], f_tgt.shape[3])
  J = jnt.shape[1]
  per = sum(dims) + J
  dxc = dx.reshape(N, 2, 2, per)
  d_obs = dxc[..., :dims[0]]
  d_dyn = dxc[..., dims[0]:dims[0] + dims[1]]
  d_tgt = dxc[..., dims[0] + dims[1] + J:]
  encoder_bwd(d_obs, c_obs, P, scope + '/ConvEncoder', G)
  encoder_bwd(d_dyn, c_dyn, P, scope + '/DynBuffEncoder', G)
  encoder_bwd(d_tgt, c_tgt, P, scope + '/DynDiffEncoder', G)
  ep = dict(out, dynbuff=dynb, dyndiff=dynd, flat_representation=x, fc1=fc1, lstm_out=m,
            lstm_state=np.concatenate([c, m], axis=1), conv8_obs=f_obs, conv8_dyn=f_dyn, conv8_diff=f_tgt)
  return L, ep, G
