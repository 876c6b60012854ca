"""Functional ops of the e2evmc graph on CUDA tensors (thin wrappers over the C-ABI).

Counterparts of the variable-free / single-layer graph functions of the reference
(src/models/e2evmc/graph.py): `dynimg` (:30-55) and one `tf.layers.conv2d` of `conv_encoder`
(:76-115) with its gradients.  Used by the parity tests and by the rank-pooling sweep in bench.py.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _stream(t):
  return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _p(t):
  return C.c_void_p(t.data_ptr()) if t is not None else None


def _req(t, name):
  if not (torch.is_tensor(t) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
    raise ValueError("%s must be a contiguous float32 CUDA tensor" % name)
  return t


def dynimg(rgb_frames, alpha=None, cluster=0, out=None):
  """graph.py:30-55.  rgb_frames [N,K,H,W,C] float32 CUDA in [0,1] -> [N,H,W,C].
  cluster: 0 = automatic cluster size, >0 = forced cluster size, -1 = two-pass kernels."""
  lib = _lib.load()
  x = _req(rgb_frames, 'rgb_frames')
  if x.dim() != 5:
    raise ValueError("rgb_frames must be [N,K,H,W,C]")
  N, K, H, W, Cc = x.shape
  if out is None:
    out = torch.empty((N, H, W, Cc), dtype=torch.float32, device=x.device)
  a = None
  if alpha is not None:
    a = (C.c_float * K)(*[float(v) for v in np.asarray(alpha, dtype=np.float32)])
  scratch = torch.empty(2 * max(N, 1), dtype=torch.float32, device=x.device)
  _lib.check(lib.geeco_dynimg(_p(x), _p(out), N, K, H, W, Cc, a, int(cluster), _p(scratch), _stream(x)))
  return out


def conv2d_same(x, w, b=None, stride=1, relu=True):
  """tf.layers.conv2d(kernel_size=3, padding='SAME') on NHWC / HWIO tensors (fp32 kernels)."""
  lib = _lib.load()
  x, w = _req(x, 'x'), _req(w, 'w')
  N, H, W, Cin = x.shape
  Cout = w.shape[3]
  Ho, Wo = -(-H // stride), -(-W // stride)
  y = torch.empty((N, Ho, Wo, Cout), dtype=torch.float32, device=x.device)
  _lib.check(lib.geeco_conv2d_same(_p(x), _p(w), _p(b), _p(y), N, H, W, Cin, Cout, stride, 1 if relu else 0, _stream(x)))
  return y


def conv2d_same_bwd(x, w, dy_pre, stride=1, relu_mask_x=None, need_dx=True):
  """Gradients of conv2d_same given dL/d(pre-activation).  Returns (dw, db, dx)."""
  lib = _lib.load()
  x, w, dy_pre = _req(x, 'x'), _req(w, 'w'), _req(dy_pre, 'dy_pre')
  N, H, W, Cin = x.shape
  Cout = w.shape[3]
  dw = torch.empty_like(w)
  db = torch.empty(Cout, dtype=torch.float32, device=x.device)
  dx = torch.zeros_like(x) if need_dx else None
  n = int(lib.geeco_conv2d_bwd_scratch_floats(N, H, W, Cin, Cout, stride))
  scratch = torch.empty(n, dtype=torch.float32, device=x.device)
  _lib.check(lib.geeco_conv2d_same_bwd(_p(x), _p(w), _p(dy_pre), _p(relu_mask_x), _p(dw), _p(db), _p(dx), _p(scratch),
                                       n, N, H, W, Cin, Cout, stride, _stream(x)))
  return dw, db, dx


def _req_bf16(t, name):
  if not (torch.is_tensor(t) and t.is_cuda and t.dtype == torch.bfloat16 and t.is_contiguous()):
    raise ValueError("%s must be a contiguous bfloat16 CUDA tensor" % name)
  return t


def conv2d_same_bf16(x, w, b=None, stride=1, relu=True, want_f32=False):
  """Tensor-core (tcgen05) conv: x bf16 [N,H,W,Cin] (Cin % 8 == 0), w fp32 HWIO [3,3,Cw,Cout] with Cw <= Cin."""
  lib = _lib.load()
  x, w = _req_bf16(x, 'x'), _req(w, 'w')
  N, H, W, Cin = x.shape
  Cw, Cout = w.shape[2], w.shape[3]
  Ho, Wo = -(-H // stride), -(-W // stride)
  y = torch.empty((N, Ho, Wo, Cout), dtype=torch.bfloat16, device=x.device)
  y32 = torch.empty((N, Ho, Wo, Cout), dtype=torch.float32, device=x.device) if want_f32 else None
  n = int(lib.geeco_conv2d_bf16_scratch_bytes(N, H, W, Cin, Cout, stride))
  scratch = torch.empty(n, dtype=torch.uint8, device=x.device)
  _lib.check(lib.geeco_conv2d_same_bf16(_p(x), _p(w), _p(b), _p(y), _p(y32), _p(scratch), n, N, H, W, Cin, Cw, Cout,
                                        stride, 1 if relu else 0, _stream(x)))
  return (y, y32) if want_f32 else y


def relu_mask_bits(y):
  """1-bit ReLU mask of a bf16 activation [..., C] (C % 16 == 0): uint16 per (pixel, 16-channel chunk)."""
  lib = _lib.load()
  y = _req_bf16(y, 'y')
  Cc = y.shape[-1]
  pixels = y.numel() // Cc
  bits = torch.empty((pixels, Cc // 16), dtype=torch.int16, device=y.device)
  _lib.check(lib.geeco_relu_mask_bits(_p(y), _p(bits), pixels, Cc, _stream(y)))
  return bits


def conv2d_same_bwd_bf16(x, w, dy_pre, stride=1, relu_mask_x=None, need_dx=True, need_dw=True, relu_mask_bits=None):
  """Tensor-core gradients of conv2d_same_bf16: returns (dw fp32, db fp32, dx bf16)."""
  lib = _lib.load()
  x, w, dy_pre = _req_bf16(x, 'x'), _req(w, 'w'), _req_bf16(dy_pre, 'dy_pre')
  N, H, W, Cin = x.shape
  Cw, Cout = w.shape[2], w.shape[3]
  dw = torch.empty_like(w) if need_dw else None
  db = torch.empty(Cout, dtype=torch.float32, device=x.device) if need_dw else None
  dx = torch.empty_like(x) if need_dx else None   # every pixel is written by one parity class
  n = int(lib.geeco_conv2d_bf16_scratch_bytes(N, H, W, Cin, Cout, stride))
  scratch = torch.empty(n, dtype=torch.uint8, device=x.device)
  if relu_mask_bits is not None:
    _lib.check(lib.geeco_conv2d_same_bwd_bf16_bits(_p(x), _p(w), _p(dy_pre), _p(relu_mask_bits), _p(dw), _p(db), _p(dx),
                                                   _p(scratch), n, N, H, W, Cin, Cw, Cout, stride, _stream(x)))
    return dw, db, dx
  _lib.check(lib.geeco_conv2d_same_bwd_bf16(_p(x), _p(w), _p(dy_pre), _p(relu_mask_x), _p(dw), _p(db), _p(dx),
                                            _p(scratch), n, N, H, W, Cin, Cw, Cout, stride, _stream(x)))
  return dw, db, dx


def _lstm_scratch(lib, N, K, xdim, Hl, device):
  n = int(lib.geeco_lstm_seq_scratch_floats(N, K, xdim, Hl))
  if n < 0:
    _lib.check(_lib.GEECO_ERR_INVALID)
  buf = torch.empty(n + 64, dtype=torch.float32, device=device)
  off = (-buf.data_ptr() // 4) % 64                      # 256-byte aligned start
  return buf[off:off + n], n


def lstm_sequence(feat_list, kernel, bias):
  """The LSTM recurrence of lstm_decoder (graph.py:212-225) over a list of K feature tensors [N, xdim] (or one
  tensor [K, N, xdim]) from the zero state.  Returns (outputs[-1] [N,Hl], state [N,2*Hl] = [c | m], saved) where
  `saved` feeds lstm_sequence_bwd."""
  lib = _lib.load()
  x = feat_list if torch.is_tensor(feat_list) else torch.stack(list(feat_list), dim=0)
  x, kernel, bias = _req(x.contiguous(), 'feat_list'), _req(kernel, 'kernel'), _req(bias, 'bias')
  if x.dim() != 3:
    raise ValueError("feat_list must be K tensors [N, xdim]")
  K, N, xdim = x.shape
  Hl = kernel.shape[1] // 4
  if tuple(kernel.shape) != (xdim + Hl, 4 * Hl) or tuple(bias.shape) != (4 * Hl,):
    raise ValueError("kernel must be [xdim + Hl, 4*Hl] and bias [4*Hl]; got %s, %s for xdim=%d"
                     % (tuple(kernel.shape), tuple(bias.shape), xdim))
  gates = torch.empty((K, N, 4 * Hl), dtype=torch.float32, device=x.device)
  c = torch.empty((K, N, Hl), dtype=torch.float32, device=x.device)
  m = torch.empty((K, N, Hl), dtype=torch.float32, device=x.device)
  scratch, n = _lstm_scratch(lib, N, K, xdim, Hl, x.device)
  _lib.check(lib.geeco_lstm_seq_fwd(_p(x), _p(kernel), _p(bias), _p(gates), _p(c), _p(m), _p(scratch), n, N, K, xdim, Hl,
                                    _stream(x)))
  return m[K - 1], torch.cat([c[K - 1], m[K - 1]], dim=1), (x, kernel, gates, c, m)


def lstm_sequence_bwd(saved, dm_last, need_dx=True):
  """Back-propagation through time of lstm_sequence: dL/d(outputs[-1]) [N,Hl] -> (dkernel, dbias, dx [K,N,xdim])."""
  lib = _lib.load()
  x, kernel, gates, c, m = saved
  dm_last = _req(dm_last.contiguous(), 'dm_last')
  K, N, xdim = x.shape
  Hl = kernel.shape[1] // 4
  if tuple(dm_last.shape) != (N, Hl):
    raise ValueError("dm_last must be [N, Hl]")
  dkernel = torch.empty_like(kernel)
  dbias = torch.empty(4 * Hl, dtype=torch.float32, device=x.device)
  dx = torch.empty_like(x) if need_dx else None
  scratch, n = _lstm_scratch(lib, N, K, xdim, Hl, x.device)
  _lib.check(lib.geeco_lstm_seq_bwd(_p(x), _p(kernel), _p(gates), _p(c), _p(m), _p(dm_last), _p(dkernel), _p(dbias),
                                    _p(dx), _p(scratch), n, N, K, xdim, Hl, _stream(x)))
  return dkernel, dbias, dx
