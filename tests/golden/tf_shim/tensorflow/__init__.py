"""Minimal eager stand-in for the TensorFlow-1.15 symbols that the reference's
`src/models/e2evmc/graph.py` touches, backed by torch (CPU, autograd on).

TEST INFRASTRUCTURE ONLY.  Purpose: TensorFlow is not installable in this image (no wheel, no
network), so the reference graph cannot run as shipped.  With this package first on sys.path,
`tests/golden/make_golden.py` imports the reference's graph.py UNMODIFIED from /root/reference and
executes it; the wiring (layer order, scopes, variable names and creation order, concat order,
strides, which heads exist, the loss functions) is then the reference's own code, and only the
per-op semantics below are ours, each written from TF-1.15's documented behaviour:

  tf.layers.conv2d      NHWC input, HWIO kernel, padding='SAME': out = ceil(in / stride),
                        pad_total = max((out-1)*stride + k - in, 0), pad_before = pad_total // 2
  tf.layers.dense       x @ kernel + bias
  tf.nn.rnn_cell.LSTMCell(state_is_tuple=False)
                        one kernel [x|h] -> 4h, gate order i, j, f, o, forget_bias = 1.0,
                        state = concat([c, m]); new_c = c*sigmoid(f+1) + sigmoid(i)*tanh(j);
                        new_m = tanh(new_c)*sigmoid(o); returns (new_m, concat([new_c, new_m]))
  tf.losses.mean_squared_error     mean over all elements (SUM_BY_NONZERO_WEIGHTS, weights=1)
  tf.losses.softmax_cross_entropy  mean over rows of -sum(onehot * log_softmax(logits))
  tf.contrib.layers.l2_regularizer(scale)(w) = scale * sum(w**2) / 2, None when scale == 0;
                        a variable_scope regularizer applies to every variable made in the scope

Variables are served from `VARIABLES` (name -> tensor), which the generator fills with the seeded
parameter set; every lookup is logged in `CREATED` (name, shape) in creation order.
"""
import contextlib

import torch
import torch.nn.functional as F

from . import contrib  # noqa: F401

float32, float64, int32, int64 = torch.float32, torch.float64, torch.int32, torch.int64
bool = torch.bool
AUTO_REUSE = 'auto_reuse'
COMPUTE_DTYPE = torch.float64      # activations / parameters; tf.float32 constants keep their dtype

VARIABLES = {}
CREATED = []
_SCOPE = []                        # [(name, regularizer)]
_COLLECTIONS = {}


class GraphKeys(object):
  LOSSES = 'losses'
  REGULARIZATION_LOSSES = 'regularization_losses'


def reset_state():
  VARIABLES.clear(); del CREATED[:]; del _SCOPE[:]; _COLLECTIONS.clear()


class _Shape(object):
  def __init__(self, t):
    self._dims = list(t.shape)

  def as_list(self):
    return list(self._dims)


torch.Tensor.get_shape = lambda self: _Shape(self)


@contextlib.contextmanager
def variable_scope(name_or_scope, regularizer=None, reuse=None):
  inherited = _SCOPE[-1][1] if _SCOPE else None
  _SCOPE.append((name_or_scope, regularizer if regularizer is not None else inherited))
  try:
    yield
  finally:
    _SCOPE.pop()


def _scope_name(leaf):
  return '/'.join([s for s, _ in _SCOPE] + [leaf])


def _get_variable(leaf, shape):
  name = _scope_name(leaf)
  if name not in VARIABLES:
    raise KeyError('graph asks for variable %s %s which the parameter set does not hold' % (name, list(shape)))
  v = VARIABLES[name]
  if list(v.shape) != list(shape):
    raise ValueError('variable %s: graph wants %s, parameter set has %s' % (name, list(shape), list(v.shape)))
  if name not in [n for n, _ in CREATED]:
    CREATED.append((name, tuple(shape)))
    reg = _SCOPE[-1][1] if _SCOPE else None
    if reg is not None:
      term = reg(v)
      if term is not None:
        _COLLECTIONS.setdefault(GraphKeys.REGULARIZATION_LOSSES, []).append(term)
  return v


def get_collection(key):
  return list(_COLLECTIONS.get(key, []))


def add_to_collection(key, value):
  _COLLECTIONS.setdefault(key, []).append(value)


# ---------------------------------------------------------------------------------- array ops
def _t(x, dtype=None):
  return x if isinstance(x, torch.Tensor) else torch.tensor(x, dtype=dtype)


def constant(v, dtype=None):
  return torch.tensor(v, dtype=dtype if dtype is not None else (float32 if isinstance(v, float) else None))


def zeros(shape, dtype=float32):
  return torch.zeros(list(shape), dtype=COMPUTE_DTYPE if dtype == float32 else dtype)


def Variable(initial_value, trainable=True, name=None):
  return initial_value


def assign(ref, value):
  return value


def identity(x, name=None):
  return x


def reshape(x, shape, name=None):
  return x.reshape(list(shape))


def concat(values, axis, name=None):
  return torch.cat(list(values), dim=axis)


def unstack(x, axis=0):
  return list(torch.unbind(x, dim=axis))


def expand_dims(x, axis):
  return x.unsqueeze(axis)


def tile(x, multiples):
  return x.repeat(*multiples)


def range(start, limit, dtype=None):          # noqa: A001  (tf.range)
  lim = float(limit) if isinstance(limit, torch.Tensor) else limit
  return torch.arange(start, lim, dtype=dtype)


def equal(a, b, name=None):
  if isinstance(a, torch.Tensor) or isinstance(b, torch.Tensor):
    return _t(a) == _t(b)
  return a == b


def cond(pred, true_fn, false_fn):
  return true_fn() if builtins_bool(pred) else false_fn()


def builtins_bool(p):
  return p.item() != 0 if isinstance(p, torch.Tensor) else (p is True or p == 1)


def map_fn(fn, elems):
  return torch.stack([_t(fn(e)) for e in elems])


def reduce_sum(x, axis=None, name=None):
  if isinstance(x, (list, tuple)):
    if len(x) == 0:
      return torch.zeros((), dtype=COMPUTE_DTYPE)
    x = torch.stack([_t(v) for v in x])
  return x.sum() if axis is None else x.sum(dim=axis)


def reduce_prod(x, axis=None, name=None):
  return x.prod() if axis is None else x.prod(dim=axis)


def reduce_min(x, axis=None):
  return x.amin(dim=tuple(axis) if isinstance(axis, (list, tuple)) else axis) if axis is not None else x.min()


def reduce_max(x, axis=None):
  return x.amax(dim=tuple(axis) if isinstance(axis, (list, tuple)) else axis) if axis is not None else x.max()


def one_hot(indices, depth):
  return F.one_hot(indices.long(), depth).to(COMPUTE_DTYPE)


def add_n(inputs, name=None):
  out = inputs[0]
  for v in inputs[1:]:
    out = out + v
  return out


def add(a, b, name=None):
  return a + b


# ---------------------------------------------------------------------------------- layers / nn
def _same_pad(size, k, s):
  out = -(-size // s)
  total = max((out - 1) * s + k - size, 0)
  return total // 2, total - total // 2


class _Layers(object):
  @staticmethod
  def conv2d(inputs, filters, kernel_size, strides=1, padding='valid', activation=None, name=None):
    assert padding.upper() == 'SAME' and name
    k, cin = int(kernel_size), inputs.shape[-1]
    w = _get_variable(name + '/kernel', (k, k, cin, filters))
    b = _get_variable(name + '/bias', (filters,))
    top, bottom = _same_pad(inputs.shape[1], k, strides)
    left, right = _same_pad(inputs.shape[2], k, strides)
    x = F.pad(inputs.permute(0, 3, 1, 2), (left, right, top, bottom))
    y = F.conv2d(x, w.permute(3, 2, 0, 1), b, stride=strides).permute(0, 2, 3, 1)
    return activation(y) if activation is not None else y

  @staticmethod
  def dense(inputs, units, activation=None, name=None):
    w = _get_variable(name + '/kernel', (inputs.shape[-1], units))
    b = _get_variable(name + '/bias', (units,))
    y = inputs @ w + b
    return activation(y) if activation is not None else y

  @staticmethod
  def flatten(inputs, name=None):
    return inputs.reshape(inputs.shape[0], -1)


layers = _Layers()


class _LSTMCell(object):
  def __init__(self, num_units, state_is_tuple=True, forget_bias=1.0):
    assert state_is_tuple is False
    self.h, self.forget_bias = num_units, forget_bias

  def zero_state(self, batch_size, dtype):
    return torch.zeros(batch_size, 2 * self.h, dtype=COMPUTE_DTYPE)

  def __call__(self, inputs, state):
    h = self.h
    c_prev, m_prev = state[:, :h], state[:, h:]
    w = _get_variable('lstm_cell/kernel', (inputs.shape[-1] + h, 4 * h))
    b = _get_variable('lstm_cell/bias', (4 * h,))
    z = torch.cat([inputs, m_prev], dim=1) @ w + b
    i, j, f, o = torch.split(z, h, dim=1)
    c = torch.sigmoid(f + self.forget_bias) * c_prev + torch.sigmoid(i) * torch.tanh(j)
    m = torch.sigmoid(o) * torch.tanh(c)
    return m, torch.cat([c, m], dim=1)


class _RnnCell(object):
  LSTMCell = _LSTMCell


class _NN(object):
  relu = staticmethod(torch.relu)
  rnn_cell = _RnnCell()


nn = _NN()


class _Losses(object):
  @staticmethod
  def mean_squared_error(labels, predictions):
    return ((predictions.to(COMPUTE_DTYPE) - labels.to(COMPUTE_DTYPE)) ** 2).mean()

  @staticmethod
  def softmax_cross_entropy(onehot_labels, logits):
    return -(onehot_labels * F.log_softmax(logits, dim=-1)).sum(dim=-1).mean()


losses = _Losses()


class _Math(object):
  rint = staticmethod(torch.round)          # both round half to even


math = _Math()                               # noqa: F811  (tf.math)


class _DTypes(object):
  @staticmethod
  def cast(x, dtype):
    return x.to(dtype)


dtypes = _DTypes()
