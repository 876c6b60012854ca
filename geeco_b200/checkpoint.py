"""TF V2 checkpoint bundles (`model.ckpt-<step>.index` + `.data-00000-of-00001`) through libgeeco_io.so.

Reference: tf.estimator writes these files from `tf.train.Saver` (scripts/train_e2evmc.py:160-161 reads their
common prefix through `tf.train.latest_checkpoint`, :178-179 copies every file that starts with it) and the
predictor restores every variable except `lstm_memory` from them (src/models/e2evmc/predictor.py:87-93).
Reading them lets checkpoints trained with the reference (the published ICRA models) drive this engine;
writing them keeps a run directory readable by TensorFlow tooling.  The `.meta` graph file is not written: it
describes a TF graph, which this build does not have.

File formats (restated from their specifications, see geeco_b200/csrc_io/geeco_io.cpp): the index is a
leveldb-format table ("" -> BundleHeaderProto, variable name -> BundleEntryProto), tensors are raw
little-endian bytes with a masked CRC-32C per entry.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _io

_DTYPES = {_io.DT_FLOAT: np.dtype('<f4'), _io.DT_INT32: np.dtype('<i4'), _io.DT_INT64: np.dtype('<i8')}
_CODES = {v: k for k, v in _DTYPES.items()}


class BundleReader(object):
  """Counterpart of `tf.train.load_checkpoint(prefix)`: `.names()`, `.shape(name)`, `.get_tensor(name)`."""

  def __init__(self, prefix):
    self._lib = _io.load()
    self._h = C.c_void_p()
    self.prefix = prefix
    _io.check(self._lib.geeco_bundle_open(prefix.encode(), C.byref(self._h)))

  def names(self):
    out = []
    for i in range(int(self._lib.geeco_bundle_num_tensors(self._h))):
      p, n = C.c_void_p(), C.c_uint64()
      _io.check(self._lib.geeco_bundle_name(self._h, i, C.byref(p), C.byref(n)))
      out.append(C.string_at(p.value, n.value).decode('utf-8', 'replace'))
    return out

  def _info(self, name):
    dt, nd, nb = C.c_int(), C.c_int(), C.c_int64()
    dims = (C.c_int64 * 8)()
    _io.check(self._lib.geeco_bundle_info(self._h, name.encode(), C.byref(dt), C.byref(nd), dims, C.byref(nb)))
    return dt.value, tuple(dims[i] for i in range(nd.value)), nb.value

  def has_tensor(self, name):
    try:
      self._info(name)
      return True
    except KeyError:
      return False

  def shape(self, name):
    return self._info(name)[1]

  def get_tensor(self, name, verify_crc=True):
    dt, shape, nbytes = self._info(name)
    if dt not in _DTYPES:
      raise ValueError("tensor '%s' has TensorFlow dtype %d: only float32 / int32 / int64 are supported" % (name, dt))
    out = np.empty(shape, dtype=_DTYPES[dt])
    if out.nbytes != nbytes:
      raise _io.DataLossError("tensor '%s': %d bytes stored, shape %s needs %d" % (name, nbytes, shape, out.nbytes))
    _io.check(self._lib.geeco_bundle_read(self._h, name.encode(), _io._ptr(out), out.nbytes, int(verify_crc)))
    return out

  def close(self):
    if self._h:
      self._lib.geeco_bundle_close(self._h)
      self._h = C.c_void_p()

  __del__ = close

  def __enter__(self):
    return self

  def __exit__(self, *exc):
    self.close()


def read_bundle(prefix):
  """{name: ndarray} of every tensor in the bundle."""
  with BundleReader(prefix) as r:
    return {n: r.get_tensor(n) for n in r.names()}


def write_bundle(prefix, tensors: dict):
  """Writes `<prefix>.index` and `<prefix>.data-00000-of-00001` holding `tensors` (float32 / int32 / int64)."""
  lib = _io.load()
  h = C.c_void_p()
  _io.check(lib.geeco_bundle_writer_create(prefix.encode(), C.byref(h)))
  try:
    for name, value in tensors.items():
      a = np.asarray(value)
      if not a.flags.c_contiguous:                            # (ascontiguousarray would promote scalars to 1-D)
        a = np.ascontiguousarray(a)
      if a.dtype == np.float64:
        a = a.astype(np.float32)
      if a.dtype.newbyteorder('<') not in _CODES and a.dtype not in _CODES:
        raise ValueError("tensor '%s': dtype %s cannot be stored (float32 / int32 / int64 only)" % (name, a.dtype))
      a = a.astype(a.dtype.newbyteorder('<'), copy=False)
      dims = (C.c_int64 * 8)(*a.shape)
      _io.check(lib.geeco_bundle_writer_add(h, name.encode(), _CODES[a.dtype], a.ndim, dims,
                                            _io._ptr(a) if a.size else None, a.nbytes))
  except BaseException:
    lib.geeco_bundle_writer_abort(h)
    raise
  _io.check(lib.geeco_bundle_writer_finish(h))

