"""TFRecord files and tf.train.SequenceExample episodes of GEECO gym (native decode through libgeeco_io.so).

Reference:
  * writer: `TfrSequenceEncoding.encode` + `TfrSequenceRecorder.finalize` (src/data/data_recorder.py:37-59, :134-156),
    value conversion `convert_to_feature` (src/data/utils/tfrecord.py:42-81): int / float / str scalars and lists,
    int32/int64 arrays -> Int64List, float32/float64/uint8 arrays -> FloatList (pixels are stored as floats 0..255);
  * reader: `tf.data.TFRecordDataset(compression_type='ZLIB')` + `tf.parse_single_sequence_example`
    (src/data/geeco_gym.py:298-301, :443-446).

`TFRecordFile` + `SequenceExample` are the native path (used by input_pipeline.decode_episode ; C++: inflate, framing + CRC-32C, protobuf index, bulk value readers);
`encode_sequence_example` / `write_tfrecord` serialise the same format so recorded data can be produced without
TensorFlow (and so the tests have episodes to read).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _io


# ------------------------------------------------------------------------------------------------
# protobuf encoding (writer side)
# ------------------------------------------------------------------------------------------------
def _varint(v: int) -> bytes:
  v &= (1 << 64) - 1
  out = bytearray()
  while v >= 0x80:
    out.append((v & 0x7f) | 0x80)
    v >>= 7
  out.append(v)
  return bytes(out)


def _field(num: int, payload: bytes) -> bytes:
  """Length-delimited field."""
  return _varint((num << 3) | 2) + _varint(len(payload)) + payload


def _feature_bytes(values) -> bytes:
  return _field(1, b''.join(_field(1, v) for v in values))


def _feature_floats(values) -> bytes:
  arr = np.ascontiguousarray(values, dtype='<f4').reshape(-1)
  return _field(2, _field(1, arr.tobytes()) if arr.size else b'')


def _feature_ints(values) -> bytes:
  packed = b''.join(_varint(int(v)) for v in np.asarray(values).reshape(-1))
  return _field(3, _field(1, packed) if packed else b'')


def convert_to_feature(value) -> bytes:
  """Serialized tf.train.Feature for `value` with the type dispatch of src/data/utils/tfrecord.py:42-81
  (same accepted types, same TypeError otherwise)."""
  t = type(value)
  if t == int:
    return _feature_ints([value])
  if t == float:
    return _feature_floats([value])
  if t == str:
    return _feature_bytes([value.encode('utf-8')])
  if t == list:
    te = type(value[0])
    if te == int:
      return _feature_ints(value)
    if te == float:
      return _feature_floats(value)
    if te == str:
      return _feature_bytes([v.encode('utf-8') for v in value])
    raise TypeError("Unsupported conversion of list type %s to tf.train.Feature!" % (te,))
  if t == np.ndarray:
    te = value.dtype
    if te == np.int32 or te == np.int64:
      return _feature_ints(value.flatten())
    if te == np.float32 or te == np.float64 or te == np.uint8:
      return _feature_floats(value.flatten())
    raise TypeError("Unsupported conversion of array type %s to tf.train.Feature!" % (te,))
  raise TypeError("Unsupported conversion of type %s to tf.train.Feature!" % (t,))


def encode_sequence_example(data: dict, context_keys, frame_keys) -> bytes:
  """Serialized tf.train.SequenceExample of one episode: `data[k]` for the context keys, `data['sequence']` a list
  of frame dicts for the frame keys (TfrSequenceEncoding.encode, data_recorder.py:37-59).  Map entries are written
  in sorted key order, as protobuf's deterministic serialisation does."""
  ctx = b''.join(_field(1, _field(1, k.encode('utf-8')) + _field(2, convert_to_feature(data[k])))
                 for k in sorted(context_keys))
  lists = b''
  for k in sorted(frame_keys):
    feature_list = b''.join(_field(1, convert_to_feature(frame[k])) for frame in data['sequence'])
    lists += _field(1, _field(1, k.encode('utf-8')) + _field(2, feature_list))
  return _field(1, ctx) + _field(2, lists)


def write_tfrecord(path: str, records, compression='auto'):
  """Writes serialized records as one TFRecord file ('none' | 'zlib' | 'gzip' | 'auto' = by suffix, the
  `.zlib` / `.gzip` suffix TfrSequenceRecorder.finalize appends)."""
  lib = _io.load()
  records = [bytes(r) for r in records]
  n = len(records)
  bufs = [np.frombuffer(r, dtype=np.uint8) for r in records]
  ptrs = (C.c_void_p * max(n, 1))(*[b.ctypes.data if b.size else None for b in bufs])
  lens = (C.c_uint64 * max(n, 1))(*[len(r) for r in records])
  _io.check(lib.geeco_tfrecord_write(path.encode(), _io.COMPRESSION[compression], n, ptrs, lens))


# ------------------------------------------------------------------------------------------------
# native reader
# ------------------------------------------------------------------------------------------------
class TFRecordFile(object):
  """All records of one TFRecord file (inflated and CRC-checked in C++)."""

  def __init__(self, path, compression='auto', verify_crc=True):
    self._lib = _io.load()
    self._h = C.c_void_p()
    _io.check(self._lib.geeco_tfrecord_open(path.encode(), _io.COMPRESSION[compression], int(verify_crc),
                                            C.byref(self._h)))

  def __len__(self):
    return int(self._lib.geeco_tfrecord_count(self._h))

  def record_view(self, i):
    """(address, length) of record i inside the handle's buffer (valid until close())."""
    p, n = C.c_void_p(), C.c_uint64()
    _io.check(self._lib.geeco_tfrecord_get(self._h, i, C.byref(p), C.byref(n)))
    return p.value or 0, int(n.value)

  def __getitem__(self, i) -> bytes:
    addr, n = self.record_view(i)
    return C.string_at(addr, n) if n else b''

  def close(self):
    if self._h:
      self._lib.geeco_tfrecord_close(self._h)
      self._h = C.c_void_p()

  __del__ = close

  def __enter__(self):
    return self

  def __exit__(self, *exc):
    self.close()


class SequenceExample(object):
  """Index over one serialized tf.train.SequenceExample; values are read in bulk into numpy arrays."""

  def __init__(self, data=None, address=None, length=None, keepalive=None):
    self._lib = _io.load()
    if data is not None:
      self._buf = np.frombuffer(bytes(data), dtype=np.uint8)
      address, length = (self._buf.ctypes.data if self._buf.size else 0), self._buf.size
    self._keep = keepalive
    self._h = C.c_void_p()
    _io.check(self._lib.geeco_seqex_parse(C.c_void_p(address), length, C.byref(self._h)))

  def keys(self, which=_io.SEQUENCE):
    out = []
    for i in range(int(self._lib.geeco_seqex_num_keys(self._h, which))):
      p, n = C.c_void_p(), C.c_uint64()
      _io.check(self._lib.geeco_seqex_key(self._h, which, i, C.byref(p), C.byref(n)))
      out.append(C.string_at(p.value, n.value).decode('utf-8', 'replace'))
    return out

  def info(self, name, which=_io.SEQUENCE):
    """(kind, frames, values per frame or -1 if ragged)."""
    k, f, p = C.c_int(), C.c_int64(), C.c_int64()
    _io.check(self._lib.geeco_seqex_info(self._h, which, name.encode(), C.byref(k), C.byref(f), C.byref(p)))
    return k.value, f.value, p.value

  def _shape(self, name, which, want_kind, per_frame):
    kind, frames, per = self.info(name, which)
    if kind != want_kind and not (kind == _io.KIND_NONE and per == 0):
      raise ValueError("feature '%s' has kind %d, expected %d" % (name, kind, want_kind))
    if per < 0 or (per_frame is not None and per != per_frame):
      # tf.FixedLenSequenceFeature / FixedLenFeature reject frames of another length
      raise ValueError("feature '%s': %s values per frame, expected %s" % (name, 'ragged' if per < 0 else per, per_frame))
    return frames, per

  def floats(self, name, which=_io.SEQUENCE, per_frame=None, out=None):
    frames, per = self._shape(name, which, _io.KIND_FLOAT, per_frame)
    if out is None:
      out = np.empty((frames, per), dtype=np.float32)
    _io.check(self._lib.geeco_seqex_read_f32(self._h, which, name.encode(), _io._ptr(out), frames * per))
    return out

  def ints(self, name, which=_io.SEQUENCE, per_frame=None):
    frames, per = self._shape(name, which, _io.KIND_INT64, per_frame)
    out = np.empty((frames, per), dtype=np.int64)
    _io.check(self._lib.geeco_seqex_read_i64(self._h, which, name.encode(), _io._ptr(out), frames * per))
    return out

  def pixel_bytes(self, name, which=_io.SEQUENCE, per_frame=None, out=None):
    """Float-encoded pixels as the recorded uint8 values; returns (array [frames, per], inexact count)."""
    frames, per = self._shape(name, which, _io.KIND_FLOAT, per_frame)
    if out is None:
      out = np.empty((frames, per), dtype=np.uint8)
    bad = C.c_int64()
    _io.check(self._lib.geeco_seqex_read_u8(self._h, which, name.encode(), _io._ptr(out), frames * per, C.byref(bad)))
    return out, int(bad.value)

  def strings(self, name, which=_io.CONTEXT, frame=0):
    _, _, per = self.info(name, which)
    out = []
    for j in range(max(per, 0)):
      p, n = C.c_void_p(), C.c_uint64()
      _io.check(self._lib.geeco_seqex_bytes(self._h, which, name.encode(), frame, j, C.byref(p), C.byref(n)))
      out.append(C.string_at(p.value, n.value) if n.value else b'')
    return out

  def close(self):
    if self._h:
      self._lib.geeco_seqex_free(self._h)
      self._h = C.c_void_p()

  __del__ = close


def window_gather(src: np.ndarray, window_size: int, first_window=0, num_windows=None, out=None):
  """_window_v3 (geeco_gym.py:614-631) for one sequence tensor: out[i] = src[first+i : first+i+K]."""
  src = np.ascontiguousarray(src)
  frames = src.shape[0]
  if num_windows is None:
    num_windows = frames - window_size + 1 - first_window
  frame_bytes = src.dtype.itemsize * int(np.prod(src.shape[1:], dtype=np.int64))
  if out is None:
    out = np.empty((num_windows, window_size) + src.shape[1:], dtype=src.dtype)
  if frame_bytes == 0 or num_windows == 0:
    return out
  _io.check(_io.load().geeco_io_window_gather(_io._ptr(src), frames, frame_bytes, window_size, first_window,
                                              num_windows, _io._ptr(out)))
  return out
