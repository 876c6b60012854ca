"""ctypes binding of libgeeco_io.so (the host-side C-ABI declared in include/geeco_io.h).

Host code only: TFRecord / SequenceExample decoding, sliding windows, TF V2 checkpoint bundles.  As with the
CUDA library there is no Python fallback: a missing library or a failing call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libgeeco_io.so')

OK, ERR_ARG, ERR_FILE, ERR_FORMAT, ERR_CRC, ERR_MISSING, ERR_SIZE = 0, -1, -2, -3, -4, -5, -6
COMPRESSION = {'auto': -1, 'none': 0, '': 0, None: 0, 'zlib': 1, 'ZLIB': 1, 'gzip': 2, 'GZIP': 2}
KIND_NONE, KIND_BYTES, KIND_FLOAT, KIND_INT64 = 0, 1, 2, 3
CONTEXT, SEQUENCE = 0, 1
DT_FLOAT, DT_INT32, DT_INT64 = 1, 3, 9

_P, _I, _I64, _U64, _U32 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_uint32
_PP = C.POINTER(C.c_void_p)
_CP = C.c_char_p

# every symbol include/geeco_io.h declares: name -> (restype, argtypes)
SYMBOLS = {
    'geeco_io_last_error': (_CP, []),
    'geeco_io_version': (_I, []),
    'geeco_io_crc32c': (_U32, [_P, C.c_size_t]),
    'geeco_io_crc32c_extend': (_U32, [_U32, _P, C.c_size_t]),
    'geeco_io_crc32c_mask': (_U32, [_U32]),
    'geeco_io_crc32c_unmask': (_U32, [_U32]),
    'geeco_tfrecord_open': (_I, [_CP, _I, _I, _PP]),
    'geeco_tfrecord_count': (_I64, [_P]),
    'geeco_tfrecord_get': (_I, [_P, _I64, _PP, C.POINTER(_U64)]),
    'geeco_tfrecord_close': (None, [_P]),
    'geeco_tfrecord_write': (_I, [_CP, _I, _I64, _PP, C.POINTER(_U64)]),
    'geeco_seqex_parse': (_I, [_P, _U64, _PP]),
    'geeco_seqex_free': (None, [_P]),
    'geeco_seqex_num_keys': (_I64, [_P, _I]),
    'geeco_seqex_key': (_I, [_P, _I, _I64, _PP, C.POINTER(_U64)]),
    'geeco_seqex_info': (_I, [_P, _I, _CP, C.POINTER(_I), C.POINTER(_I64), C.POINTER(_I64)]),
    'geeco_seqex_read_f32': (_I, [_P, _I, _CP, _P, _I64]),
    'geeco_seqex_read_i64': (_I, [_P, _I, _CP, _P, _I64]),
    'geeco_seqex_read_u8': (_I, [_P, _I, _CP, _P, _I64, C.POINTER(_I64)]),
    'geeco_seqex_bytes': (_I, [_P, _I, _CP, _I64, _I64, _PP, C.POINTER(_U64)]),
    'geeco_io_window_gather': (_I, [_P, _I64, _I64, _I64, _I64, _I64, _P]),
    'geeco_bundle_open': (_I, [_CP, _PP]),
    'geeco_bundle_close': (None, [_P]),
    'geeco_bundle_num_tensors': (_I64, [_P]),
    'geeco_bundle_name': (_I, [_P, _I64, _PP, C.POINTER(_U64)]),
    'geeco_bundle_info': (_I, [_P, _CP, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I64), C.POINTER(_I64)]),
    'geeco_bundle_read': (_I, [_P, _CP, _P, _I64, _I]),
    'geeco_bundle_writer_create': (_I, [_CP, _PP]),
    'geeco_bundle_writer_add': (_I, [_P, _CP, _I, _I, C.POINTER(_I64), _P, _I64]),
    'geeco_bundle_writer_finish': (_I, [_P]),
    'geeco_bundle_writer_abort': (None, [_P]),
}

_lib = None


def load():
  """Loads the shared library; raises if it is absent or a declared symbol is missing."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise RuntimeError("%s not found: build it with `python -m geeco_b200.build` (g++, zlib)." % LIB_PATH)
  lib = C.CDLL(LIB_PATH)
  for name, (res, args) in SYMBOLS.items():
    fn = getattr(lib, name)
    fn.restype = res
    fn.argtypes = args
  _lib = lib
  return lib


class DataLossError(IOError):
  """Checksum mismatch or malformed file (what TensorFlow reports as DataLossError)."""


def check(rc: int):
  if rc == OK:
    return
  msg = load().geeco_io_last_error().decode('utf-8', 'replace')
  if rc == ERR_MISSING:
    raise KeyError(msg)
  if rc in (ERR_ARG, ERR_SIZE):
    raise ValueError(msg)
  if rc == ERR_FILE:
    raise FileNotFoundError(msg) if msg.startswith('cannot open') else IOError(msg)
  raise DataLossError(msg)


def _ptr(a: np.ndarray):
  return a.ctypes.data_as(C.c_void_p)


def crc32c(data, crc=0) -> int:
  buf = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data)
  return int(load().geeco_io_crc32c_extend(crc, _ptr(buf) if buf.size else None, buf.nbytes))


def masked_crc32c(data) -> int:
  return int(load().geeco_io_crc32c_mask(crc32c(data)))
