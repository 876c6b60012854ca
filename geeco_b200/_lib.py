"""ctypes binding of libgeeco_b200.so (the C-ABI declared in include/geeco_b200.h).

There is no CPU fallback: if the shared library is missing or a compute entry point fails the
caller gets an exception, never a silently different code path.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('GEECO_LIB_PATH') or os.path.join(_HERE, 'libgeeco_b200.so')

GEECO_OK, GEECO_ERR_INVALID, GEECO_ERR_CUDA, GEECO_ERR_WORKSPACE, GEECO_ERR_STATE = 0, 1, 2, 3, 4
GEECO_FP32, GEECO_BF16 = 0, 1


class GeecoConfig(C.Structure):
  _fields_ = [(n, C.c_int32) for n in (
      'img_height', 'img_width', 'img_channels', 'dim_jnt_state', 'window_size', 'dim_s_obs', 'dim_s_dyn',
      'dim_s_diff', 'dim_h_lstm', 'dim_h_fc', 'num_grp_states', 'batch_size', 'precision', 'carry_state',
      'training', 'goal_condition', 'proc_obs', 'proc_tgt', 'control_mode', 'dim_grp_command')] + [
          (n, C.c_float) for n in ('lr', 'lambda_aux', 'l2_regularizer', 'adam_beta1', 'adam_beta2', 'adam_eps')]


# values of the graph switches (include/geeco_b200.h)
GOAL_CONDITION = {'target': 0, 'none': 1}
PROC_OBS = {'dynimg': 0, 'sequence': 1}
PROC_TGT = {'dyndiff': 0, 'constant': 1, 'residual': 2}
CONTROL_MODE = {'cartesian': 0, 'velocity': 1}
NUM_LOSS_SLOTS = 12


class GeecoSizes(C.Structure):
  _fields_ = [('arena_floats', C.c_int64), ('workspace_bytes', C.c_int64), ('num_params', C.c_int32),
              ('num_buckets', C.c_int32)]


class GeecoParamDesc(C.Structure):
  _fields_ = [('name', C.c_char * 96), ('offset', C.c_int64), ('numel', C.c_int64), ('ndim', C.c_int32),
              ('reserved0', C.c_int32), ('shape', C.c_int64 * 4)]


class GeecoBatch(C.Structure):
  _fields_ = [(n, C.c_void_p) for n in ('rgb', 'target_rgb', 'jnt_state', 'ee_state', 'obj_state', 'cmd', 'vel_target',
                                        'ee_target', 'grp_target', 'reset_mask', 'frame_index', 'target_index')] + [
      ('frame_format', C.c_int32), ('ring_start', C.c_int32)]


FRAMES_F32, FRAMES_U8 = 0, 1


class GeecoOutputs(C.Structure):
  _fields_ = [(n, C.c_void_p) for n in ('heads', 'fc1', 'dynbuff', 'dyndiff', 'lstm_state', 'losses')]


# every symbol include/geeco_b200.h declares: name -> (restype, argtypes)
_P, _I32, _I64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SYMBOLS = {
    'geeco_last_error': (C.c_char_p, []),
    'geeco_version': (C.c_int, []),
    'geeco_query_sizes': (C.c_int, [C.POINTER(GeecoConfig), C.POINTER(GeecoSizes)]),
    'geeco_create': (C.c_int, [C.POINTER(GeecoConfig), C.POINTER(_P)]),
    'geeco_destroy': (C.c_int, [_P]),
    'geeco_bind': (C.c_int, [_P, _P, _P, _P, _P, _P, _I64]),
    'geeco_param_info': (C.c_int, [_P, _I32, C.POINTER(GeecoParamDesc)]),
    'geeco_head_columns': (C.c_int, [C.POINTER(GeecoConfig)]),
    'geeco_grad_bucket': (C.c_int, [_P, _I32, C.POINTER(_I64), C.POINTER(_I64)]),
    'geeco_params_changed': (C.c_int, [_P, _P]),
    'geeco_set_step': (C.c_int, [_P, _I64, _P]),
    'geeco_set_lstm_state': (C.c_int, [_P, _P, _P]),
    'geeco_dynimg': (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _I32, _P, _I32, _P, _P]),
    'geeco_alpha_table': (C.c_int, [_I32, C.POINTER(C.c_float)]),
    'geeco_conv2d_same': (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    'geeco_conv2d_bwd_scratch_floats': (_I64, [_I32, _I32, _I32, _I32, _I32, _I32]),
    'geeco_conv2d_same_bwd': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    'geeco_conv2d_bf16_scratch_bytes': (_I64, [_I32, _I32, _I32, _I32, _I32, _I32]),
    'geeco_conv2d_same_bf16': (C.c_int, [_P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    'geeco_conv2d_same_bwd_bf16': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    'geeco_relu_mask_bits': (C.c_int, [_P, _P, _I64, _I32, _P]),
    'geeco_conv2d_same_bwd_bf16_bits': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    'geeco_lstm_seq_scratch_floats': (_I64, [_I32, _I32, _I32, _I32]),
    'geeco_lstm_seq_fwd': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _P]),
    'geeco_lstm_seq_bwd': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _P]),
    'geeco_forward': (C.c_int, [_P, C.POINTER(GeecoBatch), C.POINTER(GeecoOutputs), _P]),
    'geeco_train_step': (C.c_int, [_P, C.POINTER(GeecoBatch), C.POINTER(GeecoOutputs), _F, _P]),
    'geeco_step_forward': (C.c_int, [_P, C.POINTER(GeecoBatch), C.POINTER(GeecoOutputs), _P]),
    'geeco_step_backward': (C.c_int, [_P, _I32, _P]),
    'geeco_step_update': (C.c_int, [_P, _F, _P]),
    'geeco_step_update_buckets': (C.c_int, [_P, _F, _I32, _I32, _P]),
    'geeco_ring_push': (C.c_int, [_P, _P, _P, _I32, _I32, _I64, _I32, _P]),
    'geeco_debug_buffer': (C.c_int, [_P, C.c_char_p, C.POINTER(_P), C.POINTER(_I64), C.POINTER(_I32)]),
    'geeco_profile_kernel': (C.c_int, [_P, C.c_char_p, _P]),
    'geeco_launch_count': (_I64, [_I32]),
}

_lib = None


def load():
  """Loads the shared library (building nothing); raises if it is absent or a symbol is missing."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise RuntimeError(
        "%s not found: build it with `python -m geeco_b200.build` (nvcc, sm_100a). "
        "geeco_b200 has no CPU fallback." % LIB_PATH)
  lib = C.CDLL(LIB_PATH)
  for name, (res, args) in SYMBOLS.items():
    fn = getattr(lib, name)     # AttributeError if the library does not export it
    fn.restype = res
    fn.argtypes = args
  _lib = lib
  return lib


def last_error() -> str:
  return load().geeco_last_error().decode('utf-8', 'replace')


def check(rc: int):
  """Maps status codes to the exception types the reference raises (SURVEY 8b 'Error conventions')."""
  if rc == GEECO_OK:
    return
  msg = last_error()
  if rc == GEECO_ERR_INVALID:
    raise ValueError(msg)
  raise RuntimeError("geeco_b200 error %d: %s" % (rc, msg))


def alpha_table(K: int):
  import numpy as np
  buf = (C.c_float * K)()
  check(load().geeco_alpha_table(K, buf))
  return np.array(list(buf), dtype=np.float32)
