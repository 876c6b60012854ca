"""Builds geeco_b200/libgeeco_b200.so (CUDA, sm_100a only, nvcc) and geeco_b200/libgeeco_io.so (host data
formats, g++ + zlib) in-tree."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libgeeco_b200.so')
SOURCES = ['geeco_api.cu', 'rankpool.cu', 'conv_fp32.cu', 'tail.cu', 'step_bf16.cu', 'conv_tc.cu', 'conv12_fused.cu', 'conv21_bwd_fused.cu', 'conv2_wgrad_fused.cu', 'lstm_seq.cu', 'lstm_persistent.cu']
NVCC_FLAGS = (os.environ.get('GEECO_NVCC_EXTRA', '').split()) + ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _nvcc():
  for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
    if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
      return cand
  raise RuntimeError('nvcc not found')


def needs_build():
  if not os.path.exists(LIB):
    return True
  t = os.path.getmtime(LIB)
  deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, '..', 'include', 'geeco_b200.h')]
  return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_library(force=False, verbose=False):
  """Compiles every .cu under csrc/ to an object and links the shared library."""
  if not force and not needs_build():
    return LIB
  nvcc = _nvcc()
  objdir = os.path.join(HERE, 'build')
  os.makedirs(objdir, exist_ok=True)
  objs, procs = [], []
  for src in SOURCES:
    path = os.path.join(CSRC, src)
    if not os.path.exists(path):
      continue
    obj = os.path.join(objdir, src.replace('.cu', '.o'))
    objs.append(obj)
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', path, '-o', obj]
    procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
  for src, p in procs:
    out, _ = p.communicate()
    if verbose or p.returncode:
      sys.stderr.write(out)
    if p.returncode:
      raise RuntimeError('nvcc failed on %s' % src)
  cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-lcudart']
  r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
  if r.returncode:
    sys.stderr.write(r.stdout)
    raise RuntimeError('link failed')
  return LIB


IO_SRC = os.path.join(HERE, 'csrc_io', 'geeco_io.cpp')
IO_LIB = os.path.join(HERE, 'libgeeco_io.so')


def build_io_library(force=False):
  """Compiles the host-side format library (include/geeco_io.h): no CUDA, links zlib."""
  deps = [IO_SRC, os.path.join(HERE, '..', 'include', 'geeco_io.h')]
  if not force and os.path.exists(IO_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(IO_LIB) for d in deps):
    return IO_LIB
  cmd = [os.environ.get('CXX', 'g++'), '-O3', '-std=c++17', '-fPIC', '-shared', '-Wall', '-Wextra', '-o', IO_LIB,
         IO_SRC, '-lz', '-pthread']
  r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
  if r.returncode:
    sys.stderr.write(r.stdout)
    raise RuntimeError('g++ failed on %s' % IO_SRC)
  return IO_LIB


if __name__ == '__main__':
  print(build_library(force='--force' in sys.argv, verbose='-v' in sys.argv))
  print(build_io_library(force='--force' in sys.argv))
