import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope='session', autouse=True)
def _host_io_library():
  """libgeeco_io.so is a git-ignored build product: (re)build it when the source is newer (g++, a few seconds).
  The CUDA library is built by __graft_entry__.build() / tests/test_host_cpu.py."""
  from geeco_b200.build import build_io_library
  build_io_library()


@pytest.fixture(scope='session')
def cuda_device():
  import torch
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  return torch.device('cuda:0')
