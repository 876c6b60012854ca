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
