"""Data-parallel plumbing (one process per GPU, torch.distributed / NCCL over NVLink).

The reference has no distributed path at all; the hot path shards by batch with ONE
exchange step: a summing all-reduce of the flat gradient arena, issued bucket by bucket (late layers
first) while the remaining backward kernels run, and divided by the world size inside the fused
Adam kernel.  Every loss is a batch mean and nothing else couples samples, so with equal per-rank
batches the result equals the global-batch gradient exactly.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .data import rank_slice  # noqa: F401  (re-exported: the stream partition of SURVEY 8e)


def world_info():
  if dist.is_available() and dist.is_initialized():
    return dist.get_rank(), dist.get_world_size()
  return 0, 1


def allreduce_bucket(flat_grad: torch.Tensor, bucket, async_op=True):
  """Summing all-reduce of arena floats [offset, offset+numel).  Returns the work handle (or None)."""
  off, cnt = bucket
  if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
    return None
  return dist.all_reduce(flat_grad[off:off + cnt], op=dist.ReduceOp.SUM, async_op=async_op)


def allreduce_sums(values, device=None):
  """Sums a short list of Python numbers over all ranks (float64); identity for a single process.  Used for the
  streaming evaluation metrics, whose accumulators (sums and counts) add across the ranks' shards."""
  rank, world = world_info()
  if world == 1:
    return [float(v) for v in values]
  if device is None:
    device = 'cuda' if dist.get_backend() == 'nccl' else 'cpu'
  t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
  dist.all_reduce(t, op=dist.ReduceOp.SUM)
  return [float(v) for v in t.cpu()]


def data_parallel_step(engine, features, labels):
  """forward -> [backward bucket b ; all-reduce bucket b (async)]* -> wait -> Adam with 1/world."""
  rank, world = world_info()
  if world == 1:
    return engine.train_step(features, labels)
  engine.step_forward(features, labels)
  works = []
  for b in range(len(engine.buckets)):
    engine.step_backward(b)
    works.append(allreduce_bucket(engine.grad, engine.buckets[b], async_op=True))
  nb = len(works)
  split = getattr(engine, 'step_update_buckets', None)
  if split is None or nb < 2:
    for w in works:
      if w is not None:
        w.wait()
    engine.step_update(1.0 / world)
    return engine.out_losses
  # the last bucket (conv2 + conv1) can only be reduced after the last backward kernel: update every other bucket while
  # its small all-reduce is in flight instead of leaving it exposed in front of the whole optimizer step
  for w in works[:-1]:
    if w is not None:
      w.wait()
  split(1.0 / world, 0, nb - 2)
  if works[-1] is not None:
    works[-1].wait()
  split(1.0 / world, nb - 1, nb - 1)
  return engine.out_losses


def broadcast_parameters(engine, src=0):
  """All replicas start from rank `src`'s parameters (and Adam slots)."""
  rank, world = world_info()
  if world == 1:
    return
  for t in (engine.theta, engine.adam_m, engine.adam_v):
    if t is not None:
      dist.broadcast(t, src=src)
  engine.params_changed()


def pin_to_gpu_numa_node(local_rank, local_world=1):
  """Binds this process to the CPU cores next to its GPU (the GPU's NUMA node, from sysfs), split evenly between the
  ranks that share the node, and returns the chosen core list (None when the topology cannot be read -- nothing is
  changed then).  One process per GPU uploads ~25 GB/s of pinned batches at full rate: with all ranks on one NUMA
  node the uploads of 8 GPUs share one memory controller and the far GPUs cross the socket link (round-1 end-to-end
  scaling: 0.74 at 8 GPUs).  Call before the first pinned allocation so that first-touch places it on the right node."""
  import os
  try:
    p = torch.cuda.get_device_properties(local_rank)
    dev = '%04x:%02x:%02x.0' % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    base = '/sys/bus/pci/devices/' + dev
    with open(base + '/numa_node') as fp:
      node = int(fp.read().strip())
    with open(base + '/local_cpulist') as fp:
      spec = fp.read().strip()
    cpus = set()
    for part in spec.split(','):
      if '-' in part:
        a, b = part.split('-')
        cpus.update(range(int(a), int(b) + 1))
      elif part:
        cpus.add(int(part))
    allowed = sorted(cpus & os.sched_getaffinity(0))
    if not allowed:
      return None
    # ranks whose GPUs sit on the same node share its cores
    peers = []
    for r in range(local_world):
      q = torch.cuda.get_device_properties(r)
      try:
        with open('/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node' % (q.pci_domain_id, q.pci_bus_id, q.pci_device_id)) as fp:
          if int(fp.read().strip()) == node:
            peers.append(r)
      except OSError:
        pass
    if local_rank in peers and len(peers) > 1 and len(allowed) >= len(peers):
      per = len(allowed) // len(peers)
      i = peers.index(local_rank)
      allowed = allowed[i * per:(i + 1) * per]
    os.sched_setaffinity(0, allowed)
    return allowed
  except Exception:      # noqa: BLE001  (sysfs layout, permissions, cpusets: never fatal)
    return None


def tune_for_data_parallel(world):
  """Environment defaults for one-process-per-GPU training, set before NCCL and the CUDA library initialise.  The
  backward kernels are persistent grids sized for every SM; NCCL's all-reduce CTAs (up to 32 by default) share the
  SMs with them for ~0.2 ms per step.  Measured on 8 x B200 (profiles/r02_scaling_notes.md): NCCL capped at 4 CTAs
  and the grids sized for 144 of the 148 SMs: 2.463 ms per step against 2.505 ms with the defaults (8 CTAs / 140 SMs,
  16 / 132 and 8 / 148 were all slower: the 30 MB all-reduce hides behind the backward even at 4 CTAs)."""
  import os
  if world > 1:
    os.environ.setdefault('NCCL_MAX_CTAS', '4')
    os.environ.setdefault('GEECO_NUM_SMS', '144')
