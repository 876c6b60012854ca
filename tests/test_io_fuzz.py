"""Mutation fuzzing of the native parsers (libgeeco_io.so): corrupted SequenceExamples and checkpoint index tables
must be rejected (or parsed) -- never crash the process or read out of bounds.  Block checksums are recomputed after
the mutation where needed, so the table / snappy / BundleEntryProto parsers themselves see the damaged bytes."""
import os

import numpy as np

from geeco_b200 import _io, data as gdata, tfrecord as tfr
from geeco_b200.checkpoint import BundleReader
from tests import test_checkpoint_bundle as T


def _mutate(b, rng, max_edits=5):
  b = bytearray(b)
  for _ in range(int(rng.integers(1, max_edits + 1))):
    if not b:
      break
    pos, mode = int(rng.integers(0, len(b))), int(rng.integers(0, 3))
    if mode == 0:
      b[pos] = int(rng.integers(0, 256))
    elif mode == 1:
      del b[pos:pos + int(rng.integers(1, 40))]
    else:
      b[pos:pos] = bytes(rng.integers(0, 256, size=int(rng.integers(1, 20)), dtype=np.uint8))
  return bytes(b)


def test_sequence_example_parser_survives_mutations():
  rng = np.random.default_rng(0)
  data = gdata.synthetic_episode(episode_length=3, height=4, width=4, seed=1)
  blob = tfr.encode_sequence_example(data, *gdata.encoding_keys_v4(data))
  parsed = rejected = 0
  for _ in range(1500):
    try:
      s = tfr.SequenceExample(_mutate(blob, rng))
    except IOError:
      rejected += 1
      continue
    parsed += 1
    for which in (_io.CONTEXT, _io.SEQUENCE):
      for k in s.keys(which):
        try:
          kind, _, per = s.info(k, which)
          if kind == _io.KIND_FLOAT and per >= 0:
            s.floats(k, which); s.pixel_bytes(k, which)
          elif kind == _io.KIND_INT64 and per >= 0:
            s.ints(k, which)
          elif kind == _io.KIND_BYTES:
            s.strings(k, which)
        except (ValueError, KeyError, IOError):
          pass
  assert parsed > 20 and rejected > 500


def test_checkpoint_index_parser_survives_mutations(tmp_path, monkeypatch):
  rng = np.random.default_rng(1)
  Header, Entry = T._bundle_protos()
  prefix = str(tmp_path / 'f')
  tensors = {('GoalVMC/v%03d' % i): rng.standard_normal(i % 5 + 1).astype(np.float32) for i in range(25)}
  blob, items = bytearray(), [(b'', Header(num_shards=1).SerializeToString())]
  for name in sorted(tensors):
    a = tensors[name]
    e = Entry(dtype=1, offset=len(blob), size=a.nbytes, crc32c=_io.load().geeco_io_crc32c_mask(_io.crc32c(a.tobytes())))
    e.shape.SetInParent()
    for d in a.shape:
      e.shape.dim.add().size = d
    blob += a.tobytes()
    items.append((name.encode(), e.SerializeToString()))
  open(prefix + '.data-00000-of-00001', 'wb').write(bytes(blob))
  snappy = T._snappy_literal_and_copies
  opened = rejected = 0
  for it in range(900):
    compress = bool(it & 1)
    if it % 4 < 2:      # damaged keys / entry protos inside structurally valid, correctly checksummed blocks
      its = [(k if rng.random() > 0.02 else _mutate(k, rng, 2), v if rng.random() > 0.03 else _mutate(v, rng, 3))
             for k, v in items]
      monkeypatch.setattr(T, '_snappy_literal_and_copies', snappy)
    else:               # damaged snappy streams (checksummed after the damage) / damaged raw blocks
      its = items
      monkeypatch.setattr(T, '_snappy_literal_and_copies', (lambda b: _mutate(snappy(b), rng, 3)) if compress else snappy)
    T._py_write_table(prefix + '.index', its, compress=compress, block_entries=int(rng.integers(1, 9)))
    if it % 4 >= 2 and not compress:
      f = bytearray(open(prefix + '.index', 'rb').read())
      f[int(rng.integers(0, len(f)))] ^= 0xff
      open(prefix + '.index', 'wb').write(bytes(f))
    try:
      with BundleReader(prefix) as r:
        for n in r.names():
          try:
            r.get_tensor(n)
          except (ValueError, KeyError, IOError, MemoryError):
            pass
      opened += 1
    except IOError:
      rejected += 1
  assert opened > 50 and rejected > 300
