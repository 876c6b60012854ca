"""CPU tests of the TF V2 checkpoint bundle reader / writer (libgeeco_io.so, geeco_b200/checkpoint.py).

Independent checker: a pure-Python restatement of the leveldb table format (table_format.md) and of
tensor_bundle.proto decoded with the real protobuf runtime -- it reads what the native writer wrote and writes
(including snappy-compressed blocks) what the native reader must read.
"""
import os
import struct

import numpy as np
import pytest

from geeco_b200 import _io
from geeco_b200.checkpoint import BundleReader, read_bundle, write_bundle

MAGIC = 0xdb4775248b80fb57


def _bundle_protos():
  from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
  F = descriptor_pb2.FieldDescriptorProto
  fd = descriptor_pb2.FileDescriptorProto(name='geeco_test_bundle.proto', package='tbtest', syntax='proto3')

  def msg(name, parent=None):
    m = (parent.nested_type if parent is not None else fd.message_type).add(); m.name = name; return m

  def field(m, name, num, typ, label=F.LABEL_OPTIONAL, type_name=None):
    f = m.field.add(); f.name, f.number, f.type, f.label = name, num, typ, label
    if type_name: f.type_name = '.tbtest.' + type_name
  shape = msg('TensorShapeProto')
  dim = msg('Dim', shape)
  field(dim, 'size', 1, F.TYPE_INT64); field(dim, 'name', 2, F.TYPE_STRING)
  field(shape, 'dim', 2, F.TYPE_MESSAGE, F.LABEL_REPEATED, 'TensorShapeProto.Dim'); field(shape, 'unknown_rank', 3, F.TYPE_BOOL)
  ver = msg('VersionDef')
  field(ver, 'producer', 1, F.TYPE_INT32); field(ver, 'min_consumer', 2, F.TYPE_INT32)
  head = msg('BundleHeaderProto')
  field(head, 'num_shards', 1, F.TYPE_INT32); field(head, 'endianness', 2, F.TYPE_INT32)
  field(head, 'version', 3, F.TYPE_MESSAGE, type_name='VersionDef')
  ent = msg('BundleEntryProto')
  field(ent, 'dtype', 1, F.TYPE_INT32); field(ent, 'shape', 2, F.TYPE_MESSAGE, type_name='TensorShapeProto')
  field(ent, 'shard_id', 3, F.TYPE_INT32); field(ent, 'offset', 4, F.TYPE_INT64); field(ent, 'size', 5, F.TYPE_INT64)
  field(ent, 'crc32c', 6, F.TYPE_FIXED32)
  pool = descriptor_pool.DescriptorPool(); pool.Add(fd)
  get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName('tbtest.' + n))
  return get('BundleHeaderProto'), get('BundleEntryProto')


def _varint(buf, pos):
  v = shift = 0
  while True:
    b = buf[pos]; pos += 1
    v |= (b & 0x7f) << shift; shift += 7
    if not b & 0x80:
      return v, pos


def _put_varint(v):
  out = bytearray()
  while v >= 0x80:
    out.append((v & 0x7f) | 0x80); v >>= 7
  out.append(v)
  return bytes(out)


def _py_block_entries(block):
  n_restarts = struct.unpack('<I', block[-4:])[0]
  end = len(block) - 4 - 4 * n_restarts
  pos, key, out = 0, b'', []
  while pos < end:
    shared, pos = _varint(block, pos); non_shared, pos = _varint(block, pos); vlen, pos = _varint(block, pos)
    key = key[:shared] + block[pos:pos + non_shared]; pos += non_shared
    out.append((key, block[pos:pos + vlen])); pos += vlen
  return out


def _py_read_table(path):
  """[(key, value)] of a leveldb-format table with uncompressed blocks, checking every block trailer."""
  f = open(path, 'rb').read()
  assert struct.unpack('<Q', f[-8:])[0] == MAGIC
  footer = f[-48:-8]
  _, p = _varint(footer, 0); _, p = _varint(footer, p)
  ioff, p = _varint(footer, p); isize, p = _varint(footer, p)

  def block(off, size):
    assert f[off + size] == 0
    assert _io.load().geeco_io_crc32c_unmask(struct.unpack('<I', f[off + size + 1:off + size + 5])[0]) == _io.crc32c(f[off:off + size + 1])
    return f[off:off + size]
  out = []
  for _, handle in _py_block_entries(block(ioff, isize)):
    off, p = _varint(handle, 0); size, p = _varint(handle, p)
    out += _py_block_entries(block(off, size))
  return out


def _snappy_literal_and_copies(data):
  """A valid snappy stream for `data` that uses literals and, where the input repeats, copy elements."""
  out = bytearray(_put_varint(len(data)))
  pos = 0
  while pos < len(data):
    # try a copy of 4..11 bytes with a 1-byte-offset element when the previous 4 bytes repeat
    if pos >= 4 and data[pos:pos + 4] == data[pos - 4:pos]:
      n = 4
      while n < 11 and pos + n < len(data) and data[pos + n] == data[pos + n - 4]:
        n += 1
      out.append(1 | ((n - 4) << 2) | ((4 >> 8) << 5)); out.append(4)
      pos += n
      continue
    n = min(70, len(data) - pos)
    if n <= 60:
      out.append((n - 1) << 2)
    else:
      out.append(60 << 2); out.append(n - 1)
    out += data[pos:pos + n]; pos += n
  return bytes(out)


def _py_write_table(path, items, compress=False, block_entries=3):
  """Writes a leveldb-format table with `block_entries` entries per data block (no prefix sharing beyond what
  the restart interval of 2 allows), optionally snappy-compressed."""
  f = bytearray()

  def build(entries):
    buf, restarts, last = bytearray(), [], b''
    for i, (k, v) in enumerate(entries):
      shared = 0
      if i % 2 == 0:
        restarts.append(len(buf))
      else:
        while shared < min(len(k), len(last)) and k[shared] == last[shared]:
          shared += 1
      buf += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
      last = k
    restarts = restarts or [0]
    return bytes(buf) + b''.join(struct.pack('<I', r) for r in restarts) + struct.pack('<I', len(restarts))

  def emit(block):
    body, typ = (_snappy_literal_and_copies(block), 1) if compress else (block, 0)
    handle = _put_varint(len(f)) + _put_varint(len(body))
    f.extend(body); f.append(typ)
    f.extend(struct.pack('<I', _io.load().geeco_io_crc32c_mask(_io.crc32c(body + bytes([typ])))))
    return handle
  index = []
  for i in range(0, len(items), block_entries):
    part = items[i:i + block_entries]
    index.append((part[-1][0] + b'\xff', emit(build(part))))          # any separator >= the block's last key
  meta = emit(build([]))
  idx = emit(build(index))
  footer = meta + idx
  f.extend(footer + bytes(40 - len(footer)) + struct.pack('<Q', MAGIC))
  open(path, 'wb').write(bytes(f))


def _tensors():
  rng = np.random.default_rng(0)
  t = {'GoalVMC/ConvEncoder/conv1/kernel': rng.standard_normal((3, 3, 3, 32)).astype(np.float32),
       'GoalVMC/ConvEncoder/conv1/bias': np.zeros(32, np.float32),
       'GoalVMC/ConvEncoder/conv1/kernel/Adam': rng.standard_normal((3, 3, 3, 32)).astype(np.float32),
       'GoalVMC/LSTMDecoder/lstm_memory': np.zeros((4, 256), np.float32),
       'beta1_power': np.array(0.9, np.float32), 'global_step': np.array(1234567, np.int64),
       'counts': np.arange(-3, 4, dtype=np.int32), 'empty': np.zeros((0, 5), np.float32)}
  for i in range(40):                                             # enough names to exercise restart points
    t['GoalVMC/fill/var_%03d' % i] = rng.standard_normal(i % 7 + 1).astype(np.float32)
  return t


def test_bundle_roundtrip(tmp_path):
  t = _tensors()
  prefix = str(tmp_path / 'model.ckpt-7')
  write_bundle(prefix, t)
  assert sorted(os.listdir(tmp_path)) == ['model.ckpt-7.data-00000-of-00001', 'model.ckpt-7.index']
  with BundleReader(prefix) as r:
    assert r.names() == sorted(t)
    assert r.shape('GoalVMC/ConvEncoder/conv1/kernel') == (3, 3, 3, 32) and r.shape('beta1_power') == ()
    assert r.has_tensor('global_step') and not r.has_tensor('nope')
    with pytest.raises(KeyError):
      r.get_tensor('nope')
  back = read_bundle(prefix)
  for k, v in t.items():
    assert back[k].dtype == v.dtype and back[k].shape == v.shape
    np.testing.assert_array_equal(back[k], v)


def test_native_writer_read_by_python_restatement(tmp_path):
  Header, Entry = _bundle_protos()
  t = _tensors()
  prefix = str(tmp_path / 'ck')
  write_bundle(prefix, t)
  items = _py_read_table(prefix + '.index')
  assert [k for k, _ in items] == [b''] + [n.encode() for n in sorted(t)]     # sorted keys, header first
  head = Header.FromString(items[0][1])
  assert head.num_shards == 1 and head.endianness == 0 and head.version.producer == 1
  data = open(prefix + '.data-00000-of-00001', 'rb').read()
  offset = 0
  for k, v in items[1:]:
    e = Entry.FromString(v)
    want = t[k.decode()]
    assert e.dtype == {np.dtype('float32'): 1, np.dtype('int32'): 3, np.dtype('int64'): 9}[want.dtype]
    assert [d.size for d in e.shape.dim] == list(want.shape)
    assert e.shard_id == 0 and e.offset == offset and e.size == want.nbytes
    raw = data[e.offset:e.offset + e.size]
    assert raw == want.tobytes()
    assert _io.load().geeco_io_crc32c_unmask(e.crc32c) == _io.crc32c(raw)
    offset += e.size
  assert offset == len(data)


@pytest.mark.parametrize('compress', [False, True])
def test_native_reader_reads_python_written_table(tmp_path, compress):
  """Foreign writer: several data blocks, other restart interval, separator index keys, snappy blocks."""
  Header, Entry = _bundle_protos()
  t = _tensors()
  prefix = str(tmp_path / 'foreign')
  data, items = bytearray(), []
  head = Header(num_shards=1); head.version.producer = 1
  items.append((b'', head.SerializeToString()))
  for name in sorted(t):
    a = t[name]
    e = Entry(dtype={np.dtype('float32'): 1, np.dtype('int32'): 3, np.dtype('int64'): 9}[a.dtype], offset=len(data),
              size=a.nbytes, crc32c=_io.load().geeco_io_crc32c_mask(_io.crc32c(a.tobytes())))
    e.shape.SetInParent()
    for d in a.shape:
      e.shape.dim.add().size = d
    data += a.tobytes()
    items.append((name.encode(), e.SerializeToString()))
  _py_write_table(prefix + '.index', items, compress=compress)
  open(prefix + '.data-00000-of-00001', 'wb').write(bytes(data))
  back = read_bundle(prefix)
  assert sorted(back) == sorted(t)
  for k, v in t.items():
    np.testing.assert_array_equal(back[k], v)


def test_bundle_corruption_is_detected(tmp_path):
  t = {'a': np.arange(10, dtype=np.float32), 'b': np.ones((2, 2), np.float32)}
  prefix = str(tmp_path / 'c')
  write_bundle(prefix, t)
  dpath, ipath = prefix + '.data-00000-of-00001', prefix + '.index'
  good_data, good_index = open(dpath, 'rb').read(), open(ipath, 'rb').read()
  bad = bytearray(good_data); bad[5] ^= 0x40
  open(dpath, 'wb').write(bytes(bad))
  with BundleReader(prefix) as r:
    with pytest.raises(_io.DataLossError, match='checksum'):
      r.get_tensor('a')
    np.testing.assert_array_equal(r.get_tensor('b'), t['b'])           # the other tensor is intact
    assert r.get_tensor('a', verify_crc=False)[2] == 2.0
  open(dpath, 'wb').write(good_data[:-4])
  with BundleReader(prefix) as r:
    with pytest.raises(IOError, match='short read'):
      r.get_tensor('b')
  open(dpath, 'wb').write(good_data)
  bad = bytearray(good_index); bad[3] ^= 1
  open(ipath, 'wb').write(bytes(bad))
  with pytest.raises(_io.DataLossError, match='block checksum'):
    BundleReader(prefix)
  open(ipath, 'wb').write(good_index[:-1])
  with pytest.raises(_io.DataLossError, match='bad magic'):
    BundleReader(prefix)
  with pytest.raises(FileNotFoundError):
    BundleReader(str(tmp_path / 'missing'))


def test_writer_argument_checks(tmp_path):
  with pytest.raises(ValueError, match='dtype'):
    write_bundle(str(tmp_path / 'x'), {'a': np.zeros(3, np.uint8)})
  write_bundle(str(tmp_path / 'y'), {'a': np.zeros(3, np.float64)})       # float64 is stored as float32
  assert read_bundle(str(tmp_path / 'y'))['a'].dtype == np.float32


# ------------------------------------------------------------------------------------------------
# the Estimator's checkpoint protocol over both file formats (CPU: a stand-in for the Engine's arenas)
# ------------------------------------------------------------------------------------------------
class _ArenaEngine(object):
  """The part of geeco_b200.engine.Engine that save/restore_checkpoint touch, over CPU tensors."""

  def __init__(self, seed):
    import collections
    import torch
    self.param_table = collections.OrderedDict([('GoalVMC/ConvEncoder/conv1/kernel', (0, 54, (3, 3, 3, 2))),
                                                ('GoalVMC/ConvEncoder/conv1/bias', (56, 2, (2,))),
                                                ('GoalVMC/LSTMDecoder/fc1/kernel', (60, 12, (4, 3)))])
    g = torch.Generator().manual_seed(seed)
    self.theta, self.adam_m, self.adam_v = (torch.randn(72, generator=g) for _ in range(3))
    self.training, self.global_step, self.N = True, 0, 4

    class Cfg: dim_h_lstm = 8
    self.cfg = Cfg()

  def view(self, name, arena=None):
    off, cnt, shape = self.param_table[name]
    return (self.theta if arena is None else arena)[off:off + cnt].view(shape)

  def param_names(self):
    return list(self.param_table)

  def set_params(self, named):
    import torch
    for n, a in named.items():
      self.view(n).copy_(torch.as_tensor(np.asarray(a, dtype=np.float32)))

  def set_global_step(self, t):
    self.global_step = int(t)


@pytest.mark.parametrize('fmt', ['npz', 'bundle'])
def test_estimator_checkpoint_protocol(tmp_path, fmt):
  import torch
  from geeco_b200.estimator import latest_checkpoint, restore_checkpoint, save_checkpoint
  md = str(tmp_path)
  a = _ArenaEngine(1)
  assert latest_checkpoint(md) is None
  for step in (5, 10, 15):
    a.global_step = step
    a.theta += 1.0
    prefix = save_checkpoint(a, md, keep_max=2, fmt=fmt)
    assert os.path.basename(prefix) == 'model.ckpt-%d' % step and latest_checkpoint(md) == prefix
  files = sorted(os.listdir(md))
  per = ['.npz'] if fmt == 'npz' else ['.data-00000-of-00001', '.index']
  assert files == ['checkpoint'] + ['model.ckpt-%d%s' % (s, e) for s in (10, 15) for e in per]    # keep_checkpoint_max
  header = open(os.path.join(md, 'checkpoint')).read().splitlines()
  assert header == ['model_checkpoint_path: "model.ckpt-15"', 'all_model_checkpoint_paths: "model.ckpt-10"',
                    'all_model_checkpoint_paths: "model.ckpt-15"']
  b = _ArenaEngine(2)
  assert restore_checkpoint(b, latest_checkpoint(md)) == 15
  for n in a.param_names():
    assert torch.equal(a.view(n), b.view(n))
    assert torch.equal(a.view(n, a.adam_m), b.view(n, b.adam_m)) and torch.equal(a.view(n, a.adam_v), b.view(n, b.adam_v))
  if fmt == 'bundle':                                             # the names tf.train.Saver would have written
    names = BundleReader(latest_checkpoint(md)).names()
    assert 'GoalVMC/ConvEncoder/conv1/kernel/Adam_1' in names and 'beta2_power' in names and 'global_step' in names
    assert read_bundle(latest_checkpoint(md))['GoalVMC/LSTMDecoder/lstm_memory'].shape == (4, 16)
    # tf.train.AdamOptimizer: the accumulators start at beta and are multiplied after each step -> beta ** (t + 1)
    rb = read_bundle(latest_checkpoint(md))
    assert rb['beta1_power'] == np.float32(0.9 ** 16) and rb['beta2_power'] == np.float32(0.999 ** 16)
  # an inference engine restores weights only, from either format
  c = _ArenaEngine(3); c.training = False
  m0 = c.adam_m.clone()
  restore_checkpoint(c, latest_checkpoint(md))
  assert torch.equal(c.view('GoalVMC/LSTMDecoder/fc1/kernel'), a.view('GoalVMC/LSTMDecoder/fc1/kernel')) and torch.equal(c.adam_m, m0)
  with pytest.raises(ValueError, match='checkpoint_format'):
    save_checkpoint(a, md, fmt='hdf5')


@pytest.mark.parametrize('fmt', ['npz', 'bundle'])
def test_verify_checkpoint_names_the_mismatch(tmp_path, fmt):
  from geeco_b200 import create_e2evmc_config
  from geeco_b200.estimator import verify_checkpoint
  from geeco_b200.graph import variable_table
  seq = create_e2evmc_config(dict(proc_obs='sequence', proc_tgt='residual'))
  geecof = create_e2evmc_config(dict(proc_obs='dynimg', proc_tgt='dyndiff'))
  # tiny stand-ins with the right names: shapes are checked too, so use the real ones for a few small variables
  arrays = {n: np.zeros(s, np.float32) for n, s in variable_table(seq) if int(np.prod(s)) < 20000}
  prefix = str(tmp_path / 'model.ckpt-1')
  if fmt == 'bundle':
    write_bundle(prefix, arrays)
  else:
    np.savez(prefix + '.npz', **arrays)
  with pytest.raises(ValueError, match="no variable 'GoalVMC/ConvEncoder/conv3/kernel'"):
    verify_checkpoint(prefix, seq)                              # (the large kernels were left out above)
  full = {n: np.zeros(s, np.float32) for n, s in variable_table(seq)}
  if fmt == 'bundle':
    write_bundle(prefix, full)
  else:
    np.savez(prefix + '.npz', **full)
  verify_checkpoint(prefix, seq)
  with pytest.raises(ValueError, match='DynBuffEncoder'):
    verify_checkpoint(prefix, geecof)
