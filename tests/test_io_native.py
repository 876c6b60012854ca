"""CPU tests of libgeeco_io.so: CRC-32C, TFRecord framing, SequenceExample decode, windows, input pipeline.

Independent checkers used here (none of them is on the product path):
  * RFC 3720 B.4 / iSCSI CRC-32C known answers;
  * Python's `zlib` + `struct` for the TFRecord container;
  * the real protobuf runtime (`google.protobuf`) over the published example.proto / feature.proto schema,
    declared programmatically below, for the SequenceExample wire format;
  * a literal NumPy transcription of the reference's pipeline stages (src/data/geeco_gym.py:291-399, :598-631).
"""
import os
import struct
import zlib

import numpy as np
import pytest

from geeco_b200 import _io, data as gdata, input_pipeline as ip, tfrecord as tfr


# ------------------------------------------------------------------------------------------------
# the C-ABI library loads and exports what include/geeco_io.h declares
# ------------------------------------------------------------------------------------------------
def test_header_symbols_exported():
  import re
  header = open(os.path.join(os.path.dirname(__file__), '..', 'include', 'geeco_io.h')).read()
  declared = set(re.findall(r'\b(geeco_[a-z0-9_]+)\s*\(', header))
  lib = _io.load()
  assert declared == set(_io.SYMBOLS), declared ^ set(_io.SYMBOLS)
  for name in declared:
    assert hasattr(lib, name)
  assert lib.geeco_io_version() == 1


# ------------------------------------------------------------------------------------------------
# CRC-32C
# ------------------------------------------------------------------------------------------------
def test_crc32c_known_answers():
  assert _io.crc32c(b'') == 0
  assert _io.crc32c(b'123456789') == 0xE3069283
  assert _io.crc32c(bytes(32)) == 0x8A9136AA                      # RFC 3720 B.4
  assert _io.crc32c(b'\xff' * 32) == 0x62A8AB43
  assert _io.crc32c(bytes(range(32))) == 0x46DD794E
  assert _io.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C


def _crc32c_bitwise(data):
  crc = 0xFFFFFFFF
  for b in data:
    crc ^= b
    for _ in range(8):
      crc = (crc >> 1) ^ (0x82F63B78 if crc & 1 else 0)
  return crc ^ 0xFFFFFFFF


def test_crc32c_against_bitwise_and_extend():
  rng = np.random.default_rng(0)
  for n in (1, 7, 8, 9, 63, 64, 65, 1000, 4099):
    buf = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
    assert _io.crc32c(buf) == _crc32c_bitwise(buf)
    for cut in (0, 1, n // 2, n):                                 # unaligned starts, incremental form
      assert _io.crc32c(buf[cut:], _io.crc32c(buf[:cut])) == _io.crc32c(buf)


def test_crc_mask_roundtrip():
  lib = _io.load()
  for c in (0, 1, 0xE3069283, 0xFFFFFFFF, 0x12345678):
    m = lib.geeco_io_crc32c_mask(c)
    assert m == ((((c >> 15) | (c << 17)) + 0xa282ead8) & 0xFFFFFFFF)
    assert lib.geeco_io_crc32c_unmask(m) == c


# ------------------------------------------------------------------------------------------------
# TFRecord container
# ------------------------------------------------------------------------------------------------
def _frame_records(records):
  out = b''
  for r in records:
    head = struct.pack('<Q', len(r))
    out += head + struct.pack('<I', _io.masked_crc32c(head)) + r + struct.pack('<I', _io.masked_crc32c(r))
  return out


@pytest.mark.parametrize('compression,suffix', [('none', '.tfrecord'), ('zlib', '.tfrecord.zlib'), ('gzip', '.tfrecord.gzip')])
def test_tfrecord_roundtrip_and_python_zlib(tmp_path, compression, suffix):
  rng = np.random.default_rng(1)
  records = [b'', b'a', rng.integers(0, 256, size=70000, dtype=np.uint8).tobytes(), b'xyz' * 1000]
  path = str(tmp_path / ('f' + suffix))
  tfr.write_tfrecord(path, records, 'auto')
  raw = open(path, 'rb').read()
  if compression == 'zlib':
    raw = zlib.decompress(raw)
  elif compression == 'gzip':
    raw = zlib.decompress(raw, 16 + 15)
  assert raw == _frame_records(records)                           # writer == independent framing
  with tfr.TFRecordFile(path) as f:                               # reader, suffix-detected compression
    assert len(f) == len(records)
    assert [f[i] for i in range(len(f))] == records
  with tfr.TFRecordFile(path, compression=compression) as f:
    assert [f[i] for i in range(len(f))] == records


def test_tfrecord_reads_foreign_zlib_stream(tmp_path):
  """A file compressed by Python's zlib in several flushed pieces (as a streaming writer produces)."""
  records = [os.urandom(5000), b'hello']
  comp = zlib.compressobj(6)
  raw = _frame_records(records)
  packed = comp.compress(raw[:3000]) + comp.flush(zlib.Z_SYNC_FLUSH) + comp.compress(raw[3000:]) + comp.flush()
  path = str(tmp_path / 'x.tfrecord.zlib')
  open(path, 'wb').write(packed)
  with tfr.TFRecordFile(path) as f:
    assert [f[i] for i in range(len(f))] == records


def test_tfrecord_corruption_is_detected(tmp_path):
  records = [b'0123456789' * 10, b'abcdef']
  raw = bytearray(_frame_records(records))
  p = str(tmp_path / 'ok.tfrecord')
  open(p, 'wb').write(bytes(raw))
  assert len(tfr.TFRecordFile(p)) == 2
  bad = bytearray(raw); bad[20] ^= 1                              # payload bit flip
  open(p, 'wb').write(bytes(bad))
  with pytest.raises(_io.DataLossError, match='data checksum'):
    tfr.TFRecordFile(p)
  assert len(tfr.TFRecordFile(p, verify_crc=False)) == 2          # tf's reader can skip the check as well
  bad = bytearray(raw); bad[0] ^= 1                               # length bit flip
  open(p, 'wb').write(bytes(bad))
  with pytest.raises(_io.DataLossError, match='length checksum'):
    tfr.TFRecordFile(p)
  open(p, 'wb').write(bytes(raw[:-3]))                            # truncated
  with pytest.raises(_io.DataLossError, match='truncated'):
    tfr.TFRecordFile(p)
  open(p, 'wb').write(b'')
  assert len(tfr.TFRecordFile(p)) == 0
  z = str(tmp_path / 'bad.tfrecord.zlib')
  open(z, 'wb').write(zlib.compress(bytes(raw))[:-5])
  with pytest.raises(_io.DataLossError, match='compressed stream'):
    tfr.TFRecordFile(z)
  with pytest.raises(FileNotFoundError):
    tfr.TFRecordFile(str(tmp_path / 'missing.tfrecord'))


# ------------------------------------------------------------------------------------------------
# SequenceExample wire format against the protobuf runtime
# ------------------------------------------------------------------------------------------------
def _example_protos():
  """tensorflow/core/example/{feature,example}.proto declared through descriptor_pb2 (published schema)."""
  from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
  F = descriptor_pb2.FieldDescriptorProto
  fd = descriptor_pb2.FileDescriptorProto(name='geeco_test_example.proto', package='tftest', syntax='proto3')

  def msg(name):
    m = fd.message_type.add(); m.name = name; return m

  def field(m, name, num, typ, label=F.LABEL_OPTIONAL, type_name=None, packed=None, oneof=None):
    f = m.field.add(); f.name, f.number, f.type, f.label = name, num, typ, label
    if type_name: f.type_name = '.tftest.' + type_name
    if packed is not None: f.options.packed = packed
    if oneof is not None: f.oneof_index = oneof
    return f

  field(msg('BytesList'), 'value', 1, F.TYPE_BYTES, F.LABEL_REPEATED)
  field(msg('FloatList'), 'value', 1, F.TYPE_FLOAT, F.LABEL_REPEATED, packed=True)
  field(msg('Int64List'), 'value', 1, F.TYPE_INT64, F.LABEL_REPEATED, packed=True)
  field(msg('FloatListUnpacked'), 'value', 1, F.TYPE_FLOAT, F.LABEL_REPEATED, packed=False)
  field(msg('Int64ListUnpacked'), 'value', 1, F.TYPE_INT64, F.LABEL_REPEATED, packed=False)
  feat = msg('Feature'); feat.oneof_decl.add().name = 'kind'
  field(feat, 'bytes_list', 1, F.TYPE_MESSAGE, type_name='BytesList', oneof=0)
  field(feat, 'float_list', 2, F.TYPE_MESSAGE, type_name='FloatList', oneof=0)
  field(feat, 'int64_list', 3, F.TYPE_MESSAGE, type_name='Int64List', oneof=0)
  ufeat = msg('FeatureUnpacked'); ufeat.oneof_decl.add().name = 'kind'
  field(ufeat, 'bytes_list', 1, F.TYPE_MESSAGE, type_name='BytesList', oneof=0)
  field(ufeat, 'float_list', 2, F.TYPE_MESSAGE, type_name='FloatListUnpacked', oneof=0)
  field(ufeat, 'int64_list', 3, F.TYPE_MESSAGE, type_name='Int64ListUnpacked', oneof=0)
  field(msg('FeatureList'), 'feature', 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, type_name='Feature')
  field(msg('FeatureListUnpacked'), 'feature', 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, type_name='FeatureUnpacked')

  def map_field(parent, name, value_type):
    entry = parent.nested_type.add(); entry.name = name.title().replace('_', '') + 'Entry'; entry.options.map_entry = True
    field(entry, 'key', 1, F.TYPE_STRING)
    field(entry, 'value', 2, F.TYPE_MESSAGE, type_name=value_type)
    f = parent.field.add(); f.name, f.number, f.type, f.label = name, 1, F.TYPE_MESSAGE, F.LABEL_REPEATED
    f.type_name = '.tftest.%s.%s' % (parent.name, entry.name)

  map_field(msg('Features'), 'feature', 'Feature')
  map_field(msg('FeatureLists'), 'feature_list', 'FeatureList')
  map_field(msg('FeatureListsUnpacked'), 'feature_list', 'FeatureListUnpacked')
  se = msg('SequenceExample')
  field(se, 'context', 1, F.TYPE_MESSAGE, type_name='Features')
  field(se, 'feature_lists', 2, F.TYPE_MESSAGE, type_name='FeatureLists')
  seu = msg('SequenceExampleUnpacked')
  field(seu, 'context', 1, F.TYPE_MESSAGE, type_name='Features')
  field(seu, 'feature_lists', 2, F.TYPE_MESSAGE, type_name='FeatureListsUnpacked')
  pool = descriptor_pool.DescriptorPool()
  pool.Add(fd)
  get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName('tftest.' + n))
  return get('SequenceExample'), get('SequenceExampleUnpacked')


def _fill(ex, frames=5):
  rng = np.random.default_rng(7)
  ex.context.feature['episode_length'].int64_list.value.append(frames)
  ex.context.feature['names'].bytes_list.value.extend([b'alpha', b'', 'béta'.encode()])
  ex.context.feature['gain'].float_list.value.append(0.25)
  ex.context.feature['unset'].SetInParent()
  want = {'step': [], 'neg': [], 'vec': [], 'scalar': []}
  for t in range(frames):
    fl = ex.feature_lists.feature_list
    fl['step'].feature.add().int64_list.value.append(t)
    neg = [-1, -(2 ** 63), 2 ** 63 - 1, 300 * t]
    fl['neg'].feature.add().int64_list.value.extend(neg)
    vec = rng.standard_normal(9).astype(np.float32)
    fl['vec'].feature.add().float_list.value.extend(vec.tolist())
    fl['scalar'].feature.add().float_list.value.append(float(np.float32(t) / 3))
    fl['ragged'].feature.add().float_list.value.extend([1.0] * (t % 3))
    fl['empty'].feature.add().float_list.SetInParent()
    want['step'].append([t]); want['neg'].append(neg); want['vec'].append(vec); want['scalar'].append([np.float32(t) / 3])
  return want


@pytest.mark.parametrize('unpacked', [False, True])
def test_sequence_example_decoded_like_protobuf_runtime(unpacked):
  cls = _example_protos()[1 if unpacked else 0]
  ex = cls()
  want = _fill(ex)
  blob = ex.SerializeToString()
  s = tfr.SequenceExample(blob)
  assert sorted(s.keys(_io.CONTEXT)) == ['episode_length', 'gain', 'names', 'unset']
  assert s.keys(_io.SEQUENCE) == ['empty', 'neg', 'ragged', 'scalar', 'step', 'vec']
  assert s.info('episode_length', _io.CONTEXT) == (_io.KIND_INT64, 1, 1)
  assert s.info('unset', _io.CONTEXT) == (_io.KIND_NONE, 1, 0)
  assert s.info('vec') == (_io.KIND_FLOAT, 5, 9)
  assert s.info('ragged') == (_io.KIND_FLOAT, 5, -1)
  assert s.info('empty') == (_io.KIND_FLOAT, 5, 0)
  assert s.ints('episode_length', _io.CONTEXT).tolist() == [[5]]
  assert s.strings('names') == [b'alpha', b'', 'béta'.encode()]
  assert s.floats('gain', _io.CONTEXT).tolist() == [[0.25]]
  np.testing.assert_array_equal(s.ints('step'), np.array(want['step']))
  np.testing.assert_array_equal(s.ints('neg'), np.array(want['neg'], dtype=np.int64))
  np.testing.assert_array_equal(s.floats('vec'), np.stack(want['vec']))           # bit-exact floats
  np.testing.assert_array_equal(s.floats('scalar'), np.array(want['scalar'], dtype=np.float32))
  assert s.floats('empty').shape == (5, 0)
  with pytest.raises(ValueError, match='ragged'):                 # FixedLenSequenceFeature would reject it too
    s.floats('ragged')
  with pytest.raises(ValueError, match='expected 4'):
    s.floats('vec', per_frame=4)
  with pytest.raises(KeyError):
    s.floats('nope')
  with pytest.raises(ValueError):
    s.floats('step')                                              # kind mismatch


def test_python_encoder_parses_with_protobuf_runtime():
  """The writer side: bytes from encode_sequence_example are a valid SequenceExample with the same content."""
  cls = _example_protos()[0]
  data = gdata.synthetic_episode(episode_length=3, height=4, width=5, seed=3)
  ck, fk = gdata.encoding_keys_v4(data)
  blob = tfr.encode_sequence_example(data, ck, fk)
  ex = cls.FromString(blob)
  assert sorted(ex.context.feature.keys()) == sorted(ck)
  assert sorted(ex.feature_lists.feature_list.keys()) == sorted(fk)
  assert list(ex.context.feature['monitored_joints'].bytes_list.value) == [j.encode() for j in data['monitored_joints']]
  assert list(ex.context.feature['img_width'].int64_list.value) == [5]
  for t, frame in enumerate(data['sequence']):
    got = ex.feature_lists.feature_list['rgb'].feature[t].float_list.value
    np.testing.assert_array_equal(np.array(got, dtype=np.float32), frame['rgb'].reshape(-1).astype(np.float32))
    assert list(ex.feature_lists.feature_list['step'].feature[t].int64_list.value) == [t]
    k = 'joint_qpos-robot0:elbow_flex_joint'
    assert ex.feature_lists.feature_list[k].feature[t].float_list.value[0] == np.float32(frame[k])
  # deterministic protobuf serialisation of the parsed message reproduces the bytes (sorted map keys)
  assert ex.SerializeToString(deterministic=True) == blob


def test_convert_to_feature_type_dispatch():
  with pytest.raises(TypeError):
    tfr.convert_to_feature(np.zeros(3, dtype=np.uint16))
  with pytest.raises(TypeError):
    tfr.convert_to_feature([b'x'])
  with pytest.raises(TypeError):
    tfr.convert_to_feature(None)


def test_malformed_sequence_example_is_rejected():
  cls = _example_protos()[0]
  ex = cls(); _fill(ex)
  blob = ex.SerializeToString()
  for cut in (1, 5, len(blob) // 2, len(blob) - 1):
    with pytest.raises(_io.DataLossError):
      tfr.SequenceExample(blob[:cut])
  assert tfr.SequenceExample(b'').keys() == []


def test_pixel_bytes_exactness():
  data = {'sequence': [{'rgb': np.array([0, 255, 17, 128], dtype=np.uint8)},
                       {'rgb': np.array([0.5, 256.0, -1.0, 3.0], dtype=np.float32)}]}
  s = tfr.SequenceExample(tfr.encode_sequence_example(data, [], ['rgb']))
  px, inexact = s.pixel_bytes('rgb')
  assert px[0].tolist() == [0, 255, 17, 128] and px[1, 3] == 3
  assert inexact == 3


# ------------------------------------------------------------------------------------------------
# sliding windows
# ------------------------------------------------------------------------------------------------
def test_window_gather_matches_slicing():
  from hypothesis import given, settings, strategies as st

  @settings(max_examples=60, deadline=None)
  @given(st.integers(1, 30), st.integers(1, 8), st.integers(0, 5), st.integers(0, 3), st.data())
  def prop(frames, K, inner, first, draw):
    if K > frames:
      return
    src = np.arange(frames * max(inner, 1) * 2, dtype=np.float32).reshape(frames, max(inner, 1), 2)
    nw = frames - K + 1
    first = min(first, nw - 1)
    cnt = draw.draw(st.integers(0, nw - first))
    got = tfr.window_gather(src, K, first, cnt)
    want = np.stack([src[i:i + K] for i in range(first, first + cnt)]) if cnt else np.empty((0, K) + src.shape[1:], np.float32)
    np.testing.assert_array_equal(got, want)

  prop()
  with pytest.raises(ValueError, match='reach past'):
    tfr.window_gather(np.zeros((5, 2), np.float32), 4, 1, 2)


# ------------------------------------------------------------------------------------------------
# the input pipeline against a literal transcription of the reference's stages
# ------------------------------------------------------------------------------------------------
def _reference_stream(episodes, meta, K, fetch_target):
  """geeco_gym.py:291-399 + :598-631 in NumPy, one (feature, label) per window, in stream order."""
  H, W = meta.img_height, meta.img_width
  out = []
  for data in episodes:
    seq = data['sequence']
    ex = {k: np.stack([np.asarray(f[k], dtype=np.float32).reshape(-1) for f in seq]) for k in seq[0] if k != 'step'}
    ex['step'] = np.array([f['step'] for f in seq], dtype=np.int64)
    ex['rgb'] = ex['rgb'].reshape(-1, H, W, 3) / np.float32(255.0)                   # _parse_v4
    ex['depth'] = ex['depth'].reshape(-1, H, W, 1)
    p = {k: ex[k] for k in ('step', 'rgb', 'depth', 'cmd', 'ctrl')}
    p['ts'] = ex['ts'][:, 0]
    p['ee_state'], p['goal_state'], p['obj_state'] = ex['mocap_qpos-robot0:mocap'], ex['goal_qpos'], ex['obj_qpos']
    p['jnt_state'] = np.stack([ex['joint_qpos-%s' % j][:, 0] for j in ip.ARM_JOINTS], axis=1)   # _preprocess_states_v4
    p['vel_state'] = np.stack([ex['joint_qvel-%s' % j][:, 0] for j in ip.ARM_JOINTS], axis=1)
    p['grp_state'] = np.stack([ex['joint_qpos-%s' % j][:, 0] for j in ip.FINGER_JOINTS], axis=1)
    tgt = {'target_rgb': p['rgb'][-1], 'target_depth': p['depth'][-1]}
    for n in ('vel', 'ee', 'grp'):                                                   # _preprocess_targets_v3
      p['%s_target' % n] = np.roll(p['%s_state' % n], -1, axis=0)
    p = {k: v[:-1] for k, v in p.items()}
    S = meta.episode_length - 1
    for w in range(S - K + 1):                                                       # _window_v3 + _prepare_v4
      win = {k: v[w:w + K] for k, v in p.items()}
      feat = {k: win[k] for k in ip.FEATURE_KEYS}
      if fetch_target:
        feat.update(tgt)
      out.append((feat, {k: win[k][-1] for k in ip.LABEL_KEYS}))
  return out


@pytest.fixture(scope='module')
def dataset(tmp_path_factory):
  d = str(tmp_path_factory.mktemp('ds'))
  eps = gdata.write_synthetic_dataset(d, episodes=3, episode_length=11, height=8, width=6, seed=5, eval_episodes=2)
  return d, eps


def _assert_batch_equals(batch, ref_rows):
  feats, labels = batch
  assert set(feats) == set(ref_rows[0][0]) and set(labels) == set(ref_rows[0][1])
  for k in feats:
    want = np.stack([r[0][k] for r in ref_rows])
    assert feats[k].dtype == want.dtype and feats[k].shape == want.shape, k
    np.testing.assert_array_equal(feats[k], want, err_msg=k)                         # bit-exact
  for k in labels:
    np.testing.assert_array_equal(labels[k], np.stack([r[1][k] for r in ref_rows]), err_msg=k)


@pytest.mark.parametrize('K,B,fetch', [(4, 3, True), (2, 5, False), (1, 4, True), (10, 2, True)])
def test_pipeline_equals_reference_stages(dataset, K, B, fetch):
  d, eps = dataset
  meta = ip.get_meta_v4(d)
  ref = _reference_stream(eps[3:], meta, K, fetch)                                   # eval split = last two episodes
  assert len(ref) == 2 * (11 - 1 - K + 1)
  it = ip.pickplace_input_fn(d, 'default', 'eval', window_size=K, fetch_target=fetch, batch_size=B, num_threads=2)
  batches = list(it)
  assert len(batches) == len(it) == -(-len(ref) // B)                                # partial last batch is kept
  for i, b in enumerate(batches):
    _assert_batch_equals(b, ref[i * B:(i + 1) * B])
  # step holds the frame indices of the index contract
  steps = np.concatenate([b[0]['step'] for b in batches])
  for g in range(len(ref)):
    _, w, cur, _ = gdata.locate(g, 11, K)
    assert steps[g].tolist() == list(range(w, w + K)) and steps[g, -1] == cur


def test_pipeline_uint8_frames_and_epochs(dataset):
  d, eps = dataset
  meta = ip.get_meta_v4(d)
  ref = _reference_stream(eps[3:], meta, 4, True) * 2                                # repeat(num_epochs=2)
  f32 = list(ip.pickplace_input_fn_v4(d, 'default', 'eval', 4, True, batch_size=4, num_epochs=2, drop_remainder=True))
  u8 = list(ip.pickplace_input_fn_v4(d, 'default', 'eval', 4, True, batch_size=4, num_epochs=2, drop_remainder=True,
                                      frame_format='uint8'))
  assert len(f32) == len(u8) == len(ref) // 4
  for i, (a, b) in enumerate(zip(f32, u8)):
    _assert_batch_equals(a, ref[i * 4:(i + 1) * 4])
    assert b[0]['rgb'].dtype == np.uint8 and b[0]['target_rgb'].dtype == np.uint8
    np.testing.assert_array_equal(b[0]['rgb'].astype(np.float32) / np.float32(255.0), a[0]['rgb'])
    np.testing.assert_array_equal(b[0]['target_rgb'].astype(np.float32) / np.float32(255.0), a[0]['target_rgb'])
    np.testing.assert_array_equal(b[0]['jnt_state'], a[0]['jnt_state'])


@pytest.mark.parametrize('B', [4, 7])
def test_pipeline_pool_layout_expands_to_the_window_layout(dataset, B):
  """layout='pool' (frame pool + index, include/geeco_b200.h: geeco_batch.frame_index): gathering the pool through the
  index gives the window tensors of the default layout bit for bit; every frame a batch touches is stored once per
  episode piece; everything that is not an image is unchanged."""
  d, eps = dataset
  kw = dict(batch_size=B, num_epochs=1, frame_format='uint8')
  win = list(ip.pickplace_input_fn_v4(d, 'default', 'train', 4, True, seed=3, **kw))
  pool = list(ip.pickplace_input_fn_v4(d, 'default', 'train', 4, True, seed=3, layout='pool', **kw))
  assert len(win) == len(pool) > 0
  total_pool = total_win = 0
  for (fw, lw), (fp, lp) in zip(win, pool):
    idx, tix = fp['rgb_index'], fp['target_index']
    assert idx.dtype == np.int32 and idx.shape == fw['rgb'].shape[:2] and tix.shape == (fw['rgb'].shape[0],)
    assert fp['rgb'].ndim == 4 and fp['rgb'].shape[0] <= idx.size           # frames are shared (a 1-window rest: K frames)
    np.testing.assert_array_equal(fp['rgb'][idx], fw['rgb'])
    np.testing.assert_array_equal(fp['depth'][idx], fw['depth'])
    np.testing.assert_array_equal(fp['target_rgb'][tix], fw['target_rgb'])
    assert np.all(idx[:, 1:] == idx[:, :-1] + 1)                            # a window is K consecutive pool frames
    for k in fw:
      if k not in ('rgb', 'depth', 'target_rgb', 'target_depth'):
        np.testing.assert_array_equal(fp[k], fw[k], err_msg=k)
    for k in lw:
      np.testing.assert_array_equal(lp[k], lw[k], err_msg=k)
    total_pool += fp['rgb'].shape[0]; total_win += idx.size
  assert total_pool < 0.6 * total_win                                    # (B + K - 1) / (B * K) per piece
  with pytest.raises(ValueError, match='layout'):
    ip.pickplace_input_fn_v4(d, 'default', 'train', 4, True, layout='ring', **kw)
  with pytest.raises(ValueError, match='pool'):
    ip.pickplace_input_fn_v4(d, 'default', 'train', 4, True, layout='pool', device='cpu', **kw)


def test_pipeline_train_shuffles_episodes_only(dataset):
  d, eps = dataset
  meta = ip.get_meta_v4(d)
  it = ip.pickplace_input_fn_v4(d, 'default', 'train', 4, True, batch_size=7, seed=3)
  order = [int(os.path.basename(p)[:6]) for p in it.paths]
  assert sorted(order) == [0, 1, 2]
  ref = _reference_stream([eps[e] for e in order], meta, 4, True)
  got = list(it)
  for i, b in enumerate(got):
    _assert_batch_equals(b, ref[i * 7:(i + 1) * 7])
  again = ip.pickplace_input_fn_v4(d, 'default', 'train', 4, True, batch_size=7, seed=3)
  assert again.paths == it.paths                                                     # all ranks agree given a seed


@pytest.mark.parametrize('world', [2, 4])
def test_pipeline_rank_shards_are_slices_of_the_global_batch(dataset, world):
  d, eps = dataset
  B = 2
  G = B * world
  whole = list(ip.pickplace_input_fn_v4(d, 'default', 'eval', 4, True, batch_size=G, drop_remainder=True))
  for r in range(world):
    part = list(ip.pickplace_input_fn_v4(d, 'default', 'eval', 4, True, batch_size=B, rank=r, world=world))
    assert len(part) == len(whole)
    for gb, pb in zip(whole, part):
      for k in gb[0]:
        np.testing.assert_array_equal(pb[0][k], gb[0][k][r * B:(r + 1) * B], err_msg=k)
      for k in gb[1]:
        np.testing.assert_array_equal(pb[1][k], gb[1][k][r * B:(r + 1) * B], err_msg=k)
      lo, hi = gdata.rank_slice(0, G, r, world)
      assert (lo, hi) == (r * B, (r + 1) * B)


def test_pipeline_errors(dataset, tmp_path):
  d, eps = dataset
  meta = ip.get_meta_v4(d)
  with pytest.raises(ValueError, match='window_size'):
    ip.WindowBatches([], meta, window_size=11)
  with pytest.raises(NotImplementedError):
    ip.pickplace_input_fn(d, 'default', 'eval', encoding='v2')
  with pytest.raises(ValueError, match='frame_format'):
    ip.decode_episode(os.path.join(d, 'data', '000000.tfrecord.zlib'), meta, frame_format='bf16')
  # meta that disagrees with the file: FixedLenSequenceFeature shape check / episode length check
  wrong = meta._replace(img_width=7)
  with pytest.raises(ValueError, match='values per frame'):
    ip.decode_episode(os.path.join(d, 'data', '000000.tfrecord.zlib'), wrong)
  longer = meta._replace(episode_length=12)
  with pytest.raises(ValueError, match='episode_length'):
    list(ip.WindowBatches([os.path.join(d, 'data', '000000.tfrecord.zlib')], longer, batch_size=2))
  # an episode without a required feature
  data = gdata.synthetic_episode(episode_length=11, height=8, width=6, seed=1)
  ck, fk = gdata.encoding_keys_v4(data)
  fk = [k for k in fk if k != 'goal_qpos']
  p = str(tmp_path / 'x.tfrecord.zlib')
  tfr.write_tfrecord(p, [tfr.encode_sequence_example(data, ck, fk)])
  with pytest.raises(KeyError, match='goal_qpos'):
    list(ip.WindowBatches([p], meta, batch_size=2))
  # early exit of the consumer does not hang the producer thread
  it = iter(ip.pickplace_input_fn_v4(d, 'default', 'eval', 4, True, batch_size=1, prefetch_size=1))
  next(it); it.close()


def test_episode_cache_serves_identical_batches(dataset, tmp_path):
  d, eps = dataset
  cache = str(tmp_path / 'cache')
  kw = dict(window_size=4, fetch_target=True, batch_size=5, frame_format='uint8')
  plain = list(ip.pickplace_input_fn_v4(d, 'default', 'eval', **kw))
  first = list(ip.pickplace_input_fn_v4(d, 'default', 'eval', cache_dir=cache, **kw))
  entries = sorted(os.listdir(cache))
  assert len(entries) == 2 and all(e.endswith('.uint8.t1.d1.ep') for e in entries)
  # the second pass must not touch the records: make them unreadable and read again
  recs = [os.path.join(d, 'data', f) for f in ('000003.tfrecord.zlib', '000004.tfrecord.zlib')]
  saved = [(p, open(p, 'rb').read(), os.stat(p)) for p in recs]
  try:
    for p, raw, st in saved:
      open(p, 'wb').write(b'x' * len(raw)); os.utime(p, (st.st_atime, st.st_mtime))
    second = list(ip.pickplace_input_fn_v4(d, 'default', 'eval', cache_dir=cache, **kw))
  finally:
    for p, raw, st in saved:
      open(p, 'wb').write(raw); os.utime(p, (st.st_atime, st.st_mtime))
  for a, b, c in zip(plain, first, second):
    for k in a[0]:
      np.testing.assert_array_equal(a[0][k], b[0][k]); np.testing.assert_array_equal(a[0][k], c[0][k])
      assert a[0][k].dtype == c[0][k].dtype
    for k in a[1]:
      np.testing.assert_array_equal(a[1][k], c[1][k])
  # a different decode configuration gets its own entries
  list(ip.pickplace_input_fn_v4(d, 'default', 'eval', cache_dir=cache, window_size=4, fetch_target=True, batch_size=5))
  assert len(os.listdir(cache)) == 4


def test_recorder_writes_what_the_pipeline_reads(tmp_path):
  """TfrSequenceRecorder + PickAndPlaceEncodingV4 (data_recorder.py:71-156, geeco_gym.py:54-115) -> decode_episode."""
  from geeco_b200.data_recorder import PickAndPlaceEncodingV4, TfrSequenceRecorder
  data = gdata.synthetic_episode(episode_length=7, height=8, width=6, seed=9)
  meta = ip.PickAndPlaceMetaV4(**{k: data[k] for k in ip.PickAndPlaceMetaV4._fields})
  enc = PickAndPlaceEncodingV4(meta)
  assert (enc.context_keys, enc.frame_keys) == gdata.encoding_keys_v4(data)
  ctx_spec, seq_spec = enc.decode()
  assert set(seq_spec) == set(enc.frame_keys) and seq_spec['rgb'] == ('float', 8 * 6 * 3) and set(ctx_spec) == set(enc.context_keys)
  rec = TfrSequenceRecorder(enc, {k: data[k] for k in enc.context_keys}, str(tmp_path), record_name='000042')
  assert rec.record_path == str(tmp_path / '000042.tfrecord')
  for frame in data['sequence']:
    rec.feed(frame)
  with pytest.raises(ValueError, match='expected data fields'):
    rec.feed({k: v for k, v in data['sequence'][0].items() if k != 'ts'})
  with pytest.raises(KeyError):
    rec.finalize(compression='lz4')
  path = rec.finalize(compression='zlib')
  assert path.endswith('000042.tfrecord.zlib')
  ep = ip.decode_episode(path, meta, fetch_target=True, frame_format='uint8')
  want = np.stack([f['rgb'] for f in data['sequence']])
  np.testing.assert_array_equal(ep['rgb'], want[:-1])
  np.testing.assert_array_equal(ep['target_rgb'], want[-1])
  np.testing.assert_array_equal(ep['cmd'], np.stack([f['cmd'] for f in data['sequence']])[:-1])
  with tfr.TFRecordFile(path) as f:                                 # context as recorded
    s = tfr.SequenceExample(f[0])
    assert s.strings('task_object') == [b'cube2'] and s.ints('episode_length', _io.CONTEXT).tolist() == [[7]]
    assert s.strings('monitored_joints') == [j.encode() for j in data['monitored_joints']]
  plain = TfrSequenceRecorder(enc, {k: data[k] for k in enc.context_keys}, str(tmp_path), record_name='p')
  for frame in data['sequence']:
    plain.feed(frame)
  assert len(tfr.TFRecordFile(plain.finalize())) == 1               # uncompressed .tfrecord


def test_batch_ranges_partition_the_stream_property():
  """Every stream position is read exactly once across ranks (up to the dropped incomplete global batch), pieces
  never cross an episode, and a rank's ranges follow SURVEY 8e's contiguous-slice rule."""
  from hypothesis import given, settings, strategies as st
  Meta = ip.PickAndPlaceMetaV4

  @settings(max_examples=80, deadline=None)
  @given(st.integers(2, 30), st.integers(1, 6), st.integers(1, 9), st.integers(1, 4), st.integers(0, 5), st.integers(1, 3),
         st.booleans())
  def prop(L, K, B, world, files, epochs, drop):
    if K > L - 1:
      return
    meta = Meta(L, 4, 4, [], [], [], [], 4, 2)
    its = [ip.WindowBatches(['f%d' % i for i in range(files)], meta, window_size=K, batch_size=B, num_epochs=epochs,
                            drop_remainder=drop, rank=r, world=world) for r in range(world)]
    nw = L - 1 - K + 1
    total = files * epochs * nw
    seen = []
    for r, it in enumerate(its):
      for lo, hi in it.batch_ranges():
        b = lo // (B * world)
        assert lo == b * B * world + r * B and lo < hi <= lo + B          # rank_slice of global batch b
        cover = 0
        for e, w0, cnt in it.pieces(lo, hi):
          assert 0 <= w0 and cnt >= 1 and w0 + cnt <= nw and e * nw + w0 == lo + cover
          cover += cnt
        assert cover == hi - lo
        seen += list(range(lo, hi))
    G = B * world
    kept = total if (world == 1 and not drop) else (total // G) * G
    assert sorted(seen) == list(range(kept))
    assert len({len(it) for it in its}) == 1                                # ranks step together

  prop()


@pytest.mark.parametrize('fmt', ['uint8', 'float32'])
def test_device_resident_frames_equal_host_batches(dataset, fmt):
  """device mode (frames uploaded once per episode, windows gathered by index on the device) yields the same batches
  as the host path; exercised here with device='cpu' -- the same torch code path, minus the DMA."""
  import torch
  d, eps = dataset
  kw = dict(window_size=4, fetch_target=True, batch_size=5, frame_format=fmt, num_epochs=2)
  host = list(ip.pickplace_input_fn_v4(d, 'default', 'eval', **kw))
  it = ip.pickplace_input_fn_v4(d, 'default', 'eval', device='cpu', **kw)
  got = list(it)
  assert len(got) == len(host) == 6                                      # 2 episodes x 2 epochs x 7 windows in fives
  for (hf, hl), (df, dl) in zip(host, got):
    assert set(hf) == set(df) and set(hl) == set(dl)
    for k in hf:
      if k in ip.BULK_KEYS + ip.TARGET_KEYS:
        assert torch.is_tensor(df[k]) and df[k].is_contiguous()
        np.testing.assert_array_equal(df[k].numpy(), hf[k], err_msg=k)
        assert df[k].numpy().dtype == hf[k].dtype
      else:
        np.testing.assert_array_equal(df[k], hf[k], err_msg=k)
    for k in hl:
      np.testing.assert_array_equal(dl[k], hl[k])
  # each (episode, epoch) was uploaded exactly once: 4 uploads of rgb + depth + targets, not one per window
  px = 10 * 8 * 6                                                        # S = 10 frames of 8 x 6 pixels
  per_episode = px * 3 * (1 if fmt == 'uint8' else 4) * (1 + 0.1) + px * 4 * (1 + 0.1)
  assert it.uploaded_bytes == int(round(4 * per_episode))
  assert len(it._resident) == 0                                          # released when the iterator ends


def test_device_resident_frames_rank_shards(dataset):
  d, eps = dataset
  whole = list(ip.pickplace_input_fn_v4(d, 'default', 'eval', 4, True, batch_size=4, drop_remainder=True, frame_format='uint8'))
  for r in range(2):
    part = list(ip.pickplace_input_fn_v4(d, 'default', 'eval', 4, True, batch_size=2, rank=r, world=2, frame_format='uint8',
                                         device='cpu'))
    assert len(part) == len(whole)
    for gb, pb in zip(whole, part):
      np.testing.assert_array_equal(pb[0]['rgb'].numpy(), gb[0]['rgb'][r * 2:(r + 1) * 2])
      np.testing.assert_array_equal(pb[0]['target_rgb'].numpy(), gb[0]['target_rgb'][r * 2:(r + 1) * 2])
      np.testing.assert_array_equal(pb[0]['jnt_state'], gb[0]['jnt_state'][r * 2:(r + 1) * 2])


class _StubEngine(object):
  """What Estimator.train / evaluate need from an Engine, on the CPU: staging is the identity, a 'train step' folds
  the batch into a checksum so the test can tell exactly which rows were fed, in which order."""
  training = True

  def __init__(self, batch):
    import torch
    self.N, self.global_step, self.fed = batch, 0, []
    self._torch = torch

  def stage(self, features, labels, slot):
    return features, labels, None

  def wait_staged(self, event, slot):
    pass

  def release_staged(self, slot):
    pass

  def _vector(self, features, labels):
    rgb = np.asarray(features['rgb'])
    self.fed.append((np.asarray(features['step'])[:, -1].tolist(), int(rgb.astype(np.int64).sum()), rgb.dtype.name,
                     np.asarray(labels['cmd'])[:, 0].tolist()))
    n = rgb.shape[0]
    return self._torch.tensor([1.0, 0.0, 2.0, 3.0, 0.0, 10.0 + len(self.fed), float(n), float(n)])

  def train_step(self, features, labels, grad_scale=1.0):
    self.global_step += 1
    return self._vector(features, labels)

  def forward(self, features, labels=None, want_dyn=False):
    return {'losses': self._vector(features, labels)}

  def read_losses_async(self, losses, slot):
    class Done:
      def synchronize(self):
        pass
    return losses, Done()

  def losses_dict(self, losses=None):
    return dict(zip(('loss_cmd_ee', 'loss_cmd_grp', 'loss_pos_ee', 'loss_pos_obj', 'loss_reg', 'loss'),
                    [float(x) for x in losses[:6]]))


def test_estimator_consumes_the_pipeline(dataset, tmp_path):
  """Estimator.train / evaluate over WindowBatches (host logic only, stub engine): every full batch is fed once, in
  stream order, as recorded bytes; a partial batch is refused like the reference's static batch size would."""
  from geeco_b200 import create_e2evmc_config
  from geeco_b200.estimator import Estimator, RunConfig, goal_e2evmc_model_fn
  d, eps = dataset
  cfg = create_e2evmc_config(dict(proc_obs='dynimg', proc_tgt='dyndiff', batch_size=4, img_height=8, img_width=6))
  est = Estimator(goal_e2evmc_model_fn, str(tmp_path), RunConfig(save_checkpoints_steps=0), {'e2evmc_config': cfg, 'log_steps': 1,
                                                                                              'save_final_checkpoint': False},
                  batch_size=4)
  est._engine = _StubEngine(4)
  inp = lambda **kw: (lambda: ip.pickplace_input_fn_v4(d, 'default', 'eval', 4, True, batch_size=4, frame_format='uint8', **kw))
  est.train(inp(drop_remainder=True))
  meta = ip.get_meta_v4(d)
  ref = _reference_stream(eps[3:], meta, 4, True)                     # 14 windows -> 3 full batches
  assert est.engine.global_step == 3 and len(est.last_train_losses) == 3
  for b, (steps, pixel_sum, dtype, cmd0) in enumerate(est.engine.fed):
    rows = ref[4 * b:4 * b + 4]
    assert steps == [int(r[0]['step'][-1]) for r in rows] and dtype == 'uint8'
    assert pixel_sum == int(round(sum(float(r[0]['rgb'].astype(np.float64).sum()) for r in rows) * 255.0))
    assert cmd0 == [float(r[1]['cmd'][0]) for r in rows]
  assert [round(l['loss']) for _, l in est.last_train_losses] == [11, 12, 13]
  ev = est.evaluate(inp(drop_remainder=True))
  assert ev['cmd_ee'] == 1.0 and ev['pos_obj'] == 3.0 and ev['cmd_grp'] == 1.0 and ev['global_step'] == 3
  with pytest.raises(ValueError, match='batch of 2 rows'):            # the trailing partial batch (14 = 3 * 4 + 2)
    est.train(inp())
