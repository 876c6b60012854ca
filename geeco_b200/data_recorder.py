"""Recorder side of the dataset format: structured frames -> one tf.train.SequenceExample per episode -> TFRecord.

Counterpart of src/data/data_recorder.py (`TfrSequenceEncoding` :17-66, `TfrSequenceRecorder` :71-156) and of
`PickAndPlaceEncodingV4` (src/data/geeco_gym.py:54-162) without TensorFlow: `encode` returns the serialized
SequenceExample (what `example.SerializeToString()` yields there), `finalize` writes `<name>.tfrecord[.zlib|.gzip]`
through libgeeco_io.so.  Same method names, arguments, key checks and error types.
"""
from __future__ import annotations

import datetime
import os
import time

from .tfrecord import encode_sequence_example, write_tfrecord


class TfrSequenceEncoding(object):
  """Base class: context keys + frame keys of one episode encoding (data_recorder.py:17-66)."""

  def __init__(self):
    self._context_keys = []
    self._frame_keys = []

  @property
  def context_keys(self):
    return self._context_keys

  @property
  def frame_keys(self):
    return self._frame_keys

  def encode(self, data: dict) -> bytes:
    """data[k] for every context key, data['sequence'] = list of frame dicts holding every frame key
    (missing keys raise KeyError, unsupported value types TypeError, as in the reference)."""
    return encode_sequence_example(data, self._context_keys, self._frame_keys)

  def decode(self):
    raise NotImplementedError()


class PickAndPlaceEncodingV4(TfrSequenceEncoding):
  """Key layout of geeco_gym.py:93-115 for a `PickAndPlaceMetaV4`."""

  def __init__(self, meta):
    super().__init__()
    self._context_keys = list(meta._asdict().keys()) + ['task_goal', 'task_object']
    frame_keys = ['step', 'ts', 'rgb', 'depth', 'cmd', 'ctrl', 'goal_qpos', 'obj_qpos']
    for jnt_name in meta.monitored_joints:
      frame_keys.append('joint_qpos-%s' % (jnt_name,))
      frame_keys.append('joint_qvel-%s' % (jnt_name,))
    for mcp_name in meta.monitored_mocaps:
      frame_keys.append('mocap_qpos-%s' % (mcp_name,))
    for obj_jnt_name in meta.monitored_objects:
      frame_keys.append('object_qpos-%s' % (obj_jnt_name,))
    self._frame_keys = frame_keys
    self._meta = meta

  def decode(self):
    """The reference returns tf.FixedLen(Sequence)Feature specs here (geeco_gym.py:117-162); the same shape
    contract as plain data: ({context key: (kind, count)}, {frame key: (kind, values per frame)})."""
    m = self._meta
    ctx = {'episode_length': ('int64', 1), 'img_height': ('int64', 1), 'img_width': ('int64', 1),
           'monitored_joints': ('bytes', len(m.monitored_joints)), 'actuated_joints': ('bytes', len(m.actuated_joints)),
           'monitored_mocaps': ('bytes', len(m.monitored_mocaps)), 'monitored_objects': ('bytes', len(m.monitored_objects)),
           'dim_cmd': ('int64', 1), 'dim_ctrl': ('int64', 1), 'task_goal': ('bytes', 1), 'task_object': ('bytes', 1)}
    seq = {'step': ('int64', 1), 'ts': ('float', 1), 'rgb': ('float', m.img_height * m.img_width * 3),
           'depth': ('float', m.img_height * m.img_width), 'cmd': ('float', m.dim_cmd), 'ctrl': ('float', m.dim_ctrl),
           'obj_qpos': ('float', 7), 'goal_qpos': ('float', 7)}
    for k in self._frame_keys:
      if k.startswith('joint_qpos-') or k.startswith('joint_qvel-'):
        seq[k] = ('float', 1)
      elif k.startswith('mocap_qpos-') or k.startswith('object_qpos-'):
        seq[k] = ('float', 7)
    return ctx, seq


class TfrSequenceRecorder(object):
  """Buffers the frames of one episode and writes them as one TFRecord file (data_recorder.py:71-156)."""

  def __init__(self, encoding: TfrSequenceEncoding, sequence_context: dict, record_dir, record_name: str = None):
    self._encoding = encoding
    self._sequence_context = sequence_context
    self._sequence_frames = []
    self._record_dir = record_dir
    if record_name is None:
      record_name = datetime.datetime.fromtimestamp(time.time()).strftime('%Y%m%d_%H%M%S')
    self._record_name = record_name
    self._record_path = self._get_record_path()

  @property
  def record_name(self):
    return self._record_name

  @property
  def record_path(self):
    return self._record_path

  def _get_record_path(self):
    return os.path.join(self._record_dir, "%s.tfrecord" % (self._record_name,))

  def _has_valid_format(self, frame):
    return set(frame.keys()) == set(self._encoding.frame_keys)

  def feed(self, frame):
    """Appends a data frame; its keys must be exactly the encoding's frame keys (ValueError otherwise)."""
    if not self._has_valid_format(frame):
      raise ValueError("The given frame does not match the expected data fields!\n"
                       "Given data fields: %s\nExpected data fields: %s"
                       % (set(frame.keys()), set(self._encoding.frame_keys)))
    self._sequence_frames.append(frame)

  def finalize(self, compression='none'):
    """Encodes context + frames and writes the record; compression: none | gzip | zlib (suffix appended)."""
    if compression not in ('none', 'gzip', 'zlib'):
      raise KeyError(compression)
    data = {}
    data.update(self._sequence_context)
    data['sequence'] = self._sequence_frames
    path = self._get_record_path()
    if compression != 'none':
      path = path + '.%s' % (compression,)
    write_tfrecord(path, [self._encoding.encode(data)], compression)
    return path
