"""Synthetic (features, labels) batches in the layout the reference's input pipeline emits.

The reference's `pickplace_input_fn_v4` (src/data/geeco_gym.py:401-474) is a tf.data
pipeline over zlib TFRecords; datasets are not available offline and the pipeline itself is
outside the hot path, but its *layout and index contract* is the
input boundary of the train step:

  features: step [N,K] i64, rgb [N,K,H,W,3] f32 in [0,1] (stored uint8/255, geeco_gym.py:310),
            jnt_state [N,K,7], ee_state [N,K,7], obj_state [N,K,7], target_rgb [N,H,W,3]
  labels:   cmd [N,4] = cmd of the last frame of the window (geeco_gym.py:394)

Index contract (geeco_gym.py:598-631): episode length L, S = L-1 usable frames, window w in
[0, S-K] covers frames w..w+K-1, target = frame L-1, stream position g = e*(S-K+1) + w, batches
are consecutive slices of the stream (no window-level shuffle, :447-448).
"""
from __future__ import annotations

import numpy as np

EPISODE_LENGTH = 100   # src/geeco_gym/pickplace.py:157


def num_windows(episode_length: int, window_size: int) -> int:
  """_window_v3 (geeco_gym.py:617-618): (L-1) - K + 1."""
  return episode_length - 1 - window_size + 1


def window_frame_indices(episode_length: int, window_size: int) -> np.ndarray:
  """int64 [num_windows, K]: frame indices of every sliding window of one episode."""
  nw = num_windows(episode_length, window_size)
  return np.arange(nw, dtype=np.int64).reshape(nw, 1) + np.arange(window_size, dtype=np.int64).reshape(1, -1)


def locate(g: int, episode_length: int, window_size: int):
  """Stream position -> (episode, window, cur_frame, target_frame)."""
  e, w = divmod(int(g), num_windows(episode_length, window_size))
  return e, w, w + window_size - 1, episode_length - 1


def rank_slice(batch_index: int, global_batch: int, rank: int, world: int):
  """Contiguous share of global batch `batch_index` owned by `rank`:
  stream positions [b*G + r*G/n, b*G + (r+1)*G/n)."""
  if global_batch % world:
    raise ValueError("global batch %d not divisible by world size %d" % (global_batch, world))
  per = global_batch // world
  lo = batch_index * global_batch + rank * per
  return lo, lo + per


def synthetic_batch(n, window_size=4, height=256, width=256, channels=3, seed=0, first_stream_pos=0,
                    structured=True, episode_length=EPISODE_LENGTH, frame_format='float32'):
  """One (features, labels) tuple of numpy arrays, seeded, uint8-quantised pixels.

  frame_format='float32': pixels divided by 255 as the reference's input pipeline hands them to model_fn
  (src/data/geeco_gym.py:310); frame_format='uint8': the recorded bytes, same values before that division.

  structured=True gives every window temporal structure (frame k = base image rolled by k
  pixels plus fresh noise) so K-frame buffers are never static -- the dynimg min/max
  normalisation is ill-conditioned on identical frames.
  """
  rng = np.random.default_rng(seed)
  K = window_size
  if structured:
    base = rng.integers(0, 256, size=(n, height, width, channels))
    frames = []
    for k in range(K):
      noise = rng.integers(-24, 25, size=(n, height, width, channels))
      frames.append(np.clip(np.roll(base, shift=(k, 2 * k), axis=(1, 2)) + noise, 0, 255))
    rgb = np.stack(frames, axis=1)
  else:
    rgb = rng.integers(0, 256, size=(n, K, height, width, channels))
  tgt = rng.integers(0, 256, size=(n, height, width, channels))
  if frame_format == 'uint8':
    rgb, tgt = rgb.astype(np.uint8), tgt.astype(np.uint8)
  elif frame_format == 'float32':
    rgb, tgt = (rgb / 255.0).astype(np.float32), (tgt / 255.0).astype(np.float32)
  else:
    raise ValueError("frame_format must be 'float32' or 'uint8', got %r" % (frame_format,))
  jnt = rng.uniform(-np.pi, np.pi, size=(n, K, 7)).astype(np.float32)
  ee = np.zeros((n, K, 7), dtype=np.float32)
  ee[..., :3] = np.array([1.34, 0.75, 0.55], dtype=np.float32) + rng.uniform(-0.15, 0.15, size=(n, K, 3))
  ee[..., 3:] = np.array([1.0, 0.0, 1.0, 0.0], dtype=np.float32)
  obj = np.zeros((n, K, 7), dtype=np.float32)
  obj[..., 0] = rng.uniform(1.075, 1.425, size=(n, K))
  obj[..., 1] = rng.uniform(0.35, 1.15, size=(n, K))
  obj[..., 2] = 0.307
  obj[..., 3] = 1.0
  step = np.empty((n, K), dtype=np.int64)
  for i in range(n):
    _, w, _, _ = locate(first_stream_pos + i, episode_length, K)
    step[i] = w + np.arange(K)
  cmd = np.empty((n, 4), dtype=np.float32)
  cmd[:, :3] = rng.uniform(-2.0, 2.0, size=(n, 3))
  cmd[:, 3] = rng.integers(-1, 2, size=n)
  features = {'step': step, 'rgb': rgb, 'jnt_state': jnt, 'ee_state': ee, 'obj_state': obj, 'target_rgb': tgt}
  labels = {'cmd': cmd}
  # labels of --control_mode velocity (geeco_gym.py:392-398), from their own generator so that the values above
  # do not depend on them
  rv = np.random.default_rng([int(seed), 7919])
  labels['vel_target'] = rv.uniform(-1.0, 1.0, size=(n, 7)).astype(np.float32)
  ee_t = np.zeros((n, 7), dtype=np.float32)
  ee_t[:, :3] = np.array([1.34, 0.75, 0.55], dtype=np.float32) + rv.uniform(-0.15, 0.15, size=(n, 3))
  ee_t[:, 3:] = np.array([1.0, 0.0, 1.0, 0.0], dtype=np.float32)
  labels['ee_target'] = ee_t
  labels['grp_target'] = rv.uniform(0.0, 0.05, size=(n, 2)).astype(np.float32)
  return features, labels


def to_pool_layout(features, pieces=2):
  """Re-expresses a dense batch whose windows are CONSECUTIVE windows of `pieces` episodes (what the input pipeline
  emits: batches are consecutive stream positions, geeco_gym.py:471-473) in the frame-pool layout of
  include/geeco_b200.h: `rgb` [F,H,W,C] holds every distinct frame once, `rgb_index` [N,K] int32 addresses it,
  `target_rgb` [pieces,H,W,C] + `target_index` [N].  Only the layout changes; `synthetic_pool_batch` builds batches
  that really have this structure."""
  rgb, tgt = np.asarray(features['rgb']), np.asarray(features['target_rgb'])
  n, K = rgb.shape[:2]
  per = -(-n // pieces)
  frames, index, tix, targets = [], np.empty((n, K), np.int32), np.empty((n,), np.int32), []
  base = 0
  for p in range(pieces):
    lo, hi = p * per, min(n, (p + 1) * per)
    if hi <= lo:
      break
    cnt = hi - lo
    # window i of the piece = frames i .. i+K-1 of the piece's frame run
    run = [rgb[lo, k] for k in range(K - 1)] + [rgb[lo + i, K - 1] for i in range(cnt)]
    frames += run
    index[lo:hi] = base + np.arange(cnt, dtype=np.int32)[:, None] + np.arange(K, dtype=np.int32)[None, :]
    targets.append(tgt[lo])
    tix[lo:hi] = p
    base += len(run)
  out = dict(features)
  out['rgb'], out['rgb_index'] = np.stack(frames), index
  out['target_rgb'], out['target_index'] = np.stack(targets), tix
  return out


def expand_pool_layout(features):
  """The dense [N,K,H,W,C] / [N,H,W,C] tensors a frame-pool batch stands for."""
  out = {k: v for k, v in features.items() if k not in ('rgb_index', 'target_index')}
  out['rgb'] = np.asarray(features['rgb'])[np.asarray(features['rgb_index'])]
  if 'target_index' in features:
    out['target_rgb'] = np.asarray(features['target_rgb'])[np.asarray(features['target_index'])]
  return out


def synthetic_pool_batch(n, window_size=4, pieces=2, seed=0, frame_format='uint8', **kw):
  """A synthetic batch of `n` windows that are consecutive windows of `pieces` episodes, in the frame-pool layout
  (see to_pool_layout).  Returns (features, labels); `expand_pool_layout(features)` gives the equivalent dense batch."""
  feats, labels = synthetic_batch(n, window_size=window_size, seed=seed, frame_format=frame_format, **kw)
  return to_pool_layout(feats, pieces), labels


# ------------------------------------------------------------------------------------------------
# synthetic datasets on disk, in the recorder's format (src/data/data_recorder.py, PickAndPlaceEncodingV4)
# ------------------------------------------------------------------------------------------------
def synthetic_episode(episode_length=EPISODE_LENGTH, height=256, width=256, seed=0, monitored_joints=None):
  """One episode as the dict `TfrSequenceEncoding.encode` takes (context keys + 'sequence' list of frames), with
  the frame keys of PickAndPlaceEncodingV4 (geeco_gym.py:105-113).  Pixels are uint8 arrays, stored by the
  recorder as floats 0..255 (utils/tfrecord.py:75-76)."""
  from .input_pipeline import ARM_JOINTS, FINGER_JOINTS, MOCAP
  rng = np.random.default_rng(seed)
  joints = list(monitored_joints or (ARM_JOINTS + FINGER_JOINTS))
  context = {
      'episode_length': int(episode_length), 'img_height': int(height), 'img_width': int(width),
      'monitored_joints': joints, 'actuated_joints': list(FINGER_JOINTS), 'monitored_mocaps': [MOCAP],
      'monitored_objects': ['object0:joint'], 'dim_cmd': 4, 'dim_ctrl': 2, 'task_goal': 'pad2',
      'task_object': 'cube2'}
  base = rng.integers(0, 256, size=(height, width, 3))
  frames = []
  for t in range(episode_length):
    rgb = np.clip(np.roll(base, shift=(t, 2 * t), axis=(0, 1)) + rng.integers(-24, 25, size=base.shape), 0, 255)
    frame = {
        'step': t, 'ts': float(np.float32(0.04 * t)), 'rgb': rgb.astype(np.uint8),
        'depth': rng.uniform(0.5, 2.0, size=(height, width, 1)).astype(np.float32),
        'cmd': np.concatenate([rng.uniform(-2.0, 2.0, size=3), [float(rng.integers(-1, 2))]]).astype(np.float32),
        'ctrl': rng.uniform(-1.0, 1.0, size=2).astype(np.float32),
        'goal_qpos': rng.uniform(0.3, 1.5, size=7).astype(np.float32),
        'obj_qpos': rng.uniform(0.3, 1.5, size=7).astype(np.float32),
        'mocap_qpos-%s' % MOCAP: rng.uniform(0.3, 1.5, size=7).astype(np.float32),
        'object_qpos-object0:joint': rng.uniform(0.3, 1.5, size=7).astype(np.float32),
    }
    for j in joints:
      frame['joint_qpos-%s' % j] = float(rng.uniform(-np.pi, np.pi))
      frame['joint_qvel-%s' % j] = float(rng.uniform(-1.0, 1.0))
    frames.append(frame)
  data = dict(context)
  data['sequence'] = frames
  return data


def encoding_keys_v4(data):
  """(context_keys, frame_keys) of PickAndPlaceEncodingV4.__init__ (geeco_gym.py:102-113) for one episode dict."""
  context_keys = ['episode_length', 'img_height', 'img_width', 'monitored_joints', 'actuated_joints',
                  'monitored_mocaps', 'monitored_objects', 'dim_cmd', 'dim_ctrl', 'task_goal', 'task_object']
  frame_keys = ['step', 'ts', 'rgb', 'depth', 'cmd', 'ctrl', 'goal_qpos', 'obj_qpos']
  for j in data['monitored_joints']:
    frame_keys += ['joint_qpos-%s' % j, 'joint_qvel-%s' % j]
  frame_keys += ['mocap_qpos-%s' % m for m in data['monitored_mocaps']]
  frame_keys += ['object_qpos-%s' % o for o in data['monitored_objects']]
  return context_keys, frame_keys


def write_synthetic_dataset(dataset_dir, episodes=2, episode_length=EPISODE_LENGTH, height=256, width=256, seed=0,
                            split_name='default', eval_episodes=1):
  """Writes a dataset with the directory layout pickplace_input_fn_v4 expects (geeco_gym.py:414-431):
  meta/meta_info.json, data/<name>.tfrecord.zlib (one episode each), splits/<split>/{train,eval}.txt.
  Returns the list of episode dicts (train episodes first)."""
  import json
  import os
  from .tfrecord import encode_sequence_example, write_tfrecord
  os.makedirs(os.path.join(dataset_dir, 'meta'), exist_ok=True)
  os.makedirs(os.path.join(dataset_dir, 'data'), exist_ok=True)
  os.makedirs(os.path.join(dataset_dir, 'splits', split_name), exist_ok=True)
  all_eps, names = [], []
  for e in range(episodes + eval_episodes):
    data = synthetic_episode(episode_length, height, width, seed=seed * 1000 + e)
    ck, fk = encoding_keys_v4(data)
    name = '%06d.tfrecord.zlib' % e
    write_tfrecord(os.path.join(dataset_dir, 'data', name), [encode_sequence_example(data, ck, fk)], 'zlib')
    all_eps.append(data)
    names.append(name)
  meta = {k: all_eps[0][k] for k in ('episode_length', 'img_height', 'img_width', 'monitored_joints',
                                     'actuated_joints', 'monitored_mocaps', 'monitored_objects', 'dim_cmd', 'dim_ctrl')}
  with open(os.path.join(dataset_dir, 'meta', 'meta_info.json'), 'w') as fp:
    json.dump(meta, fp)
  for mode, part in (('train', names[:episodes]), ('eval', names[episodes:]), ('test', names[episodes:])):
    with open(os.path.join(dataset_dir, 'splits', split_name, '%s.txt' % mode), 'w') as fp:
      fp.write('\n'.join(part) + '\n')
  return all_eps
