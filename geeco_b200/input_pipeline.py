"""Native input pipeline: recorded GEECO-gym episodes -> (features, labels) batches for the train step.

Counterpart of `pickplace_input_fn_v4` (src/data/geeco_gym.py:401-474): the same dataset layout
(`meta/meta_info.json`, `data/*.tfrecord.zlib`, `splits/<split>/<mode>.txt`), the same stages and the same
output -- only executed differently.  The reference chains tf.data ops on 4 CPU threads and materialises every
episode as float32 (78 MB of pixels per 100-frame episode); here

  * a worker pool decodes whole episodes in C++ (libgeeco_io.so: inflate, TFRecord framing + CRC-32C,
    SequenceExample index, bulk value reads; ctypes releases the GIL so the workers run in parallel),
  * pixels come back as the recorded uint8 values (`frame_format='uint8'`, 4x fewer bytes over PCIe; the
    `/= 255.0` of geeco_gym.py:310 then runs inside the rank-pooling kernel and yields bit-identical network
    input) or as float32 already divided by 255 (`frame_format='float32'`, the reference's exact tensors),
  * windows of consecutive stream positions are gathered straight into the batch buffers (`pin_memory=True`:
    page-locked torch tensors that the engine uploads without a staging copy),
  * batches are assembled `prefetch_size` ahead by a background thread,
  * `cache_dir` keeps every decoded episode as uncompressed arrays (uint8 pixels: 19.7 MB instead of the 100 MB
    protobuf per 100-frame episode), so only the first epoch pays for inflating the recorder's float encoding
    (a single zlib stream per file: ~0.6 s per episode and thread, the format's own limit).

Stage by stage (all file:line into src/data/geeco_gym.py):
  _parse_v4 :291-316            reshape rgb [L,H,W,3], depth [L,H,W,1]; rgb /= 255; target = last frame
  _preprocess_states_v4 :318-371  jnt_state / vel_state = the 7 arm joints in the order of :339-345 / :353-359,
                                grp_state = the two finger joints, ee_state = mocap_qpos-robot0:mocap,
                                goal_state = goal_qpos, obj_state = obj_qpos
  _preprocess_targets_v3 :598-612 vel/ee/grp targets = next frame's state (roll by -1), then every sequence
                                loses its last frame (targets stay): S = L - 1
  _window_v3 :614-631           num_windows = S - K + 1 windows of K consecutive frames, targets tiled
  unbatch + _prepare_v4 :373-399  one example per window; labels = values of the window's last frame
  repeat(num_epochs).batch(batch_size).prefetch(prefetch_size) :470-473
Train mode shuffles the episode order only (`np.random.shuffle(tfrecord_paths)`, :436-437); windows keep
their order inside an episode and every file holds one episode, so the stream position of window w of the
e-th file is g = e * num_windows + w (geeco_b200.data.locate).

Device-resident frames (`device='cuda'`): consecutive windows share K-1 of their K frames, so a batch of B windows
holds only about B + K - 1 distinct frames.  In this mode every episode's frames are uploaded ONCE (pinned host ->
device, 19.4 MB of uint8 per 100-frame episode instead of 77 MB of windows), stay resident while batches read them,
and the [B,K,H,W,C] window tensor is gathered on the device (an index-select: data movement only); the host never
copies a pixel per batch.  Features `rgb` / `depth` / `target_*` are then device tensors, which the engine uses in
place; everything else is unchanged.

Data parallel: with `world > 1` a global batch is `batch_size * world` consecutive stream
positions and rank r takes rows [r * batch_size, (r+1) * batch_size); each rank decodes only the episodes
its rows touch.  Incomplete trailing global batches are dropped there (ranks must step together).
"""
from __future__ import annotations

import collections
import concurrent.futures
import json
import os
import queue
import threading

import numpy as np

from .tfrecord import SequenceExample, TFRecordFile, window_gather

PickAndPlaceMetaV4 = collections.namedtuple('PickAndPlaceMetaV4', [
    'episode_length', 'img_height', 'img_width', 'monitored_joints', 'actuated_joints', 'monitored_mocaps',
    'monitored_objects', 'dim_cmd', 'dim_ctrl'])           # geeco_gym.py:34-49

ARM_JOINTS = ('robot0:shoulder_pan_joint', 'robot0:shoulder_lift_joint', 'robot0:upperarm_roll_joint',
              'robot0:elbow_flex_joint', 'robot0:forearm_roll_joint', 'robot0:wrist_flex_joint',
              'robot0:wrist_roll_joint')                     # geeco_gym.py:339-345
FINGER_JOINTS = ('robot0:l_gripper_finger_joint', 'robot0:r_gripper_finger_joint')   # :366-367
MOCAP = 'robot0:mocap'                                     # :331

FEATURE_KEYS = ('step', 'ts', 'rgb', 'depth', 'jnt_state', 'vel_state', 'ee_state', 'grp_state', 'goal_state',
                'obj_state', 'cmd', 'ctrl')                # _prepare_v4 :375-388
TARGET_KEYS = ('target_rgb', 'target_depth')               # :389-391
LABEL_KEYS = ('cmd', 'ctrl', 'vel_target', 'ee_target', 'grp_target')   # :392-398
BULK_KEYS = ('rgb', 'depth')                               # per-frame images: kept on the device in device mode


def get_meta_v4(dataset_dir):
  """_get_meta_v4 (geeco_gym.py:283-289)."""
  with open(os.path.join(dataset_dir, 'meta', 'meta_info.json'), 'r') as fp:
    return PickAndPlaceMetaV4(**json.load(fp))


def collect_tfrecords_v2(data_dir, split_name, mode):
  """_collect_tfrecords_v2 (geeco_gym.py:780-793): the split file's entries, or every record of the dataset when
  split_name and mode are None (directory order, as os.listdir returns it)."""
  record_dir = os.path.join(data_dir, 'data')
  if split_name is None and mode is None:
    record_files = [fn for fn in os.listdir(record_dir) if fn.endswith('.tfrecord.zlib')]
  else:
    with open(os.path.join(data_dir, 'splits', split_name, '%s.txt' % (mode,))) as fp:
      record_files = fp.read().split('\n')
  return [os.path.join(record_dir, fn) for fn in record_files if fn.endswith('.tfrecord.zlib')]


def _load_rgbd(rgb_path, depth_path):
  from PIL import Image
  rgb = np.array(Image.open(rgb_path), dtype=np.float32) / 255.0
  if depth_path is None:
    return rgb
  return np.concatenate([rgb, np.expand_dims(np.load(depth_path), axis=-1)], axis=-1)


def load_target_frame(dataset_dir, tfrecord_name, load_depth=True):
  """Target frame of an episode as float32 RGB(-D) in [0,1] (geeco_gym.py:184-198):
  `images/targets/rgb/<name>.png` (+ `images/targets/depth/<name>.npy`)."""
  filename = os.path.basename(tfrecord_name).split('.')[0]
  rgb_path = os.path.join(dataset_dir, 'images', 'targets', 'rgb', filename + '.png')
  depth_path = os.path.join(dataset_dir, 'images', 'targets', 'depth', filename + '.npy') if load_depth else None
  frame = _load_rgbd(rgb_path, depth_path)
  print("Read %s" % filename)
  return frame


def load_keyframes(dataset_dir, tfrecord_name):
  """All key frames `images/keyframes/{rgb,depth}/<name>*` of an episode as RGB-D tensors, sorted by file name
  (geeco_gym.py:200-217)."""
  filename = os.path.basename(tfrecord_name).split('.')[0]
  rgb_dir = os.path.join(dataset_dir, 'images', 'keyframes', 'rgb')
  depth_dir = os.path.join(dataset_dir, 'images', 'keyframes', 'depth')
  rgb_files = sorted(f for f in os.listdir(rgb_dir) if f.startswith(filename))
  depth_files = sorted(f for f in os.listdir(depth_dir) if f.startswith(filename))
  return [_load_rgbd(os.path.join(rgb_dir, r), os.path.join(depth_dir, d)) for r, d in zip(rgb_files, depth_files)]


def load_target_frames(dataset_dir, tfrecord_name, load_depth=True):
  """Key frames when `data/key_frames_<id>.json` exists, else the single target frame (geeco_gym.py:168-182)."""
  import re
  record_id = re.search(r'\d+', tfrecord_name).group(0)
  if os.path.exists(os.path.join(dataset_dir, 'data', 'key_frames_%s.json' % (record_id,))):
    return load_keyframes(dataset_dir, tfrecord_name)
  return [load_target_frame(dataset_dir, tfrecord_name, load_depth)]


def decode_episode(path, meta, fetch_target=True, frame_format='uint8', verify_crc=True, want_depth=True):
  """One episode file -> dict of per-frame arrays after _parse_v4, _preprocess_states_v4 and
  _preprocess_targets_v3 (every sequence has S = L - 1 frames; targets are single frames).

  `rgb` is uint8 [S,H,W,3] (`frame_format='uint8'`) or float32 in [0,1]; with 'uint8' a file whose pixel floats
  are not exact bytes is rejected rather than silently quantised."""
  if frame_format not in ('uint8', 'float32'):
    raise ValueError("frame_format must be 'float32' or 'uint8', got %r" % (frame_format,))
  H, W = int(meta.img_height), int(meta.img_width)
  with TFRecordFile(path, compression='zlib', verify_crc=verify_crc) as rec:
    if len(rec) != 1:
      raise ValueError("%s holds %d records, expected one episode" % (path, len(rec)))
    addr, n = rec.record_view(0)
    ex = SequenceExample(address=addr, length=n, keepalive=rec)
    try:
      seq = {}
      seq['step'] = ex.ints('step', per_frame=1)[:, 0]
      L = seq['step'].shape[0]
      seq['ts'] = ex.floats('ts', per_frame=1)[:, 0]
      rgb8, inexact = ex.pixel_bytes('rgb', per_frame=H * W * 3)
      if rgb8.shape[0] != L:
        raise ValueError("%s: rgb has %d frames, step has %d" % (path, rgb8.shape[0], L))
      if frame_format == 'uint8':
        if inexact:
          raise ValueError("%s: %d rgb values are not bytes; read it with frame_format='float32'" % (path, inexact))
        seq['rgb'] = rgb8.reshape(L, H, W, 3)
      elif inexact == 0:
        # fp32 division of exact bytes == the reference's `rgb /= 255.0` on the float-decoded pixels
        seq['rgb'] = rgb8.reshape(L, H, W, 3).astype(np.float32) / np.float32(255.0)
      else:
        seq['rgb'] = ex.floats('rgb', per_frame=H * W * 3).reshape(L, H, W, 3) / np.float32(255.0)
      if want_depth:
        seq['depth'] = ex.floats('depth', per_frame=H * W).reshape(L, H, W, 1)
      seq['cmd'] = ex.floats('cmd', per_frame=int(meta.dim_cmd))
      seq['ctrl'] = ex.floats('ctrl', per_frame=int(meta.dim_ctrl))
      seq['jnt_state'] = np.stack([ex.floats('joint_qpos-%s' % j, per_frame=1)[:, 0] for j in ARM_JOINTS], axis=1)
      seq['vel_state'] = np.stack([ex.floats('joint_qvel-%s' % j, per_frame=1)[:, 0] for j in ARM_JOINTS], axis=1)
      seq['grp_state'] = np.stack([ex.floats('joint_qpos-%s' % j, per_frame=1)[:, 0] for j in FINGER_JOINTS], axis=1)
      seq['ee_state'] = ex.floats('mocap_qpos-%s' % MOCAP, per_frame=7)
      seq['goal_state'] = ex.floats('goal_qpos', per_frame=7)
      seq['obj_state'] = ex.floats('obj_qpos', per_frame=7)
    finally:
      ex.close()
  for k, v in seq.items():
    if v.shape[0] != L:
      raise ValueError("%s: feature %s has %d frames, step has %d" % (path, k, v.shape[0], L))
  out = {}
  if fetch_target:                                          # _parse_v4 :312-315: the episode's last frame
    out['target_rgb'] = seq['rgb'][L - 1].copy()
    if want_depth:
      out['target_depth'] = seq['depth'][L - 1].copy()
  for name in ('vel', 'ee', 'grp'):                         # _preprocess_targets_v3 :600-606
    seq['%s_target' % name] = np.roll(seq['%s_state' % name], shift=-1, axis=0)
  for k, v in seq.items():                                  # :607-612 drop the last frame
    out[k] = v[:L - 1]
  return out


def _pinned_empty(shape, dtype):
  """Page-locked torch tensor (torch's caching host allocator keeps a block alive until the asynchronous copies
  that read it have run, which a bare numpy view of the same memory would not get)."""
  import torch
  return torch.empty(tuple(int(s) for s in shape), dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)


def _host_tensor(arr, pin):
  """One copy of an array (possibly a read-only memory map of the episode cache) into a torch host tensor,
  page-locked when it is going to be uploaded."""
  import torch
  t = torch.empty(tuple(arr.shape), dtype=getattr(torch, np.dtype(arr.dtype).name), pin_memory=bool(pin))
  np.copyto(t.numpy(), arr)
  return t


def _gather_frames(frames, idx):
  """frames[idx] along dim 0, moving 8-byte words when a frame's byte size allows it (a uint8 index_select moves one
  byte per element; 256x256x3 frames are 24576 words)."""
  import torch
  flat = frames.reshape(frames.shape[0], -1)
  row_bytes = flat.shape[1] * flat.element_size()
  if flat.element_size() < 8 and row_bytes % 8 == 0 and flat.shape[0] > 0:
    out = flat.view(torch.int64).index_select(0, idx).view(frames.dtype)
  else:
    out = flat.index_select(0, idx)
  return out.reshape((idx.shape[0],) + tuple(frames.shape[1:]))


class WindowBatches(object):
  """Iterable over (features, labels) batches of consecutive window stream positions (see module docstring)."""

  def __init__(self, tfrecord_paths, meta, window_size=4, fetch_target=False, batch_size=1, num_epochs=1,
               num_threads=4, prefetch_size=4, frame_format='float32', drop_remainder=False, rank=0, world=1,
               pin_memory=False, verify_crc=True, want_depth=True, cache_dir=None, device=None, layout='windows'):
    if window_size < 1 or window_size > meta.episode_length - 1:
      raise ValueError("window_size %d does not fit episodes of %d frames" % (window_size, meta.episode_length))
    if not 0 <= rank < world:
      raise ValueError("rank %d outside world of %d" % (rank, world))
    self.paths = list(tfrecord_paths)
    self.meta = meta
    self.K = int(window_size)
    self.fetch_target = bool(fetch_target)
    self.B = int(batch_size)
    self.epochs = int(num_epochs)
    self.threads = max(1, int(num_threads))
    self.prefetch = max(1, int(prefetch_size))
    self.frame_format = frame_format
    self.drop_remainder = bool(drop_remainder) or world > 1
    self.rank, self.world = int(rank), int(world)
    self.pin = bool(pin_memory)
    self.verify_crc = bool(verify_crc)
    self.want_depth = bool(want_depth)
    self.device = device                                   # None: host batches; 'cuda' / 'cpu': resident frames
    if layout not in ('windows', 'pool'):
      raise ValueError("layout must be 'windows' or 'pool', got %r" % (layout,))
    if layout == 'pool' and device is not None:
      raise ValueError("layout='pool' is a host-batch layout; device-resident frames already share frames between windows")
    # 'pool': the image features of a batch are the DISTINCT frames its windows touch ([F,H,W,C], F = windows + K - 1
    # per episode piece) plus an int32 index [B,K] into them, and one goal frame per piece with an index [B]: the
    # frame-pool layout of include/geeco_b200.h (geeco_batch.frame_index).  Consecutive windows share K-1 frames, so
    # the host copies and the upload move ~K times fewer bytes; the step reads the same pixels.
    self.layout = layout
    self._resident = collections.OrderedDict()             # stream episode -> {key: tensor on self.device}
    self._cuda_index = None                                # the consumer's CUDA device, adopted by the helper threads
    self.uploaded_bytes = 0
    self.cache_dir = cache_dir
    if cache_dir:
      os.makedirs(cache_dir, exist_ok=True)
    self.nw = meta.episode_length - 1 - self.K + 1          # _window_v3 :616-617
    self.total = len(self.paths) * self.epochs * self.nw    # windows in the whole stream

  # -- index math ---------------------------------------------------------------------------------
  def batch_ranges(self):
    """[(lo, hi)] stream-position ranges of this rank's batches, in order."""
    G = self.B * self.world
    out, b = [], 0
    while b * G < self.total:
      lo = b * G + self.rank * self.B
      hi = min(lo + self.B, self.total)
      if self.drop_remainder and (b + 1) * G > self.total:
        break
      if hi > lo:
        out.append((lo, hi))
      b += 1
    return out

  def pieces(self, lo, hi):
    """[(stream episode index, first window, count)] covering stream positions [lo, hi)."""
    out = []
    g = lo
    while g < hi:
      e, w = divmod(g, self.nw)
      n = min(hi - g, self.nw - w)
      out.append((e, w, n))
      g += n
    return out

  def __len__(self):
    return len(self.batch_ranges())

  # -- assembly -----------------------------------------------------------------------------------
  def _decode(self, stream_episode):
    path = self.paths[stream_episode % len(self.paths)]
    self._adopt_cuda_device()
    ep = self._cached(path)
    if ep['step'].shape[0] != self.meta.episode_length - 1:
      raise ValueError("%s holds %d frames but meta_info.json says episode_length=%d"
                       % (path, ep['step'].shape[0] + 1, self.meta.episode_length))
    if self.device is not None and str(self.device).startswith('cuda'):
      # page-locked copies, so that the one upload per episode is a DMA
      ep['_pinned'] = {k: _host_tensor(ep[k], pin=True) for k in BULK_KEYS + TARGET_KEYS if k in ep}
    return ep

  def _cached(self, path):
    """decode_episode, through the episode cache when one is configured.  A cache entry is keyed by the record's
    name, size and mtime and by what was decoded; it is a directory written under a temporary name and renamed, so
    concurrent ranks sharing a cache directory never read a partial entry."""
    args = (self.meta, self.fetch_target, self.frame_format, self.verify_crc, self.want_depth)
    if not self.cache_dir:
      return decode_episode(path, *args)
    st = os.stat(path)
    key = '%s.%d.%d.%s.t%d.d%d.ep' % (os.path.basename(path), st.st_size, int(st.st_mtime), self.frame_format,
                                      self.fetch_target, self.want_depth)
    entry = os.path.join(self.cache_dir, key)
    if os.path.isdir(entry):
      # one .npy per array, memory-mapped: nothing is copied or checksummed until a window (or the one upload
      # of device mode) reads it from the page cache
      return {f[:-4]: np.load(os.path.join(entry, f), mmap_mode='r') for f in os.listdir(entry) if f.endswith('.npy')}
    ep = decode_episode(path, *args)
    tmp = '%s.%d.%d.tmp' % (entry, os.getpid(), threading.get_ident())
    os.makedirs(tmp)
    for k, v in ep.items():
      np.save(os.path.join(tmp, k + '.npy'), v)
    try:
      os.rename(tmp, entry)
    except OSError:                                         # another rank / thread finished the same entry first
      import shutil
      shutil.rmtree(tmp, ignore_errors=True)
    return ep

  def _alloc(self, shape, dtype):
    if not self.pin:
      return np.empty(shape, dtype=dtype)
    t = _pinned_empty(shape, dtype)
    self._pinned.append(t)
    return t.numpy()

  def _assemble(self, lo, hi, episodes):
    n = hi - lo
    feats, labels = {}, {}
    self._pinned = []
    at = 0
    on_device = self.device is not None
    pool = self.layout == 'pool'
    if pool:
      pieces = self.pieces(lo, hi)
      F = sum(cnt + self.K - 1 for _, _, cnt in pieces)
      feats['rgb_index'] = self._alloc((n, self.K), np.int32)
      if self.fetch_target:
        feats['target_index'] = self._alloc((n,), np.int32)
      fat = 0
    for pi, (e, w0, cnt) in enumerate(self.pieces(lo, hi)):
      ep = episodes[e]
      if pool:
        nf = cnt + self.K - 1
        for k in BULK_KEYS:
          if k not in ep:
            continue
          if k not in feats:
            feats[k] = self._alloc((F,) + ep[k].shape[1:], ep[k].dtype)
          feats[k][fat:fat + nf] = ep[k][w0:w0 + nf]        # the frames windows w0 .. w0+cnt-1 touch, once each
        feats['rgb_index'][at:at + cnt] = fat + np.arange(cnt, dtype=np.int32)[:, None] + np.arange(self.K, dtype=np.int32)[None, :]
        if self.fetch_target:
          for k in TARGET_KEYS:
            if k not in ep:
              continue
            if k not in feats:
              feats[k] = self._alloc((len(pieces),) + ep[k].shape, ep[k].dtype)
            feats[k][pi] = ep[k]
          feats['target_index'][at:at + cnt] = pi
        fat += nf
      for k in FEATURE_KEYS:
        if k not in ep or ((on_device or pool) and k in BULK_KEYS):
          continue
        if k not in feats:
          feats[k] = self._alloc((n, self.K) + ep[k].shape[1:], ep[k].dtype)
        window_gather(ep[k], self.K, w0, cnt, out=feats[k][at:at + cnt])
      if self.fetch_target and not on_device and not pool:
        for k in TARGET_KEYS:
          if k not in ep:
            continue
          if k not in feats:
            feats[k] = self._alloc((n,) + ep[k].shape, ep[k].dtype)
          feats[k][at:at + cnt] = ep[k]                     # tf.tile over the windows (:624-625)
      last = slice(w0 + self.K - 1, w0 + self.K - 1 + cnt)  # the window's last frame (_prepare_v4 :393-397)
      for k in LABEL_KEYS:
        if k not in labels:
          labels[k] = np.empty((n,) + ep[k].shape[1:], dtype=ep[k].dtype)
        labels[k][at:at + cnt] = ep[k][last]
      at += cnt
    if self.pin:      # hand out the pinned tensors themselves (features are filled through their numpy views)
      by_ptr = {t.data_ptr(): t for t in self._pinned}
      feats = {k: by_ptr[v.ctypes.data] for k, v in feats.items()}
    if on_device:     # the consumer thread uploads / gathers the images (see _device_images)
      return feats, labels, self.pieces(lo, hi), {e: episodes[e] for e, _, _ in self.pieces(lo, hi)}
    return feats, labels

  def _device_images(self, feats, plan, episodes):
    """Adds the image features of one batch as tensors on self.device: uploads the episodes that are not resident
    yet (once each), gathers the windows by index, drops the episodes no later batch can read.  Runs in the
    consumer's thread on its current stream, so uploads, gathers and the step that follows are ordered."""
    import torch
    dev = torch.device(self.device)
    for e in [e for e in self._resident if e < plan[0][0]]:
      del self._resident[e]                                 # the stream only moves forward
    parts = collections.defaultdict(list)
    steps = torch.arange(self.K, device=dev)
    for e, w0, cnt in plan:
      if e not in self._resident:
        ep, res = episodes[e], {}
        for k in BULK_KEYS + (TARGET_KEYS if self.fetch_target else ()):
          if k not in ep:
            continue
          src = ep['_pinned'][k] if '_pinned' in ep else _host_tensor(ep[k], pin=False)
          res[k] = src.to(dev, non_blocking=True)
          self.uploaded_bytes += src.numel() * src.element_size()
        res['_host'] = ep.get('_pinned')                    # keeps the page-locked source alive with its copy
        self._resident[e] = res
      res = self._resident[e]
      idx = (torch.arange(w0, w0 + cnt, device=dev).unsqueeze(1) + steps.unsqueeze(0)).reshape(-1)
      for k in BULK_KEYS:
        if k in res:
          parts[k].append(_gather_frames(res[k], idx).reshape((cnt, self.K) + tuple(res[k].shape[1:])))
      for k in TARGET_KEYS:
        if k in res:
          parts[k].append(res[k].unsqueeze(0).expand((cnt,) + tuple(res[k].shape)))
    for k, ps in parts.items():
      feats[k] = (torch.cat(ps, dim=0) if len(ps) > 1 else ps[0]).contiguous()
    return feats

  def _adopt_cuda_device(self):
    """Page-locked allocations made in a helper thread must belong to the consumer's GPU, not to device 0 (one
    process per GPU: a thread that never called set_device would open a context on GPU 0 from every rank)."""
    if self._cuda_index is not None:
      import torch
      torch.cuda.set_device(self._cuda_index)

  def _produce(self, out_q, stop):
    try:
      self._adopt_cuda_device()
      ranges = self.batch_ranges()
      need = [[e for e, _, _ in self.pieces(lo, hi)] for lo, hi in ranges]
      order = sorted({e for es in need for e in es})        # stream episodes this rank reads, ascending
      with concurrent.futures.ThreadPoolExecutor(self.threads) as pool:
        pending, nxt = {}, 0
        for i, ((lo, hi), es) in enumerate(zip(ranges, need)):
          if stop.is_set():
            break
          # submit what this batch needs plus `threads` episodes of lookahead
          while nxt < len(order) and (order[nxt] <= es[-1] or len(pending) < len(es) + self.threads):
            pending[order[nxt]] = pool.submit(self._decode, order[nxt])
            nxt += 1
          batch = self._assemble(lo, hi, {e: pending[e].result() for e in es})
          keep_from = need[i + 1][0] if i + 1 < len(need) else None
          for e in [e for e in pending if keep_from is None or e < keep_from]:
            del pending[e]                                  # no later batch of this rank reads them
          out_q.put(batch)
      out_q.put(None)
    except BaseException as exc:                            # surfaces in the consumer
      out_q.put(exc)

  def __iter__(self):
    if self.pin or (self.device is not None and str(self.device).startswith('cuda')):
      import torch
      if torch.cuda.is_available():
        self._cuda_index = torch.cuda.current_device()
    out_q = queue.Queue(maxsize=self.prefetch)
    stop = threading.Event()
    worker = threading.Thread(target=self._produce, args=(out_q, stop), daemon=True)
    worker.start()
    try:
      while True:
        item = out_q.get()
        if item is None:
          return
        if isinstance(item, BaseException):
          raise item
        if self.device is not None:
          feats, labels, plan, episodes = item
          item = (self._device_images(feats, plan, episodes), labels)
        yield item
    finally:
      stop.set()
      self._resident.clear()
      while worker.is_alive():                              # unblock a producer waiting on a full queue
        try:
          out_q.get(timeout=0.05)
        except queue.Empty:
          pass


def pickplace_input_fn_v4(dataset_dir, split_name, mode, window_size=4, fetch_target=False, shuffle_buffer=128,
                          batch_size=1, num_epochs=1, num_threads=4, prefetch_size=4, seed=None,
                          frame_format='float32', drop_remainder=False, rank=0, world=1, pin_memory=False,
                          cache_dir=None, want_depth=True, device=None, layout='windows'):
  """Same signature and defaults as the reference (geeco_gym.py:401-412) plus the execution keywords after
  `seed`.  `shuffle_buffer` is accepted and unused, as in the reference (its window-level shuffle is commented
  out, :446-448); `mode == 'train'` shuffles the episode order with numpy's global generator (:436-437), or
  with `seed` when one is given so that all ranks of a data-parallel job agree on the order."""
  del shuffle_buffer
  meta = get_meta_v4(dataset_dir)
  paths = collect_tfrecords_v2(dataset_dir, split_name, mode)
  if mode == 'train':
    (np.random if seed is None else np.random.RandomState(seed)).shuffle(paths)
  print("[pickplace_input_fn_v4] #tfrecords: %d" % len(paths))
  return WindowBatches(paths, meta, window_size=window_size, fetch_target=fetch_target, batch_size=batch_size,
                       num_epochs=num_epochs, num_threads=num_threads, prefetch_size=prefetch_size,
                       frame_format=frame_format, drop_remainder=drop_remainder, rank=rank, world=world,
                       pin_memory=pin_memory, cache_dir=cache_dir, want_depth=want_depth, device=device, layout=layout)


def pickplace_input_fn(dataset_dir, split_name, mode, encoding='v4', window_size=4, fetch_target=False,
                       shuffle_buffer=128, batch_size=1, num_epochs=1, num_threads=4, prefetch_size=4, seed=None,
                       **execution):
  """Dispatcher of geeco_gym.py:234-279; only the V4 encoding (the published datasets) is implemented."""
  if encoding != 'v4':
    raise NotImplementedError("data encoding %r: only 'v4' is available" % (encoding,))
  return pickplace_input_fn_v4(dataset_dir, split_name, mode, window_size, fetch_target, shuffle_buffer, batch_size,
                               num_epochs, num_threads, prefetch_size, seed, **execution)
