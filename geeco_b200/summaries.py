"""TensorBoard event files with the reference's tag set.

Reference: every member of `tf.GraphKeys.LOSSES` gets a `tf.summary.scalar(name=loss.name, ...)` saved every
`log_steps` by a `SummarySaverHook` (src/models/e2evmc/estimator.py:262-265, :305-313), next to the `loss` and
`global_step/sec` scalars that tf.estimator writes itself; training events go to `model_dir`, evaluation events to
`model_dir/eval`.  A summary's tag is the op name with the characters TensorFlow does not allow in a tag (the ':' of
the output index) replaced by '_', so `CartesianCmdLoss/mean_squared_error/value:0` is logged as
`CartesianCmdLoss/mean_squared_error/value_0` (TF-1.15 behaviour stated from its documented naming: `tf.losses.*`
builds its result in an op called `value` inside the scope `mean_squared_error` / `softmax_cross_entropy_loss`; repeated
scopes get the suffixes `_1`, `_2`, ...).

The writer is tensorboard's own `EventFileWriter` (no TensorFlow needed); files are readable by TensorBoard and by
`tensorboard.backend.event_processing.event_accumulator`.
"""
from __future__ import annotations

import time

# loss-vector key (engine.losses_dict) -> TensorBoard tag, per control mode
LOSS_TAGS = {
    'cartesian': [                                             # graph.py:452-500, estimator.py:218-223
        ('loss_reg', 'RegularizationLoss/l2_reg_loss_0'),      # estimator.py:201-204 (added to LOSSES first)
        ('loss_cmd_ee', 'CartesianCmdLoss/mean_squared_error/value_0'),
        ('loss_cmd_grp', 'GripperCmdLoss/softmax_cross_entropy_loss/value_0'),
        ('loss_pos_ee', 'EEPoseAuxLoss/mean_squared_error/value_0'),
        ('loss_pos_obj', 'ObjPoseAuxLoss/mean_squared_error/value_0'),
    ],
    'velocity': [                                              # mse_loss, graph.py:430-450 over _PREDICTION_KEYS
        ('loss_reg', 'RegularizationLoss/l2_reg_loss_0'),
        ('loss_cmd_vel', 'MSELoss/mean_squared_error/value_0'),
        ('loss_cmd_ee', 'MSELoss/mean_squared_error_1/value_0'),
        ('loss_cmd_grp', 'MSELoss/mean_squared_error_2/value_0'),
        ('loss_pos_ee', 'MSELoss/mean_squared_error_3/value_0'),
        ('loss_pos_obj', 'MSELoss/mean_squared_error_4/value_0'),
    ],
}


class SummaryWriter(object):
  """Scalar summaries into `logdir/events.out.tfevents.*`."""

  def __init__(self, logdir):
    from tensorboard.summary.writer.event_file_writer import EventFileWriter
    self._w = EventFileWriter(logdir)

  def scalars(self, step, values: dict):
    from tensorboard.compat.proto import event_pb2, summary_pb2
    summ = summary_pb2.Summary(value=[summary_pb2.Summary.Value(tag=t, simple_value=float(v)) for t, v in values.items()])
    self._w.add_event(event_pb2.Event(wall_time=time.time(), step=int(step), summary=summ))

  def flush(self):
    self._w.flush()

  def close(self):
    self._w.close()


def loss_scalars(losses: dict, control_mode: str):
  """{tag: value} for one logged step: the loss parts under the reference's tags plus tf.estimator's own `loss`."""
  out = {tag: losses[k] for k, tag in LOSS_TAGS[control_mode] if k in losses}
  out['loss'] = losses['loss']
  return out


def read_scalars(logdir):
  """{tag: [(step, value)]} of every event file under `logdir` (used by the tests and tools)."""
  from tensorboard.backend.event_processing import event_accumulator
  acc = event_accumulator.EventAccumulator(logdir, size_guidance={event_accumulator.SCALARS: 0})
  acc.Reload()
  return {t: [(e.step, e.value) for e in acc.Scalars(t)] for t in acc.Tags()['scalars']}
