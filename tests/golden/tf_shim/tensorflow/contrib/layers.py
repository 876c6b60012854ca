"""tf.contrib.layers.l2_regularizer (see ../__init__.py)."""


def l2_regularizer(scale):
  if not scale:
    return lambda w: None
  return lambda w: scale * (w ** 2).sum() / 2
