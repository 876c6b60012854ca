"""`data.geeco_gym` of the reference (src/data/geeco_gym.py), for the names its scripts import:
scripts/train_e2evmc.py:13 (`pickplace_input_fn`), scripts/gym_pickplace.py:41 / gym_pushing.py:38 (`load_target_frame`)."""
from geeco_b200.input_pipeline import (PickAndPlaceMetaV4, collect_tfrecords_v2 as _collect_tfrecords_v2,  # noqa: F401
                                       get_meta_v4 as _get_meta_v4, load_keyframes, load_target_frame,
                                       load_target_frames, pickplace_input_fn, pickplace_input_fn_v4)
from geeco_b200.data_recorder import PickAndPlaceEncodingV4  # noqa: F401,E402
