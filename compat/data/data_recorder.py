"""`data.data_recorder` of the reference (src/data/data_recorder.py)."""
from geeco_b200.data_recorder import TfrSequenceEncoding, TfrSequenceRecorder  # noqa: F401
