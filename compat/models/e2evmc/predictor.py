from geeco_b200.predictor import GoalE2EVMCPredictor, BatchedGoalPredictor, TOL_FRAME_RANGE  # noqa: F401
from geeco_b200.predictor import E2EVMCPredictor  # noqa: F401
