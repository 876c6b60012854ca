from geeco_b200.estimator import goal_e2evmc_model_fn, Estimator, RunConfig, ModeKeys  # noqa: F401
from geeco_b200.estimator import e2evmc_model_fn  # noqa: F401
