from geeco_b200.estimator import goal_e2evmc_model_fn, Estimator, RunConfig, ModeKeys  # noqa: F401
