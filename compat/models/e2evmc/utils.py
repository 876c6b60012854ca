from geeco_b200.params import load_model_config, save_model_config  # noqa: F401
