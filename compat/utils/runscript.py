from geeco_b200.runscript import save_run_command  # noqa: F401
