from geeco_b200.graph import dynimg, conv_encoder, goal_e2evmc  # noqa: F401
