from geeco_b200.params import *  # noqa: F401,F403
from geeco_b200.params import E2EVMCConfig, E2E_VMC_DEFAULT_CONFIG, E2E_VMC_DEFAULT_PARAM_DICT, create_e2evmc_config  # noqa: F401
