"""geeco_b200: B200-native implementation of the GEECO e2evmc controller's train / inference step.

Python host over a C-ABI shared library of hand-written sm_100a CUDA kernels
(include/geeco_b200.h).  See DESIGN.md for the scope and INTEGRATION.md for the drop-in boundary.
"""
from .params import (E2EVMCConfig, E2E_VMC_DEFAULT_CONFIG, E2E_VMC_DEFAULT_PARAM_DICT, create_e2evmc_config,
                     load_model_config, save_model_config)

__all__ = ['E2EVMCConfig', 'E2E_VMC_DEFAULT_CONFIG', 'E2E_VMC_DEFAULT_PARAM_DICT', 'create_e2evmc_config',
           'load_model_config', 'save_model_config']
