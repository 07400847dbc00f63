"""B200-native batched simulator for the Multi-UAV-TA "Windowed Pop-up Strike" step path.

Public surface:
  agentEnvOptions, CASE_SPECS, WPS_ENV_FLAGS, make_config, wps_config   (config.py)
  BatchedMultiUAVEnv, AllocSpec                                         (batched_env.py; needs the CUDA library)
"""
from .config import CASE_SPECS, WPS_ENV_FLAGS, agentEnvOptions, burst_scaled_spec, make_config, wps_config  # noqa: F401


from . import sharding  # noqa: E402,F401


def __getattr__(name):
    # torch / CUDA dependent parts are imported lazily so that config-only users stay light
    if name in ("BatchedMultiUAVEnv", "AllocSpec"):
        from . import batched_env

        return getattr(batched_env, name)
    if name in ("MultiUAVEnv", "HungarianAllocator"):
        from . import env

        return getattr(env, name)
    raise AttributeError(name)
