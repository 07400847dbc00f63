"""B200-native batched simulator for the Multi-UAV-TA "Windowed Pop-up Strike" step path.

Public surface:
  agentEnvOptions, CASE_SPECS, WPS_ENV_FLAGS, make_config, wps_config   (config.py: every scenario of paper_scenarios.py)
  BatchedMultiUAVEnv, AllocSpec                                         (batched_env.py; needs the CUDA library)
  MultiUAVEnv, HungarianAllocator, PerformanceImpact, CBBAReplan, CBBA  (env.py: the single-environment drop-in facade)
  core_sim                                                              (core_sim.py: the PyO3 module's stand-in)
  scorers                 AttPair / MLPPair / ContextPair / GNN / Commit / Coalition networks, fused Att-Pair kernel
  collectors              batched IL / RL data collectors of train_pair_cost.py
  evaluate                batched wps_eval.py / escort_eval.py driver (python -m multi_uav_ta_gym_env_b200.evaluate)
  replay                  replay documents of generate_simulation_replay.py
  sharding                env-index sharding across ranks + the end-of-episode metric all-reduce
"""
from .config import CASE_SPECS, WPS_ENV_FLAGS, agentEnvOptions, burst_scaled_spec, make_config, wps_config  # noqa: F401


from . import sharding  # noqa: E402,F401


def __getattr__(name):
    # torch / CUDA dependent parts are imported lazily so that config-only users stay light
    if name in ("BatchedMultiUAVEnv", "AllocSpec"):
        from . import batched_env

        return getattr(batched_env, name)
    if name in ("MultiUAVEnv", "HungarianAllocator", "PerformanceImpact", "CBBAReplan", "CBBA"):
        from . import env

        return getattr(env, name)
    raise AttributeError(name)
