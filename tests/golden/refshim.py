"""Import shim that lets the UNMODIFIED reference tree (/root/reference) run in
the authoring container, where gymnasium / pettingzoo / matplotlib / seaborn /
tianshou and the Rust `core_sim` crate are absent.

Used ONLY by tests/golden/gen_golden.py (fixture generation) and by the
optional `-m "not gpu"` cross-checks that skip when /root/reference is missing.
Nothing under multi_uav_ta_gym_env_b200/, bench.py or smoke() imports this.

Stubs follow SURVEY.md Appendix D:
  * gymnasium.spaces.{Dict,Box,Discrete,MultiDiscrete} are only constructed
    (DroneEnv.py:298-308,499).
  * pettingzoo.utils.agent_selector semantics: next() advances and returns
    order[cur-1]; reset() = reinit + next()  (DroneEnv.py:142-143,597-598,754,787).
  * core_sim.SimCore.avoid_obstacles restates core_sim/src/sim_core.rs:24-59
    (Rust `%` == C fmod).
"""
from __future__ import annotations

import math
import os
import sys
import types

REF_ROOT = os.environ.get("MUAV_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "mUAV_TA"))


def _mod(name):
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


def install():
    if "mUAV_TA.DroneEnv" in sys.modules:
        return
    if not reference_available():
        raise RuntimeError("reference tree not present")

    if "gymnasium" not in sys.modules:
        gym = _mod("gymnasium")
        spaces = _mod("gymnasium.spaces")

        class _Space:
            def __init__(self, *a, **k):
                self.shape = k.get("shape", None)
                self.args = a
                self.kwargs = k

        class Dict(dict):
            def __init__(self, d=None, **kw):
                super().__init__(d or {}, **kw)

        class Box(_Space):
            pass

        class Discrete(_Space):
            pass

        class MultiDiscrete(_Space):
            pass

        spaces.Dict, spaces.Box, spaces.Discrete, spaces.MultiDiscrete = Dict, Box, Discrete, MultiDiscrete
        gym.spaces = spaces

    if "pettingzoo" not in sys.modules:
        pz = _mod("pettingzoo")

        class ParallelEnv:
            def __init__(self, *a, **k):
                pass

        pz.ParallelEnv = ParallelEnv
        utils = _mod("pettingzoo.utils")
        utils.parallel_to_aec = lambda e: e
        wrappers = _mod("pettingzoo.utils.wrappers")
        wrappers.OrderEnforcingWrapper = lambda e: e
        utils.wrappers = wrappers
        sel = _mod("pettingzoo.utils.agent_selector")

        class agent_selector:
            def __init__(self, order):
                self.reinit(order)

            def reinit(self, order):
                self.agent_order = order
                self._current_agent = 0
                self.selected_agent = 0

            def reset(self):
                self.reinit(self.agent_order)
                return self.next()

            def next(self):
                self._current_agent = (self._current_agent + 1) % len(self.agent_order)
                self.selected_agent = self.agent_order[self._current_agent - 1]
                return self.selected_agent

        sel.agent_selector = agent_selector
        utils.agent_selector = sel
        pz.utils = utils

    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        if name not in sys.modules:
            _mod(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

    if "core_sim" not in sys.modules:
        cs = _mod("core_sim")

        class SimCore:
            @staticmethod
            def avoid_obstacles(agent_pos, obstacles, movement):
                ax = ay = 0.0
                for ob in obstacles:
                    dx = ob[0] - agent_pos[0]
                    dy = ob[1] - agent_pos[1]
                    d = math.sqrt(dx * dx + dy * dy)
                    dz = d - ob[2]
                    if dz < 40.0:
                        nx, ny = dx / dz, dy / dz
                        f = 0.5 / (1.0 - math.log(max(1.05, dz)))
                        ang = math.atan2(movement[1], movement[0]) - math.atan2(dy, dx)
                        ang = math.fmod(ang + math.pi, 2.0 * math.pi) - math.pi
                        if ang > 0.0:
                            rx, ry = ny, -nx
                        else:
                            rx, ry = -ny, nx
                        ax += rx * f
                        ay += ry * f
                return [ax, ay]

        cs.SimCore = SimCore

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


def make_config(spec, env_flags, **overrides):
    """Same kwargs mapping as experiments/paper_eval.py:42-82 (restated so that
    tianshou does not have to be importable)."""
    install()
    from mUAV_TA.MultiDroneEnvUtils import agentEnvOptions

    kw = dict(
        render_speed=-1,
        simulation_frame_rate=0.01,
        max_time_steps=150,
        action_mode="TaskAssign",
        agents=dict(spec["agents"]),
        tasks=dict(spec["tasks"]),
        random_init_pos=False,
        num_obstacles=0,
        multiple_tasks_per_agent=False,
        multiple_agents_per_task=True,
        fail_rate=float(spec.get("fail_rate", 0.0)),
        threats_list=list(spec.get("threats_list") or []),
        fixed_seed=-1,
        early_terminate=bool(env_flags.get("early_terminate", True)),
        capability_mask=bool(env_flags.get("capability_mask", False)),
        saturate_mask=bool(env_flags.get("saturate_mask", False)),
        reward_weights=env_flags.get("reward_weights"),
        arrival_rate=float(spec.get("arrival_rate", 0.0)),
        include_time_windows=bool(env_flags.get("include_time_windows", False)),
        dynamic_idle_penalty=float(env_flags.get("dynamic_idle_penalty", 0.0)),
        sense_radius=float(spec.get("sense_radius", 0.0) or 0.0),
        threat_delay=int(spec.get("threat_delay", 0) or 0),
        hard_windows=bool(spec.get("hard_windows", False)),
        window_length=int(spec.get("window_length", 30) or 30),
        burst_mode=bool(spec.get("burst_mode", False)),
        burst_size=int(spec.get("burst_size", 3) or 3),
        miss_penalty=float(spec.get("miss_penalty", 25.0) or 0.0),
        on_time_bonus=float(spec.get("on_time_bonus", 10.0) or 0.0),
        dual_region_bursts=bool(spec.get("dual_region_bursts", False)),
        share_knowledge=bool(spec.get("share_knowledge", True)),
        commit_horizon=int(spec.get("commit_horizon", 0) or 0),
        reassign_penalty=float(spec.get("reassign_penalty", 0.0) or 0.0),
        escort_enabled=bool(spec.get("escort_enabled", False)),
        escort_radius=float(spec.get("escort_radius", 70.0) or 70.0),
        escort_requirement=float(spec.get("escort_requirement", 1.2) or 1.2),
        escort_intercept_radius=float(spec.get("escort_intercept_radius", 100.0) or 100.0),
        mutual_support_radius=float(spec.get("mutual_support_radius", 80.0) or 80.0),
        escort_agent_types=tuple(spec.get("escort_agent_types", ("F1", "F2")) or ("F1", "F2")),
    )
    kw.update(overrides)
    return agentEnvOptions(**kw)


def wps_config(case_id, **overrides):
    """The configuration run_wps_episode / run_escort_episode build
    (experiments/wps_eval.py:91-97, escort_eval.py:95-101)."""
    install()
    from experiments.paper_scenarios import CASE_SPECS, WPS_ENV_FLAGS

    flags = dict(WPS_ENV_FLAGS)
    flags["capability_mask"] = False
    flags["saturate_mask"] = False
    cfg = make_config(CASE_SPECS[case_id], flags, **overrides)
    cfg.multiple_tasks_per_agent = True
    return cfg
