"""Configuration objects of the batched WPS environment.

`agentEnvOptions` keeps the keyword surface of the reference's config class
(mUAV_TA/MultiDroneEnvUtils.py:5-105); `CASE_SPECS` / `WPS_ENV_FLAGS` carry the
scenario registry values of experiments/paper_scenarios.py:59-266,353-366 and
`make_config` / `wps_config` mirror experiments/paper_eval.py:42-82 and
experiments/wps_eval.py:91-97.
"""
from __future__ import annotations

import copy
from typing import Any, Dict

UAV_TYPES = ["R1", "R2", "E1", "F1", "F2", "T1", "T2"]
TASK_TYPES = ["Hold", "Rec", "Att", "Def", "Int", "Det"]
EVENT_TAGS = ["Reset_Allocation", "Agent_Fail", "New_Threat", "Escort_Created", "Escort_Retired"]

# mUAV_TA/MultiDroneEnvData.py:8-83
GAME_AREA = (1200, 700)
CONTACT_LINE = 550
BASE = (400, 680)
CAP_TABLE = {
    "R1": [0.1, 1.0, 0.0, 0.2, 0.0, 0.0],
    "R2": [0.1, 0.6, 0.0, 0.1, 0.0, 0.0],
    "E1": [0.1, 0.8, 0.0, 0.2, 0.0, 1.0],
    "F1": [0.1, 0.0, 0.7, 1.0, 1.0, 1.0],
    "F2": [0.1, 0.0, 1.0, 0.6, 0.8, 1.0],
    "T1": [0.0, 0.0, 0.2, 0.5, 1.0, 1.0],
    "T2": [0.0, 0.0, 0.2, 0.4, 0.8, 0.8],
}
MAX_SPEEDS = {"F1": 20.0, "F2": 15.0, "R1": 5.0, "R2": 8.0, "E1": 5.0, "T1": 14.0, "T2": 12.0}
ENGAGE_RANGE = {"F1": 40.0, "F2": 30.0, "R1": 0.0, "R2": 0.0, "E1": 0.0, "T1": 35.0, "T2": 25.0}
FAIL_TABLE = {"F1": 1.5, "F2": 0.8, "R1": 1.2, "R2": 0.8, "E1": 1.5, "T1": 1.8, "T2": 1.0}
TASK_DURATION = {"Hold": 1, "Rec": 10, "Att": 5, "Def": 5, "Int": 0, "Det": 1}

_DEFAULT_RW = {
    "action": 0.0, "distance": 1.0, "quality": 1.0, "s_quality": 1.0,
    "time": 0.0, "alloc": 0.0, "time_penaulty": 0.0, "step": 0.0,
}
RW_KEYS = ["action", "distance", "quality", "s_quality", "time", "alloc", "time_penaulty", "step"]


class agentEnvOptions:
    """Same keyword arguments and attribute names as the reference's agentEnvOptions."""

    def __init__(self, render_mode="human", render_speed=-1, simulation_frame_rate=0.01,
                 action_mode="TaskAssign", simulator_module="Internal", max_time_steps=150,
                 agents=None, tasks=None, multiple_tasks_per_agent=False, multiple_agents_per_task=True,
                 random_init_pos=False, num_obstacles=0, hidden_obstacles=False, fail_rate=0.0,
                 threats_list=None, fixed_seed=-1, info="No Info", early_terminate=False,
                 capability_mask=False, saturate_mask=False, reward_weights=None, arrival_rate=0.0,
                 include_time_windows=False, dynamic_idle_penalty=0.0, sense_radius=0.0, threat_delay=0,
                 hard_windows=False, window_length=30, burst_mode=False, burst_size=3, miss_penalty=25.0,
                 on_time_bonus=10.0, dual_region_bursts=False, share_knowledge=True, commit_horizon=0,
                 reassign_penalty=0.0, escort_enabled=False, escort_radius=70.0, escort_requirement=1.2,
                 escort_intercept_radius=100.0, mutual_support_radius=80.0, escort_agent_types=("F1", "F2")):
        self.render_mode = render_mode
        self.render_speed = render_speed
        self.simulation_frame_rate = simulation_frame_rate
        self.action_mode = action_mode
        self.simulator_module = simulator_module
        self.max_time_steps = max_time_steps
        self.random_init_pos = random_init_pos
        self.agents = dict(agents) if agents is not None else {"F1": 0, "F2": 0, "R1": 1, "R2": 1}
        self.tasks = dict(tasks) if tasks is not None else {"Att": 0, "Rec": 2, "Hold": 0}
        self.multiple_tasks_per_agent = multiple_tasks_per_agent
        self.multiple_agents_per_task = multiple_agents_per_task
        self.num_obstacles = num_obstacles
        self.hidden_obstacles = hidden_obstacles
        self.fail_rate = fail_rate
        self.threats_list = list(threats_list) if threats_list is not None else [("T1", 4), ("T2", 2)]
        self.fixed_seed = fixed_seed
        self.info = info
        self.early_terminate = early_terminate
        self.capability_mask = capability_mask
        self.saturate_mask = saturate_mask
        self.reward_weights = reward_weights or dict(_DEFAULT_RW)
        self.arrival_rate = arrival_rate
        self.include_time_windows = include_time_windows
        self.dynamic_idle_penalty = dynamic_idle_penalty
        self.sense_radius = sense_radius
        self.threat_delay = threat_delay
        self.hard_windows = hard_windows
        self.window_length = window_length
        self.burst_mode = burst_mode
        self.burst_size = burst_size
        self.miss_penalty = miss_penalty
        self.on_time_bonus = on_time_bonus
        self.dual_region_bursts = dual_region_bursts
        self.share_knowledge = share_knowledge
        self.commit_horizon = int(commit_horizon or 0)
        self.reassign_penalty = float(reassign_penalty or 0.0)
        self.escort_enabled = bool(escort_enabled)
        self.escort_radius = float(escort_radius or 70.0)
        self.escort_requirement = float(escort_requirement or 1.2)
        self.escort_intercept_radius = float(escort_intercept_radius or 100.0)
        self.mutual_support_radius = float(mutual_support_radius or 80.0)
        self.escort_agent_types = tuple(escort_agent_types or ("F1", "F2"))


def _wps(agents, tasks, fail_rate, threats, arrival, sense, delay, window, burst_mode, burst_size,
         miss=None, bonus=None, **extra):
    spec = {
        "agents": dict(zip(("F1", "F2", "R1", "R2"), agents)),
        "tasks": {"Att": tasks[0], "Rec": tasks[1], "Hold": 0},
        "fail_rate": fail_rate,
        "threats_list": [("T1", threats[0]), ("T2", threats[1])],
        "arrival_rate": arrival,
        "sense_radius": sense,
        "threat_delay": delay,
        "hard_windows": True,
        "window_length": window,
        "burst_mode": burst_mode,
        "burst_size": burst_size,
    }
    if miss is not None:
        spec["miss_penalty"] = miss
    if bonus is not None:
        spec["on_time_bonus"] = bonus
    spec.update(extra)
    return spec


_DUAL = dict(dual_region_bursts=True, share_knowledge=False)
_ESCORT = dict(escort_enabled=True, escort_radius=70.0, escort_requirement=1.2, escort_intercept_radius=100.0,
               mutual_support_radius=80.0, escort_agent_types=("F1", "F2"))

CASE_SPECS: Dict[str, Dict[str, Any]] = {
    "WPS_easy": _wps((2, 2, 2, 2), (4, 6), 0.05, (4, 3), 0.08, 250.0, 8, 40, False, 2, 25.0, 10.0),
    "WPS_hard": _wps((2, 2, 2, 2), (3, 5), 0.08, (5, 4), 0.12, 120.0, 15, 25, True, 3, 30.0, 12.0),
    "WPS_burst": _wps((2, 2, 2, 2), (2, 4), 0.1, (6, 4), 0.15, 150.0, 12, 20, True, 4, 35.0, 15.0),
    "WPS_attn": _wps((4, 2, 4, 2), (4, 8), 0.08, (8, 6), 0.18, 90.0, 18, 22, True, 4, 30.0, 12.0, **_DUAL),
    "WPS_attn_XL": _wps((14, 6, 14, 6), (13, 26), 0.08, (26, 20), 0.18, 90.0, 18, 22, True, 4, 30.0, 12.0, **_DUAL),
    "WPS_commit": _wps((4, 2, 4, 2), (4, 8), 0.08, (8, 6), 0.18, 90.0, 18, 22, True, 4, 30.0, 12.0,
                       commit_horizon=25, reassign_penalty=2.0, **_DUAL),
    "WPS_escort": _wps((5, 3, 4, 2), (2, 6), 0.03, (4, 6), 0.15, 100.0, 15, 28, True, 3, 30.0, 12.0,
                       commit_horizon=20, reassign_penalty=2.0, **_DUAL, **_ESCORT),
}


def _attn_variant(**over):
    s = copy.deepcopy(CASE_SPECS["WPS_attn"])
    s.update(over)
    return s


def _static(agents, tasks, fail_rate=0.0, threats=(), arrival=0.0):
    return {"agents": dict(zip(("F1", "F2", "R1", "R2"), agents)), "tasks": {"Att": tasks[0], "Rec": tasks[1], "Hold": 0},
            "fail_rate": fail_rate, "threats_list": [(n, c) for n, c in zip(("T1", "T2"), threats)], "arrival_rate": arrival}


# the remaining registered scenarios of experiments/paper_scenarios.py: common-operating-picture sweeps of WPS_attn
# (:126-160, 274-300), its larger fleets (:161-215) and the legacy static / dynamic cases (:14-57)
CASE_SPECS.update({
    "WPS_attn_AWACS": _attn_variant(sense_radius=0.0, threat_delay=0, share_knowledge=True),
    "WPS_attn_L": _attn_variant(agents={"F1": 10, "F2": 5, "R1": 10, "R2": 5}, tasks={"Att": 10, "Rec": 20, "Hold": 0},
                                threats_list=[("T1", 20), ("T2", 15)]),
    "WPS_attn_OS18": _attn_variant(agents={"F1": 6, "F2": 3, "R1": 6, "R2": 3}),
    "WPS_attn_OS24": _attn_variant(agents={"F1": 8, "F2": 4, "R1": 8, "R2": 4}),
    "static_strike": _static((0, 2, 0, 0), (15, 0)),
    "scal_None": _static((0, 2, 0, 0), (15, 0)),
    "recon_strike_mix": _static((2, 0, 4, 0), (6, 12)),
    "train_mixed": _static((2, 0, 4, 0), (6, 12)),
    "agent_scaling_mid": _static((3, 0, 6, 0), (6, 24)),
    "scal_Agents_mid": _static((3, 0, 6, 0), (6, 24)),
    "D1_attrition": _static((2, 0, 4, 0), (6, 12), fail_rate=0.1),
    "D2_popup_threats": _static((2, 2, 2, 2), (4, 8), threats=(3, 2)),
    "D3_combined": _static((2, 2, 2, 2), (4, 8), fail_rate=0.1, threats=(3, 2), arrival=0.02),
})
for _r in (60, 90, 150, 250):
    CASE_SPECS[f"WPS_attn_COP_R{_r}"] = _attn_variant(sense_radius=float(_r), threat_delay=18, share_knowledge=False)
for _d in (0, 6, 12, 18):
    CASE_SPECS[f"WPS_attn_COP_d{_d}"] = _attn_variant(sense_radius=90.0, threat_delay=_d, share_knowledge=False)
    CASE_SPECS[f"WPS_attn_COP_cue_d{_d}"] = _attn_variant(sense_radius=0.0, threat_delay=_d, share_knowledge=True)


# display names used by the result tables (experiments/paper_scenarios.py "label")
CASE_LABELS = {'D1_attrition': 'Attrition (fail_rate=0.1)',
 'D2_popup_threats': 'Pop-up Threats',
 'D3_combined': 'Attrition+Threats',
 'WPS_attn': 'WPS Attn stress (multi-front)',
 'WPS_attn_AWACS': 'WPS Attn + full COP (AWACS/ground)',
 'WPS_attn_COP_R150': 'WPS Attn COP R=150 d=18',
 'WPS_attn_COP_R250': 'WPS Attn COP R=250 d=18',
 'WPS_attn_COP_R60': 'WPS Attn COP R=60 d=18',
 'WPS_attn_COP_R90': 'WPS Attn COP R=90 d=18',
 'WPS_attn_COP_cue_d0': 'WPS Attn COP cueing d=0 (share)',
 'WPS_attn_COP_cue_d12': 'WPS Attn COP cueing d=12 (share)',
 'WPS_attn_COP_cue_d18': 'WPS Attn COP cueing d=18 (share)',
 'WPS_attn_COP_cue_d6': 'WPS Attn COP cueing d=6 (share)',
 'WPS_attn_COP_d0': 'WPS Attn COP R=90 d=0',
 'WPS_attn_COP_d12': 'WPS Attn COP R=90 d=12',
 'WPS_attn_COP_d18': 'WPS Attn COP R=90 d=18',
 'WPS_attn_COP_d6': 'WPS Attn COP R=90 d=6',
 'WPS_attn_L': 'WPS Attn L (~30 agents)',
 'WPS_attn_OS18': 'WPS Attn oversized 18 agents (1.5x)',
 'WPS_attn_OS24': 'WPS Attn oversized 24 agents (2x)',
 'WPS_attn_XL': 'WPS Attn XL (~40 agents)',
 'WPS_burst': 'WPS Burst stress',
 'WPS_commit': 'WPS Commit (dual-front + rematch)',
 'WPS_easy': 'WPS Easy (windows+delay)',
 'WPS_escort': 'WPS Escort (coalition protect)',
 'WPS_hard': 'WPS Hard (tight+local+burst)',
 'agent_scaling_mid': 'Agent Scaling',
 'recon_strike_mix': 'Recon-Strike Mix',
 'scal_Agents_mid': 'Agent Scaling',
 'scal_None': 'Static Strike',
 'static_strike': 'Static Strike',
 'train_mixed': 'Recon-Strike Mix'}
for _k, _v in CASE_LABELS.items():
    CASE_SPECS[_k]["label"] = _v


def burst_scaled_spec(k: int) -> Dict[str, Any]:
    """BASELINE config 5: WPS_burst with agents/tasks/threats scaled by k."""
    s = copy.deepcopy(CASE_SPECS["WPS_burst"])
    s["agents"] = {"F1": 2 * k, "F2": 2 * k, "R1": 2 * k, "R2": 2 * k}
    s["tasks"] = {"Att": 2 * k, "Rec": 4 * k, "Hold": 0}
    s["threats_list"] = [("T1", 6 * k), ("T2", 4 * k)]
    return s


_E3_RW = {"action": 0.0, "distance": 1.0, "quality": 1.0, "s_quality": 1.0,
          "time": 0.0, "alloc": 0.0, "time_penaulty": 0.0, "step": 0.0}

WPS_ENV_FLAGS = {
    "early_terminate": False,
    "capability_mask": True,
    "saturate_mask": True,
    "include_time_windows": True,
    "dynamic_idle_penalty": 0.05,
    "reward_weights": _E3_RW,
}


def make_config(spec: Dict[str, Any], env_flags: Dict[str, Any], **overrides) -> agentEnvOptions:
    kw = dict(
        render_speed=-1, simulation_frame_rate=0.01, max_time_steps=150, action_mode="TaskAssign",
        agents=dict(spec["agents"]), tasks=dict(spec["tasks"]), random_init_pos=False, num_obstacles=0,
        multiple_tasks_per_agent=False, multiple_agents_per_task=True,
        fail_rate=float(spec.get("fail_rate", 0.0)), threats_list=list(spec.get("threats_list") or []),
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


def wps_config(case_id, **overrides) -> agentEnvOptions:
    """Configuration of run_wps_episode / run_escort_episode: WPS flags without the action masks and
    with per-agent task queues (wps_eval.py:91-97, escort_eval.py:95-101)."""
    spec = case_id if isinstance(case_id, dict) else CASE_SPECS[case_id]
    flags = dict(WPS_ENV_FLAGS)
    flags["capability_mask"] = False
    flags["saturate_mask"] = False
    cfg = make_config(spec, flags)
    cfg.multiple_tasks_per_agent = True
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg
