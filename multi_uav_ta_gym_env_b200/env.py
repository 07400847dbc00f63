"""Single-environment drop-in surface over the CUDA path.

  MultiUAVEnv          PettingZoo ParallelEnv-shaped reset/step/observations of
                       mUAV_TA/DroneEnv.py:70-1206 (same constructor argument, same dict outputs, same
                       attribute surface used by allocators and hybrids: SURVEY.md section 8(b)).
  HungarianAllocator   TaskAllocation/OptimizationBased/HungarianAllocator.py:14-208 (same signature).

State lives in one HBM record stepped by the kernels behind include/muav.h with E = 1; the Python
objects below are proxies refreshed from a host copy of that record after every call.  Proxies are
stable per agent / task id, so `task in env.last_tasks_info` and `.index(task)` behave as in the
reference (experiments/wps_eval.py:55-61).
"""
from __future__ import annotations

import contextlib
import ctypes as C
import random
import sys
from typing import List, Optional

import numpy as np

from . import _lib, state as _state
from .config import (CAP_TABLE, ENGAGE_RANGE, EVENT_TAGS, MAX_SPEEDS, TASK_DURATION, TASK_TYPES, UAV_TYPES,
                     agentEnvOptions)

MAX_INT = sys.maxsize


# --------------------------------------------------------------------------- proxies
class TaskProxy:
    """Read-only view of one task (DroneEnvComponents.Task attribute names)."""

    def __init__(self, env, tid):
        self._env = env
        self.id = int(tid)
        self.info = None
        self.relative_threat = None
        self.protected_task = None
        self.protected_agent = None
        self.kind = None
        self.eligible_agent_types = None
        self.required_agents = 0
        self.allocationDetails = {}
        self.max_time_steps = getattr(env, "max_time_steps", None)
        if tid == 0:
            self.type = "Hold"
            self.typeIdx = 0
            self.position = np.array([0, 0])
            self.status = 0
            z = np.zeros(6)
            self.orgReqs, self.currentReqs, self.allocatedReqs, self.doneReqs = z.copy(), z.copy(), z.copy(), z.copy()
            self.task_duration = TASK_DURATION["Hold"]
            self.initTime = self.doneTime = -1
            self.created_at = 0
            self.final_quality = -1

    def _refresh(self, s, k):
        env = self._env
        ti = int(s["k_type"][k])
        self.typeIdx = ti
        self.type = TASK_TYPES[ti]
        self.position = s["k_pos"][k].copy()
        self.status = int(s["k_status"][k])
        self.currentReqs = s["k_cur"][k].copy()
        self.allocatedReqs = s["k_alloc"][k].copy()
        org = np.zeros(6)
        org[ti] = s["k_org_ti"][k]
        self.orgReqs = org
        done = np.zeros(6)
        done[ti] = s["k_done_ti"][k]
        self.doneReqs = done
        self.task_duration = TASK_DURATION[self.type]
        self.initTime = float(s["k_init_time"][k])
        self.doneTime = float(s["k_done_time"][k])
        self.created_at = int(s["k_created_at"][k])
        self.final_quality = float(s["k_final_quality"][k])
        dl = int(s["k_deadline"][k])
        if dl >= 0:
            self.hard_deadline = dl          # attribute exists only once a window is set (DroneEnv.py:1493-1495)
            self.task_window = (self.created_at, dl)
        elif "hard_deadline" in self.__dict__:
            del self.__dict__["hard_deadline"]
        self.kind = "Escort" if s["k_kind"][k] == 1 else None
        self.required_agents = int(s["k_required_agents"][k])
        el = int(s["k_elig"][k])
        self.eligible_agent_types = None if el == 0 else {UAV_TYPES[i] for i in range(7) if (el >> i) & 1}
        pa = int(s["k_prot_agent"][k])
        self.protected_agent = env.agents_obj[pa] if pa >= 0 else None
        pt = int(s["k_prot_task"][k])
        self.protected_task = env._task(pt) if pt > 0 else None
        det = s["k_det_time"][k]
        self.allocationDetails = {int(a): (None, float(det[a])) for a in range(len(det)) if det[a] >= 0}
        if self.status == 2:
            # a closed task keeps the entries it had when it closed (removeAgentCap is a no-op then); the device keeps
            # their number only -- the one thing the reference reads from them (the replay's assigned_agents)
            self.allocationDetails = {f"stale{i}": (None, -1.0) for i in range(int(s["k_det_frozen"][k]))}
        self._wps_outcome_counted = bool(s["k_counted"][k])

    def __repr__(self):
        return f"<Task {self.id} {self.type} status={self.status}>"


class AgentProxy:
    """View of one UAV (DroneEnvComponents.UAV attribute names); commit_until is writable."""

    def __init__(self, env, aid, name, utype):
        self._env = env
        self.env = env
        self.id = int(aid)
        self.name = name
        self.type = utype
        self.typeIdx = UAV_TYPES.index(utype)
        self.max_speed = MAX_SPEEDS[utype] / env.simulation_frame_rate * 0.02
        self.engage_range = ENGAGE_RANGE[utype]
        self.initialCap2Task = np.array(CAP_TABLE[utype], dtype=np.float64)
        self._commit_until = 0

    def _refresh(self, s):
        env = self._env
        a = self.id
        self.position = s["a_pos"][a].copy()
        self.state = int(s["a_state"][a])
        self.currentCap2Task = s["a_caps"][a].copy()
        self.attackCap = int(s["a_ammo"][a])
        self.task_start = int(s["a_task_start"][a])
        self.fail_event = int(s["a_fail_event"][a])
        self.next_free_time = float(s["a_nft"][a])
        self.next_free_position = s["a_nfp"][a].copy()
        self.re_eval = bool(s["a_re_eval"][a])
        lt = int(s["a_last_task"][a])
        self.last_task = None if lt < 0 else env._task(lt)
        self._commit_until = int(s["a_commit_until"][a])
        q = [int(t) for t in s["a_queue"][a][: int(s["a_qlen"][a])]]
        self.tasks = [env._task(t) for t in q] if q else [env.task_idle]

    @property
    def commit_until(self):
        return self._commit_until

    @commit_until.setter
    def commit_until(self, value):
        self._commit_until = int(value)
        self._env._backend.patch_field("a_commit", self.id, int(value))

    def __repr__(self):
        return f"<UAV {self.name} id={self.id} state={self.state}>"


class ThreatProxy:
    def __init__(self, hid):
        self.id = hid

    def _refresh(self, env, s):
        h = self.id
        self.position = s["h_pos"][h].copy()
        self.status = int(s["h_status"][h])
        self.threat_type = UAV_TYPES[int(s["h_type"][h])]
        self.threat_group = int(s["h_group"][h])
        self.attackCap = int(s["h_ammo"][h])
        tg = int(s["h_target"][h])
        self.target_agent = env.agents_obj[tg] if tg >= 0 else None
        ms = int(s["h_mission"][h])
        self.mission_target_agent = env.agents_obj[ms] if ms >= 0 else None
        ic = int(s["h_intercept"][h])
        self.intercepting_agent = env.agents_obj[ic] if ic >= 0 else None
        tk = int(s["h_task"][h])
        self.relative_task = env._task(tk) if tk > 0 else None


class _AgentSelector:
    """pettingzoo.utils.agent_selector semantics (only feeds infos['selected'], DroneEnv.py:787,1194)."""

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


# --------------------------------------------------------------------------- CUDA backend (E = 1)
class _CudaBackend:
    def __init__(self, config, device, task_cap="all", queue_cap=16):
        # one slot per task id: the facade keeps the whole task history like env.tasks of the reference (closed tasks
        # stay readable for the replay generator / the web UI); recycling slots only pays for large batches
        from .batched_env import BatchedMultiUAVEnv

        self.b = BatchedMultiUAVEnv(config, 1, device=device, task_cap=task_cap, queue_cap=queue_cap)
        self.lib = self.b.lib
        self.cfg = self.b.cfg
        self.codec = self.b.codec

    def reset(self, seed):
        self.b.reset([seed])
        return self.b.scenarios[0]

    def record(self):
        return self.b.record_host(0)

    def step(self, actions):
        import torch

        self.b.step_batched(torch.from_numpy(actions[None]))
        return (float(self.b.reward[0].item()), bool(self.b.terminated[0].item()), bool(self.b.truncated[0].item()),
                self.b.events_of(0))

    def allocate(self, spec, scores, priorities, reserved, order, cbba_seed=None):
        import torch

        O, keep = self.b._alloc_opts(spec, None, None if priorities is None else torch.from_numpy(priorities[None]),
                                     None if reserved is None else torch.from_numpy(reserved[None]))
        if cbba_seed is not None:
            sd = torch.tensor([int(cbba_seed)], dtype=torch.int32, device=self.b.device)
            keep.append(sd)
            O.d_cbba_seed = sd.data_ptr()
        if scores is not None:
            sc = torch.from_numpy(scores[None]).to(self.b.device)
            keep.append(sc)
            O.d_edge_scores = sc.data_ptr()
            O.score_rows, O.score_cols, O.score_f64 = scores.shape[0], scores.shape[1], 1
        if order is not None:
            od = torch.from_numpy(order[None]).to(self.b.device)
            keep.append(od)
            O.d_task_order = od.data_ptr()
        rc = self.lib.dll.muav_allocate(C.byref(self.cfg), self.b.records.data_ptr(), C.byref(O), C.byref(self.b._out),
                                        None, 1, self.b._stream())
        _lib.check(rc, "muav_allocate")
        return self.b.pairs_of(0)

    def bundle_pairs(self):
        return self.b.bundle_pairs_of(0)

    def observe(self, max_rows):
        return {k: v[0].cpu().numpy() for k, v in self.b.observe(max_rows).items()}

    def metrics(self):
        return self.b.metrics()[0].cpu().numpy()

    def patch_field(self, name, index, value):
        import torch

        off, cnt, dt = self.codec.F[name]
        raw = np.array([value], dtype=dt).view(np.uint8)
        pos = off + index * dt.itemsize
        self.b.records[0, pos:pos + dt.itemsize] = torch.from_numpy(raw).to(self.b.device)


# --------------------------------------------------------------------------- environment
class MultiUAVEnv:
    metadata = {"render_modes": ["human"], "name": "multi_agent_env_v0"}

    def __init__(self, config=None, device="cuda:0"):
        self.config = config if config is not None else agentEnvOptions()
        cfg = self.config
        g = lambda n, d=None: getattr(cfg, n, d)
        self.fixed_seed = g("fixed_seed", -1)
        self._seed = self.fixed_seed if self.fixed_seed != -1 else 0
        self.area_width, self.area_height = 1200, 700
        self.max_coord = max(self.area_height, self.area_width)
        self.bases = [np.array([400, 680])]
        self.max_time_steps = cfg.max_time_steps
        self.simulation_frame_rate = cfg.simulation_frame_rate
        self.agents_config = cfg.agents
        self.n_agents = sum(cfg.agents.values())
        self.max_agents = max(48, self.n_agents + 8)
        self.possible_agents = [f"{t[0:2]}_agent{i}" for t, n in cfg.agents.items() for i in range(n)]
        self.agents = self.possible_agents
        self.agent_selection = self.possible_agents[0]
        self.n_tasks = sum(cfg.tasks.values()) + 1
        self.max_tasks = self.n_tasks + 28
        self.multiple_tasks_per_agent = cfg.multiple_tasks_per_agent
        self.multiple_agents_per_task = cfg.multiple_agents_per_task
        self.num_obstacles = cfg.num_obstacles
        self.fail_rate = cfg.fail_rate
        self.early_terminate = g("early_terminate", False)
        self.capability_mask = g("capability_mask", False)
        self.saturate_mask = g("saturate_mask", False)
        self.include_time_windows = bool(g("include_time_windows", False))
        self.arrival_rate = float(g("arrival_rate", 0.0) or 0.0)
        self.sense_radius = float(g("sense_radius", 0.0) or 0.0)
        self.threat_delay = int(g("threat_delay", 0) or 0)
        self.hard_windows = bool(g("hard_windows", False))
        self.window_length = int(g("window_length", 30) or 30)
        self.miss_penalty = float(g("miss_penalty", 25.0) or 0.0)
        self.on_time_bonus = float(g("on_time_bonus", 10.0) or 0.0)
        self.burst_mode = bool(g("burst_mode", False))          # read by build_rah_state (ReserveAwareHybrid.py:56)
        self.burst_size = int(g("burst_size", 3) or 3)
        self.dual_region_bursts = bool(g("dual_region_bursts", False))
        self.share_knowledge = bool(g("share_knowledge", True))
        self.commit_horizon = int(g("commit_horizon", 0) or 0)
        self.reassign_penalty = float(g("reassign_penalty", 0.0) or 0.0)
        self.escort_enabled = bool(g("escort_enabled", False))
        self.escort_radius = float(g("escort_radius", 70.0) or 70.0)
        self.escort_requirement = float(g("escort_requirement", 1.2) or 1.2)
        self.escort_intercept_radius = float(g("escort_intercept_radius", 100.0) or 100.0)
        self.mutual_support_radius = float(g("mutual_support_radius", 80.0) or 80.0)
        self.escort_agent_types = tuple(g("escort_agent_types", ("F1", "F2")) or ("F1", "F2"))
        self.agent_selector = _AgentSelector(self.possible_agents)
        self.current_agent = self.agent_selector.next()
        self.agents_obj: Optional[List[AgentProxy]] = None
        self.tasks: List[TaskProxy] = []
        self.threats: List[ThreatProxy] = []
        self.last_tasks_info = None
        self.task_idle = TaskProxy(self, 0)
        self.time_steps = 0
        self.event_list = []
        self._tasks_by_id = {}
        self._obs_cache = None
        self._backend = self._make_backend(cfg, device)
        self.rewards = {a: 0 for a in self.possible_agents}
        self.terminations = {a: False for a in self.possible_agents}
        self.truncations = {a: False for a in self.possible_agents}
        self.infos = {a: {} for a in self.possible_agents}

    def _make_backend(self, cfg, device):
        """The state holder behind the facade: the CUDA library on `device`.  Raises without libmuav_b200.so or a GPU --
        there is no CPU fallback in the product (tests/helpers.py subclasses the facade over the CPU build of the kernel
        sources to run the unmodified reference planners side by side in the GPU-less container)."""
        return _CudaBackend(cfg, device)

    # ---- spaces (the reference's declared spaces do not match its emitted observations; kept as shapes)
    def observation_space(self, agent):
        return {"agent_position": (2,), "agent_state": (5,), "agent_type": (6,), "next_free_time": (1,),
                "position_after_last_task": (2,), "tasks_info": (self.max_tasks * 12,)}

    def action_space(self, agent):
        return {a: (self.max_tasks,) for a in self.possible_agents}

    # ---- proxies
    def _task(self, tid):
        if tid == 0:
            return self.task_idle
        t = self._tasks_by_id.get(tid)
        if t is None:
            t = TaskProxy(self, tid)
            self._tasks_by_id[tid] = t
        return t

    def _sync(self):
        rec = self._backend.record()
        s = self._backend.codec.snapshot(rec)
        self._snap = s
        self._rec = rec
        self.time_steps = s["t"]
        for a in self.agents_obj:
            pass
        T = s["n_tasks"]
        self.tasks = [self._task(k + 1) for k in range(T)]
        for a in self.agents_obj:
            a._refresh(s)
        for k, t in enumerate(self.tasks):
            t._refresh(s, k)
        self.threats = []
        for hid in s["h_order"]:
            th = self._threats_all[int(hid)]
            th._refresh(self, s)
            self.threats.append(th)
        for k, t in enumerate(self.tasks):
            h = int(s["k_threat"][k])
            t.relative_threat = self._threats_all[h] if h >= 0 else None
        open_ids = self._backend.codec.open_task_ids(rec)
        self.last_tasks_info = [self._task(t) for t in open_ids]
        self.agent_known_tasks = {a.name: {k + 1 for k in range(T) if s["known"][a.id, k]} for a in self.agents_obj}
        self.event_list = [[EVENT_TAGS[int(tag)], int(arg)] for tag, arg in s["events"]]
        for name, key in (("F_Reward", "F_Reward"), ("total_distance", "total_distance"), ("n_on_time", "n_on_time"),
                          ("n_missed_windows", "n_missed"), ("n_windowed_tasks", "n_windowed"),
                          ("n_task_switches", "n_task_switches"), ("n_reallocations", "n_reallocations"),
                          ("n_arrivals", "n_arrivals"), ("conclusion_time", "conclusion_time"),
                          ("escort_requests", "escort_requests"), ("escort_completed", "escort_completed"),
                          ("escort_failed", "escort_failed"), ("escort_required_steps", "escort_required_steps"),
                          ("escort_covered_steps", "escort_covered_steps"), ("protection_breaches", "protection_breaches"),
                          ("threats_intercepted", "threats_intercepted"), ("recon_losses", "recon_losses"),
                          ("escort_losses", "escort_losses"), ("mutual_support_engagements", "mutual_support"),
                          ("protected_rec_completed", "protected_rec_completed"), ("_idle_reserve_steps", "idle_reserve_steps")):
            setattr(self, name, s[key])
        self._pending_reset = bool(s["pending_reset"])
        self._escort_by_recon = {a.name: self._task(int(s["a_escort"][a.id])) for a in self.agents_obj
                                 if s["a_escort"][a.id] != 0}
        self.agent_distances = s["a_dist"].copy()
        self._obs_cache = None

    # ---- ParallelEnv API
    def seed(self, seed):
        self.reset(seed=seed)

    def reset(self, seed=None, return_info=True, options=None):
        self._seed = random.randint(0, MAX_INT) if seed is None else seed
        if self.fixed_seed != -1:
            self._seed = self.fixed_seed
        sc = self._backend.reset(self._seed)
        self.agents = self.possible_agents.copy()
        self.agent_selector = _AgentSelector(self.agents)
        self.current_agent = self.agent_selector.next()
        self.agents_obj = [AgentProxy(self, aid, sc.agent_names[aid], UAV_TYPES[sc.agent_type[aid]])
                           for aid in range(self.n_agents)]
        self.agent_by_name = {a.name: a for a in self.agents_obj}
        self._tasks_by_id = {}
        self._threats_all = [ThreatProxy(h) for h in range(self._backend.cfg.n_threats)]
        self.mission_areas = list(sc.mission_areas)
        self.reward_norm_factor = sc.reward_norm_factor
        self.obstacles = list(sc.obstacles)
        self.current_agent = self.agent_selector.reset()
        self._sync()
        self.rewards = {a.name: 0 for a in self.agents_obj}
        self.terminations = {a.name: False for a in self.agents_obj}
        self.truncations = {a.name: False for a in self.agents_obj}
        self.infos = {a.name: {} for a in self.agents_obj}
        return self.observations, self.infos

    def step(self, actions):
        self.agent_selection = self.agent_selector.next()
        self.current_agent = self.agent_selection
        A = self.n_agents
        act = np.full((A, 2), -1, dtype=np.int32)
        n = 0
        if isinstance(actions, dict):
            for name, idxs in actions.items():
                aid = self.agent_by_name[name].id
                if not isinstance(idxs, list):
                    idxs = [idxs]
                for idx in idxs:
                    if n >= A:
                        raise ValueError("more than n_agents actions in one step are not supported")
                    act[n] = (aid, int(idx))
                    n += 1
        reward, term, trunc, events = self._backend.step(act)
        self._sync()
        self.rewards = {a.name: reward for a in self.agents_obj}
        self.terminations = {a.name: term for a in self.agents_obj}
        self.truncations = {a.name: trunc for a in self.agents_obj}
        self.infos = {a.name: {} for a in self.agents_obj}
        self.infos["selected"] = self.agent_selection
        self.infos["events"] = events
        if term or trunc:
            self.infos["metrics"] = self.calculate_metrics()
        return self.observations, self.rewards, self.terminations, self.truncations, self.infos

    @property
    def observations(self):
        """Per-agent observation dicts (DroneEnv.py:468-492), built lazily from the muav_observe tensors."""
        if self._obs_cache is None:
            self._obs_cache = self._build_observations()
        return self._obs_cache

    def _build_observations(self):
        rows = max(self.max_tasks, len(self.last_tasks_info), 1)
        o = self._backend.observe(rows)
        ti = o["tasks_info"]
        n_rows = int(o["n_rows"])
        shared = []
        for r in range(n_rows):
            row = ti[r]
            d = {"id": int(row[0]), "position": row[1:3].copy(), "status": int(row[3]),
                 "current_reqs": row[4:10].copy(), "alloc_reqs": row[10:16].copy()}
            if len(self.last_tasks_info) > 0:
                if self.include_time_windows:
                    d["init_time"] = float(row[16])
                    d["end_time"] = float(row[17])
                    d["type_idx"] = float(row[18])
                d["unmet"] = float(row[19])
                d["age"] = float(row[20])
            shared.append(d)
        pad = max(self.max_tasks - n_rows, 0)
        shared.extend([{"status": -1} for _ in range(pad)])
        mask = [True] * n_rows + [False] * pad
        obs = {}
        ev = o["event_flags"].astype(np.float32)
        for a in self.agents_obj:
            legal = [bool(x) for x in o["legal_mask"][a.id][:n_rows]] + [False] * pad
            obs[a.name] = {
                "agent_position": a.position / self.max_coord,
                "agent_caps": a.currentCap2Task,
                "alloc_task": a.tasks[0].id,
                "tasks_info": shared,
                "mask": mask,
                "legal_mask": legal,
                "event_flags": ev.copy(),
            }
        return obs

    def observe(self, agent):
        ob = self.observations[agent]
        ob["agent_id"] = agent
        return ob

    # ---- surface used by allocators / hybrids
    def get_live_agents(self):
        return [a for a in self.agents_obj if a.state != -1]

    def agent_visibility_map(self):
        if not self.sense_radius and not self.threat_delay:
            return None
        return {name: set(ids) for name, ids in self.agent_known_tasks.items()}

    def known_tasks_for(self, agent_name=None):
        if agent_name is not None:
            ids = self.agent_known_tasks.get(agent_name, set())
            return [t for t in self.tasks if t.id in ids or t.id == 0]
        known = set()
        for s in self.agent_known_tasks.values():
            known |= s
        if not self.sense_radius and not self.threat_delay:
            return list(self.tasks)
        return [t for t in self.tasks if t.id in known or t.id == 0]

    def _is_task_action_valid(self, agent, task):
        if task is None or task.status == 2:
            return False
        if len(agent.tasks) > 0 and agent.tasks[0].id == task.id:
            return True
        eligible = getattr(task, "eligible_agent_types", None)
        if eligible is not None:
            if isinstance(eligible, str):
                eligible = {eligible}
            if agent.type not in eligible:
                return False
        if self.capability_mask and agent.currentCap2Task[task.typeIdx] <= 0:
            return False
        if self.saturate_mask and task.allocatedReqs[task.typeIdx] >= task.orgReqs[task.typeIdx]:
            return False
        return True

    def compute_s_wps(self) -> float:
        return float(self._metric("S_WPS"))

    def compute_s_esc(self) -> float:
        return float(self._metric("S_ESC"))

    def _metric(self, name):
        m = self._backend.metrics()
        return m[self._metric_names().index(name)]

    def _metric_names(self):
        if not hasattr(self, "_mnames"):
            d = self._backend.lib.dll
            d.muav_metric_name.restype = C.c_char_p
            d.muav_metric_name.argtypes = [C.c_int]
            self._mnames = [d.muav_metric_name(i).decode() for i in range(_lib.N_METRICS)]
        return self._mnames

    def calculate_metrics(self):
        m = self._backend.metrics()
        names = self._metric_names()
        ints = {"Losses", "Kills", "n_reallocations", "n_task_switches", "n_arrivals", "n_tasks_final", "n_reached",
                "n_missed_windows", "n_on_time", "n_windowed_tasks", "protected_rec_completed", "recon_losses",
                "escort_losses", "threats_intercepted", "mutual_support_engagements", "protection_breaches",
                "escort_requests", "escort_completed", "escort_failed"}
        return {n: (int(v) if n in ints else float(v)) for n, v in zip(names, m)}

    def _escort_fighters_near(self, protected_agent, radius=None):
        """_escort_fighters_near (DroneEnv.py:1746-1764) over the proxies (used by escort hybrids' tokens)."""
        if protected_agent is None:
            return []
        escort = self._escort_by_recon.get(protected_agent.name)
        if escort is None or escort.status == 2:
            return []
        r = float(self.escort_radius if radius is None else radius)
        near = []
        for a in self.agents_obj:
            if a.state == -1 or a.type not in self.escort_agent_types:
                continue
            if not a.tasks or a.tasks[0].id != escort.id:
                continue
            d = float(np.linalg.norm(a.position - protected_agent.position))
            if d <= r:
                near.append((d, a))
        near.sort(key=lambda x: x[0])
        return [a for _, a in near]


def env(config=None):
    return MultiUAVEnv(config)


def raw_env(config=None):
    return MultiUAVEnv(config)


# --------------------------------------------------------------------------- allocator
@contextlib.contextmanager
def _known_override(env, agent_known_ids):
    """`agent_known_ids` of an allocator call (HungarianAllocator.py:125-147, CBBA.py:115-134, PerformanceImpact.py:94-123:
    agent name -> set of task ids, a name missing from the map knows nothing).  The device allocator filters by the
    known-task bitmask of the record; when the caller's map is not the environment's own (env.agent_visibility_map()),
    the differing mask words are written into the record for the duration of the call and put back afterwards."""
    own = env.agent_visibility_map() or {}
    if agent_known_ids is None or all(set(agent_known_ids.get(n, ())) == ids for n, ids in own.items()):
        yield
        return
    be = env._backend
    n_agents = len(env.agents_obj)
    old = np.array(be.codec.field(be.record(), "known"), dtype=np.uint32)
    n_words = old.size // n_agents          # layout [word, agent], bit k of word k >> 5 = task id k + 1
    new = np.zeros_like(old)
    for a in env.agents_obj:
        for tid in agent_known_ids.get(a.name, ()):
            k = int(tid) - 1
            if 0 <= k < 32 * n_words:
                new[(k >> 5) * n_agents + a.id] |= np.uint32(1 << (k & 31))
    changed = [i for i in range(old.size) if new[i] != old[i]]
    for i in changed:
        be.patch_field("known", i, int(new[i]))
    try:
        yield
    finally:
        for i in changed:
            be.patch_field("known", i, int(old[i]))


class HungarianAllocator:
    """Same constructor, attributes and allocate_tasks signature as the reference class; the cost matrix,
    the LSAP rounds and the acceptance rule run in the CUDA allocator (muav_allocate)."""

    def __init__(self, replan_interval: int = 20, max_coord: float = 1000.0):
        self.replan_interval = max(1, int(replan_interval))
        self.max_coord = max_coord
        self.last_plan_step = -10**9
        self.n_replans = 0
        self.n_calls = 0

    def should_replan(self, time_step: int, events=None) -> bool:
        if time_step - self.last_plan_step >= self.replan_interval:
            return True
        if events:
            for ev in events:
                tag = ev[0] if isinstance(ev, (list, tuple)) and ev else ev
                if tag in EVENT_TAGS:
                    return True
        return False

    def allocate_tasks(self, agents, tasks, time_step: int = 0, events=None, force: bool = False, task_priorities=None,
                       reserved_agent_names=None, agent_known_ids=None, edge_scores=None):
        from .batched_env import AllocSpec

        self.n_calls += 1
        if not force and not self.should_replan(time_step, events):
            return []
        agents = list(agents)
        tasks = list(tasks)
        env = None
        for obj in agents + tasks:
            env = getattr(obj, "_env", None)
            if env is not None:
                break
        if env is None:
            return []
        if int(time_step) != int(env.time_steps):
            raise ValueError("time_step must be env.time_steps (the device allocator reads the env clock)")
        ids = [a.id for a in agents]
        if ids != sorted(ids):
            # the rows of the cost matrix -- and with them SciPy's choice among equal-cost assignments and the order of the
            # result -- follow the `agents` list (HungarianAllocator.py:128-141); the device builds the rows in agent-id
            # order, which is what env.get_live_agents() and every order-preserving filter of it give
            raise ValueError("agents must be listed in environment order (env.get_live_agents() or a filtered copy of it)")
        cfg = env._backend.cfg
        A, TC = cfg.n_agents, max(cfg.id_cap, cfg.task_cap)  # per-task arguments are indexed by task id
        reserved_names = set(reserved_agent_names or [])
        given = {a.id for a in agents}
        reserved = np.zeros(A, dtype=np.uint8)
        for a in env.agents_obj:
            if a.id not in given or a.name in reserved_names:
                reserved[a.id] = 1
        order = np.full(TC, -1, dtype=np.int32)
        n_ord = 0
        for t in tasks:
            if t.id != 0 and n_ord < TC:
                order[n_ord] = t.id - 1
                n_ord += 1
        use_vis = agent_known_ids is not None
        pri = None
        if task_priorities:
            pri = np.zeros(TC, dtype=np.float64)
            for tid, v in task_priorities.items():
                if 0 < int(tid) <= TC:
                    pri[int(tid) - 1] = float(v)
        scores = None
        if edge_scores:
            scores = np.zeros((A, TC), dtype=np.float64)
            by_name = env.agent_by_name
            for (name, tid), v in edge_scores.items():
                if name in by_name and 0 < int(tid) <= TC:
                    scores[by_name[name].id, int(tid) - 1] = float(v)
        spec = AllocSpec(3, self.replan_interval, 0, use_vis, False, float(self.max_coord))
        before = env._backend.codec.header(env._backend.record(), "N_REPLANS")
        with _known_override(env, agent_known_ids):
            pairs = env._backend.allocate(spec, scores, pri, reserved, order)
        after = env._backend.codec.header(env._backend.record(), "N_REPLANS")
        if after == before:
            return []  # no live agent or no open task: the reference returns before touching its counters
        self.last_plan_step = time_step
        self.n_replans += 1
        return [(env.agents_obj[a].name, env._task(tid)) for a, tid in pairs]


def _in_caller_order(result, agents):
    """The market allocators list their plan agent by agent in the order of the `agents` argument
    (PerformanceImpact.py:205-218, CBBA.py:88-108,199-211); the device lists it in agent-id order, which is the same thing
    for env.get_live_agents() -- every reference driver's argument -- and is re-ordered here for any other list.  (Exact
    ties between indistinguishable agents -- same type, same place -- are still broken in agent-id order on the device.)"""
    pos = {a.name: i for i, a in enumerate(agents)}
    return sorted(result, key=lambda item: pos.get(item[0], len(pos)))


class PerformanceImpact:
    """Same constructor, attributes and allocate_tasks signature as the reference's market baseline
    (TaskAllocation/MarketBased/PerformanceImpact.py:27-224); the slot expansion, the IPI / RPI costs and the
    inclusion loop run in the CUDA allocator (muav_allocate, muav_alloc_opts.planner = 6).  Returns
    (agent_name, [tasks in path order]) items like the reference.  max_tasks_per_agent = 1 is what every reference driver
    passes (experiments/wps_eval.py:147-156, escort_eval.py:162-170); bundles of up to 4 tasks per agent are built by the
    same allocator (pi_bundles, csrc/muav_alloc.cuh); larger values raise."""

    def __init__(self, max_coord: float = 1000.0, seed: int = 0, replan_interval: int = 12, max_iters: int = 40):
        self.max_coord = float(max_coord)
        self.seed = int(seed)
        self.replan_interval = max(1, int(replan_interval))
        self.max_iters = max(4, int(max_iters))
        self.last_plan_step = -10**9
        self.n_replans = 0
        self.n_calls = 0

    should_replan = HungarianAllocator.should_replan

    def allocate_tasks(self, agents, tasks, time_step: int = 0, events=None, force: bool = False, agent_known_ids=None,
                       reserved_agent_names=None, max_tasks_per_agent: int = 1):
        from .batched_env import AllocSpec

        if not 1 <= int(max_tasks_per_agent) <= 4:
            raise NotImplementedError("the device PI allocator builds bundles of 1..4 tasks per agent")
        self.n_calls += 1
        if not force and not self.should_replan(time_step, events):
            return []
        agents = list(agents)
        tasks = list(tasks)
        # every call that passes the rule counts as a replan, also the empty ones (PerformanceImpact.py:80-93,222-223)
        self.last_plan_step = time_step
        self.n_replans += 1
        env = None
        for obj in agents + tasks:
            env = getattr(obj, "_env", None)
            if env is not None:
                break
        if env is None or not agents or not tasks:
            return []
        if int(time_step) != int(env.time_steps):
            raise ValueError("time_step must be env.time_steps (the device allocator reads the env clock)")
        cfg = env._backend.cfg
        A, TC = cfg.n_agents, max(cfg.id_cap, cfg.task_cap)
        reserved_names = set(reserved_agent_names or [])
        given = {a.id for a in agents}
        reserved = np.zeros(A, dtype=np.uint8)
        for a in env.agents_obj:
            if a.id not in given or a.name in reserved_names:
                reserved[a.id] = 1
        order = np.full(TC, -1, dtype=np.int32)
        n_ord = 0
        for t in tasks:
            if t.id != 0 and n_ord < TC:
                order[n_ord] = t.id - 1
                n_ord += 1
        use_vis = agent_known_ids is not None
        spec = AllocSpec(3, self.replan_interval, 0, use_vis, False, float(self.max_coord), planner=6,
                         max_tasks_per_agent=int(max_tasks_per_agent))
        with _known_override(env, agent_known_ids):
            pairs = env._backend.allocate(spec, None, None, reserved, order)
        if max_tasks_per_agent > 1:   # (name, [tasks in path order]) like the reference (PerformanceImpact.py:207-220)
            out = []
            for a, tid in env._backend.bundle_pairs():
                if out and out[-1][0] == env.agents_obj[a].name:
                    out[-1][1].append(env._task(tid))
                else:
                    out.append((env.agents_obj[a].name, [env._task(tid)]))
            return _in_caller_order(out, agents)
        return _in_caller_order([(env.agents_obj[a].name, [env._task(tid)]) for a, tid in pairs], agents)


class CBBAReplan:
    """Same constructor, attributes and allocate_tasks signature as the reference's CBBA with periodic / event-triggered
    replan (TaskAllocation/MarketBased/CBBA_Replan.py:15-69 around MarketBased/CBBA.py:68-324); the auction (slot
    expansion, MT19937 shuffles, bids over insertion points, consensus) runs in the CUDA allocator
    (muav_alloc_opts.planner = 7, csrc/muav_cbba.cuh).  The reference's auction order starts from a set of strings, so its
    result depends on the interpreter's string hash: this class reproduces the reference run under PYTHONHASHSEED=0.
    max_tasks_per_agent = 1 is what every reference driver passes; bundles of up to 4 tasks per agent are built by the same
    auction (CBBA.py:128-147 with a longer bundle); larger values raise."""

    def __init__(self, agents=None, tasks=None, max_coord: float = 1000.0, seed: int = 0, replan_interval: int = 20):
        self.max_coord = max_coord
        self.seed = int(seed)
        self.replan_interval = max(1, int(replan_interval))
        self.last_plan_step = -10**9
        self.n_replans = 0
        self.n_calls = 0

    should_replan = HungarianAllocator.should_replan

    def allocate_tasks(self, agents, tasks, time_step: int = 0, events=None, force: bool = False, agent_known_ids=None,
                       reserved_agent_names=None, max_tasks_per_agent: int = 1):
        from .batched_env import AllocSpec

        if not 1 <= int(max_tasks_per_agent) <= 4:
            raise NotImplementedError("the device CBBA allocator builds bundles of 1..4 tasks per agent")
        self.n_calls += 1
        if not force and not self.should_replan(time_step, events):
            return []
        agents = list(agents)
        tasks = list(tasks)
        self.last_plan_step = time_step
        self.n_replans += 1
        env = None
        for obj in agents + tasks:
            env = getattr(obj, "_env", None)
            if env is not None:
                break
        if env is None or not agents or not tasks:
            return []
        if int(time_step) != int(env.time_steps):
            raise ValueError("time_step must be env.time_steps (the device allocator reads the env clock)")
        cfg = env._backend.cfg
        A, TC = cfg.n_agents, max(cfg.id_cap, cfg.task_cap)
        reserved_names = set(reserved_agent_names or [])
        given = {a.id for a in agents}
        reserved = np.zeros(A, dtype=np.uint8)
        for a in env.agents_obj:
            if a.id not in given or a.name in reserved_names:
                reserved[a.id] = 1
        order = np.full(TC, -1, dtype=np.int32)
        n_ord = 0
        for t in tasks:
            if t.id != 0 and n_ord < TC:
                order[n_ord] = t.id - 1
                n_ord += 1
        use_vis = agent_known_ids is not None
        spec = AllocSpec(3, self.replan_interval, 0, use_vis, False, float(self.max_coord), planner=7,
                         max_tasks_per_agent=int(max_tasks_per_agent))
        # the device seeds its generator with d_cbba_seed + N_REPLANS of the record (incremented by this call): hand it the
        # difference so that the generator is Random(self.seed + self.n_replans), whatever else replanned on this record
        n_dev = int(env._backend.codec.header(env._backend.record(), "N_REPLANS"))
        with _known_override(env, agent_known_ids):
            pairs = env._backend.allocate(spec, None, None, reserved, order, cbba_seed=self.seed + self.n_replans - (n_dev + 1))
        if max_tasks_per_agent > 1:   # (name, [tasks in the order they were won]) like the reference (CBBA.py:192-204)
            out = []
            for a, tid in env._backend.bundle_pairs():
                if out and out[-1][0] == env.agents_obj[a].name:
                    out[-1][1].append(env._task(tid))
                else:
                    out.append((env.agents_obj[a].name, [env._task(tid)]))
            return _in_caller_order(out, agents)
        return _in_caller_order([(env.agents_obj[a].name, [env._task(tid)]) for a, tid in pairs], agents)


class CBBA:
    """The bare auction class (TaskAllocation/MarketBased/CBBA.py:68-324): same constructor and allocate_tasks signature;
    one auction with the generator `random.Random(seed)`, which is what `CBBAReplan` builds for every replan
    (CBBA_Replan.py:62-69) and what the reference's own tests do with it (test_escort.py:166-189).  The reference keeps
    drawing from the same generator when allocate_tasks is called again on one instance (the legacy "CBBA" algorithm of
    paper_eval.py:187-195); the device generator is seeded per call, so a second call on the same instance raises instead
    of answering with another stream.  Reproduces the reference under PYTHONHASHSEED=0 (see CBBAReplan)."""

    def __init__(self, drones=None, tasks=None, max_dist: float = 1000.0, seed: int = 0):
        self.max_dist = max_dist
        self.seed = int(seed)
        self._auctions = 0

    def allocate_tasks(self, agents, tasks, Qs=None, agent_known_ids=None, reserved_agent_names=None, time_step: int = 0,
                       max_tasks_per_agent: int = 1):
        agents = list(agents)
        tasks = list(tasks)
        env = next((getattr(o, "_env", None) for o in agents + tasks if getattr(o, "_env", None) is not None), None)
        if env is None or not agents or not tasks:
            return []
        if self._auctions:
            raise NotImplementedError("one auction per CBBA instance: the device generator is seeded per call "
                                      "(use CBBAReplan, or a fresh CBBA(seed) per call)")
        self._auctions += 1
        # CBBAReplan runs CBBA(seed + n_replans) with n_replans = 1 on its first replan
        inner = CBBAReplan(None, None, self.max_dist, seed=self.seed - 1, replan_interval=1)
        return inner.allocate_tasks(agents, tasks, time_step=env.time_steps, force=True, agent_known_ids=agent_known_ids,
                                    reserved_agent_names=reserved_agent_names, max_tasks_per_agent=max_tasks_per_agent)
