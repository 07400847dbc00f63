"""Replay files for the reference's web UI (server/api.py + frontend/): the JSON document of
experiments/generate_simulation_replay.py (metadata / frames / events / final_metrics, :119-314), produced from an
environment of this package.  SURVEY.md section 8(f) row 4.

`record_replay(env, plan, ...)` works on anything with the MultiUAVEnv attribute surface (this package's facade on its
CUDA backend, or the reference environment itself -- which is how tests pin the format: the same planner on both gives
byte-identical documents, and the reference's own generator run on the facade gives the same file too).
`plan(env, events) -> (pairs, new_commit_names)` is the caller's planner hook.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Callable, Iterable

REPLAN_TAGS = ("Reset_Allocation", "New_Threat", "Agent_Fail", "Escort_Created", "Escort_Retired")
TITLES = {
    "WPS_escort": ("WPS_escort: protect recon with fighter coalitions", "Urgency-Coalition + Coalition-Hungarian"),
    None: ("WPS_commit: dual-front dynamic mission", "Urgency-Commit + Local-Hungarian"),
}
DYNAMICS = (("arrival_rate", float), ("fail_rate", float), ("sense_radius", float), ("threat_delay", int),
            ("hard_windows", bool), ("window_length", int), ("burst_mode", bool), ("burst_size", int),
            ("dual_region_bursts", bool), ("share_knowledge", bool), ("commit_horizon", int), ("reassign_penalty", float))


def replan_due(env, events, interval=15) -> bool:
    """generate_simulation_replay._should_replan (:21-37)."""
    if env.time_steps == 0 or env.time_steps % interval == 0:
        return True
    return any((ev[0] if isinstance(ev, (list, tuple)) and ev else ev) in REPLAN_TAGS for ev in events)


def _event(ev, t):
    if isinstance(ev, (list, tuple)):
        return {"time": t, "type": str(ev[0]) if ev else "Unknown", "detail": [str(x) for x in ev[1:]]}
    return {"time": t, "type": str(ev), "detail": []}


def _xy(p):
    return [float(p[0]), float(p[1])]


def snapshot_frame(env, events: Iterable, replanned: bool, new_commits: list) -> dict:
    """One frame (:120-216): what every UAV, task and threat looks like now, plus the running scores."""
    vis = env.agent_visibility_map()
    if not vis:
        everything = {t.id for t in env.tasks if t.id != 0}
        vis = {a.name: everything for a in env.get_live_agents()}
    seen_by = {}
    for ids in vis.values():
        for tid in ids:
            seen_by[tid] = seen_by.get(tid, 0) + 1
    agents = []
    for a in env.agents_obj:
        head = a.tasks[0] if a.tasks else env.task_idle
        agents.append({"id": int(a.id), "name": a.name, "type": a.type, "position": _xy(a.position), "state": int(a.state),
                       "task_id": int(head.id), "commit_until": int(getattr(a, "commit_until", 0) or 0),
                       "known_tasks": len(vis.get(a.name, set()))})
    tasks = []
    for t in env.tasks:
        if t.id == 0:
            continue
        dl = getattr(t, "hard_deadline", None)
        kind = getattr(t, "kind", None)
        prot = getattr(t, "protected_agent", None)
        tasks.append({
            "id": int(t.id), "type": t.type, "kind": kind, "position": _xy(t.position), "status": int(t.status),
            "created_at": int(getattr(t, "created_at", 0) or 0), "deadline": None if dl is None else int(dl),
            "required": float(t.currentReqs[t.typeIdx]), "allocated": float(t.allocatedReqs[t.typeIdx]),
            "known_by": int(seen_by.get(t.id, 0)), "is_dynamic": dl is not None, "is_escort": kind == "Escort",
            "required_agents": int(getattr(t, "required_agents", 0) or 0),
            "assigned_agents": int(len(getattr(t, "allocationDetails", {}) or {})),
            "protected_agent": None if prot is None else str(prot.name),
            "protected_position": None if prot is None else _xy(prot.position)})
    threats = []
    for h in env.threats:
        tgt = getattr(h, "mission_target_agent", None)
        icp = getattr(h, "intercepting_agent", None)
        threats.append({"id": int(h.id), "position": _xy(h.position), "status": int(h.status), "group": int(h.threat_group),
                        "threat_type": getattr(h, "threat_type", None),
                        "mission_target": None if tgt is None else str(tgt.name),
                        "intercepting": None if icp is None else str(icp.name)})
    t_now = int(env.time_steps)
    cover = float(getattr(env, "escort_covered_steps", 0) / max(getattr(env, "escort_required_steps", 0), 1))
    return {
        "time": t_now, "agents": agents, "tasks": tasks, "threats": threats,
        "events": [_event(ev, env.time_steps) for ev in events],
        "decision": {"replanned": replanned, "new_commits": new_commits},
        "metrics": {
            "s_wps": float(env.compute_s_wps()),
            "s_esc": float(env.compute_s_esc()) if hasattr(env, "compute_s_esc") else float(env.compute_s_wps()),
            "on_time": int(env.n_on_time), "missed": int(env.n_missed_windows), "switches": int(env.n_task_switches),
            "distance": float(env.total_distance),
            "active_agents": sum(1 for a in env.agents_obj if a.state != -1),
            "open_tasks": sum(1 for t in env.tasks if t.id != 0 and t.status != 2),
            "escort_coverage": cover, "recon_losses": int(getattr(env, "recon_losses", 0)),
            "protected_rec": int(getattr(env, "protected_rec_completed", 0)),
            "mutual_support": int(getattr(env, "mutual_support_engagements", 0))},
    }


def derived_events(prev: dict, cur: dict) -> list:
    """Reviewer-facing events that the environment does not emit itself (:59-117)."""
    t = cur["time"]
    out = []
    before_agents = {a["name"]: a for a in prev["agents"]}
    before_tasks = {(k["type"], int(k["id"])): k for k in prev["tasks"]}
    before_threats = {h["id"] for h in prev["threats"]}
    for a in cur["agents"]:
        old = before_agents.get(a["name"])
        if old and old["state"] != -1 and a["state"] == -1:
            out.append({"time": t, "type": "Agent_Fail", "detail": [a["name"]]})
    for k in cur["tasks"]:
        old = before_tasks.get((k["type"], int(k["id"])))
        label = f"{k['type']}{k['id']}"
        if old is None:
            out.append({"time": t, "type": "Task_Arrival", "detail": [label, "left" if k["position"][0] < 600 else "right"]})
        elif old["status"] != 2 and k["status"] == 2:
            late = k["deadline"] is not None and t > k["deadline"]
            out.append({"time": t, "type": "Window_Missed" if late else "Task_Completed", "detail": [label]})
        if old and old["known_by"] == 0 and k["known_by"] > 0:
            out.append({"time": t, "type": "Task_Discovered", "detail": [label, f"by {k['known_by']} UAV(s)"]})
    for h in cur["threats"]:
        if h["id"] not in before_threats:
            out.append({"time": t, "type": "Threat_Spawn", "detail": [str(h["id"])]})
    for name in cur["decision"]["new_commits"]:
        out.append({"time": t, "type": "Agent_Commit", "detail": [name]})
    if cur["decision"]["replanned"]:
        out.append({"time": t, "type": "Replan", "detail": []})
    return out


def record_replay(env, info, plan: Callable, config, scenario: str, seed: int) -> dict:
    """Roll one episode from a freshly reset `env` (`info` = the infos of reset) and return the replay document."""
    title, algorithm = TITLES.get(scenario, TITLES[None])
    frames = [snapshot_frame(env, [], False, [])]
    log = []
    done = {a: False for a in env.agents}
    trunc = {a: False for a in env.agents}
    while not all(done.values()) and not all(trunc.values()):
        seen = list(info.get("events") or []) if isinstance(info, dict) else []
        actions, replanned, commits = {}, False, []
        if replan_due(env, seen):
            pairs, commits = plan(env, seen)
            for name, task in pairs:
                if env.last_tasks_info and task in env.last_tasks_info:
                    actions[name] = env.last_tasks_info.index(task)
            replanned = True
        _, _, done, trunc, info = env.step(actions)
        now = list(info.get("events") or []) if isinstance(info, dict) else []
        frame = snapshot_frame(env, now, replanned, list(commits))
        entries = [_event(ev, env.time_steps) for ev in now]
        extra = derived_events(frames[-1], frame)
        frame["events"].extend(extra)
        log.extend(entries + extra)
        frames.append(frame)
    dyn = {k: cast(getattr(config, k)) for k, cast in DYNAMICS}
    dyn["escort_enabled"] = bool(getattr(config, "escort_enabled", False))
    dyn["escort_radius"] = float(getattr(config, "escort_radius", 0.0) or 0.0)
    return {"metadata": {"title": title, "scenario": scenario, "algorithm": algorithm, "seed": seed,
                         "max_time_steps": int(config.max_time_steps),
                         "area": [float(env.area_width), float(env.area_height)], "dynamics": dyn},
            "events": log, "frames": frames, "final_metrics": frames[-1]["metrics"]}


def write_replay(doc: dict, path) -> None:
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    path.write_text(json.dumps(doc, indent=2), encoding="utf-8")
