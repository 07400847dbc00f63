"""TEST INFRASTRUCTURE (oracle) -- not part of the product path.

Restates, over the oracle's flat state (oracle/sim.py):
  build_att_tokens (raw=False)     TaskAllocation/Hybrid/AttentionRAH.py:50-173  (_urgency :29-34, _scarcity :37-41,
                                   _known_by_count :44-47)
  build_pair_tokens                TaskAllocation/Hybrid/PairCostHybrid.py:31-65
  edge_score_dict + plan(scores=)  TaskAllocation/Hybrid/PairCostHybrid.py:280-291, 308-328
  hybrid replan cadence            experiments/wps_eval.py:64-73
  _generate_observations / get_task_info / _event_flag_vector   mUAV_TA/DroneEnv.py:365-492
  build_att_tokens (raw=True), build_context_summary / build_context_pair_tokens
                                   AttentionRAH.py:86-97,140-146, TaskAllocation/Hybrid/ContextPairHybrid.py:33-78
  _expert_mask / _selected_mask    experiments/train_pair_cost.py:53-70, PairCostHybrid.py:293-306
  _open_tasks_residual / _threat_stats / _task_priority_key / build_escort_tokens
                                   TaskAllocation/Hybrid/AttentionEscort.py:31-241
"""
from __future__ import annotations

import numpy as np

from .fparith import norm2
from .sim import UAV_TYPES, T_ATT, T_INT, T_REC, EV_FAIL, EV_RESET, EV_THREAT


def urgency(env, k, t):
    dl = env.k_deadline[k]
    if dl < 0:
        return 0.0
    remaining = max(dl - t, 0)
    return 1.0 - min(remaining / 40.0, 1.0)


def hybrid_should_replan(env, events, interval=15):
    return env.t == 0 or env.t % interval == 0 or any(ev[0] in (EV_RESET, EV_THREAT, EV_FAIL) for ev in events)


def build_pair_tokens(env, max_tasks=32, max_agents=16, raw=False):
    """raw=True: per-entity attributes only (task 9 / agent 11 features, AttentionRAH.py:86-97,140-146)."""
    A = env.n_agents
    T = len(env.k_pos)
    max_coord = float(env.max_coord)
    horizon = max(env.max_time_steps, 1)
    mid_x = float(env.area_width) * 0.5
    vis = env.visibility()
    live = env.live_agents()
    n_agents = max(len(live), 1)
    specialists = [a for a in live if UAV_TYPES[env.a_type[a]] == "F2"]
    open_tasks = [k for k in range(T)
                  if env.k_status[k] != 2 and env.k_alloc[k][env.k_type[k]] < env.k_cur[k][env.k_type[k]]]
    task_feats = np.zeros((max_tasks, 9 if raw else 13), dtype=np.float32)
    task_mask = np.ones(max_tasks, dtype=bool)
    task_ids = []
    thr = 1.0 - 12.0 / 40.0
    for i, k in enumerate(open_tasks[:max_tasks]):
        ti = env.k_type[k]
        urg = urgency(env, k, env.t)
        if vis is None:
            scar = 0.0
            n_know = 1.0
        else:
            cnt = sum(1 for a in range(A) if vis[a][k])
            scar = 1.0 - min(cnt / max(n_agents, 1), 1.0)
            n_know = float(cnt)
        rem = max(float(env.k_cur[k][ti] - env.k_alloc[k][ti]), 0.0)
        is_dynamic = 1.0 if env.k_deadline[k] >= 0 else 0.0
        tp = env.k_pos[k]
        if specialists:
            d_spec = min(norm2(env.a_pos[a][0] - tp[0], env.a_pos[a][1] - tp[1]) for a in specialists)
        else:
            d_spec = max_coord
        region = 0.0 if float(tp[0]) < mid_x else 1.0
        if raw:
            dl = env.k_deadline[k]
            t_left = 1.0 if dl < 0 else min(max(dl - env.t, 0) / horizon, 1.0)
            task_feats[i] = [
                float(tp[0]) / max_coord, float(tp[1]) / max_coord, float(ti) / 8.0,
                1.0 if ti == T_ATT else 0.0, 1.0 if ti == T_REC else 0.0, 1.0 if ti == T_INT else 0.0,
                t_left, min(rem / 4.0, 1.0), is_dynamic,
            ]
        else:
            task_feats[i] = [
                float(tp[0]) / max_coord, float(tp[1]) / max_coord, float(ti) / 8.0,
                1.0 if ti == T_ATT else 0.0, 1.0 if ti == T_REC else 0.0, 1.0 if ti == T_INT else 0.0,
                urg, scar, min(rem / 4.0, 1.0), is_dynamic, min(n_know / max(n_agents, 1), 1.0),
                min(d_spec / max_coord, 1.0), region,
            ]
        task_mask[i] = False
        task_ids.append(k + 1)
    agent_feats = np.zeros((max_agents, 11 if raw else 12), dtype=np.float32)
    agent_mask = np.ones(max_agents, dtype=bool)
    for i, a in enumerate(live[:max_agents]):
        caps = env.a_caps[a]
        atype = UAV_TYPES[env.a_type[a]]
        n_known_urgent = 0
        for k in open_tasks:
            if vis is not None and not vis[a][k]:
                continue
            if urgency(env, k, env.t) >= thr and env.k_deadline[k] >= 0:
                n_known_urgent += 1
        base = [
            float(env.a_pos[a][0]) / max_coord, float(env.a_pos[a][1]) / max_coord,
            1.0 if atype.startswith("F") else 0.0, 1.0 if atype.startswith("R") else 0.0,
            1.0 if not env.a_queue[a] else 0.0,
            min(float(caps[2]) / 2.0, 1.0), min(float(caps[3]) / 2.0, 1.0), min(float(caps[1]) / 2.0, 1.0),
            float(env.a_state[a]) / 5.0, float(env.t) / horizon,
        ]
        if raw:
            agent_feats[i] = base + [1.0 if atype == "F2" else 0.0]
        else:
            agent_feats[i] = base + [min(n_known_urgent / max(len(open_tasks), 1), 1.0), 1.0 if atype == "F2" else 0.0]
        agent_mask[i] = False
    kept = open_tasks[:max_tasks]
    edge_valid = np.zeros((max_agents, max_tasks), dtype=np.float32)
    for i, a in enumerate(live[:max_agents]):
        for j, k in enumerate(kept):
            if vis is not None and not vis[a][k]:
                continue
            el = env.k_elig[k]
            if el != 0 and not (el >> env.a_type[a]) & 1:
                continue
            if float(env.a_caps[a][env.k_type[k]]) <= 0:
                continue
            edge_valid[i, j] = 1.0
    ids = np.zeros(max_tasks, dtype=np.int32)
    ids[: len(task_ids)] = task_ids
    return {"task_feats": task_feats, "task_mask": task_mask, "agent_feats": agent_feats, "agent_mask": agent_mask,
            "edge_valid": edge_valid, "task_ids": ids, "open_tasks": [k + 1 for k in kept], "live": live, "vis": vis}


def build_context_pair_tokens(env, max_tasks=32, max_agents=16, raw=False):
    """build_context_pair_tokens (ContextPairHybrid.py:73-78): pair tokens + the team / situation vector of
    build_context_summary (:33-70)."""
    tok = build_pair_tokens(env, max_tasks, max_agents, raw=raw)
    horizon = max(env.max_time_steps, 1)
    if raw:
        tok["context"] = np.asarray([float(env.t) / horizon], dtype=np.float32)
        return tok
    mid_x = float(env.area_width) * 0.5
    live = tok["live"]
    tasks = [tid - 1 for tid in tok["open_tasks"]]
    n_agents = max(len(live), 1)
    n_tasks = max(len(tasks), 1)
    n_urgent = left = right = 0
    for k in tasks:
        if urgency(env, k, env.t) >= (1.0 - 12.0 / 40.0) and env.k_deadline[k] >= 0:
            n_urgent += 1
        if float(env.k_pos[k][0]) < mid_x:
            left += 1
        else:
            right += 1
    free = sum(1 for a in live if not env.a_queue[a])
    fighters = sum(1 for a in live if UAV_TYPES[env.a_type[a]].startswith("F"))
    tok["context"] = np.asarray([
        n_urgent / n_tasks, min(len(tasks) / float(n_agents), 4.0) / 4.0, free / n_agents, fighters / n_agents,
        left / n_tasks, right / n_tasks, abs(left - right) / n_tasks, float(env.t) / horizon], dtype=np.float32)
    return tok


def commit_tokens(env, max_tasks=32, max_agents=16):
    """enrich_commit_tokens(build_att_tokens(env)) (AttentionCommit.py:49-62): agent feature 13 = commit-lock remainder."""
    tok = build_pair_tokens(env, max_tasks, max_agents)
    af = tok["agent_feats"]
    horizon = max(int(env.commit_horizon or 25), 1)
    extra = np.zeros((af.shape[0], 1), dtype=np.float32)
    for i, a in enumerate(tok["live"][: af.shape[0]]):
        rem = max(float(env.a_commit_until[a] or 0) - float(env.t), 0.0)
        extra[i, 0] = min(rem / horizon, 1.0)
    out = dict(tok)
    out["agent_feats"] = np.concatenate([af, extra], axis=1)
    out.pop("edge_valid")
    return out


def open_tasks_residual(env):
    """_open_tasks_residual (AttentionEscort.py:31-43) -> 0-based task indices."""
    out = []
    for k in range(len(env.k_pos)):
        if env.k_status[k] == 2:
            continue
        if env.k_kind[k] == 1 or float(env.k_required_agents[k] or 0) > 0:
            if float(env.k_required_agents[k] or 1) - len(env.k_details[k]) > 0:
                out.append(k)
        elif env.k_alloc[k][env.k_type[k]] < env.k_cur[k][env.k_type[k]]:
            out.append(k)
    return out


def threat_stats(env, k):
    """_threat_stats (AttentionEscort.py:46-65): (pressure, nearest threat distance / max_coord, fighter pressure)."""
    mc = float(env.max_coord)
    anchor = env.k_pos[k]
    prot = env.k_prot_agent[k]
    if prot >= 0:
        anchor = env.a_pos[prot]
    best = mc
    n_near = 0
    for hid in env.h_order:
        if env.h_status[hid] == 2:
            continue
        d = norm2(env.h_pos[hid][0] - anchor[0], env.h_pos[hid][1] - anchor[1])
        best = min(best, d)
        if d < 150.0:
            n_near += 1
    return 1.0 - min(best / mc, 1.0), min(best / mc, 1.0), min(n_near / 4.0, 1.0)


def escort_priority_key(env, k):
    """_task_priority_key (AttentionEscort.py:68-73)."""
    urg = urgency(env, k, env.t)
    pressure = threat_stats(env, k)[0]
    is_escort = 1.0 if env.k_kind[k] == 1 else 0.0
    is_int = 1.0 if env.k_type[k] == T_INT else 0.0
    return -(1.5 * urg + 1.2 * pressure + 0.8 * is_escort + 0.5 * is_int)


def build_escort_tokens(env, max_tasks=48, max_agents=16):
    """build_escort_tokens (AttentionEscort.py:76-241).  task_ids / open_tasks are in the priority-sorted order."""
    A = env.n_agents
    mc = float(env.max_coord)
    mid_x = float(env.area_width) * 0.5
    vis = env.visibility()
    live = env.live_agents()
    n_agents = max(len(live), 1)
    specialists = [a for a in live if UAV_TYPES[env.a_type[a]] == "F2"]
    open_all = open_tasks_residual(env)
    if vis is None:
        open_tasks = list(open_all)
    else:
        open_tasks = [k for k in open_all if any(vis[a][k] for a in live)]
        if not open_tasks:
            open_tasks = list(open_all)
    open_tasks.sort(key=lambda k: escort_priority_key(env, k))
    horizon = max(int(env.commit_horizon or 20), 1)
    t_now = float(env.t)
    task_feats = np.zeros((max_tasks, 22), dtype=np.float32)
    task_mask = np.ones(max_tasks, dtype=bool)
    kept = open_tasks[:max_tasks]
    for i, k in enumerate(kept):
        ti = env.k_type[k]
        urg = urgency(env, k, env.t)
        if vis is None:
            scar, n_know = 0.0, 0.0
        else:
            cnt = sum(1 for a in range(A) if vis[a][k])
            scar = 1.0 - min(cnt / max(n_agents, 1), 1.0)
            n_know = float(cnt)
        if env.k_kind[k] == 1 or float(env.k_required_agents[k] or 0) > 0:
            rem = max(float(env.k_required_agents[k] or 1) - len(env.k_details[k]), 0.0)
            req_agents = float(env.k_required_agents[k] or 1)
        else:
            rem = max(float(env.k_cur[k][ti] - env.k_alloc[k][ti]), 0.0)
            req_agents = 1.0
        is_dynamic = 1.0 if env.k_deadline[k] >= 0 else 0.0
        tp = env.k_pos[k]
        if specialists:
            d_spec = min(norm2(env.a_pos[a][0] - tp[0], env.a_pos[a][1] - tp[1]) for a in specialists)
        else:
            d_spec = mc
        region = 0.0 if float(tp[0]) < mid_x else 1.0
        deficit = min(rem / 4.0, 1.0)
        pressure, threat_dist, fighter_pressure = threat_stats(env, k)
        prot = env.k_prot_agent[k]
        prot_alive = 0.0
        if prot >= 0:
            prot_x, prot_y = float(env.a_pos[prot][0]) / mc, float(env.a_pos[prot][1]) / mc
            prot_alive = 0.0 if env.a_state[prot] == -1 else 1.0
        else:
            prot_x, prot_y = float(tp[0]) / mc, float(tp[1]) / mc
        task_feats[i] = [
            float(tp[0]) / mc, float(tp[1]) / mc, float(ti) / 8.0,
            1.0 if ti == T_ATT else 0.0, 1.0 if ti == T_REC else 0.0, 1.0 if ti == T_INT else 0.0,
            urg, scar, deficit, is_dynamic, min(n_know / max(n_agents, 1), 1.0), min(d_spec / mc, 1.0), region,
            1.0 if env.k_kind[k] == 1 else 0.0, deficit, pressure, prot_x, prot_y, min(req_agents / 4.0, 1.0),
            threat_dist, prot_alive, fighter_pressure,
        ]
        task_mask[i] = False
    agent_feats = np.zeros((max_agents, 16), dtype=np.float32)
    agent_mask = np.ones(max_agents, dtype=bool)
    edge_valid = np.zeros((max_agents, max_tasks), dtype=np.float32)
    thr = 1.0 - 12.0 / 40.0
    for i, a in enumerate(live[:max_agents]):
        caps = env.a_caps[a]
        atype = UAV_TYPES[env.a_type[a]]
        n_known_urgent = 0
        n_known_tasks = 0 if vis is None else int(sum(1 for v in vis[a] if v))
        for k in open_all:
            if vis is not None and not vis[a][k]:
                continue
            if urgency(env, k, env.t) >= thr and env.k_deadline[k] >= 0:
                n_known_urgent += 1
        is_escorting, dist_prot, near_escort = 0.0, 1.0, 0.0
        if env.a_queue[a] and env.k_kind[env.a_queue[a][0] - 1] == 1:
            is_escorting = 1.0
            prot = env.k_prot_agent[env.a_queue[a][0] - 1]
            if prot >= 0:
                dist_prot = min(norm2(env.a_pos[a][0] - env.a_pos[prot][0], env.a_pos[a][1] - env.a_pos[prot][1]) / mc, 1.0)
                near_escort = 1.0 - dist_prot
        rem_commit = max(float(env.a_commit_until[a] or 0) - t_now, 0.0)
        agent_feats[i] = [
            float(env.a_pos[a][0]) / mc, float(env.a_pos[a][1]) / mc,
            1.0 if atype.startswith("F") else 0.0, 1.0 if atype.startswith("R") else 0.0,
            1.0 if not env.a_queue[a] else 0.0,
            min(float(caps[2]) / 2.0, 1.0), min(float(caps[3]) / 2.0, 1.0), min(float(caps[1]) / 2.0, 1.0),
            float(env.a_state[a]) / 5.0, float(env.t) / max(env.max_time_steps, 1),
            min(n_known_urgent / 8.0, 1.0), 1.0 if atype == "F2" else 0.0, is_escorting, dist_prot,
            min(rem_commit / horizon, 1.0), min(near_escort + n_known_tasks / 16.0, 1.0),
        ]
        agent_mask[i] = False
        for j, k in enumerate(kept):
            if vis is not None and not vis[a][k]:
                continue
            el = env.k_elig[k]
            if el != 0 and not (el >> env.a_type[a]) & 1:
                continue
            edge_valid[i, j] = 1.0
    ids = np.zeros(max_tasks, dtype=np.int32)
    ids[: len(kept)] = [k + 1 for k in kept]
    return {"task_feats": task_feats, "task_mask": task_mask, "agent_feats": agent_feats, "agent_mask": agent_mask,
            "edge_valid": edge_valid, "task_ids": ids, "open_tasks": [k + 1 for k in kept], "live": live, "vis": vis}


def pair_mask(tok, pairs, require_valid):
    """pairs: [(agent_id, task_id)] of the allocator.  require_valid: _expert_mask, else _selected_mask."""
    mask = np.zeros(tok["edge_valid"].shape, dtype=np.float32)
    row_of = {a: i for i, a in enumerate(tok["live"][: mask.shape[0]]) if not tok["agent_mask"][i]}
    col_of = {int(tid): j for j, tid in enumerate(tok["open_tasks"])}
    for a, tid in pairs:
        i, j = row_of.get(a), col_of.get(int(tid))
        if i is None or j is None:
            continue
        if require_valid and tok["edge_valid"][i, j] < 0.5:
            continue
        mask[i, j] = 1.0
    return mask


def pair_plan(env, hung, scores, max_tasks=32, max_agents=16):
    """PairCostHybrid.plan(scores=...) -> ordered [(agent_id, task_id)]."""
    tok = build_pair_tokens(env, max_tasks, max_agents)
    edge = {}
    for i, a in enumerate(tok["live"][:max_agents]):
        for j, tid in enumerate(tok["open_tasks"]):
            if tok["edge_valid"][i, j] < 0.5:
                continue
            edge[(a, int(tid))] = float(scores[i, j])
    return hung.allocate(env, agents=env.live_agents(), tasks=tok["open_tasks"], time_step=env.t,
                         events=env.last_events, force=True, known=tok["vis"], edge_scores=edge)


def observe(env, max_rows=None):
    """Tensor form of _generate_observations (layout documented in include/muav.h, muav_observe)."""
    A = env.n_agents
    T = len(env.k_pos)
    mc = float(env.max_coord)
    mt = max(env.max_time_steps, 1)
    max_rows = int(max_rows or env.max_tasks)
    ti_out = np.zeros((max_rows, 21))
    ti_out[:, 3] = -1.0
    pad = np.zeros(max_rows, dtype=bool)
    legal = np.zeros((A, max_rows), dtype=bool)
    open_k = [k for k in range(T) if env.k_status[k] != 2]
    rows = open_k[:max_rows]
    for r, k in enumerate(rows):
        tt = env.k_type[k]
        o = ti_out[r]
        o[0] = k + 1
        o[1] = env.k_pos[k][0] / mc
        o[2] = env.k_pos[k][1] / mc
        o[3] = env.k_status[k]
        o[4:10] = env.k_cur[k]
        o[10:16] = env.k_alloc[k]
        o[16] = (env.k_init_time[k] - env.t) / mt
        o[17] = (env.k_done_time[k] - env.t) / mt
        o[18] = float(tt) / 6.0
        unmet = float(max(env.k_cur[k][tt] - env.k_alloc[k][tt], 0.0))
        o[19] = unmet / max(float(env.k_org[k][tt]), 1e-6)
        o[20] = min((env.t - float(env.k_created_at[k] or 0)) / mt, 1.0)
        pad[r] = True
    if not open_k:
        ti_out[0, 3] = 0.0
        pad[0] = True
        for a in range(A):
            head = env.a_queue[a][0] if env.a_queue[a] else 0
            legal[a, 0] = (head == 0) if env.a_state[a] == 2 else True
        n_rows = 1
    else:
        for a in range(A):
            head = env.a_queue[a][0] if env.a_queue[a] else 0
            m = [env._is_valid(a, k + 1) for k in rows]
            if not any(m):
                for r, k in enumerate(rows):
                    if k + 1 == head:
                        m[r] = True
                        break
                else:
                    m[0] = True
            if env.a_state[a] == 2:
                m = [(k + 1) == head for k in rows]
            legal[a, : len(rows)] = m
        n_rows = len(open_k)
    ao = np.zeros((A, 9))
    for a in range(A):
        ao[a, 0] = env.a_pos[a][0] / mc
        ao[a, 1] = env.a_pos[a][1] / mc
        ao[a, 2:8] = env.a_caps[a]
        ao[a, 8] = env.a_queue[a][0] if env.a_queue[a] else 0
    fail = threat = rst = 0.0
    for ev in env.events:
        if ev[0] == EV_FAIL:
            fail = 1.0
        elif ev[0] == EV_THREAT:
            threat = 1.0
        elif ev[0] == EV_RESET:
            rst = 1.0
    ef = np.asarray([fail, threat, rst, env.t / mt, len(open_k) / max(env.max_tasks, 1)], dtype=np.float32)
    return {"tasks_info": ti_out, "mask": pad, "legal_mask": legal, "agent_obs": ao, "event_flags": ef, "n_rows": n_rows}
