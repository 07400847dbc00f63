"""TEST INFRASTRUCTURE (oracle) -- not part of the product path.

Deterministic hybrid planners restated over the oracle state:
  committed_names / apply_agent_commits   TaskAllocation/Hybrid/AttentionCommit.py:24-46
  UrgencyCommit.plan                      TaskAllocation/Hybrid/AttentionCommit.py:310-357
  _open_tasks_residual / _threat_stats    TaskAllocation/Hybrid/AttentionEscort.py:31-65
  UrgencyCoalition.plan                   TaskAllocation/Hybrid/AttentionEscort.py:720-767
  urgency_edge_scores / UrgencyPair.plan  TaskAllocation/Hybrid/PairCostHybrid.py:68-86, 520-550
  AttentionCommit._plan_from_scores       TaskAllocation/Hybrid/AttentionCommit.py:266-300
  AttentionEscort._plan_from_scores       TaskAllocation/Hybrid/AttentionEscort.py:500-517 (edge_score_dict :478-489)
"""
from __future__ import annotations

from .fparith import norm2
from .hungarian import is_coalition
from .sim import T_INT, T_REC, UAV_TYPES
from .tokens import urgency


def committed(env):
    return {a for a in env.live_agents() if int(env.a_commit_until[a] or 0) > env.t}


def apply_commits(env, agents, horizon):
    if horizon <= 0:
        return
    until = env.t + int(horizon)
    for a in agents:
        if env.a_state[a] == -1:
            continue
        if env.a_queue[a]:
            env.a_commit_until[a] = until


def urgency_commit_plan(env, hung, commit_fraction=0.35):
    A = env.n_agents
    T = len(env.k_pos)
    vis = env.visibility()
    live = env.live_agents()
    open_known = [k + 1 for k in range(T)
                  if env.k_status[k] != 2 and env.k_alloc[k][env.k_type[k]] < env.k_cur[k][env.k_type[k]]]
    reserved = committed(env)
    n = max(len(live), 1)
    pri = {}
    for tid in open_known:
        k = tid - 1
        urg = urgency(env, k, env.t)
        if vis is None:
            scar = 0.0
        else:
            cnt = sum(1 for a in range(A) if vis[a][k])
            scar = 1.0 - min(cnt / max(n, 1), 1.0)
        pri[tid] = 0.6 * urg + 0.4 * scar
    result = hung.allocate(env, agents=live, tasks=open_known, time_step=env.t, events=env.last_events, force=True,
                           task_priorities=pri, reserved=reserved, known=vis)
    assigned = {a for a, _ in result}
    free_assigned = [a for a in live if a in assigned and a not in reserved]
    scores = []
    thr = 1.0 - 12.0 / 40.0
    for a in free_assigned:
        urgent = [tid - 1 for tid in open_known
                  if (vis is None or vis[a][tid - 1]) and env.k_deadline[tid - 1] >= 0 and urgency(env, tid - 1, env.t) >= thr]
        if urgent:
            dmin = min(norm2(env.a_pos[a][0] - env.k_pos[k][0], env.a_pos[a][1] - env.k_pos[k][1]) for k in urgent)
        else:
            dmin = 0.0
        bonus = 500.0 if UAV_TYPES[env.a_type[a]] == "F2" else 0.0
        scores.append((dmin + bonus, env.a_name[a], a))
    scores.sort(key=lambda x: (x[0], x[1]), reverse=True)
    n_lock = max(1, int(round(commit_fraction * max(len(free_assigned), 1))))
    to_commit = [a for _, _, a in scores[:n_lock]]
    apply_commits(env, to_commit, int(env.commit_horizon or 25))
    return result


def threat_pressure(env, k):
    anchor = env.k_pos[k]
    prot = env.k_prot_agent[k]
    if prot >= 0:
        anchor = env.a_pos[prot]
    best = float(env.max_coord)
    for hid in env.h_order:
        if env.h_status[hid] == 2:
            continue
        d = norm2(env.h_pos[hid][0] - anchor[0], env.h_pos[hid][1] - anchor[1])
        best = min(best, d)
    return 1.0 - min(best / float(env.max_coord), 1.0)


def urgency_coalition_plan(env, hung):
    T = len(env.k_pos)
    max_coord = float(env.max_coord)
    open_tasks = []
    for k in range(T):
        if env.k_status[k] == 2:
            continue
        if is_coalition(env, k):
            if float(env.k_required_agents[k] or 1) - len(env.k_details[k]) > 0:
                open_tasks.append(k + 1)
        elif env.k_alloc[k][env.k_type[k]] < env.k_cur[k][env.k_type[k]]:
            open_tasks.append(k + 1)
    live = env.live_agents()
    edge = {}
    for a in live:
        atype = UAV_TYPES[env.a_type[a]]
        for tid in open_tasks:
            k = tid - 1
            el = env.k_elig[k]
            if el != 0 and not (el >> env.a_type[a]) & 1:
                continue
            ti = env.k_type[k]
            urg = urgency(env, k, env.t)
            pressure = threat_pressure(env, k)
            is_escort = 1.0 if env.k_kind[k] == 1 else 0.0
            cap = float(env.a_caps[a][ti]) if env.a_caps[a][ti] > 0 else 0.0
            dist = norm2(env.a_pos[a][0] - env.k_pos[k][0], env.a_pos[a][1] - env.k_pos[k][1]) / max_coord
            score = 0.45 * urg + 0.35 * pressure * (0.5 + 0.5 * is_escort) + 0.3 * min(cap, 1.0) - 0.25 * dist
            if atype.startswith("F") and (is_escort or ti == T_INT):
                score += 0.2
            if atype.startswith("R") and ti == T_REC:
                score += 0.2
            edge[(a, tid)] = float(min(max(score, 0.0), 1.0))
    reserved = committed(env)
    result = hung.allocate(env, agents=live, tasks=open_tasks, time_step=env.t, events=env.last_events, force=True,
                           reserved=reserved, known=env.visibility(), edge_scores=edge)
    apply_commits(env, [a for a, tid in result if tid != 0], int(env.commit_horizon or 0))
    return result


def att_commit_plan_from_scores(env, hung, pri_vec, com_vec, commit_threshold=0.5, max_tasks=32, max_agents=16):
    """Priorities / commit gates of a commit network -> Local-Hungarian on the free agents -> commit locks."""
    A = env.n_agents
    T = len(env.k_pos)
    vis = env.visibility()
    live = env.live_agents()
    open_known = [k + 1 for k in range(T)
                  if env.k_status[k] != 2 and env.k_alloc[k][env.k_type[k]] < env.k_cur[k][env.k_type[k]]]
    reserved = committed(env)
    n = max(len(live), 1)
    pri = {}
    for i, tid in enumerate(open_known[:max_tasks]):
        k = tid - 1
        urg = urgency(env, k, env.t)
        if vis is None:
            scar = 0.0
        else:
            cnt = sum(1 for a in range(A) if vis[a][k])
            scar = 1.0 - min(cnt / max(n, 1), 1.0)
        pri[tid] = 0.35 * urg + 0.40 * float(pri_vec[i]) + 0.25 * scar
    result = hung.allocate(env, agents=live, tasks=open_known, time_step=env.t, events=env.last_events, force=True,
                           task_priorities=pri, reserved=reserved, known=vis)
    assigned = {a for a, _ in result}
    to_commit = [a for i, a in enumerate(live[:max_agents])
                 if a not in reserved and a in assigned and float(com_vec[i]) >= commit_threshold]
    apply_commits(env, to_commit, int(env.commit_horizon or 25))
    return result


def att_escort_plan_from_scores(env, hung, scores, max_tasks=48, max_agents=16):
    """Edge scores in build_escort_tokens' layout -> Coalition-Hungarian over the token's (sorted, truncated)
    task list -> every assigned agent is locked."""
    from .tokens import build_escort_tokens

    tok = build_escort_tokens(env, max_tasks, max_agents)
    edge = {}
    for i, a in enumerate(tok["live"][:max_agents]):
        for j, tid in enumerate(tok["open_tasks"]):
            edge[(a, int(tid))] = float(scores[i, j])
    reserved = committed(env)
    result = hung.allocate(env, agents=env.live_agents(), tasks=tok["open_tasks"], time_step=env.t,
                           events=env.last_events, force=True, reserved=reserved, known=env.visibility(),
                           edge_scores=edge)
    apply_commits(env, [a for a, tid in result if tid != 0], int(env.commit_horizon or 0))
    return result


def urgency_pair_plan(env, hung, max_tasks=32, max_agents=16):
    """UrgencyPair.plan: hand-made edge residuals clip(0.5 urg + 0.3 scar - 0.4 dist, +-0.35) stored as float32."""
    import numpy as np

    from .tokens import build_pair_tokens

    tok = build_pair_tokens(env, max_tasks, max_agents)
    A = env.n_agents
    vis = tok["vis"]
    live = tok["live"]
    n_agents = max(len(live), 1)
    mc = max(float(env.max_coord), 1.0)
    edge = {}
    for i, a in enumerate(live[:max_agents]):
        for j, tid in enumerate(tok["open_tasks"]):
            if tok["edge_valid"][i, j] < 0.5:
                continue
            k = tid - 1
            urg = urgency(env, k, env.t)
            if vis is None:
                scar = 0.0
            else:
                cnt = sum(1 for b in range(A) if vis[b][k])
                scar = 1.0 - min(cnt / max(n_agents, 1), 1.0)
            dist = norm2(env.a_pos[a][0] - env.k_pos[k][0], env.a_pos[a][1] - env.k_pos[k][1]) / mc
            raw = 0.5 * urg + 0.3 * scar - 0.4 * dist
            edge[(a, int(tid))] = float(np.float32(min(max(raw, -0.35), 0.35)))
    return hung.allocate(env, agents=env.live_agents(), tasks=tok["open_tasks"], time_step=env.t, events=env.last_events,
                         force=True, known=vis, edge_scores=edge)
