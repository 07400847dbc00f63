"""TEST INFRASTRUCTURE (oracle) -- not part of the product path.

Restates the Performance-Impact market allocator TaskAllocation/MarketBased/PerformanceImpact.py:27-311
(with the slot expansion and eligibility test it imports from MarketBased/CBBA.py:10-65) over the oracle's flat
state (oracle/sim.py).  Pinned by tests/golden/wps_{hard,commit,escort}_pi.json.gz (max_tasks_per_agent=1) and
wps_{hard,commit,escort}_pi2.json.gz (bundles of two), generated from the unmodified
reference class under the episode loops of experiments/wps_eval.py:147-159 / escort_eval.py:162-174.

The reference's other market baseline, CBBA / CBBAReplan (MarketBased/CBBA.py:68-324), is restated in oracle/cbba.py:
its auction order starts from `list(remaining)` of a *set of strings* (CBBA.py:116,128), i.e. from CPython's string
hash, so it is pinned for PYTHONHASHSEED=0 (oracle/pyset.py restates the hash and the set order).

Arithmetic (float64, one rounding per operation):
    start = max(next_free_time, t) + ||pos - task_pos|| / max(speed, 1e-6)             (_schedule :227-241)
    cost  = sum over the path of  start [+ 200 + (start - deadline) if start > deadline] - 5 * cap'   (_path_cost :243-261)
    cap'  = max(cap, 0.5) for coalition tasks, cap otherwise
"""
from __future__ import annotations

import math

from .fparith import norm2
from .hungarian import REPLAN_TAGS, is_coalition, residual_demand
from .sim import DURATION

INF = float("inf")


def expand_slot_keys(env, tasks):
    """expand_slot_keys (CBBA.py:46-65): [(key, task id)], one virtual auction slot per residual unit."""
    slots = []
    for tid in tasks:
        k = tid - 1
        if tid == 0 or env.k_status[k] == 2:
            continue
        rem = residual_demand(env, k)
        if rem <= 0:
            continue
        if is_coalition(env, k):
            for j in range(int(math.ceil(rem))):
                slots.append((f"{tid}#c{j}", tid))
        else:
            for j in range(max(1, int(math.ceil(min(rem, 4.0))))):
                slots.append((f"{tid}#r{j}", tid))
    return slots


def agent_eligible(env, a, tid, known):
    """agent_eligible (CBBA.py:27-43); known = the agent's row of the visibility table or None."""
    k = tid - 1
    if env.a_state[a] == -1:
        return False
    if known is not None and not known[k]:
        return False
    el = env.k_elig[k]
    if el != 0 and not (el >> env.a_type[a]) & 1:
        return False
    if a in env.k_details[k]:
        return False
    if is_coalition(env, k):
        return True
    return float(env.a_caps[a][env.k_type[k]]) > 0


class OraclePI:
    def __init__(self, max_coord=1000.0, seed=0, replan_interval=12, max_iters=40):
        self.max_coord = float(max_coord)
        self.replan_interval = max(1, int(replan_interval))
        self.max_iters = max(4, int(max_iters))
        self.last_plan_step = -10**9
        self.n_replans = 0
        self.n_calls = 0

    def should_replan(self, time_step, events=None):
        if time_step - self.last_plan_step >= self.replan_interval:
            return True
        return any(ev[0] in REPLAN_TAGS for ev in (events or ()))

    # ---- schedule / cost helpers over a path of (key, tid) pairs
    def _starts(self, env, a, path, t):
        pos = env.a_pos[a]
        now = max(float(env.a_nft[a] or 0), float(t))
        speed = max(float(env.a_speed[a] or 1.0), 1e-6)
        out = []
        for _key, tid in path:
            tp = env.k_pos[tid - 1]
            start = now + norm2(pos[0] - tp[0], pos[1] - tp[1]) / speed
            out.append(start)
            pos = tp
            now = start + float(DURATION[env.k_type[tid - 1]])
        return out

    def _path_cost(self, env, a, path, t):
        cost = 0.0
        for (_key, tid), start in zip(path, self._starts(env, a, path, t)):
            k = tid - 1
            cost += start
            dl = env.k_deadline[k]
            if dl >= 0 and start > float(dl):
                cost += 200.0 + (start - float(dl))
            cap = float(env.a_caps[a][env.k_type[k]])
            cost -= 5.0 * (max(cap, 0.5) if is_coalition(env, k) else cap)
        return cost

    def _feasible_prefix(self, env, a, path, t):
        keep = []
        for item, start in zip(path, self._starts(env, a, path, t)):
            dl = env.k_deadline[item[1] - 1]
            if dl >= 0 and start > float(dl) + 1e-6:
                break
            keep.append(item)
        return keep

    def _best_inclusion(self, env, a, path, tid, t):
        base = self._path_cost(env, a, path, t)
        best, at = INF, 0
        for i in range(len(path) + 1):
            mapped = path[:i] + [(f"{tid}#ins", tid)] + path[i:]
            if len(self._feasible_prefix(env, a, mapped, t)) != len(mapped):
                continue
            ipi = self._path_cost(env, a, mapped, t) - base
            if ipi < best - 1e-9:
                best, at = ipi, i
        return best, at

    def _removal_impact(self, env, a, path, key, t):
        if all(kk != key for kk, _ in path):
            return -INF
        return self._path_cost(env, a, path, t) - self._path_cost(env, a, [it for it in path if it[0] != key], t)

    def allocate(self, env, agents=None, tasks=None, time_step=0, events=None, force=False, known=None, reserved=None,
                 max_tasks_per_agent=1):
        """Returns ordered [(agent_id, task_id)] (the reference's (name, [tasks]) list, flattened)."""
        from .hungarian import open_tasks

        self.n_calls += 1
        if not force and not self.should_replan(time_step, events):
            return []
        agents = env.live_agents() if agents is None else agents
        tasks = open_tasks(env) if tasks is None else tasks
        reserved = set(reserved or ())
        live = [a for a in agents if env.a_state[a] != -1 and a not in reserved]
        slots = expand_slot_keys(env, tasks) if (live and tasks) else []
        if not slots:
            self.last_plan_step = time_step
            self.n_replans += 1
            return []
        t = time_step
        paths = {a: [] for a in live}
        winners = {key: (None, -INF) for key, _ in slots}
        assigned = set()
        single = max_tasks_per_agent <= 1
        for _ in range(len(slots) * max(len(live), 1)):
            best = None
            for a in live:
                if (a in assigned and single) or len(paths[a]) >= max_tasks_per_agent:
                    continue
                kn = None if known is None else known[a]
                owned = {tid for _, tid in paths[a]}
                for key, tid in slots:
                    cur_w, cur_rpi = winners[key]
                    if (cur_w is not None and cur_w == a) or tid in owned or not agent_eligible(env, a, tid, kn):
                        continue
                    ipi, at = self._best_inclusion(env, a, paths[a], tid, t)
                    if not math.isfinite(ipi):
                        continue
                    ins = f"{tid}#ins"
                    prov = self._removal_impact(env, a, paths[a][:at] + [(ins, tid)] + paths[a][at:], ins, t)
                    if cur_w is not None:
                        if prov < cur_rpi - 1e-9:
                            continue
                        if abs(prov - cur_rpi) <= 1e-9 and a >= cur_w:
                            continue
                    cand = (ipi, a, key, at)
                    if best is None or cand < best:
                        best = cand
            if best is None:
                break
            _ipi, a, key, at = best
            tid = dict(slots)[key]
            prev = winners[key][0]
            if prev is not None and prev != a:
                paths[prev] = [it for it in paths[prev] if it[0] != key]
                if single:
                    assigned.discard(prev)
            paths[a].insert(at, (key, tid))
            winners[key] = (a, self._removal_impact(env, a, paths[a], key, t))
            if single:
                assigned.add(a)
        # consensus clean-up (:168-205): unique winner per slot by max RPI, then drop infeasible tails
        for _ in range(self.max_iters):
            changed = False
            claimed = {key: [] for key, _ in slots}
            for a in live:
                for key, _tid in list(paths[a]):
                    claimed[key].append((a, self._removal_impact(env, a, paths[a], key, t)))
            for key, cl in claimed.items():
                if len(cl) <= 1:
                    if cl:
                        winners[key] = cl[0]
                    continue
                cl.sort(key=lambda x: (-x[1], x[0]))
                winners[key] = cl[0]
                for a, _r in cl[1:]:
                    if any(kk == key for kk, _ in paths[a]):
                        paths[a] = [it for it in paths[a] if it[0] != key]
                        changed = True
            for a in live:
                feas = self._feasible_prefix(env, a, paths[a], t)
                if feas != paths[a]:
                    for key, _tid in paths[a][len(feas):]:
                        if winners[key][0] == a:
                            winners[key] = (None, -INF)
                    paths[a] = feas
                    changed = True
            if not changed:
                break
        out = []
        for a in live:
            seen = set()
            for _key, tid in paths[a]:
                if tid not in seen:
                    seen.add(tid)
                    out.append((a, tid))
        self.last_plan_step = time_step
        self.n_replans += 1
        return out
