"""TEST INFRASTRUCTURE (oracle) -- not part of the product path.

Restates the reference's CBBA market baseline over the oracle's flat state (oracle/sim.py):
  CBBA.allocate_tasks and its scoring helpers    TaskAllocation/MarketBased/CBBA.py:68-324
  CBBAReplan (periodic / event-triggered replan, a fresh CBBA(seed + n_replans) per replan)
                                                 TaskAllocation/MarketBased/CBBA_Replan.py:15-69
under the loops of experiments/wps_eval.py:134-146 (Local-CBBA-Replan, interval 20) and escort_eval.py:149-161
(Local-CBBA-Coalition, interval 12), both with max_tasks_per_agent = 1; bundles (max_tasks_per_agent > 1) as
CBBA.py:128-190 builds them.

The reference is deterministic only under a pinned string hash: every auction round starts from
`ordered_keys = list(remaining)` of a SET of slot-key strings (CBBA.py:116,128) and shuffles that list with its own
seeded generator.  oracle/pyset.py restates CPython's str hash and set order for PYTHONHASHSEED=0
(tests/test_pyset.py compares them with the interpreter), and the fixtures tests/golden/wps_*_cbba.json.gz (one task per
agent), wps_*_cbba2, wps_hard_cbba3 and wps_commit_cbba4 (bundles of two, three and four) were generated from the unmodified reference in an interpreter started with PYTHONHASHSEED=0 (tests/golden/gen_golden.py).

Arithmetic (float64, one rounding per operation, CBBA.py:288-309):
    time  = temp_time + dist / max(speed, 1e-6)
    score = -50                                                    if the task has a deadline and time > deadline
          = ((-2.5 * dist) / max(max_dist, 1)) + 160 * quality + 2 * (makespan - time)      if time < makespan
          = ((-2.5 * dist) / max(max_dist, 1)) + 160 * quality - 2 * (time - makespan)      otherwise
    quality = cap[task type], at least 1 for coalition tasks
"""
from __future__ import annotations

import random

from .fparith import norm2
from .hungarian import REPLAN_TAGS, is_coalition
from .market import agent_eligible, expand_slot_keys
from .pyset import PySet
from .sim import DURATION

INF = float("inf")


class OracleCBBA:
    def __init__(self, max_dist, seed=0):
        self.max_dist = max_dist
        self.rnd = random.Random(seed)
        self.makespan = 0

    # ---- scoring (CBBA.py:239-324) over paths of task ids (0 = the tentative task marker is resolved by the caller)
    def _task_score(self, env, a, tid, pos, time):
        k = tid - 1
        tp = env.k_pos[k]
        dist = norm2(pos[0] - tp[0], pos[1] - tp[1])
        quality = float(env.a_caps[a][env.k_type[k]])
        if is_coalition(env, k):
            quality = max(quality, 1.0)
        speed = max(float(env.a_speed[a] or 1.0), 1e-6)
        time = float(time) + dist / speed
        dl = env.k_deadline[k]
        if dl >= 0 and time > float(dl):
            return -50.0
        if time < self.makespan:
            return -2.5 * dist / max(self.max_dist, 1.0) + 160.0 * quality + 2.0 * (self.makespan - time)
        return -2.5 * dist / max(self.max_dist, 1.0) + 160.0 * quality - 2.0 * (time - self.makespan)

    def _score_path(self, env, a, tids):
        score = 0.0
        pos = env.a_pos[a]
        time = float(env.a_nft[a] or 0)
        for tid in tids:
            score += self._task_score(env, a, tid, pos, time)
            tp = env.k_pos[tid - 1]
            dist = norm2(pos[0] - tp[0], pos[1] - tp[1])
            speed = max(float(env.a_speed[a] or 1.0), 1e-6)
            pos = tp
            time += dist / speed + float(DURATION[env.k_type[tid - 1]])
        return score

    def _bid(self, env, a, tid, path_tids):
        best = -INF
        for i in range(len(path_tids) + 1):
            s = self._score_path(env, a, path_tids[:i] + [tid] + path_tids[i:])
            if s > best:
                best = s
        return best - self._score_path(env, a, path_tids)

    def _insertion(self, env, a, tid, path_tids):
        best, at = -INF, 0
        for i in range(len(path_tids) + 1):
            s = self._score_path(env, a, path_tids[:i] + [tid] + path_tids[i:])
            if s > best:
                best, at = s, i
        return at

    def _total_time(self, env, a, path_tids):
        pos = env.a_pos[a]
        time = float(env.a_nft[a] or 0)
        speed = max(float(env.a_speed[a] or 1.0), 1e-6)
        for tid in path_tids:
            tp = env.k_pos[tid - 1]
            dist = norm2(pos[0] - tp[0], pos[1] - tp[1])
            pos = tp
            time += dist / speed + float(DURATION[env.k_type[tid - 1]])
        return time

    def allocate(self, env, agents, tasks, known=None, reserved=None, max_tasks_per_agent=1):
        """Ordered [(agent_id, task_id)] (the reference's (name, [tasks]) list, flattened)."""
        reserved = set(reserved or ())
        live = [a for a in agents if env.a_state[a] != -1 and a not in reserved]
        if not live or not tasks:
            return []
        slots = expand_slot_keys(env, tasks)
        if not slots:
            return []
        slot_keys = [k for k, _ in slots]
        slot_task = dict(slots)
        remaining = PySet(slot_keys)
        bundles = {a: [] for a in live}
        paths = {a: [] for a in live}          # slot keys, in visiting order
        self.makespan = 0
        bids = {k: (None, -INF) for k in slot_keys}
        owned = {a: set() for a in live}
        tids_of = lambda a: [slot_task[k] for k in paths[a]]
        for _ in range(max(8, len(slot_keys) * 2)):
            if not remaining:
                break
            changed = False
            ordered = list(remaining)
            self.rnd.shuffle(ordered)
            agent_order = list(live)
            self.rnd.shuffle(agent_order)
            for key in ordered:
                tid = slot_task[key]
                for a in agent_order:
                    kn = None if known is None else known[a]
                    if not agent_eligible(env, a, tid, kn):
                        continue
                    if tid in owned[a]:
                        continue
                    if len(bundles[a]) >= max_tasks_per_agent and key not in bundles[a]:
                        continue
                    if key in bundles[a]:
                        continue
                    bid = self._bid(env, a, tid, tids_of(a))
                    if bid <= bids[key][1]:
                        continue
                    changed = True
                    prev = bids[key][0]
                    if prev is not None:
                        if key in paths[prev]:
                            paths[prev].remove(key)
                        if key in bundles[prev]:
                            bundles[prev].remove(key)
                        owned[prev].discard(tid)
                    bids[key] = (a, bid)
                    paths[a].insert(self._insertion(env, a, tid, tids_of(a)), key)
            if not changed:
                break
            to_remove = []
            for key in slot_keys:
                winner = bids[key][0]
                if winner is None:
                    continue
                tid = slot_task[key]
                for a in live:
                    if key in bundles[a] and a != winner:
                        bundles[a].remove(key)
                        if key in paths[a]:
                            paths[a].remove(key)
                        owned[a].discard(tid)
                if key not in bundles[winner]:
                    if tid in owned[winner] or len(bundles[winner]) >= max_tasks_per_agent:
                        bids[key] = (None, -INF)
                        if key in paths[winner]:
                            paths[winner].remove(key)
                        continue
                    bundles[winner].append(key)
                    owned[winner].add(tid)
                    to_remove.append(key)
            for key in to_remove:
                remaining.discard(key)
            self.makespan = max((self._total_time(env, a, tids_of(a)) for a in live), default=0.0)
        out = []
        for a in live:
            seen = set()
            for key in bundles[a]:
                tid = slot_task[key]
                if tid not in seen:
                    seen.add(tid)
                    out.append((a, tid))
        return out


class OracleCBBAReplan:
    def __init__(self, max_coord, seed=0, replan_interval=20):
        self.max_coord = max_coord
        self.seed = seed
        self.replan_interval = max(1, int(replan_interval))
        self.last_plan_step = -10**9
        self.n_replans = 0
        self.n_calls = 0

    def should_replan(self, time_step, events=None):
        if time_step - self.last_plan_step >= self.replan_interval:
            return True
        return any(ev[0] in REPLAN_TAGS for ev in (events or ()))

    def allocate(self, env, agents=None, tasks=None, time_step=0, events=None, force=False, known=None, reserved=None,
                 max_tasks_per_agent=1):
        from .hungarian import open_tasks

        self.n_calls += 1
        if not force and not self.should_replan(time_step, events):
            return []
        self.last_plan_step = time_step
        self.n_replans += 1
        cbba = OracleCBBA(self.max_coord, seed=self.seed + self.n_replans)
        return cbba.allocate(env, env.live_agents() if agents is None else agents,
                             open_tasks(env) if tasks is None else tasks, known=known, reserved=reserved,
                             max_tasks_per_agent=max_tasks_per_agent)
