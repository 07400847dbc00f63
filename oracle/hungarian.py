"""TEST INFRASTRUCTURE (oracle) -- not part of the product path.

Restates TaskAllocation/OptimizationBased/HungarianAllocator.py:14-208 over the
oracle's flat state (oracle/sim.py), plus the driver glue
experiments/paper_eval.py:85-101 (_task_residual, _open_tasks) and
experiments/wps_eval.py:55-61 (_apply_assign: first pair per agent wins).

Cost expression, one rounding per operation (HungarianAllocator.py:65-70,177):
    c = (((dist/max(mc,1)) - 0.5*min(cap, missing)) - 0.4*pri) - 0.6*urg ; c = c - score
"""
from __future__ import annotations

from .fparith import norm2
from .lsap import lsap

REPLAN_TAGS = (0, 1, 2, 3, 4)  # every tag the reference emits triggers a replan (:33-39)


def is_coalition(env, k):
    """is_escort (HungarianAllocator.py:94-98)."""
    return env.k_kind[k] == 1 or float(env.k_required_agents[k] or 0) > 0


def residual_demand(env, k):
    """residual_demand (HungarianAllocator.py:100-111) == _task_residual (paper_eval.py:85-93)."""
    if is_coalition(env, k):
        required = float(env.k_required_agents[k] or 1)
        return max(required - len(env.k_details[k]), 0.0)
    ti = env.k_type[k]
    return max(float(env.k_cur[k][ti] - env.k_alloc[k][ti]), 0.0)


def open_tasks(env):
    """_open_tasks (paper_eval.py:96-101) -> task ids."""
    return [k + 1 for k in range(len(env.k_pos)) if env.k_status[k] != 2 and residual_demand(env, k) > 0]


class OracleHungarian:
    def __init__(self, replan_interval=20, max_coord=1000.0):
        self.replan_interval = max(1, int(replan_interval))
        self.max_coord = max_coord
        self.last_plan_step = -10**9
        self.n_replans = 0
        self.n_calls = 0
        self.trace = None  # optional list collecting (cost, nr, nc, rows, cols)

    def should_replan(self, time_step, events=None):
        if time_step - self.last_plan_step >= self.replan_interval:
            return True
        if events:
            for ev in events:
                if ev[0] in REPLAN_TAGS:
                    return True
        return False

    def allocate(self, env, agents=None, tasks=None, time_step=0, events=None, force=False,
                 task_priorities=None, reserved=None, known=None, edge_scores=None):
        """agents: agent ids (default live); tasks: task ids (default open_tasks);
        known: None or [A][T] 0/1 table; edge_scores: dict (agent_id, task_id) -> float;
        task_priorities: dict task_id -> float; reserved: set of agent ids.
        Returns ordered [(agent_id, task_id)]."""
        self.n_calls += 1
        if not force and not self.should_replan(time_step, events):
            return []
        if agents is None:
            agents = env.live_agents()
        if tasks is None:
            tasks = open_tasks(env)
        reserved = set(reserved or ())
        live = [a for a in agents if env.a_state[a] != -1 and a not in reserved]
        open_t = [tid for tid in tasks if env.k_status[tid - 1] != 2 and residual_demand(env, tid - 1) > 0]
        if not live or not open_t:
            return []
        pri = task_priorities or {}
        scores = edge_scores or {}
        residuals = {tid: residual_demand(env, tid - 1) for tid in open_t}
        free = list(live)
        actions = []
        mc = max(self.max_coord, 1.0)
        while free:
            round_tasks = [tid for tid in open_t if residuals[tid] > 1e-9]
            if not round_tasks:
                break
            nr, nc = len(free), len(round_tasks)
            cost = [1e6] * (nr * nc)
            for i, a in enumerate(free):
                ap = env.a_pos[a]
                for j, tid in enumerate(round_tasks):
                    k = tid - 1
                    if known is not None and not known[a][k]:
                        continue
                    el = env.k_elig[k]
                    if el != 0 and not (el >> env.a_type[a]) & 1:
                        continue
                    urgency = 0.0
                    dl = env.k_deadline[k]
                    if dl >= 0:
                        remaining = max(dl - time_step, 0)
                        urgency = 1.0 - min(remaining / 40.0, 1.0)
                    ti = env.k_type[k]
                    delivered = 1.0 if is_coalition(env, k) else float(env.a_caps[a][ti])
                    # _cost (HungarianAllocator.py:43-70)
                    if delivered <= 0:
                        base = 1e6
                    else:
                        tp = env.k_pos[k]
                        dist = norm2(ap[0] - tp[0], ap[1] - tp[1])
                        missing = max(float(residuals[tid]), 1e-6)
                        base = (dist / mc - 0.5 * min(delivered, missing)
                                - 0.4 * float(pri.get(tid, 0.0)) - 0.6 * float(urgency))
                    if base < 1e5 / 2:
                        cost[i * nc + j] = base - float(scores.get((a, tid), 0.0))
            rows, cols = lsap(cost, nr, nc)
            if self.trace is not None:
                self.trace.append((list(cost), nr, nc, list(rows), list(cols)))
            accepted = []
            for r, c in zip(rows, cols):
                if cost[r * nc + c] >= 1e5 / 2:
                    continue
                a = free[r]
                tid = round_tasks[c]
                k = tid - 1
                delivered = 1.0 if is_coalition(env, k) else float(env.a_caps[a][env.k_type[k]])
                actions.append((a, tid))
                residuals[tid] = max(residuals[tid] - delivered, 0.0)
                accepted.append(a)
            if not accepted:
                break
            acc = set(accepted)
            free = [a for a in free if a not in acc]
        self.last_plan_step = time_step
        self.n_replans += 1
        return actions


def apply_assign(env, pairs):
    """_apply_assign (wps_eval.py:55-61): ordered [(agent_id, idx into env.last_open)]."""
    actions = []
    seen = set()
    for a, tid in pairs:
        if env.last_open and tid in env.last_open:
            if a not in seen:
                seen.add(a)
                actions.append((a, env.last_open.index(tid)))
    return actions
