"""Side-by-side run of the live reference and the oracle (authoring container only).

    python tests/golden/crosscheck.py WPS_hard local_hungarian 0 20

Steps both with the reference's actions and reports the first differing field;
also checks that the oracle allocator emits the reference's pairs.
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import refshim  # noqa: E402
import refsnap  # noqa: E402
import gen_golden  # noqa: E402


def crosscheck(case, driver, seed, verbose=False, over=None):
    refshim.install()
    from mUAV_TA.DroneEnv import MultiUAVEnv
    from TaskAllocation.OptimizationBased.HungarianAllocator import HungarianAllocator
    from oracle.sim import OracleEnv
    from oracle.hungarian import OracleHungarian, apply_assign

    cfg = refshim.wps_config(case)
    for k, v in (over or {}).items():
        setattr(cfg, k, v)
    env = MultiUAVEnv(cfg)
    obs, info = env.reset(seed=seed)
    orc = OracleEnv(cfg).reset(seed)
    d = refsnap.diff(refsnap.snapshot(env), orc.snapshot())
    if d:
        return f"reset: {d}"
    interval = 12 if driver == "coalition" else 20
    hung = HungarianAllocator(replan_interval=interval, max_coord=env.max_coord)
    ohung = OracleHungarian(replan_interval=interval, max_coord=orc.max_coord)
    import random
    rnd = random.Random(seed * 7919 + 13)
    while True:
        events = list(info.get("events") or []) if isinstance(info, dict) else []
        oevents = orc.last_events
        pairs = []
        if driver in ("local_hungarian", "coalition", "global_hungarian"):
            vis = env.agent_visibility_map() if driver != "global_hungarian" else None
            pairs = hung.allocate_tasks(env.get_live_agents(), gen_golden.ref_open_tasks(env),
                                        time_step=env.time_steps, events=events, agent_known_ids=vis)
            ovis = orc.visibility() if driver != "global_hungarian" else None
            opairs = ohung.allocate(orc, time_step=orc.t, events=oevents, known=ovis)
            rp = [(env.agent_by_name[n].id, t.id) for n, t in pairs]
            if rp != opairs:
                return f"t={env.time_steps}: allocator pairs differ ref={rp} oracle={opairs}"
        actions = {}
        if driver == "random_actions":
            n_open = len(env.last_tasks_info)
            for a in env.agents_obj:
                u = rnd.random()
                if u < 0.08:
                    actions[a.name] = rnd.randrange(0, max(n_open, 1))
                elif u < 0.09:
                    actions[a.name] = n_open + rnd.randrange(0, 3)
        else:
            for name, task in pairs:
                if env.last_tasks_info and task in env.last_tasks_info and name not in actions:
                    actions[name] = env.last_tasks_info.index(task)
        oact = [(env.agent_by_name[n].id, i) for n, i in actions.items()]
        if driver != "random_actions":
            oa2 = apply_assign(orc, opairs)
            if oa2 != oact:
                return f"t={env.time_steps}: apply_assign differs {oact} vs {oa2}"
        obs, rew, term, trunc, info = env.step(actions)
        r, oterm, otrunc, oev = orc.step(oact)
        d = refsnap.diff(refsnap.snapshot(env), orc.snapshot())
        r0 = next(iter(rew.values()))
        if r0 != r:
            d.append(f"reward {r0!r} != {r!r}")
        if d:
            return f"t={env.time_steps}: " + "; ".join(d[:6])
        if all(term.values()) or all(trunc.values()):
            break
    m = info["metrics"]
    om = orc.calculate_metrics()
    bad = [k for k in m if not (m[k] == om[k] or (m[k] != m[k] and om[k] != om[k]))]
    if bad:
        return "metrics differ: " + ", ".join(f"{k}: {m[k]!r} vs {om[k]!r}" for k in bad)
    return None


if __name__ == "__main__":
    case, driver = sys.argv[1], sys.argv[2]
    s0, s1 = int(sys.argv[3]), int(sys.argv[4])
    nbad = 0
    for seed in range(s0, s1):
        res = crosscheck(case, driver, seed)
        if res:
            nbad += 1
            print(f"seed {seed}: {res}")
    print(f"{case} {driver} seeds [{s0},{s1}): {nbad} mismatching episodes")
