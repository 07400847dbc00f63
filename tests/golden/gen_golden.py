"""Generate golden fixtures by running the UNMODIFIED reference (/root/reference,
imported read-only through tests/golden/refshim.py) in the authoring container.

    python tests/golden/gen_golden.py            # writes tests/golden/*.json.gz

Each fixture holds, per episode: the ordered actions fed to env.step, the
allocator's ordered (agent_id, task_id) pairs, the drained events, the step
reward (float hex), termination flags, a 64-bit digest of the full canonical
state after every step (tests/golden/refsnap.py) and the terminal metrics dict.
Floats are stored as float.hex() strings, so fixtures are bit-exact.

Drivers (all restate the reference's own episode loops):
  local_hungarian   experiments/wps_eval.py:123-133 (interval 20, visibility map)
  coalition         experiments/escort_eval.py:137-148 (interval 12)
  global_hungarian  experiments/wps_eval.py:117-122 (no visibility mask)
  pair_injected     experiments/wps_eval.py:226-230 via PairCostHybrid.plan(scores=...) with
                    deterministic injected edge scores (PairCostHybrid.py:308-328), replan rule :64-73
  random_actions    env.step driven by a seeded random policy (valid, stale and out-of-range indices)
  urgency_commit    UrgencyCommit.plan (AttentionCommit.py:310-357) under the hybrid cadence wps_eval.py:64-73
  urgency_coalition UrgencyCoalition.plan (AttentionEscort.py:720-767) under escort_eval.py:52-58, interval 12
  urgency_pair      UrgencyPair.plan (PairCostHybrid.py:520-550) under the hybrid cadence wps_eval.py:64-73
  context_injected  ContextPairHybrid.plan(scores=...) (ContextPairHybrid.py:210-233 over PairCostHybrid.plan) with injected edge
                    scores; the reference's build_context_pair_tokens tensors (raw=False and raw=True) of every 4th plan
                    are stored too
  att_commit_injected  AttentionCommit.plan (AttentionCommit.py:260-300) with the network replaced by injected
                    (priority, commit) vectors, hybrid cadence wps_eval.py:64-73
  att_escort_injected  AttentionEscort.plan (AttentionEscort.py:519-524: build_escort_tokens -> act -> _plan_from_scores)
                    with the network replaced by injected logits, cadence escort_eval.py:52-58 (interval 12);
                    the reference tokens of every 4th plan are stored too
  local_pi / pi_coalition  PerformanceImpact.allocate_tasks(max_tasks_per_agent=1) (MarketBased/PerformanceImpact.py:59-224)
                    under experiments/wps_eval.py:147-159 (interval 20) / escort_eval.py:162-174 (interval 12)
"""
from __future__ import annotations

import gzip
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import refshim  # noqa: E402
import refsnap  # noqa: E402


def fhex(x):
    return float(x).hex()


def ref_open_tasks(env):
    def res(t):
        if getattr(t, "kind", None) == "Escort" or float(getattr(t, "required_agents", 0) or 0) > 0:
            return max(float(getattr(t, "required_agents", 1) or 1) - len(t.allocationDetails), 0.0)
        return max(float(t.currentReqs[t.typeIdx] - t.allocatedReqs[t.typeIdx]), 0.0)

    return [t for t in env.tasks if t.id != 0 and t.status != 2 and res(t) > 0]


def injected_scores(seed, t, n_rows, n_cols):
    """Deterministic pseudo-random edge scores in [-0.35, 0.35] (float32)."""
    i = np.arange(n_rows, dtype=np.uint64)[:, None]
    j = np.arange(n_cols, dtype=np.uint64)[None, :]
    x = (np.uint64(seed) * np.uint64(1000003) + np.uint64(t) * np.uint64(7919)
         + i * np.uint64(104729) + j * np.uint64(1299709) + np.uint64(12345))
    x = (x * np.uint64(2654435761)) % np.uint64(2001)
    return ((x.astype(np.float64) - 1000.0) / 1000.0 * 0.35).astype(np.float32)


def injected_commit_vectors(seed, t):
    """(pri_vec [32], com_vec [16]) in [0, 1] (float32), standing in for AttCommitNet's sigmoid heads."""
    m = injected_scores(seed, t, 2, 32)
    v = ((m.astype(np.float64) / 0.35 + 1.0) * 0.5).astype(np.float32)
    return v[0], v[1, :16]


def injected_logits(seed, t, n_rows, n_cols):
    """Pair logits in [-3.5, 3.5] (float32), standing in for AttCoalitionNet."""
    return (injected_scores(seed, t, n_rows, n_cols) * np.float32(10.0)).astype(np.float32)


class _InjectedNet:
    """Stands in for planner.net: returns what the episode loop put into `.out` (torch tensors, batch of 1)."""

    def __init__(self):
        self.out = None

    def eval(self):
        return self

    def __call__(self, *a, **k):
        return self.out


def tok_dump(tok):
    return {k: np.asarray(tok[k]).astype(np.float64).tolist() if k in ("task_feats", "agent_feats", "edge_valid")
            else [int(x) for x in np.asarray(tok[k]).reshape(-1)]
            for k in ("task_feats", "task_mask", "agent_feats", "agent_mask", "edge_valid", "task_ids") if k in tok}


def ctx_dump(tok):
    d = tok_dump(tok)
    d["context"] = np.asarray(tok["context"]).astype(np.float64).tolist()
    return d


def hybrid_should_replan(env, events, interval=15):
    return (env.time_steps == 0 or env.time_steps % interval == 0
            or any(ev[0] in ("Reset_Allocation", "New_Threat", "Agent_Fail") for ev in events))


CBBA_DRIVERS = ("cbba_replan", "cbba_coalition", "cbba2_replan", "cbba2_coalition", "cbba3_replan", "cbba4_replan")   # cbba<N>: bundles of N


def run_episode(case, seed, driver, overrides=None):
    refshim.install()
    from mUAV_TA.DroneEnv import MultiUAVEnv
    from TaskAllocation.OptimizationBased.HungarianAllocator import HungarianAllocator

    cfg = refshim.wps_config(case, **(overrides or {}))
    env = MultiUAVEnv(cfg)
    obs, info = env.reset(seed=seed)
    interval = 12 if driver == "coalition" else 20
    hung = HungarianAllocator(replan_interval=interval, max_coord=env.max_coord)
    pair = None
    if driver == "pair_injected":
        from TaskAllocation.Hybrid.PairCostHybrid import PairCostHybrid
        pair = PairCostHybrid(use_attention=False, device="cpu")
    planner = None
    if driver == "urgency_commit":
        from TaskAllocation.Hybrid.AttentionCommit import UrgencyCommit
        planner = UrgencyCommit()
    elif driver == "urgency_coalition":
        from TaskAllocation.Hybrid.AttentionEscort import UrgencyCoalition
        planner = UrgencyCoalition()
        hung = HungarianAllocator(replan_interval=10**9, max_coord=env.max_coord)
    elif driver == "urgency_pair":
        from TaskAllocation.Hybrid.PairCostHybrid import UrgencyPair
        planner = UrgencyPair()
    elif driver == "context_injected":
        from TaskAllocation.Hybrid.ContextPairHybrid import ContextPairHybrid, build_context_pair_tokens
        pair = ContextPairHybrid(use_attention=False, device="cpu")
    elif driver == "att_commit_injected":
        import torch
        from TaskAllocation.Hybrid.AttentionCommit import AttentionCommit
        planner = AttentionCommit(use_attention=False, device="cpu")
        planner.net = _InjectedNet()
    elif driver == "att_escort_injected":
        import torch
        from TaskAllocation.Hybrid.AttentionEscort import AttentionEscort
        planner = AttentionEscort(use_attention=False, device="cpu", d_model=16)
        planner.net = _InjectedNet()
        hung = HungarianAllocator(replan_interval=10**9, max_coord=env.max_coord)
    elif driver in ("local_pi", "pi_coalition", "local_pi2", "pi2_coalition"):
        from TaskAllocation.MarketBased.PerformanceImpact import PerformanceImpact
        planner = PerformanceImpact(max_coord=env.max_coord, seed=seed,
                                    replan_interval=20 if driver in ("local_pi", "local_pi2") else 12)
    elif driver in CBBA_DRIVERS:
        # Local-CBBA-Replan (wps_eval.py:105,134-146) / Local-CBBA-Coalition (escort_eval.py:108-112,149-161).  CBBA's
        # auction order starts from a set of strings: reproducible only with the string hash pinned
        assert os.environ.get("PYTHONHASHSEED") == "0", "generate the CBBA fixtures with PYTHONHASHSEED=0"
        from TaskAllocation.MarketBased.CBBA_Replan import CBBAReplan
        planner = CBBAReplan(env.agents_obj, env.tasks, env.max_coord, seed=seed,
                             replan_interval=12 if driver.endswith("coalition") else 20)
    n_plans = 0
    rnd = random.Random(seed * 7919 + 13)
    ep = {"case": case, "seed": seed, "driver": driver, "overrides": overrides or {},
          "agent_names": [a.name for a in env.agents_obj],
          "digest0": str(refsnap.digest(refsnap.snapshot(env))), "steps": []}
    while True:
        events = list(info.get("events") or []) if isinstance(info, dict) else []
        pairs = []
        tok_rec = None
        ctx_rec = None
        if driver in ("local_hungarian", "coalition"):
            res = hung.allocate_tasks(env.get_live_agents(), ref_open_tasks(env), time_step=env.time_steps,
                                      events=events, agent_known_ids=env.agent_visibility_map())
            pairs = res
        elif driver == "global_hungarian":
            pairs = hung.allocate_tasks(env.get_live_agents(), ref_open_tasks(env), time_step=env.time_steps,
                                        events=events)
        elif driver in ("local_pi", "pi_coalition", "local_pi2", "pi2_coalition"):
            # the *2 drivers: bundles of two tasks per agent (max_tasks_per_agent=2, PerformanceImpact.py:59-224 in full);
            # _apply_assign keeps the first task of every agent's path (wps_eval.py:55-61)
            res = planner.allocate_tasks(env.get_live_agents(), ref_open_tasks(env), time_step=env.time_steps, events=events,
                                         agent_known_ids=env.agent_visibility_map(),
                                         max_tasks_per_agent=2 if driver.endswith(("pi2", "pi2_coalition")) else 1)
            pairs = [(name, task) for name, tl in res for task in tl]   # _flatten_pairs (wps_eval.py:40-52)
        elif driver in CBBA_DRIVERS:
            res = planner.allocate_tasks(env.get_live_agents(), ref_open_tasks(env), time_step=env.time_steps, events=events,
                                         agent_known_ids=env.agent_visibility_map(),
                                         max_tasks_per_agent=int(driver[4]) if driver[4].isdigit() else 1)
            pairs = [(name, task) for name, tl in res for task in tl]
        elif driver == "pair_injected":
            if hybrid_should_replan(env, events):
                sc = injected_scores(seed, env.time_steps, pair.max_agents, pair.max_tasks)
                pairs = pair.plan(env, hung, events=events, explore=False, force=True, scores=sc)[0]
        elif driver == "urgency_commit":
            if hybrid_should_replan(env, events):
                pairs = planner.plan(env, hung, events=events, force=True)[0]
        elif driver == "urgency_coalition":
            if (env.time_steps == 0 or env.time_steps % 12 == 0 or any(
                    ev[0] in ("Reset_Allocation", "New_Threat", "Agent_Fail", "Escort_Created", "Escort_Retired")
                    for ev in events)):
                pairs = planner.plan(env, hung, events=events, force=True)
        elif driver == "urgency_pair":
            if hybrid_should_replan(env, events):
                pairs = planner.plan(env, hung, events=events, force=True)[0]
        elif driver == "context_injected":
            if hybrid_should_replan(env, events):
                sc = injected_scores(seed, env.time_steps, pair.max_agents, pair.max_tasks)
                if n_plans % 4 == 0:
                    ctx_rec = {"tok": ctx_dump(build_context_pair_tokens(env, 32, 16, raw=False)),
                               "raw": ctx_dump(build_context_pair_tokens(env, 32, 16, raw=True))}
                n_plans += 1
                pairs = pair.plan(env, hung, events=events, explore=False, force=True, scores=sc)[0]
        elif driver == "att_commit_injected":
            if hybrid_should_replan(env, events):
                pv, cv = injected_commit_vectors(seed, env.time_steps)
                planner.net.out = (torch.tensor(pv)[None], torch.tensor(cv)[None])
                pairs = planner.plan(env, hung, events=events, force=True)[0]
        elif driver == "att_escort_injected":
            if (env.time_steps == 0 or env.time_steps % 12 == 0 or any(
                    ev[0] in ("Reset_Allocation", "New_Threat", "Agent_Fail", "Escort_Created", "Escort_Retired")
                    for ev in events)):
                lg = injected_logits(seed, env.time_steps, planner.max_agents, planner.max_tasks)
                planner.net.out = (torch.tensor(lg)[None], torch.zeros(1))
                out = planner.plan(env, hung, events=events, explore=False, force=True)
                pairs = out[0]
                if n_plans % 4 == 0:
                    tok_rec = {"t": int(env.time_steps), "tok": tok_dump(out[1])}
                n_plans += 1
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
        obs, rew, term, trunc, info = env.step(actions)
        snap = refsnap.snapshot(env)
        r0 = next(iter(rew.values()))
        assert all(v == r0 for v in rew.values())
        ep["steps"].append({
            "pairs": [[env.agent_by_name[n].id, int(t.id)] for n, t in pairs],
            "actions": [[env.agent_by_name[n].id, int(i)] for n, i in actions.items()],
            "events": [[refsnap.EVENT_TAGS.index(e[0]), int(e[1])] for e in info["events"]],
            "reward": fhex(r0),
            "term": bool(all(term.values())), "trunc": bool(all(trunc.values())),
            "n_open": len(env.last_tasks_info),
            "digest": str(refsnap.digest(snap)),
        })
        if tok_rec is not None:
            ep["steps"][-1]["escort_tokens"] = tok_rec["tok"]
        if ctx_rec is not None:
            ep["steps"][-1]["context_tokens"] = ctx_rec["tok"]
            ep["steps"][-1]["context_tokens_raw"] = ctx_rec["raw"]
        if all(term.values()) or all(trunc.values()):
            break
    m = info["metrics"]
    ep["metrics"] = {k: (fhex(v) if isinstance(v, (float, np.floating)) else int(v)) for k, v in m.items()}
    ep["n_replans"] = int(planner.n_replans if driver in ("local_pi", "pi_coalition", "local_pi2", "pi2_coalition",
                                                          ) + CBBA_DRIVERS
                          else hung.n_replans)
    return ep


PLAN = [
    # (file, case, driver, seeds, overrides)
    ("wps_easy_local", "WPS_easy", "local_hungarian", range(0, 8), None),
    ("wps_hard_local", "WPS_hard", "local_hungarian", range(0, 16), None),
    ("wps_burst_local", "WPS_burst", "local_hungarian", range(0, 8), None),
    ("wps_commit_local", "WPS_commit", "local_hungarian", range(0, 8), None),
    ("wps_escort_coalition", "WPS_escort", "coalition", range(0, 8), None),
    ("wps_hard_global", "WPS_hard", "global_hungarian", range(0, 4), None),
    ("wps_hard_pair", "WPS_hard", "pair_injected", range(0, 12), None),
    ("wps_commit_pair", "WPS_commit", "pair_injected", range(0, 4), None),
    ("wps_hard_random", "WPS_hard", "random_actions", range(0, 8), None),
    ("wps_escort_random", "WPS_escort", "random_actions", range(0, 4), None),
    ("wps_attn_xl_local", "WPS_attn_XL", "local_hungarian", range(0, 2), None),
    ("wps_hard_single_task", "WPS_hard", "local_hungarian", range(0, 3), {"multiple_tasks_per_agent": False}),
    ("wps_commit_urgency", "WPS_commit", "urgency_commit", range(0, 8), None),
    ("wps_escort_urgency", "WPS_escort", "urgency_coalition", range(0, 6), None),
    ("wps_hard_obstacles", "WPS_hard", "local_hungarian", range(0, 4), {"num_obstacles": 4}),
    ("wps_hard_urgency_pair", "WPS_hard", "urgency_pair", range(0, 6), None),
    ("wps_attn_context", "WPS_attn", "context_injected", range(0, 4), None),
    ("wps_commit_attcommit", "WPS_commit", "att_commit_injected", range(0, 6), None),
    ("wps_escort_attescort", "WPS_escort", "att_escort_injected", range(0, 6), None),
    ("wps_hard_pi", "WPS_hard", "local_pi", range(0, 8), None),
    ("wps_commit_pi", "WPS_commit", "local_pi", range(0, 4), None),
    ("wps_escort_pi", "WPS_escort", "pi_coalition", range(0, 4), None),
    ("wps_hard_pi2", "WPS_hard", "local_pi2", range(0, 6), None),
    ("wps_commit_pi2", "WPS_commit", "local_pi2", range(0, 3), None),
    ("wps_escort_pi2", "WPS_escort", "pi2_coalition", range(0, 3), None),
    # CBBA: run this script with PYTHONHASHSEED=0 (set-of-strings iteration order, CBBA.py:116,128)
    ("wps_hard_cbba", "WPS_hard", "cbba_replan", range(0, 6), None),
    ("wps_commit_cbba", "WPS_commit", "cbba_replan", range(0, 3), None),
    ("wps_escort_cbba", "WPS_escort", "cbba_coalition", range(0, 3), None),
    ("wps_hard_cbba2", "WPS_hard", "cbba2_replan", range(0, 4), None),
    ("wps_commit_cbba2", "WPS_commit", "cbba2_replan", range(0, 2), None),
    ("wps_escort_cbba2", "WPS_escort", "cbba2_coalition", range(0, 2), None),
    ("wps_hard_cbba3", "WPS_hard", "cbba3_replan", range(4, 6), None),
    ("wps_commit_cbba4", "WPS_commit", "cbba4_replan", range(2, 3), None),
]


def main(only=None):
    for fname, case, driver, seeds, ov in PLAN:
        if only and fname not in only:
            continue
        eps = []
        for seed in seeds:
            cfg_over = dict(ov or {})
            ep = run_episode(case, seed, driver, None)if not cfg_over else run_episode_over(case, seed, driver, cfg_over)
            eps.append(ep)
        path = os.path.join(HERE, fname + ".json.gz")
        with gzip.open(path, "wt", compresslevel=9) as f:
            json.dump({"episodes": eps}, f, separators=(",", ":"))
        print(fname, len(eps), os.path.getsize(path))


def run_episode_over(case, seed, driver, over):
    """Overrides that wps_config applies AFTER its own multiple_tasks_per_agent=True."""
    orig = refshim.wps_config

    def patched(case_id, **kw):
        cfg = orig(case_id, **kw)
        for k, v in over.items():
            setattr(cfg, k, v)
        return cfg

    refshim.wps_config = patched
    try:
        ep = run_episode(case, seed, driver, None)
    finally:
        refshim.wps_config = orig
    ep["overrides"] = over
    return ep


if __name__ == "__main__":
    main(set(sys.argv[1:]) or None)
