"""The CUDA simulation core compiled for the CPU (tests/hostcheck) against the golden fixtures:
the same source the kernel runs, checked bit-exactly where no GPU exists.  The `-m gpu` suite
repeats these checks through the C ABI on the device."""
import numpy as np
import pytest

from helpers import (BUNDLE_DRIVERS, CBBA_DRIVERS, bundle_of, alloc_opts_for, assert_tokens_equal_reference, escort_scores_from_logits, golden_config, injected_commit_vectors, injected_logits,
                     injected_scores, load_golden)

STEP_CASES = ["wps_easy_local", "wps_hard_local", "wps_burst_local", "wps_commit_local", "wps_escort_coalition",
              "wps_hard_global", "wps_hard_pair", "wps_commit_pair", "wps_hard_random", "wps_escort_random",
              "wps_attn_xl_local", "wps_hard_single_task", "wps_hard_obstacles"]
# planner fixtures mutate commit_until between steps: they are replayed through the fused planner only
ALLOC_CASES = [c for c in STEP_CASES if "random" not in c and "obstacles" not in c] + [
    "wps_commit_urgency", "wps_escort_urgency", "wps_commit_attcommit", "wps_escort_attescort", "wps_hard_urgency_pair", "wps_attn_context",
    "wps_hard_pi", "wps_commit_pi", "wps_escort_pi", "wps_hard_cbba", "wps_commit_cbba", "wps_escort_cbba",
    "wps_hard_pi2", "wps_commit_pi2", "wps_escort_pi2", "wps_hard_cbba2", "wps_commit_cbba2", "wps_escort_cbba2", "wps_hard_cbba3", "wps_commit_cbba4"]


@pytest.mark.parametrize("name", STEP_CASES)
def test_step_with_reference_actions(hostcheck, name):
    eps = load_golden(name)
    env = hostcheck.make(golden_config(eps[0]), [ep["seed"] for ep in eps], queue_cap=16 if "random" in name else 8)
    for e, ep in enumerate(eps):
        assert str(env.digest(e)) == ep["digest0"]
    for t in range(len(eps[0]["steps"])):
        env.step_actions([[tuple(a) for a in ep["steps"][t]["actions"]] for ep in eps])
        for e, ep in enumerate(eps):
            st = ep["steps"][t]
            assert env.err(e) == 0
            assert env.reward[e] == float.fromhex(st["reward"]), (name, ep["seed"], t)
            assert env.events_of(e) == st["events"], (name, ep["seed"], t)
            assert (bool(env.term[e]), bool(env.trunc[e])) == (st["term"], st["trunc"])
            assert int(env.n_open[e]) == st["n_open"]
            assert str(env.digest(e)) == st["digest"], (name, ep["seed"], t)


@pytest.mark.parametrize("name", ALLOC_CASES)
def test_fused_allocator(hostcheck, name):
    eps = load_golden(name)
    drv = eps[0]["driver"]
    env = hostcheck.make(golden_config(eps[0]), [ep["seed"] for ep in eps])
    O = alloc_opts_for(drv)
    if drv in CBBA_DRIVERS:
        cbba_seeds = np.array([ep["seed"] for ep in eps], np.int32)
        O.d_cbba_seed = cbba_seeds.ctypes.data
    bundles = drv in BUNDLE_DRIVERS
    if bundles:   # the whole plan (every path entry) next to the step's first-task pairs
        A = len(eps[0]["agent_names"])
        bp = np.zeros((len(eps), A * bundle_of(drv)), np.int32)
        nbp = np.zeros(len(eps), np.int32)
        O.d_bundle_pairs, O.d_n_bundle_pairs = bp.ctypes.data, nbp.ctypes.data
    for t in range(len(eps[0]["steps"])):
        if drv in ("pair_injected", "context_injected"):
            sc = np.stack([injected_scores(ep["seed"], t, 16, 32) for ep in eps])
            O.d_edge_scores = sc.ctypes.data
            if any("context_tokens" in ep["steps"][t] for ep in eps):
                tok, raw = env.tokens_context(32, 16, False), env.tokens_context(32, 16, True)
                for e, ep in enumerate(eps):
                    if "context_tokens" in ep["steps"][t]:
                        assert_tokens_equal_reference(ep["steps"][t]["context_tokens"], tok, e, (ep["seed"], t))
                        assert_tokens_equal_reference(ep["steps"][t]["context_tokens_raw"], raw, e, (ep["seed"], t, "raw"))
        elif drv == "att_commit_injected":
            vec = [injected_commit_vectors(ep["seed"], t) for ep in eps]
            pv = np.stack([v[0] for v in vec])
            cv = np.stack([v[1] for v in vec])
            O.d_plan_pri, O.d_plan_commit = pv.ctypes.data, cv.ctypes.data
        elif drv == "att_escort_injected":
            tok = env.tokens_escort(48, 16)
            for e, ep in enumerate(eps):
                ref_tok = ep["steps"][t].get("escort_tokens")
                if ref_tok is None:
                    continue
                for k in ("task_feats", "agent_feats", "edge_valid"):
                    assert np.array_equal(np.asarray(ref_tok[k], np.float32), tok[k][e]), (ep["seed"], t, k)
                for k in ("task_mask", "agent_mask"):
                    assert [int(x) for x in tok[k][e]] == ref_tok[k], (ep["seed"], t, k)
                nk = len(ref_tok["task_ids"])
                assert [int(x) for x in tok["task_ids"][e][:nk]] == ref_tok["task_ids"] and not tok["task_ids"][e][nk:].any()
                assert [int(x) + 1 for x in tok["task_order"][e][:nk]] == ref_tok["task_ids"] and tok["task_order"][e][nk] == -1
            sc = np.stack([escort_scores_from_logits(injected_logits(ep["seed"], t, 16, 48), tok["edge_valid"][e],
                                                     tok["agent_mask"][e], tok["task_mask"][e]) for e, ep in enumerate(eps)])
            O.d_edge_scores, O.d_task_order = sc.ctypes.data, tok["task_order"].ctypes.data
        env.step_alloc(O)
        for e, ep in enumerate(eps):
            st = ep["steps"][t]
            if bundles:
                got = [[int(v) >> 16, int(v) & 0xFFFF] for v in bp[e, : nbp[e]]]
                assert got == st["pairs"], (name, ep["seed"], t, got, st["pairs"])
                first = []
                for a, k in st["pairs"]:
                    if a not in [p[0] for p in first]:
                        first.append([a, k])
                assert env.pairs_of(e) == first, (name, ep["seed"], t)
            else:
                assert env.pairs_of(e) == st["pairs"], (name, ep["seed"], t)
            assert env.reward[e] == float.fromhex(st["reward"])
            assert str(env.digest(e)) == st["digest"], (name, ep["seed"], t)
    if drv in ("local_hungarian", "coalition", "global_hungarian", "local_pi", "pi_coalition", "cbba_replan", "cbba_coalition",
               "local_pi2", "pi2_coalition") + CBBA_DRIVERS:
        for e, ep in enumerate(eps):
            assert env.codec.header(env.rec[e], "N_REPLANS") == ep["n_replans"]


def test_multi_step_launch_equals_single_steps(hostcheck):
    """n_steps = K in one call == K calls (the fused rollout keeps state resident)."""
    cfg = golden_config(load_golden("wps_hard_local")[0])
    a = hostcheck.make(cfg, range(4))
    b = hostcheck.make(cfg, range(4))
    O = alloc_opts_for("local_hungarian")
    a.step_alloc(O, n_steps=60)
    for _ in range(60):
        b.step_alloc(O)
    assert (a.rec == b.rec).all()


@pytest.mark.parametrize("case", ["static_strike", "D1_attrition", "D3_combined", "WPS_attn_AWACS", "WPS_attn_COP_cue_d12",
                                  "WPS_attn_OS24"])
def test_other_registered_scenarios_match_the_oracle(hostcheck, case):
    """Scenarios of experiments/paper_scenarios.py outside the WPS core set (legacy static / dynamic cases, the
    common-operating-picture sweeps and larger fleets of WPS_attn): kernel core vs oracle, Local-Hungarian, every step.
    (The oracle itself was cross-checked against the live reference on these cases, tests/golden/crosscheck.py.)"""
    from multi_uav_ta_gym_env_b200 import wps_config
    from oracle.hungarian import OracleHungarian, apply_assign
    from oracle.sim import OracleEnv
    import refsnap

    cfg = wps_config(case)
    seeds = [0, 1]
    env = hostcheck.make(cfg, seeds, queue_cap=16)
    O = alloc_opts_for("local_hungarian")
    oracles = [OracleEnv(cfg).reset(s) for s in seeds]
    hungs = [OracleHungarian(20, 1200.0) for _ in seeds]
    for t in range(150):
        env.step_alloc(O)
        for e, o in enumerate(oracles):
            pairs = hungs[e].allocate(o, time_step=o.t, events=o.last_events, known=o.visibility())
            o.step(apply_assign(o, pairs))
            assert env.err(e) == 0
            assert refsnap.digest(env.snapshot(e)) == refsnap.digest(o.snapshot()), (case, seeds[e], t)


@pytest.mark.parametrize("case,market,bundle,interval", [
    ("WPS_hard", "pi", 3, 20), ("WPS_commit", "pi", 4, 20), ("WPS_escort", "pi", 4, 12),
    ("WPS_hard", "cbba", 4, 20), ("WPS_commit", "cbba", 3, 20), ("WPS_escort", "cbba", 3, 12), ("WPS_escort", "cbba", 4, 12)])
def test_market_bundles_match_the_oracle_on_fresh_seeds(hostcheck, case, market, bundle, interval):
    """Bundles of three and four tasks per agent (the goldens hold twelve PI episodes with two and eleven CBBA episodes
    with two to four) on seeds outside the fixtures: kernel core vs the oracle that the goldens pin, whole plan (every
    bundle entry in the order allocate_tasks returns it) and state digest at every step."""
    from multi_uav_ta_gym_env_b200 import wps_config
    from oracle.cbba import OracleCBBAReplan
    from oracle.hungarian import apply_assign
    from oracle.market import OraclePI
    from oracle.sim import OracleEnv
    import refsnap

    cfg = wps_config(case)
    seeds = [2000 + 7 * bundle, 2001 + 7 * bundle, 2002 + 7 * bundle]
    if case == "WPS_escort" and market == "pi":
        seeds = seeds[:1]   # the oracle's escort PI with four-task paths takes ten seconds per episode
    env = hostcheck.make(cfg, seeds)
    O = alloc_opts_for("cbba_replan" if market == "cbba" else "local_pi")
    O.replan_interval = interval
    O.max_tasks_per_agent = bundle
    cbba_seeds = np.array(seeds, np.int32)
    O.d_cbba_seed = cbba_seeds.ctypes.data
    A = sum(cfg.agents.values())
    bp = np.zeros((len(seeds), A * bundle), np.int32)
    nbp = np.zeros(len(seeds), np.int32)
    O.d_bundle_pairs, O.d_n_bundle_pairs = bp.ctypes.data, nbp.ctypes.data
    oracles = [OracleEnv(cfg).reset(s) for s in seeds]
    planners = [OracleCBBAReplan(o.max_coord, s, interval) if market == "cbba" else OraclePI(o.max_coord, s, interval)
                for o, s in zip(oracles, seeds)]
    longest = 0
    for t in range(150):
        env.step_alloc(O)
        for e, o in enumerate(oracles):
            pairs = planners[e].allocate(o, time_step=o.t, events=o.last_events, known=o.visibility(), max_tasks_per_agent=bundle)
            got = [[int(v) >> 16, int(v) & 0xFFFF] for v in bp[e, : nbp[e]]]
            assert got == [list(p) for p in pairs], (case, market, seeds[e], t)
            for a in {p[0] for p in pairs}:
                longest = max(longest, sum(1 for p in pairs if p[0] == a))
            o.step(apply_assign(o, pairs))
            assert env.err(e) == 0
            assert refsnap.digest(env.snapshot(e)) == refsnap.digest(o.snapshot()), (case, market, seeds[e], t)
    assert longest >= 3   # the long bundles really occur
    for e in range(len(seeds)):
        assert env.codec.header(env.rec[e], "N_REPLANS") == planners[e].n_replans
