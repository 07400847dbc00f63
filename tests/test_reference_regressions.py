"""The reference's own regression tests (experiments/test_escort.py:19-319 -- its only test file, seven tests on
WPS_escort) restated for the drop-in facade.

The reference tests poke private methods and mutate Python objects (`env._create_escort_for`, `recon.position = ...`);
the facade's state lives in device records, so every test here reaches the same situation through the public surface --
reset / step / the allocator classes / the read-only proxies -- and asserts the SAME properties (ids stay unique; an
escort is a two-fighter task only F1 / F2 may take and follows its recon; the coalition Hungarian fills two slots, a
legacy task one; fighters near the recon are found; CBBA / PI respect visibility and eligibility and replan on escort
events; PI prefers the nearer fighter; escort tokens have the v2 shapes and one actor-critic update runs).  Each scenario
returns a summary that is (a) checked by asserts and (b) compared with the summary of the UNMODIFIED reference
environment and allocator classes driven by the same code (authoring container only; the reference tree does not travel).

CPU: facade over the CPU build of the kernel sources (tests/helpers.host_facade).  GPU: the same scenarios on the CUDA
backend (`-m gpu`), compared with the CPU build's summaries."""
import numpy as np
import pytest

import refshim
from helpers import host_facade

ESCORT_FLAGS_NO_MASKS = {"capability_mask": False, "saturate_mask": False}


# ----------------------------------------------------------------------------- environments under test
def facade_env(seed, cuda=False, **over):
    from multi_uav_ta_gym_env_b200 import wps_config
    from multi_uav_ta_gym_env_b200.env import MultiUAVEnv

    cfg = wps_config("WPS_escort", **over)
    env = MultiUAVEnv(cfg) if cuda else host_facade(cfg)
    _, info = env.reset(seed=seed)
    return env, info


def reference_env(seed, **over):
    refshim.install()
    from mUAV_TA.DroneEnv import MultiUAVEnv as RefEnv

    cfg = refshim.wps_config("WPS_escort")
    for k, v in over.items():
        setattr(cfg, k, v)
    env = RefEnv(cfg)
    _, info = env.reset(seed=seed)
    return env, info


def facade_classes():
    from multi_uav_ta_gym_env_b200.env import CBBA, CBBAReplan, HungarianAllocator, PerformanceImpact

    return HungarianAllocator, PerformanceImpact, CBBAReplan, CBBA


def reference_classes():
    refshim.install()
    from TaskAllocation.MarketBased.CBBA import CBBA
    from TaskAllocation.MarketBased.CBBA_Replan import CBBAReplan
    from TaskAllocation.MarketBased.PerformanceImpact import PerformanceImpact
    from TaskAllocation.OptimizationBased.HungarianAllocator import HungarianAllocator

    return HungarianAllocator, PerformanceImpact, CBBAReplan, CBBA


def open_tasks(env):
    """_open_tasks (experiments/paper_eval.py:96-101)."""
    out = []
    for t in env.tasks:
        if t.id == 0 or t.status == 2:
            continue
        if float(t.orgReqs[t.typeIdx] - t.allocatedReqs[t.typeIdx] - t.doneReqs[t.typeIdx]) > 0 or getattr(t, "kind", None) == "Escort":
            out.append(t)
    return out


def to_actions(env, pairs):
    """_apply_assign (experiments/wps_eval.py:55-61): first pair of an agent wins."""
    actions = {}
    for name, task in pairs:
        if env.last_tasks_info and task in env.last_tasks_info and name not in actions:
            actions[name] = env.last_tasks_info.index(task)
    return actions


def events_of(info):
    return list(info.get("events") or []) if isinstance(info, dict) else []


# ----------------------------------------------------------------------------- scenarios (one per reference test)
def scenario_unique_task_ids(env, info, classes):
    """test_unique_task_ids (test_escort.py:19-47): 80 forced Local-Hungarian replans, ids unique after every step."""
    hung = classes[0](replan_interval=5, max_coord=env.max_coord)
    counts = []
    for _ in range(80):
        result = hung.allocate_tasks(env.get_live_agents(), open_tasks(env), time_step=env.time_steps, events=[], force=True,
                                     agent_known_ids=env.agent_visibility_map())
        _, _, done, trunc, info = env.step(to_actions(env, result))
        ids = [t.id for t in env.tasks]
        assert len(ids) == len(set(ids)), f"duplicate task ids: {ids}"
        counts.append(len(ids))
        if all(done.values()) or all(trunc.values()):
            break
    assert counts[-1] > counts[0]   # escorts / arrivals did create tasks
    return counts


def coalition_rollout(env, info, classes, steps=150, interval=12, on_step=None):
    """Coalition-Hungarian loop of experiments/escort_eval.py:137-148."""
    hung = classes[0](replan_interval=interval, max_coord=env.max_coord)
    for _ in range(steps):
        events = events_of(info)
        result = hung.allocate_tasks(env.get_live_agents(), open_tasks(env), time_step=env.time_steps, events=events,
                                     agent_known_ids=env.agent_visibility_map())
        if on_step:
            on_step("plan", result, events)
        _, _, done, trunc, info = env.step(to_actions(env, result))
        if on_step:
            on_step("step", None, events_of(info))
        if all(done.values()) or all(trunc.values()):
            break
    return info


def scenario_escort_lifecycle_and_follow(env, info, classes):
    """test_escort_lifecycle_and_follow (test_escort.py:50-78): an escort is kind 'Escort', only F1 / F2 may take it, its
    recon cannot escort itself, it sits on its recon after every step, and it is closed (status 2) and forgotten by
    _escort_by_recon once retired."""
    log = {"created": 0, "retired": 0, "followed": 0, "seen": []}
    fighters = [a for a in env.agents_obj if a.type == "F1"]

    def on_step(phase, result, events):
        if phase != "step":
            return
        for tag, arg in events:
            if tag == "Escort_Created":
                log["created"] += 1
                log["seen"].append(int(arg))
            if tag == "Escort_Retired":
                log["retired"] += 1
                t = next(t for t in env.tasks if t.id == arg)
                assert t.status == 2
                assert all(e.id != arg for e in env._escort_by_recon.values())
        for recon_name, escort in env._escort_by_recon.items():
            if escort.status == 2:
                continue
            recon = env.agent_by_name[recon_name]
            assert escort.kind == "Escort"
            assert set(escort.eligible_agent_types) == {"F1", "F2"}
            assert recon.type.startswith("R")
            assert not env._is_task_action_valid(recon, escort)        # a recon cannot escort itself
            assert all(env._is_task_action_valid(f, escort) for f in fighters if f.state != -1)
            assert np.allclose(escort.position, recon.position)         # _sync_escorts: the escort follows its recon
            log["followed"] += 1

    coalition_rollout(env, info, classes, on_step=on_step)
    assert log["created"] > 0 and log["retired"] > 0 and log["followed"] > 0
    return log


def scenario_coalition_two_slot_and_legacy(env, info, classes):
    """test_coalition_two_slot_and_legacy (test_escort.py:81-116): the coalition Hungarian hands an escort to two fighters
    in ONE call; one fighter offered one legacy (non-coalition) task gives at most one pair."""
    log = {"two_slot_calls": 0, "max_per_legacy_task": 0, "plans": 0}
    fighter = next(a for a in env.get_live_agents() if a.type in ("F1", "F2"))
    att = next(t for t in env.tasks if t.type == "Att")
    single = classes[0](replan_interval=1, max_coord=env.max_coord).allocate_tasks([fighter], [att], time_step=env.time_steps,
                                                                                   force=True)
    assert len(single) <= 1
    log["single"] = [(n, t.id) for n, t in single]

    def on_step(phase, result, events):
        if phase == "step" and "synthetic" not in log:
            # test_escort.py:90-105: three fighters offered one escort with edge scores 1.0 -> at least two assignments;
            # here additionally with a caller-made visibility map in which the third fighter does not know the escort
            escorts = [t for t in env.tasks if getattr(t, "kind", None) == "Escort" and t.status != 2 and t.required_agents >= 2]
            fighters = [a for a in env.get_live_agents() if a.type in ("F1", "F2")][:3]
            if escorts and len(fighters) == 3:
                escort = escorts[0]
                hung = classes[0](replan_interval=1, max_coord=env.max_coord)
                scores = {(f.name, escort.id): 1.0 for f in fighters}
                both = hung.allocate_tasks(fighters, [escort], time_step=env.time_steps, force=True, edge_scores=scores)
                two = hung.allocate_tasks(fighters, [escort], time_step=env.time_steps, force=True, edge_scores=scores,
                                          agent_known_ids={f.name: {escort.id} for f in fighters[:2]})
                assert len(both) >= 2, f"expected >=2 escort assigns, got {both}"
                assert all(n != fighters[2].name for n, _ in two)
                log["synthetic"] = (env.time_steps, [(n, t.id) for n, t in both], [(n, t.id) for n, t in two])
        if phase != "plan" or not result:
            return
        log["plans"] += 1
        per_task = {}
        for name, task in result:
            per_task.setdefault(task.id, []).append(name)
        for tid, names in per_task.items():
            task = next(t for t in env.tasks if t.id == tid)
            assert len(names) == len(set(names))
            if getattr(task, "kind", None) == "Escort":
                assert all(env.agent_by_name[n].type in ("F1", "F2") for n in names)
                assert len(names) <= max(1, int(task.required_agents))
                log["two_slot_calls"] += len(names) >= 2
            elif int(getattr(task, "required_agents", 0) or 0) <= 1:
                log["max_per_legacy_task"] = max(log["max_per_legacy_task"], len(names))

    coalition_rollout(env, info, classes, on_step=on_step)
    assert log["two_slot_calls"] >= 1, "expected an escort taken by >= 2 fighters in one call"
    assert "synthetic" in log
    return log


def scenario_threat_diversion_inputs(env, info, classes):
    """test_threat_diversion_inputs (test_escort.py:119-138): with two fighters on the escort and next to the recon,
    _escort_fighters_near(recon) returns both, nearest first."""
    log = {"max_near": 0, "trace": []}

    def on_step(phase, result, events):
        if phase != "step":
            return
        for recon_name in sorted(env._escort_by_recon):
            recon = env.agent_by_name[recon_name]
            near = env._escort_fighters_near(recon)
            d = [float(np.linalg.norm(a.position - recon.position)) for a in near]
            assert d == sorted(d) and all(x <= env.escort_radius for x in d)
            assert all(a.type in ("F1", "F2") and a.tasks[0].id == env._escort_by_recon[recon_name].id for a in near)
            log["max_near"] = max(log["max_near"], len(near))
            if near:
                log["trace"].append((env.time_steps, recon_name, [a.name for a in near]))

    coalition_rollout(env, info, classes, on_step=on_step)
    assert log["max_near"] >= 2, log["max_near"]
    return log


def market_rollout(env, info, planner, steps, log, make_probe):
    """Local-PI-Coalition / Local-CBBA-Coalition loop (experiments/escort_eval.py:149-174) with the reference test's
    assertions on every plan: visibility, eligibility, the recon never escorts, one task per agent."""
    for _ in range(steps):
        events = events_of(info)
        known = env.agent_visibility_map()
        out = planner.allocate_tasks(env.get_live_agents(), open_tasks(env), time_step=env.time_steps, events=events,
                                     agent_known_ids=known, max_tasks_per_agent=1)
        assigned = [(n, t) for n, ts in out for t in ts]
        names = [n for n, _ in assigned]
        assert len(names) == len(set(names)), "no duplicate agent assignment"
        for n, t in assigned:
            assert t.id in known[n], (n, t.id)
            el = getattr(t, "eligible_agent_types", None)
            if el:
                assert env.agent_by_name[n].type in el
            if getattr(t, "kind", None) == "Escort":
                assert not n.startswith("R"), "recon must not escort"
                log["escort_assigns"] += 1
        if assigned:
            log["plans"].append([(n, t.id) for n, t in assigned])
            assert to_actions(env, assigned), "expected convertible actions"
        if "restricted" not in log:
            # test_escort.py:160-176: only the first two fighters know the escort -> nobody else may be assigned to it
            # (a caller-made agent_known_ids map, not the environment's own; a name missing from the map knows nothing)
            escorts = [t for t in env.tasks if getattr(t, "kind", None) == "Escort" and t.status != 2]
            fighters = [a for a in env.get_live_agents() if a.type in ("F1", "F2")]
            if escorts and len(fighters) >= 3:
                escort = escorts[0]
                only = {f.name: set() for f in fighters[1:]}
                only[fighters[1].name].add(escort.id)
                only[fighters[2].name].add(escort.id)
                probe = make_probe()
                res = probe.allocate_tasks(fighters, [escort], time_step=env.time_steps, force=True, agent_known_ids=only,
                                           max_tasks_per_agent=1)
                got = [(n, t.id) for n, ts in res for t in ts]
                assert all(n in (fighters[1].name, fighters[2].name) and tid == escort.id for n, tid in got), got
                log["restricted"] = (env.time_steps, escort.id, sorted(got))
                assert env.agent_visibility_map() == known   # the environment's own sets are back in place
        _, _, done, trunc, info = env.step(to_actions(env, assigned))
        if all(done.values()) or all(trunc.values()):
            break
    return info


class _BareCBBA:
    """The bare CBBA class behind the probe's call shape (its allocate_tasks has no events / force arguments)."""

    def __init__(self, cbba):
        self.cbba = cbba

    def allocate_tasks(self, agents, tasks, time_step=0, force=True, agent_known_ids=None, max_tasks_per_agent=1):
        return self.cbba.allocate_tasks(agents, tasks, agent_known_ids=agent_known_ids, max_tasks_per_agent=max_tasks_per_agent)


def scenario_pi_coalition_eligibility_visibility(env, info, classes):
    """test_cbba_pi_coalition_eligibility_visibility (test_escort.py:141-241), Performance-Impact half."""
    pi = classes[1](max_coord=env.max_coord, seed=3, replan_interval=12)
    log = {"escort_assigns": 0, "plans": []}
    market_rollout(env, info, pi, 100, log, lambda: classes[1](max_coord=env.max_coord, seed=3, replan_interval=1))
    assert len(log["restricted"][2]) >= 1
    assert log["escort_assigns"] >= 2
    assert pi.should_replan(10, [["Escort_Created", 7]]) and pi.should_replan(10, [["Escort_Retired", 7]])
    return log


def scenario_cbba_coalition_eligibility_visibility(env, info, classes):
    """test_cbba_pi_coalition_eligibility_visibility (test_escort.py:141-241), CBBA half; the plans themselves depend on
    the interpreter's string hash in the reference (CBBA.py:116,128), so only the properties are compared."""
    rp = classes[2](env.agents_obj, env.tasks, env.max_coord, seed=0, replan_interval=12)
    log = {"escort_assigns": 0, "plans": []}
    market_rollout(env, info, rp, 60, log, lambda: _BareCBBA(classes[3](env.agents_obj, env.tasks, env.max_coord, seed=1)))
    assert len(log["restricted"][2]) >= 1
    assert log["escort_assigns"] >= 2
    far = classes[2](env.agents_obj, env.tasks, env.max_coord, seed=0, replan_interval=1000)
    far.last_plan_step = 0
    assert far.should_replan(10, [["Escort_Created", 7]]) and far.should_replan(10, [["Escort_Retired", 7]])
    assert not far.should_replan(10, [])
    return log


def place(env, agent, xy):
    """`agent.position = xy` of the reference tests: on the facade the two position fields of the record are written
    and the proxies refreshed."""
    if hasattr(env, "_backend"):
        env._backend.patch_field("a_posx", agent.id, float(xy[0]))
        env._backend.patch_field("a_posy", agent.id, float(xy[1]))
        env._sync()
    else:
        agent.position = np.array([float(xy[0]), float(xy[1])])


def scenario_pi_prefers_nearer(env, info, classes):
    """test_pi_schedule_impact_prefers_nearer (test_escort.py:244-276): an F1 at 5 and an F2 at 400 from an escort, both
    idle and both knowing it (a caller-made visibility map) -> the near one is in PI's plan.  The escort (both fighter
    slots still open, as in the reference test) comes from the rollout instead of env._create_escort_for."""
    hung = classes[0](replan_interval=12, max_coord=env.max_coord)
    log = {}
    for _ in range(60):
        result = hung.allocate_tasks(env.get_live_agents(), open_tasks(env), time_step=env.time_steps, events=events_of(info),
                                     agent_known_ids=env.agent_visibility_map())
        _, _, done, trunc, info = env.step(to_actions(env, result))
        escorts = [t for t in env.tasks if getattr(t, "kind", None) == "Escort" and t.status != 2
                   and int(t.required_agents) - len(t.allocationDetails) >= 2]
        idle = [a for a in env.get_live_agents() if a.type in ("F1", "F2") and a.tasks[0].id == 0]
        near = next((a for a in idle if a.type == "F1"), None)
        far = next((a for a in idle if a.type == "F2"), None)
        if not escorts or near is None or far is None:
            continue
        escort = escorts[0]
        ex, ey = float(escort.position[0]), float(escort.position[1])
        place(env, near, (ex + 5.0, ey))
        place(env, far, (ex + 400.0 if ex < 600.0 else ex - 400.0, ey))
        known = {near.name: {escort.id}, far.name: {escort.id}}
        pi = classes[1](max_coord=env.max_coord, seed=0, replan_interval=1)
        out = pi.allocate_tasks([near, far], [escort], time_step=env.time_steps, force=True, agent_known_ids=known,
                                max_tasks_per_agent=1)
        names = [n for n, _ in out]
        assert near.name in names, f"near fighter should be chosen, got {out}"
        slots = int(escort.required_agents) - len(escort.allocationDetails)
        log = {"t": env.time_steps, "escort": escort.id, "names": names, "slots": slots,
               "near": [float(x) for x in near.position], "far": [float(x) for x in far.position]}
        break
    assert log, "no escort with two idle fighters in 60 steps"
    return log


SCENARIOS = [
    # (scenario, seed of the reference test, config overrides)
    (scenario_unique_task_ids, 7, ESCORT_FLAGS_NO_MASKS),
    (scenario_escort_lifecycle_and_follow, 3, ESCORT_FLAGS_NO_MASKS),
    (scenario_coalition_two_slot_and_legacy, 1, {}),
    (scenario_threat_diversion_inputs, 2, {}),
    (scenario_pi_coalition_eligibility_visibility, 5, {}),
    (scenario_cbba_coalition_eligibility_visibility, 5, {}),
    (scenario_pi_prefers_nearer, 9, {}),
]
IDS = [s[0].__name__[9:] for s in SCENARIOS]


@pytest.mark.parametrize("scenario,seed,over", SCENARIOS, ids=IDS)
def test_reference_regression_on_the_facade(scenario, seed, over):
    env, info = facade_env(seed, **over)
    mine = scenario(env, info, facade_classes())
    if not refshim.reference_available():
        return
    renv, rinfo = reference_env(seed, **over)
    ref = scenario(renv, rinfo, reference_classes())
    if scenario is scenario_cbba_coalition_eligibility_visibility:
        assert ref["escort_assigns"] >= 2   # plans: string-hash dependent in the reference
    else:
        assert mine == ref


@pytest.mark.gpu
@pytest.mark.parametrize("scenario,seed,over", SCENARIOS, ids=IDS)
def test_reference_regression_on_the_cuda_facade(scenario, seed, over):
    """The same seven scenarios on the CUDA backend; summaries equal the CPU build's (same kernel sources)."""
    env, info = facade_env(seed, cuda=True, **over)
    mine = scenario(env, info, facade_classes())
    henv, hinfo = facade_env(seed, **over)
    assert mine == scenario(henv, hinfo, facade_classes())


def test_att_coalition_v2_tokens_and_one_update(hostcheck):
    """test_att_coalition_v2_tokens_and_buffer (test_escort.py:279-319): escort token shapes (48 x 22, 16 x 16, 16 x 48)
    with at least one valid edge, scores / selection masks of the same shape, one actor-critic update on a filled buffer,
    and a state_dict round trip that keeps the architecture."""
    import io

    import torch
    from multi_uav_ta_gym_env_b200.scorers import AttCoalitionNet, coalition_scores
    from multi_uav_ta_gym_env_b200.training import escort_loss

    from multi_uav_ta_gym_env_b200 import wps_config

    tok = {k: v[0] for k, v in hostcheck.make(wps_config("WPS_escort"), [1]).tokens_escort(48, 16).items()}
    assert tok["task_feats"].shape == (48, 22)
    assert tok["agent_feats"].shape == (16, 16)
    assert tok["edge_valid"].shape == (16, 48)
    assert tok["edge_valid"].sum() > 0
    torch.manual_seed(0)
    net = AttCoalitionNet(max_tasks=48, max_agents=16, d_model=64, n_layers=2)
    target = AttCoalitionNet(max_tasks=48, max_agents=16, d_model=64, n_layers=2)
    target.load_state_dict(net.state_dict())
    bt = {k: torch.as_tensor(np.asarray(v))[None] for k, v in tok.items() if k in ("task_feats", "agent_feats", "edge_valid")}
    bt.update({k: torch.as_tensor(np.asarray(tok[k]))[None].bool() for k in ("task_mask", "agent_mask")})
    scores = coalition_scores(net, bt)
    assert scores.shape == (1, 16, 48)
    B = 16
    rep = lambda x: x.expand(B, *x.shape[1:]).clone()  # noqa: E731
    batch = {"tok." + k: rep(v.float()) for k, v in bt.items()}
    batch.update({"next." + k: rep(v.float()) for k, v in bt.items()})
    batch["scores"] = rep(scores.detach())
    batch["noise"] = torch.zeros(B, 16, 48)
    batch["selected"] = rep((bt["edge_valid"] > 0).float())
    batch["reward"] = torch.full((B,), 0.05)
    batch["done"] = torch.zeros(B)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    loss = escort_loss(net, target, batch)
    assert loss is not None and torch.isfinite(loss)
    opt.zero_grad()
    loss.backward()
    opt.step()
    buf = io.BytesIO()
    torch.save({"state_dict": net.state_dict(), "max_tasks": 48, "max_agents": 16, "d_model": 64, "n_layers": 2, "version": 2}, buf)
    buf.seek(0)
    ck = torch.load(buf)
    net2 = AttCoalitionNet(max_tasks=ck["max_tasks"], max_agents=ck["max_agents"], d_model=ck["d_model"], n_layers=ck["n_layers"])
    net2.load_state_dict(ck["state_dict"])
    assert (ck["max_tasks"], ck["d_model"], ck["n_layers"]) == (48, 64, 2)
    assert torch.equal(coalition_scores(net2.eval(), bt), coalition_scores(net.eval(), bt))


def test_hungarian_rejects_a_permuted_agent_list():
    """The device builds the cost rows in agent-id order; a caller list in another order would change SciPy's tie-breaks
    (HungarianAllocator.py:128-141), so the facade class refuses it instead of answering differently."""
    env, _ = facade_env(0)
    hung = facade_classes()[0](replan_interval=1, max_coord=env.max_coord)
    live = env.get_live_agents()
    assert hung.allocate_tasks(live[::2], open_tasks(env), time_step=0, force=True)          # order-preserving subset: fine
    with pytest.raises(ValueError):
        hung.allocate_tasks(live[::-1], open_tasks(env), time_step=0, force=True)
