"""The single-env drop-in facade on its real CUDA backend (E = 1), against the golden fixtures."""
import numpy as np
import pytest

from helpers import golden_config, load_golden
import refsnap

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,interval", [("wps_hard_local", 20), ("wps_escort_coalition", 12)])
def test_facade_episode_matches_reference_golden(name, interval):
    from multi_uav_ta_gym_env_b200.env import HungarianAllocator, MultiUAVEnv

    ep = load_golden(name)[1]
    env = MultiUAVEnv(golden_config(ep))
    obs, info = env.reset(seed=ep["seed"])
    assert [a.name for a in env.agents_obj] == ep["agent_names"]
    hung = HungarianAllocator(interval, env.max_coord)
    for t, st in enumerate(ep["steps"]):
        events = list(info.get("events") or []) if isinstance(info, dict) else []
        open_tasks = [tk for tk in env.tasks if tk.status != 2 and _residual(tk) > 0]
        res = hung.allocate_tasks(env.get_live_agents(), open_tasks, time_step=env.time_steps, events=events,
                                  agent_known_ids=env.agent_visibility_map())
        assert [[env.agent_by_name[n].id, tk.id] for n, tk in res] == st["pairs"], t
        actions = {}
        for n, tk in res:
            if env.last_tasks_info and tk in env.last_tasks_info and n not in actions:
                actions[n] = env.last_tasks_info.index(tk)
        assert [[env.agent_by_name[n].id, i] for n, i in actions.items()] == st["actions"]
        obs, rew, term, trunc, info = env.step(actions)
        assert next(iter(rew.values())) == float.fromhex(st["reward"])
        assert str(refsnap.digest(env._snap)) == st["digest"], t
        assert len(obs[env.agents_obj[0].name]["tasks_info"]) >= env.max_tasks
    for k, v in ep["metrics"].items():
        want = float.fromhex(v) if isinstance(v, str) else v
        got = info["metrics"][k]
        assert got == want or (got != got and want != want), k
    assert hung.n_replans == ep["n_replans"]


def _residual(t):
    if t.kind == "Escort" or float(t.required_agents or 0) > 0:
        return max(float(t.required_agents or 1) - len(t.allocationDetails), 0.0)
    return max(float(t.currentReqs[t.typeIdx] - t.allocatedReqs[t.typeIdx]), 0.0)


def test_core_sim_shim():
    from multi_uav_ta_gym_env_b200 import core_sim, wps_config
    from oracle.sim import OracleEnv

    obst = [[300.0, 300.0, 50.0], [700.0, 200.0, 80.0]]
    o = OracleEnv(wps_config("WPS_hard"))
    o.obstacles = [tuple(r) for r in obst]
    for pos, mv in (([341.0, 333.0], [1.0, 0.0]), ([650.0, 260.0], [-0.6, 0.8]), ([10.0, 10.0], [0.0, 1.0])):
        got = core_sim.SimCore.avoid_obstacles(pos, obst, mv)
        want = o._avoid(pos[0], pos[1], mv[0], mv[1])
        assert np.allclose(got, want, rtol=1e-9, atol=1e-12)
    assert core_sim.SimCore.avoid_obstacles([5.0, 5.0], [], [1.0, 0.0]) == [0.0, 0.0]


@pytest.mark.parametrize("scenario,seed", [("WPS_commit", 2), ("WPS_escort", 3)])
def test_replay_document_on_the_cuda_backend(scenario, seed):
    """replay.record_replay (the reference web UI's replay file) from the facade on its CUDA backend:
    same document as on the CPU build of the kernel core (tests/golden/replay_digests.json, made by
    tests/golden/gen_replay_digests.py; the format itself is pinned against the reference generator in
    tests/test_dropin_facade.py)."""
    import json
    import os

    import gen_replay_digests as G

    with open(os.path.join(os.path.dirname(G.__file__), "replay_digests.json")) as f:
        want = json.load(f)[f"{scenario}:{seed}"]
    got = G.replay_digest(scenario, seed)
    assert got == want
