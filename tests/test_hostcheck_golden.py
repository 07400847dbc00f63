"""The CUDA simulation core compiled for the CPU (tests/hostcheck) against the golden fixtures:
the same source the kernel runs, checked bit-exactly where no GPU exists.  The `-m gpu` suite
repeats these checks through the C ABI on the device."""
import numpy as np
import pytest

from helpers import alloc_opts_for, golden_config, injected_scores, load_golden

STEP_CASES = ["wps_easy_local", "wps_hard_local", "wps_burst_local", "wps_commit_local", "wps_escort_coalition",
              "wps_hard_global", "wps_hard_pair", "wps_commit_pair", "wps_hard_random", "wps_escort_random",
              "wps_attn_xl_local", "wps_hard_single_task", "wps_hard_obstacles"]
# planner fixtures mutate commit_until between steps: they are replayed through the fused planner only
ALLOC_CASES = [c for c in STEP_CASES if "random" not in c and "obstacles" not in c] + ["wps_commit_urgency", "wps_escort_urgency"]


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
    for t in range(len(eps[0]["steps"])):
        if drv == "pair_injected":
            sc = np.stack([injected_scores(ep["seed"], t, 16, 32) for ep in eps])
            O.d_edge_scores = sc.ctypes.data
        env.step_alloc(O)
        for e, ep in enumerate(eps):
            st = ep["steps"][t]
            assert env.pairs_of(e) == st["pairs"], (name, ep["seed"], t)
            assert env.reward[e] == float.fromhex(st["reward"])
            assert str(env.digest(e)) == st["digest"], (name, ep["seed"], t)
    if drv not in ("pair_injected", "urgency_commit", "urgency_coalition"):
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
