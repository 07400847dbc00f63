"""The oracle (oracle/) against the golden fixtures produced by the UNMODIFIED reference
(tests/golden/gen_golden.py): per-step digests of the full canonical state, rewards, drained
events, allocator pairs and terminal metrics must be identical (bit-exact)."""
import pytest

import numpy as np

from helpers import (CBBA_DRIVERS, bundle_of, assert_tokens_equal_reference, escort_scores_from_logits, injected_commit_vectors, injected_logits, injected_scores, load_golden,
                     golden_config)
import refsnap
from oracle.hungarian import OracleHungarian, apply_assign, open_tasks
from oracle.sim import OracleEnv
from oracle import tokens as otok
from oracle import planners as oplan
from oracle.cbba import OracleCBBAReplan
from oracle.market import OraclePI

# (fixture, max episodes replayed on CPU -- keeps the CPU suite short; the GPU suite replays all)
CASES = [
    ("wps_easy_local", 3), ("wps_hard_local", 4), ("wps_burst_local", 2), ("wps_commit_local", 2),
    ("wps_escort_coalition", 2), ("wps_hard_global", 2), ("wps_hard_pair", 3), ("wps_commit_pair", 1),
    ("wps_hard_random", 3), ("wps_escort_random", 1), ("wps_attn_xl_local", 1), ("wps_hard_single_task", 2),
    ("wps_commit_urgency", 3), ("wps_escort_urgency", 2), ("wps_hard_obstacles", 2),
    ("wps_hard_urgency_pair", 3), ("wps_attn_context", 2), ("wps_commit_attcommit", 3), ("wps_escort_attescort", 3),
    ("wps_hard_pi", 4), ("wps_commit_pi", 2), ("wps_escort_pi", 2),
    ("wps_hard_cbba", 6), ("wps_commit_cbba", 3), ("wps_escort_cbba", 3),
    ("wps_hard_pi2", 6), ("wps_commit_pi2", 3), ("wps_escort_pi2", 3),
    ("wps_hard_cbba2", 4), ("wps_commit_cbba2", 2), ("wps_escort_cbba2", 2), ("wps_hard_cbba3", 2), ("wps_commit_cbba4", 1),
]


def replay(ep):
    cfg = golden_config(ep)
    o = OracleEnv(cfg).reset(ep["seed"])
    assert str(refsnap.digest(o.snapshot())) == ep["digest0"]
    drv = ep["driver"]
    interval = 12 if drv == "coalition" else (10**9 if drv in ("urgency_coalition", "att_escort_injected") else 20)
    hung = OracleHungarian(interval, o.max_coord)
    pi = OraclePI(o.max_coord, ep["seed"], 12 if drv in ("pi_coalition", "pi2_coalition") else 20)
    cbba = OracleCBBAReplan(o.max_coord, ep["seed"], 12 if drv.endswith("coalition") else 20)
    for t, st in enumerate(ep["steps"]):
        if drv in ("local_hungarian", "coalition", "global_hungarian"):
            known = None if drv == "global_hungarian" else o.visibility()
            pairs = hung.allocate(o, time_step=o.t, events=o.last_events, known=known)
            assert [list(p) for p in pairs] == st["pairs"], (ep["seed"], t)
            assert [list(a) for a in apply_assign(o, pairs)] == st["actions"], (ep["seed"], t)
        elif drv in ("local_pi", "pi_coalition", "local_pi2", "pi2_coalition"):
            pairs = pi.allocate(o, time_step=o.t, events=o.last_events, known=o.visibility(),
                                max_tasks_per_agent=2 if drv.endswith(("pi2", "pi2_coalition")) else 1)
            assert [list(p) for p in pairs] == st["pairs"], (ep["seed"], t)
            assert [list(a) for a in apply_assign(o, pairs)] == st["actions"], (ep["seed"], t)
        elif drv in CBBA_DRIVERS:
            pairs = cbba.allocate(o, time_step=o.t, events=o.last_events, known=o.visibility(),
                                  max_tasks_per_agent=bundle_of(drv))
            assert [list(p) for p in pairs] == st["pairs"], (ep["seed"], t)
            assert [list(a) for a in apply_assign(o, pairs)] == st["actions"], (ep["seed"], t)
        elif drv in ("pair_injected", "context_injected"):
            pairs = []
            if "context_tokens" in st:   # build_context_pair_tokens of the reference at this step, both variants
                assert_tokens_equal_reference(st["context_tokens"], otok.build_context_pair_tokens(o, 32, 16), None, (t, "ctx"))
                assert_tokens_equal_reference(st["context_tokens_raw"], otok.build_context_pair_tokens(o, 32, 16, raw=True),
                                              None, (t, "raw"))
            if otok.hybrid_should_replan(o, o.last_events, 15):
                sc = injected_scores(ep["seed"], o.t, 16, 32)
                pairs = otok.pair_plan(o, hung, sc)
            assert [list(p) for p in pairs] == st["pairs"], (ep["seed"], t)
        elif drv == "urgency_commit":
            pairs = []
            if otok.hybrid_should_replan(o, o.last_events, 15):
                pairs = oplan.urgency_commit_plan(o, hung)
            assert [list(p) for p in pairs] == st["pairs"], (ep["seed"], t)
        elif drv == "urgency_coalition":
            pairs = []
            if o.t == 0 or o.t % 12 == 0 or len(o.last_events) > 0:
                pairs = oplan.urgency_coalition_plan(o, hung)
            assert [list(p) for p in pairs] == st["pairs"], (ep["seed"], t)
        elif drv == "urgency_pair":
            pairs = []
            if otok.hybrid_should_replan(o, o.last_events, 15):
                pairs = oplan.urgency_pair_plan(o, hung)
            assert [list(p) for p in pairs] == st["pairs"], (ep["seed"], t)
        elif drv == "att_commit_injected":
            pairs = []
            if otok.hybrid_should_replan(o, o.last_events, 15):
                pv, cv = injected_commit_vectors(ep["seed"], o.t)
                pairs = oplan.att_commit_plan_from_scores(o, hung, pv, cv)
            assert [list(p) for p in pairs] == st["pairs"], (ep["seed"], t)
        elif drv == "att_escort_injected":
            pairs = []
            if o.t == 0 or o.t % 12 == 0 or len(o.last_events) > 0:
                tok = otok.build_escort_tokens(o, 48, 16)
                if "escort_tokens" in st:   # the reference's own build_escort_tokens output at this step
                    ref_tok = st["escort_tokens"]
                    for k in ("task_feats", "agent_feats", "edge_valid"):
                        assert np.array_equal(np.asarray(ref_tok[k], np.float32), tok[k]), (ep["seed"], t, k)
                    for k in ("task_mask", "agent_mask"):
                        assert [int(x) for x in tok[k]] == ref_tok[k], (ep["seed"], t, k)
                    assert [int(x) for x in tok["task_ids"][: len(ref_tok["task_ids"])]] == ref_tok["task_ids"]
                lg = injected_logits(ep["seed"], o.t, 16, 48)
                sc = escort_scores_from_logits(lg, tok["edge_valid"], tok["agent_mask"], tok["task_mask"])
                pairs = oplan.att_escort_plan_from_scores(o, hung, sc)
            assert [list(p) for p in pairs] == st["pairs"], (ep["seed"], t)
        r, term, trunc, ev = o.step([tuple(a) for a in st["actions"]])
        assert [list(e) for e in ev] == st["events"], (ep["seed"], t)
        assert r == float.fromhex(st["reward"]), (ep["seed"], t)
        assert (term, trunc) == (st["term"], st["trunc"])
        assert len(o.last_open) == st["n_open"]
        assert str(refsnap.digest(o.snapshot())) == st["digest"], (ep["seed"], t)
    m = o.calculate_metrics()
    for k, v in ep["metrics"].items():
        want = float.fromhex(v) if isinstance(v, str) else v
        assert m[k] == want or (m[k] != m[k] and want != want), k
    if drv in ("local_hungarian", "coalition", "global_hungarian"):
        assert hung.n_replans == ep["n_replans"]
    if drv in CBBA_DRIVERS:
        assert cbba.n_replans == ep["n_replans"]
    if drv in ("local_pi", "pi_coalition", "local_pi2", "pi2_coalition"):
        assert pi.n_replans == ep["n_replans"]


@pytest.mark.parametrize("name,n", CASES)
def test_oracle_replays_reference(name, n):
    for ep in load_golden(name)[:n]:
        replay(ep)


def test_known_answers_wps_easy_seed0():
    """BASELINE.md section 2 / SURVEY.md 8(d) config 1: WPS_easy, seed 0, Local-Hungarian interval 20."""
    ep = load_golden("wps_easy_local")[0]
    assert ep["seed"] == 0
    m = ep["metrics"]
    assert (m["n_on_time"], m["n_missed_windows"], m["n_windowed_tasks"]) == (8, 4, 15)
    assert (m["Kills"], m["Losses"], m["n_arrivals"], m["n_tasks_final"], m["n_task_switches"]) == (5, 4, 8, 27, 15)
    assert abs(float.fromhex(m["S_WPS"]) - (-24.1033)) < 1e-3
    assert abs(float.fromhex(m["total_distance"]) - 12399.0425) < 1e-3
    ep = load_golden("wps_hard_local")[0]
    m = ep["metrics"]
    assert (m["n_on_time"], m["n_missed_windows"], m["n_windowed_tasks"], m["Kills"], m["Losses"]) == (8, 14, 26, 7, 4)
    assert abs(float.fromhex(m["S_WPS"]) - (-324.0840)) < 1e-3
