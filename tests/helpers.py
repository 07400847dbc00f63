"""Shared test helpers: golden-fixture loading and the CPU build of the kernel core."""
from __future__ import annotations

import ctypes as C
import gzip
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, GOLDEN):
    if p not in sys.path:
        sys.path.insert(0, p)

import refsnap  # noqa: E402
from multi_uav_ta_gym_env_b200 import _lib, config, reset, state  # noqa: E402


def load_golden(name):
    with gzip.open(os.path.join(GOLDEN, name + ".json.gz"), "rt") as f:
        return json.load(f)["episodes"]


def golden_config(ep):
    return config.wps_config(ep["case"], **ep.get("overrides", {}))


def injected_scores(seed, t, n_rows, n_cols):
    """Same generator as tests/golden/gen_golden.py (kept here so the GPU box needs no reference)."""
    i = np.arange(n_rows, dtype=np.uint64)[:, None]
    j = np.arange(n_cols, dtype=np.uint64)[None, :]
    x = (np.uint64(seed) * np.uint64(1000003) + np.uint64(t) * np.uint64(7919)
         + i * np.uint64(104729) + j * np.uint64(1299709) + np.uint64(12345))
    x = (x * np.uint64(2654435761)) % np.uint64(2001)
    return ((x.astype(np.float64) - 1000.0) / 1000.0 * 0.35).astype(np.float32)


def injected_commit_vectors(seed, t):
    """Same generator as tests/golden/gen_golden.py: (pri_vec [32], com_vec [16]) float32 in [0, 1]."""
    m = injected_scores(seed, t, 2, 32)
    v = ((m.astype(np.float64) / 0.35 + 1.0) * 0.5).astype(np.float32)
    return v[0], v[1, :16]


def injected_logits(seed, t, n_rows, n_cols):
    return (injected_scores(seed, t, n_rows, n_cols) * np.float32(10.0)).astype(np.float32)


def escort_scores_from_logits(logits, edge_valid, agent_mask, task_mask):
    """AttentionEscort.act with explore=False (AttentionEscort.py:449-466): float32 sigmoid of the clipped logits,
    zeroed on invalid edges and padded rows / columns."""
    scores = 1.0 / (1.0 + np.exp(-np.clip(logits, -20, 20)))
    scores = scores * edge_valid
    scores = scores * (~np.asarray(agent_mask, bool))[:, None] * (~np.asarray(task_mask, bool))[None, :]
    return scores.astype(np.float32)


def assert_tokens_equal_reference(ref_tok, tok, e=None, what=""):
    """ref_tok: token dump of the reference stored in a fixture; tok: dict of arrays ([E, ...] when e is given)."""
    pick = (lambda v: v[e]) if e is not None else (lambda v: v)
    for k in ("task_feats", "agent_feats", "edge_valid", "context"):
        if k in ref_tok:
            assert np.array_equal(np.asarray(ref_tok[k], np.float32), np.asarray(pick(tok[k]), np.float32)), (what, k)
    for k in ("task_mask", "agent_mask"):
        assert [int(x) for x in pick(tok[k])] == ref_tok[k], (what, k)
    nk = len(ref_tok["task_ids"])
    ids = pick(tok["task_ids"])
    assert [int(x) for x in ids[:nk]] == ref_tok["task_ids"] and not np.asarray(ids[nk:]).any(), (what, "task_ids")


CBBA_DRIVERS = ("cbba_replan", "cbba_coalition", "cbba2_replan", "cbba2_coalition", "cbba3_replan", "cbba4_replan")
BUNDLE_DRIVERS = ("local_pi2", "pi2_coalition", "cbba2_replan", "cbba2_coalition", "cbba3_replan", "cbba4_replan")


def bundle_of(driver):
    """max_tasks_per_agent of a golden driver name: local_pi2 / pi2_coalition / cbba<N>_* build bundles of N tasks."""
    for c in driver:
        if c.isdigit():
            return int(c)
    return 1


def alloc_opts_for(driver):
    if driver == "context_injected":
        driver = "pair_injected"   # ContextPairHybrid plans exactly like PairCostHybrid (it only adds the context vector)
    O = _lib.MuavAllocOpts()
    O.max_coord = 1200.0
    if driver in ("local_hungarian", "coalition", "global_hungarian"):
        O.mode = 1
        O.replan_interval = 12 if driver == "coalition" else 20
        O.event_mask = 0x1F
        O.use_visibility = 0 if driver == "global_hungarian" else 1
    elif driver in ("local_pi", "pi_coalition", "local_pi2", "pi2_coalition"):
        O.mode, O.replan_interval, O.event_mask, O.use_visibility = 1, (12 if "coalition" in driver else 20), 0x1F, 1
        O.planner = 6
        O.max_tasks_per_agent = 2 if "pi2" in driver else 1   # the caller points d_bundle_pairs / d_n_bundle_pairs somewhere
    elif driver in CBBA_DRIVERS:
        # CBBAReplan under the drivers of wps_eval.py:134-146 / escort_eval.py:149-161; the caller sets O.d_cbba_seed
        O.mode, O.replan_interval, O.event_mask, O.use_visibility = 1, (12 if "coalition" in driver else 20), 0x1F, 1
        O.planner = 7
        O.max_tasks_per_agent = bundle_of(driver)
    elif driver == "pair_injected":
        O.mode = 2
        O.replan_interval = 15
        O.event_mask = 0b111
        O.use_visibility = 1
        O.pair_tokens = 1
        O.score_rows = 16
        O.score_cols = 32
    elif driver == "urgency_commit":
        O.mode, O.replan_interval, O.event_mask, O.use_visibility = 2, 15, 0b111, 1
        O.planner, O.commit_fraction = 1, 0.35
    elif driver == "urgency_coalition":
        O.mode, O.replan_interval, O.event_mask, O.use_visibility = 2, 12, 0x1F, 1
        O.planner = 2
    elif driver == "urgency_pair":
        O.mode, O.replan_interval, O.event_mask, O.use_visibility = 2, 15, 0b111, 1
        O.planner, O.score_rows, O.score_cols = 5, 16, 32
    elif driver == "att_commit_injected":
        O.mode, O.replan_interval, O.event_mask, O.use_visibility = 2, 15, 0b111, 1
        O.planner, O.commit_threshold, O.score_rows, O.score_cols = 3, 0.5, 16, 32
    elif driver == "att_escort_injected":
        O.mode, O.replan_interval, O.event_mask, O.use_visibility = 2, 12, 0x1F, 1
        O.planner, O.score_rows, O.score_cols = 4, 16, 48
    else:
        raise ValueError(driver)
    return O


class HostCheck:
    """tests/hostcheck: the CUDA simulation core compiled for the CPU (never used by the product)."""

    def __init__(self):
        import __graft_entry__ as ge

        path = ge.build_hostcheck()
        self.lib = _lib.Lib(path)
        P = C.c_void_p
        d = self.lib.dll
        d.hostcheck_step.restype = C.c_int
        d.hostcheck_step.argtypes = [C.POINTER(_lib.MuavConfig), P, P, P, C.POINTER(_lib.MuavAllocOpts),
                                     C.POINTER(_lib.MuavStepOut), C.c_int, C.c_int]
        d.hostcheck_tokens_context.restype = C.c_int
        d.hostcheck_tokens_context.argtypes = [C.POINTER(_lib.MuavConfig), P, C.c_int, C.c_int, C.c_int, P, P, P, P, P, P, P, C.c_int]
        d.hostcheck_tokens_escort.restype = C.c_int
        d.hostcheck_tokens_escort.argtypes = [C.POINTER(_lib.MuavConfig), P, C.c_int, C.c_int, P, P, P, P, P, P, P, C.c_int]
        d.hostcheck_lsap.restype = C.c_int
        d.hostcheck_lsap.argtypes = [P, P, P, C.c_int, C.c_int, P, C.c_int]

    def make(self, opts, seeds, queue_cap=8):
        return HostEnv(self, opts, seeds, queue_cap)

    def lsap(self, cost, nr, nc):
        cost = np.ascontiguousarray(cost, dtype=np.float64)
        B, nr_max, nc_max = cost.shape
        nr = np.ascontiguousarray(nr, dtype=np.int32)
        nc = np.ascontiguousarray(nc, dtype=np.int32)
        out = np.full((B, nr_max), -7, np.int32)
        rc = self.lib.dll.hostcheck_lsap(cost.ctypes.data, nr.ctypes.data, nc.ctypes.data, nr_max, nc_max,
                                         out.ctypes.data, B)
        assert rc == 0
        return out


class HostEnv:
    def __init__(self, hc, opts, seeds, queue_cap=8):
        self.hc = hc
        self.lib = hc.lib
        self.opts = opts
        self.cfg = _lib.build_config(opts, queue_cap=queue_cap)
        self.E = len(seeds)
        sc = [reset.generate_scenario(opts, int(s), list(self.cfg.tape_words)) for s in seeds]
        self.rec, self.tapes = reset.pack_records(self.lib, self.cfg, sc)
        self.codec = state.RecordCodec(self.lib, self.cfg)
        E, A = self.E, self.cfg.n_agents
        self.reward = np.zeros(E)
        self.term = np.zeros(E, np.uint8)
        self.trunc = np.zeros(E, np.uint8)
        self.n_events = np.zeros(E, np.int32)
        self.events = np.zeros((E, self.cfg.event_cap), np.int32)
        self.n_pairs = np.zeros(E, np.int32)
        self.pairs = np.zeros((E, A), np.int32)
        self.n_open = np.zeros(E, np.int32)
        out = _lib.MuavStepOut()
        out.d_reward = self.reward.ctypes.data
        out.d_terminated = self.term.ctypes.data
        out.d_truncated = self.trunc.ctypes.data
        out.d_n_events = self.n_events.ctypes.data
        out.d_events = self.events.ctypes.data
        out.d_n_pairs = self.n_pairs.ctypes.data
        out.d_pairs = self.pairs.ctypes.data
        out.d_n_open = self.n_open.ctypes.data
        self.out = out

    def step_actions(self, actions, n_steps=1):
        """actions: per env ordered list of (agent, idx)."""
        A = self.cfg.n_agents
        act = np.full((self.E, A, 2), -1, np.int32)
        for e, lst in enumerate(actions):
            for i, (a, idx) in enumerate(lst):
                act[e, i] = (a, idx)
        rc = self.lib.dll.hostcheck_step(C.byref(self.cfg), self.rec.ctypes.data, self.tapes.ctypes.data,
                                         act.ctypes.data, None, C.byref(self.out), self.E, n_steps)
        assert rc == 0

    def step_alloc(self, O, n_steps=1):
        self.n_pairs[:] = 0
        rc = self.lib.dll.hostcheck_step(C.byref(self.cfg), self.rec.ctypes.data, self.tapes.ctypes.data, None,
                                         C.byref(O), C.byref(self.out), self.E, n_steps)
        assert rc == 0

    def tokens_context(self, max_tasks=32, max_agents=16, raw=False):
        E = self.E
        tok = {"task_feats": np.zeros((E, max_tasks, 9 if raw else 13), np.float32), "task_mask": np.zeros((E, max_tasks), np.uint8),
               "agent_feats": np.zeros((E, max_agents, 11 if raw else 12), np.float32),
               "agent_mask": np.zeros((E, max_agents), np.uint8), "edge_valid": np.zeros((E, max_agents, max_tasks), np.float32),
               "task_ids": np.zeros((E, max_tasks), np.int32), "context": np.zeros((E, 1 if raw else 8), np.float32)}
        rc = self.lib.dll.hostcheck_tokens_context(
            C.byref(self.cfg), self.rec.ctypes.data, max_tasks, max_agents, int(raw), tok["task_feats"].ctypes.data,
            tok["task_mask"].ctypes.data, tok["agent_feats"].ctypes.data, tok["agent_mask"].ctypes.data,
            tok["edge_valid"].ctypes.data, tok["task_ids"].ctypes.data, tok["context"].ctypes.data, E)
        assert rc == 0
        return tok

    def tokens_escort(self, max_tasks=48, max_agents=16):
        E = self.E
        IC = max(self.cfg.id_cap, self.cfg.task_cap)
        tok = {"task_feats": np.zeros((E, max_tasks, 22), np.float32), "task_mask": np.zeros((E, max_tasks), np.uint8),
               "agent_feats": np.zeros((E, max_agents, 16), np.float32), "agent_mask": np.zeros((E, max_agents), np.uint8),
               "edge_valid": np.zeros((E, max_agents, max_tasks), np.float32), "task_ids": np.zeros((E, max_tasks), np.int32),
               "task_order": np.zeros((E, IC), np.int32)}
        rc = self.lib.dll.hostcheck_tokens_escort(
            C.byref(self.cfg), self.rec.ctypes.data, max_tasks, max_agents, tok["task_feats"].ctypes.data,
            tok["task_mask"].ctypes.data, tok["agent_feats"].ctypes.data, tok["agent_mask"].ctypes.data,
            tok["edge_valid"].ctypes.data, tok["task_ids"].ctypes.data, tok["task_order"].ctypes.data, E)
        assert rc == 0
        return tok

    def snapshot(self, e):
        return self.codec.snapshot(self.rec[e])

    def digest(self, e):
        return refsnap.digest(self.snapshot(e))

    def pairs_of(self, e):
        return [[int(p) >> 16, int(p) & 0xFFFF] for p in self.pairs[e, : self.n_pairs[e]]]

    def events_of(self, e):
        return [[int(v) & 0xFF, (int(v) >> 8) - 1] for v in self.events[e, : self.n_events[e]]]

    def err(self, e):
        return self.codec.header(self.rec[e], "ERRFLAGS")


class HostBackend:
    """Test-only backend of multi_uav_ta_gym_env_b200.env.MultiUAVEnv over the CPU build of the kernel core,
    so the single-env facade (proxies, observation dicts, allocator class) can be exercised -- and driven by
    the UNMODIFIED reference hybrids -- in the GPU-less authoring container.  Never used by the product."""

    _hc = None

    def __init__(self, opts, queue_cap=16, task_cap="all"):
        if HostBackend._hc is None:
            HostBackend._hc = HostCheck()
        hc = HostBackend._hc
        self.lib = hc.lib
        self.opts = opts
        self.cfg = _lib.build_config(opts, queue_cap=queue_cap, task_cap=task_cap)
        self.codec = state.RecordCodec(self.lib, self.cfg)
        d = self.lib.dll
        P = C.c_void_p
        d.hostcheck_allocate.restype = C.c_int
        d.hostcheck_allocate.argtypes = [C.POINTER(_lib.MuavConfig), P, C.POINTER(_lib.MuavAllocOpts),
                                         C.POINTER(_lib.MuavStepOut), P, C.c_int]
        d.hostcheck_observe.restype = C.c_int
        d.hostcheck_observe.argtypes = [C.POINTER(_lib.MuavConfig), P, C.c_int, P, P, P, P, P, P, C.c_int]
        d.hostcheck_metrics.restype = C.c_int
        d.hostcheck_metrics.argtypes = [C.POINTER(_lib.MuavConfig), P, P, C.c_int]
        d.hostcheck_tokens_pair.restype = C.c_int
        d.hostcheck_tokens_pair.argtypes = [C.POINTER(_lib.MuavConfig), P, C.c_int, C.c_int, P, P, P, P, P, P, C.c_int]

    def reset(self, seed):
        sc = reset.generate_scenario(self.opts, int(seed), list(self.cfg.tape_words))
        self.rec, self.tapes = reset.pack_records(self.lib, self.cfg, [sc])
        return sc

    def record(self):
        return self.rec[0]

    def step(self, actions):
        A = self.cfg.n_agents
        rew = np.zeros(1)
        term = np.zeros(1, np.uint8)
        trunc = np.zeros(1, np.uint8)
        nev = np.zeros(1, np.int32)
        evs = np.zeros((1, self.cfg.event_cap), np.int32)
        out = _lib.MuavStepOut()
        out.d_reward, out.d_terminated, out.d_truncated = rew.ctypes.data, term.ctypes.data, trunc.ctypes.data
        out.d_n_events, out.d_events = nev.ctypes.data, evs.ctypes.data
        act = np.ascontiguousarray(actions.reshape(1, A, 2), dtype=np.int32)
        rc = self.lib.dll.hostcheck_step(C.byref(self.cfg), self.rec.ctypes.data, self.tapes.ctypes.data, act.ctypes.data,
                                         None, C.byref(out), 1, 1)
        assert rc == 0
        return float(rew[0]), bool(term[0]), bool(trunc[0]), state.decode_events(nev[0], evs[0])

    def allocate(self, spec, scores, priorities, reserved, order, cbba_seed=None):
        A = self.cfg.n_agents
        O = _lib.MuavAllocOpts()
        if cbba_seed is not None:
            self._cbba_seed = np.array([int(cbba_seed)], np.int32)
            O.d_cbba_seed = self._cbba_seed.ctypes.data
        O.mode, O.replan_interval, O.event_mask = spec.mode, spec.replan_interval, spec.event_mask
        O.use_visibility, O.pair_tokens, O.max_coord = int(spec.use_visibility), int(spec.pair_tokens), spec.max_coord
        O.planner, O.commit_fraction = spec.planner, spec.commit_fraction
        keep = []
        if scores is not None:
            sc = np.ascontiguousarray(scores, dtype=np.float64)
            keep.append(sc)
            O.d_edge_scores, O.score_rows, O.score_cols, O.score_f64 = sc.ctypes.data, sc.shape[0], sc.shape[1], 1
        if priorities is not None:
            pr = np.ascontiguousarray(priorities, dtype=np.float64)
            keep.append(pr)
            O.d_priorities = pr.ctypes.data
        if reserved is not None:
            rs = np.ascontiguousarray(reserved, dtype=np.uint8)
            keep.append(rs)
            O.d_reserved = rs.ctypes.data
        if order is not None:
            od = np.ascontiguousarray(order, dtype=np.int32)
            keep.append(od)
            O.d_task_order = od.ctypes.data
        mb = int(getattr(spec, "max_tasks_per_agent", 1))
        self._bundle = np.zeros((1, A * max(mb, 1)), np.int32)
        self._n_bundle = np.zeros(1, np.int32)
        if spec.planner in (6, 7) and mb > 1:
            O.max_tasks_per_agent = mb
            O.d_bundle_pairs, O.d_n_bundle_pairs = self._bundle.ctypes.data, self._n_bundle.ctypes.data
        npairs = np.zeros(1, np.int32)
        pairs = np.zeros((1, A), np.int32)
        out = _lib.MuavStepOut()
        out.d_n_pairs, out.d_pairs = npairs.ctypes.data, pairs.ctypes.data
        rc = self.lib.dll.hostcheck_allocate(C.byref(self.cfg), self.rec.ctypes.data, C.byref(O), C.byref(out), None, 1)
        assert rc == 0
        return [[int(p) >> 16, int(p) & 0xFFFF] for p in pairs[0, : npairs[0]]]

    def bundle_pairs(self):
        return [[int(p) >> 16, int(p) & 0xFFFF] for p in self._bundle[0, : self._n_bundle[0]]]

    def observe(self, max_rows):
        A = self.cfg.n_agents
        ti = np.zeros((max_rows, 21))
        pm = np.zeros(max_rows, np.uint8)
        lm = np.zeros((A, max_rows), np.uint8)
        ao = np.zeros((A, 9))
        ef = np.zeros(5, np.float32)
        nr = np.zeros(1, np.int32)
        rc = self.lib.dll.hostcheck_observe(C.byref(self.cfg), self.rec.ctypes.data, max_rows, ti.ctypes.data, pm.ctypes.data,
                                            lm.ctypes.data, ao.ctypes.data, ef.ctypes.data, nr.ctypes.data, 1)
        assert rc == 0
        return {"tasks_info": ti, "mask": pm.astype(bool), "legal_mask": lm.astype(bool), "agent_obs": ao,
                "event_flags": ef, "n_rows": nr[0]}

    def tokens_pair(self, max_tasks=32, max_agents=16):
        tf = np.zeros((max_tasks, 13), np.float32)
        tm = np.zeros(max_tasks, np.uint8)
        af = np.zeros((max_agents, 12), np.float32)
        am = np.zeros(max_agents, np.uint8)
        ev = np.zeros((max_agents, max_tasks), np.float32)
        ids = np.zeros(max_tasks, np.int32)
        rc = self.lib.dll.hostcheck_tokens_pair(C.byref(self.cfg), self.rec.ctypes.data, max_tasks, max_agents, tf.ctypes.data,
                                                tm.ctypes.data, af.ctypes.data, am.ctypes.data, ev.ctypes.data,
                                                ids.ctypes.data, 1)
        assert rc == 0
        return {"task_feats": tf, "task_mask": tm.astype(bool), "agent_feats": af, "agent_mask": am.astype(bool),
                "edge_valid": ev, "task_ids": ids}

    def metrics(self):
        m = np.zeros(30)
        rc = self.lib.dll.hostcheck_metrics(C.byref(self.cfg), self.rec.ctypes.data, m.ctypes.data, 1)
        assert rc == 0
        return m

    def patch_field(self, name, index, value):
        self.codec.field(self.rec[0], name)[index] = value


def host_facade(cfg, **kw):
    """multi_uav_ta_gym_env_b200.env.MultiUAVEnv over the CPU build of the kernel sources (HostBackend): a test-side
    subclass overriding the facade's backend hook; the product class itself only ever builds the CUDA backend."""
    from multi_uav_ta_gym_env_b200.env import MultiUAVEnv

    class HostMultiUAVEnv(MultiUAVEnv):
        def _make_backend(self, c, device):
            return HostBackend(c, **kw)

    return HostMultiUAVEnv(cfg)
