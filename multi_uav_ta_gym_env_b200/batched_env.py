"""Batched WPS environment: E independent MultiUAVEnv instances resident in HBM, stepped by the
CUDA kernels behind the C ABI (include/muav.h).  This is the "batched-tensor step added
alongside" the PettingZoo surface (multi_uav_ta_gym_env_b200/env.py wraps it with E = 1).

Reference semantics: MultiUAVEnv.reset/step (mUAV_TA/DroneEnv.py:522-762, 774-1206) and, when an
`AllocSpec` is given, HungarianAllocator.allocate_tasks + the drivers' glue
(HungarianAllocator.py:72-208, experiments/wps_eval.py:55-73,123-133, escort_eval.py:137-148).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib, reset as _reset, state as _state
from .config import EVENT_TAGS

ALL_EVENTS = 0x1F
HYBRID_EVENTS = 0b00111  # Reset_Allocation, Agent_Fail, New_Threat (wps_eval.py:64-73)


@dataclass
class AllocSpec:
    """How the fused allocator runs in front of each step (see muav_alloc_opts in include/muav.h)."""

    mode: int = 1                 # 1: HungarianAllocator.should_replan rule; 2: hybrid cadence + force
    replan_interval: int = 20
    event_mask: int = ALL_EVENTS
    use_visibility: bool = True
    pair_tokens: bool = False
    max_coord: float = 1200.0
    planner: int = 0              # 1: UrgencyCommit.plan, 2: UrgencyCoalition.plan (device-side planners),
                                  # 3: AttentionCommit._plan_from_scores, 4: AttentionEscort._plan_from_scores,
                                  # 5: UrgencyPair.plan, 6: PerformanceImpact.allocate_tasks (max_tasks_per_agent below),
                                  # 7: CBBAReplan.allocate_tasks (max_tasks_per_agent below)
    commit_fraction: float = 0.35
    commit_threshold: float = 0.5
    max_tasks_per_agent: int = 1  # planners 6 / 7: bundles of up to 4 tasks per agent (the step takes the first of each)

    @staticmethod
    def local_hungarian(interval=20):
        return AllocSpec(1, interval, ALL_EVENTS, True, False)

    @staticmethod
    def global_hungarian(interval=20):
        return AllocSpec(1, interval, ALL_EVENTS, False, False)

    @staticmethod
    def coalition_hungarian(interval=12):
        return AllocSpec(1, interval, ALL_EVENTS, True, False)

    @staticmethod
    def performance_impact(interval=20, max_tasks_per_agent=1):
        """Local-PI / Local-PI-Coalition (MarketBased/PerformanceImpact.py:59-224) under its own should_replan rule, as
        experiments/wps_eval.py:147-159 (interval 20) and escort_eval.py:162-174 (12) run it (max_tasks_per_agent=1 there).
        With max_tasks_per_agent 2..4 the allocator builds bundles; the step takes the first task of every path and the
        whole plan is in env.bundle_pairs_of(e)."""
        return AllocSpec(1, interval, ALL_EVENTS, True, False, planner=6, max_tasks_per_agent=int(max_tasks_per_agent))

    @staticmethod
    def cbba_replan(interval=20, max_tasks_per_agent=1):
        """Local-CBBA-Replan / Local-CBBA-Coalition (MarketBased/CBBA_Replan.py:15-69 around CBBA.py:68-324 with
        max_tasks_per_agent=1, a fresh CBBA(seed + n_replans) per replan; seed = the environment's reset seed) as
        experiments/wps_eval.py:134-146 (interval 20) and escort_eval.py:149-161 (12) run it.  Reproduces the reference
        run under PYTHONHASHSEED=0 (csrc/muav_cbba.cuh).  max_tasks_per_agent 2..4: bundles, as for performance_impact()."""
        return AllocSpec(1, interval, ALL_EVENTS, True, False, planner=7, max_tasks_per_agent=int(max_tasks_per_agent))

    @staticmethod
    def pair_hybrid(interval=15):
        return AllocSpec(2, interval, HYBRID_EVENTS, True, True)

    @staticmethod
    def urgency_commit(interval=15, commit_fraction=0.35):
        """UrgencyCommit under the hybrid cadence of wps_eval.py:64-73,207-213."""
        return AllocSpec(2, interval, HYBRID_EVENTS, True, False, planner=1, commit_fraction=commit_fraction)

    @staticmethod
    def urgency_coalition(interval=12):
        """UrgencyCoalition under escort_eval.py:52-58,175-179 (every event tag triggers)."""
        return AllocSpec(2, interval, ALL_EVENTS, True, False, planner=2)

    @staticmethod
    def urgency_pair(interval=15):
        """UrgencyPair.plan (PairCostHybrid.py:520-550) under the hybrid cadence; scores are computed on the device."""
        return AllocSpec(2, interval, HYBRID_EVENTS, True, True, planner=5)

    @staticmethod
    def att_commit(interval=15, commit_threshold=0.5):
        """AttentionCommit.plan under the hybrid cadence (wps_eval.py:64-73): step_allocated(..., plan_pri=, plan_commit=)
        take the network's priority / commit vectors in muav_tokens_commit layout."""
        return AllocSpec(2, interval, HYBRID_EVENTS, True, False, planner=3, commit_threshold=commit_threshold)

    @staticmethod
    def att_escort(interval=12):
        """AttentionEscort.plan under escort_eval.py:52-58: step_allocated(..., edge_scores=, task_order=) take the
        scores and the task order of tokens_escort()."""
        return AllocSpec(2, interval, ALL_EVENTS, True, False, planner=4)


class _NoCtx:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_CTX = _NoCtx()


class BatchedMultiUAVEnv:
    def __init__(self, config, n_envs: int, device="cuda:0", task_cap=None, queue_cap=16, id_cap=None):
        self.lib = _lib.cuda_lib()  # raises if the CUDA library is not built: no CPU fallback
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedMultiUAVEnv needs a CUDA device (B200); there is no CPU path")
        self.config = config
        self.n_envs = int(n_envs)
        self.device = torch.device(device)
        self.cfg = _lib.build_config(config, task_cap=task_cap, queue_cap=queue_cap, id_cap=id_cap)
        self.codec = _state.RecordCodec(self.lib, self.cfg)
        self.record_bytes = self.codec.record_bytes
        self.n_agents = self.cfg.n_agents
        self.task_cap = self.cfg.task_cap
        self.max_coord = 1200.0
        self.possible_agents = [f"{t[0:2]}_agent{i}" for t, n in config.agents.items() for i in range(n)]
        E, A = self.n_envs, self.n_agents
        dev = self.device
        self.records = None
        self.tapes = None
        self.reward = torch.zeros(E, dtype=torch.float64, device=dev)
        self.terminated = torch.zeros(E, dtype=torch.uint8, device=dev)
        self.truncated = torch.zeros(E, dtype=torch.uint8, device=dev)
        self.n_events = torch.zeros(E, dtype=torch.int32, device=dev)
        self.events = torch.zeros(E, self.cfg.event_cap, dtype=torch.int32, device=dev)
        self.n_pairs = torch.zeros(E, dtype=torch.int32, device=dev)
        self.pairs = torch.zeros(E, A, dtype=torch.int32, device=dev)
        self.n_open = torch.zeros(E, dtype=torch.int32, device=dev)
        self._out = _lib.MuavStepOut()
        self._out.d_reward = self.reward.data_ptr()
        self._out.d_terminated = self.terminated.data_ptr()
        self._out.d_truncated = self.truncated.data_ptr()
        self._out.d_n_events = self.n_events.data_ptr()
        self._out.d_events = self.events.data_ptr()
        self._out.d_n_pairs = self.n_pairs.data_ptr()
        self._out.d_pairs = self.pairs.data_ptr()
        self._out.d_n_open = self.n_open.data_ptr()
        # launch-slot order (scheduling hint of muav_step_out): two buffers used alternately, see include/muav.h
        self.group_replanners = os.environ.get("MUAV_ENV_ORDER", "1") != "0"
        self._order = torch.zeros(2, E + 2, dtype=torch.int32, device=dev)
        self._order_cur = -1  # index of the buffer holding the order for the next launch (-1: identity)
        # workspace that lets a single fused step run as allocator kernel + step kernel (muav_step_out.d_actions_ws): the
        # library takes that form when the step-only launch keeps >= 1.5 x as many environments resident per SM as the
        # fused one (big shapes: WPS_escort, burst x4 / x8), see launch rule in csrc/muav_kernels.cu
        self._actions_ws = torch.empty(E, A, 2, dtype=torch.int32, device=dev)
        self._out.d_actions_ws = self._actions_ws.data_ptr()
        self.scenarios = None
        self.agent_names = None
        self.launches = 0
        self._ctx = None  # muav_ctx handle of the host-buffer entry points (created on first use, owned by this object)

    def __del__(self):
        ctx, self._ctx = getattr(self, "_ctx", None), None
        if ctx:
            try:
                self.lib.dll.muav_ctx_destroy(ctx)
            except Exception:
                pass

    def _host_ctx(self):
        """The caller-owned handle behind step_host / allocate_host (include/muav.h: muav_ctx): staging buffers for this
        environment batch on this device."""
        if self._ctx is None:
            h = C.c_void_p()
            idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
            _lib.check(self.lib.dll.muav_ctx_create(C.byref(self.cfg), self.n_envs, int(idx), C.byref(h)), "muav_ctx_create")
            self._ctx = h
        return self._ctx

    # ------------------------------------------------------------------ reset
    def reset(self, seeds: Optional[Sequence[int]] = None):
        """Host scenario generation per seed (DroneEnv.py:522-762) + one upload."""
        if seeds is None:
            seeds = range(self.n_envs)
        seeds = list(seeds)
        if len(seeds) != self.n_envs:
            raise ValueError("need one seed per environment")
        tw = list(self.cfg.tape_words)
        # scenario generation is host work (about a millisecond per 8-agent environment): repeated seeds are generated
        # and packed once, then their rows are replicated
        uniq = sorted(set(int(s) for s in seeds))
        made = {s: _reset.generate_scenario(self.config, s, tw) for s in uniq}
        self.scenarios = [made[int(s)] for s in seeds]
        self.agent_names = [sc.agent_names for sc in self.scenarios]
        rec, tapes = _reset.pack_records(self.lib, self.cfg, [made[s] for s in uniq])
        if len(uniq) != len(seeds):
            row = {s: i for i, s in enumerate(uniq)}
            take = np.fromiter((row[int(s)] for s in seeds), dtype=np.int64, count=len(seeds))
            rec, tapes = np.ascontiguousarray(rec[take]), np.ascontiguousarray(tapes[take])
        E = self.n_envs
        dll = self.lib.dll
        assert dll.muav_state_bytes(C.byref(self.cfg), E) == rec.nbytes and dll.muav_tape_bytes(C.byref(self.cfg), E) == tapes.nbytes
        self.records = torch.empty(E, rec.shape[1], dtype=torch.uint8, device=self.device)
        self.tapes = torch.empty(E, tapes.shape[1], dtype=torch.int32, device=self.device)
        rc = dll.muav_reset_upload(C.byref(self.cfg), self.records.data_ptr(), self.tapes.data_ptr(), rec.ctypes.data,
                                   tapes.ctypes.data, E, self._stream())
        _lib.check(rc, "muav_reset_upload")
        torch.cuda.current_stream(self.device).synchronize()
        self._records0 = self.records.clone()
        self.seeds = torch.tensor([int(s) for s in seeds], dtype=torch.int32, device=self.device)   # CBBAReplan(seed=...)
        self.n_open.copy_(self.header_int("N_OPEN"))
        self._order_cur = -1   # identity order for the first launch; both buffers start with zeroed fill counters
        self._order.zero_()
        return self

    def restore(self):
        """Rewind every environment to its reset state (device-to-device copy, no host work)."""
        self.records.copy_(self._records0)

    # ------------------------------------------------------------------ fused token emission
    def enable_fused_tokens(self, max_tasks=32, max_agents=16, interval=15, event_mask=HYBRID_EVENTS, context=False,
                            commit=False, escort=False):
        """Ask the step kernel to emit pair tokens for the environments that will replan before the next
        step (muav_token_out).  Returns the token dict (tensors are updated in place by every step) with
        `need` u8[E].  commit=True: commit tokens instead (agent features [., 13], enrich_commit_tokens).
        escort=True: escort tokens (build_escort_tokens: task features [., 22], agent features [., 16]) plus
        `task_order` [E, id_cap] for AllocSpec.att_escort(); pass the escort replan rule (interval 12, every event)."""
        E, dev = self.n_envs, self.device
        tfd = 22 if escort else 13
        afd = 16 if escort else (13 if commit else 12)
        tok = {
            "task_feats": torch.zeros(E, max_tasks, tfd, dtype=torch.float32, device=dev),
            "task_mask_u8": torch.ones(E, max_tasks, dtype=torch.uint8, device=dev),
            "agent_feats": torch.zeros(E, max_agents, afd, dtype=torch.float32, device=dev),
            "agent_mask_u8": torch.ones(E, max_agents, dtype=torch.uint8, device=dev),
            "edge_valid": torch.zeros(E, max_agents, max_tasks, dtype=torch.float32, device=dev),
            "task_ids": torch.zeros(E, max_tasks, dtype=torch.int32, device=dev),
            "need": torch.zeros(E, dtype=torch.uint8, device=dev),
        }
        if context:   # build_context_summary of the same tokens (ContextPairHybrid.py:33-70)
            tok["context"] = torch.zeros(E, 8, dtype=torch.float32, device=dev)
        if escort:
            tok["task_order"] = torch.zeros(E, max(self.cfg.id_cap, self.cfg.task_cap), dtype=torch.int32, device=dev)
        T = _lib.MuavTokenOut()
        T.d_task_feats = tok["task_feats"].data_ptr()
        T.d_task_mask = tok["task_mask_u8"].data_ptr()
        T.d_agent_feats = tok["agent_feats"].data_ptr()
        T.d_agent_mask = tok["agent_mask_u8"].data_ptr()
        T.d_edge_valid = tok["edge_valid"].data_ptr()
        T.d_task_ids = tok["task_ids"].data_ptr()
        T.d_need = tok["need"].data_ptr()
        T.d_context = tok["context"].data_ptr() if context else None
        T.d_task_order = tok["task_order"].data_ptr() if escort else None
        T.max_tasks, T.max_agents, T.interval, T.event_mask = max_tasks, max_agents, interval, event_mask
        T.agent_feat_dim = afd
        self._tok = T
        self.fused_tokens = tok
        return tok

    def refresh_fused_tokens(self):
        """Fill the fused token tensors for ALL environments with the standalone kernel (after reset/restore)."""
        tok, T = self.fused_tokens, self._tok
        if T.agent_feat_dim == 16:
            rc = self.lib.dll.muav_tokens_escort(C.byref(self.cfg), self.records.data_ptr(), T.max_tasks, T.max_agents,
                                                 T.d_task_feats, T.d_task_mask, T.d_agent_feats, T.d_agent_mask,
                                                 T.d_edge_valid, T.d_task_ids, T.d_task_order, self.n_envs, self._stream())
        elif T.agent_feat_dim == 13:
            rc = self.lib.dll.muav_tokens_commit(C.byref(self.cfg), self.records.data_ptr(), T.max_tasks, T.max_agents,
                                                 T.d_task_feats, T.d_task_mask, T.d_agent_feats, T.d_agent_mask,
                                                 T.d_task_ids, self.n_envs, self._stream())
        elif T.d_context:
            rc = self.lib.dll.muav_tokens_context(C.byref(self.cfg), self.records.data_ptr(), T.max_tasks, T.max_agents, 0,
                                                  T.d_task_feats, T.d_task_mask, T.d_agent_feats, T.d_agent_mask,
                                                  T.d_edge_valid, T.d_task_ids, T.d_context, self.n_envs, self._stream())
        else:
            rc = self.lib.dll.muav_tokens_pair(C.byref(self.cfg), self.records.data_ptr(), T.max_tasks, T.max_agents,
                                               T.d_task_feats, T.d_task_mask, T.d_agent_feats, T.d_agent_mask,
                                               T.d_edge_valid, T.d_task_ids, self.n_envs, self._stream())
        _lib.check(rc, "muav_tokens_pair")
        tok["need"].fill_(1)
        self.launches += 1

    def _tok_ref(self):
        return C.byref(self._tok) if getattr(self, "_tok", None) is not None else None

    # ------------------------------------------------------------------ step
    def _order_args(self):
        """Point muav_step_out at the current / next launch-slot order.  The buffers flip in _order_commit(), i.e. only
        after a launch that really ran and filled the next order."""
        if not self.group_replanners:
            self._out.d_env_order = None
            self._out.d_env_order_next = None
            self._order_next = -1
            return
        cur = self._order_cur
        nxt = 0 if cur != 0 else 1
        self._out.d_env_order = None if cur < 0 else self._order[cur].data_ptr()
        self._out.d_env_order_next = self._order[nxt].data_ptr()
        self._order_next = nxt

    def _order_commit(self, rc, n_steps):
        if rc == 0 and n_steps > 0 and self.n_envs > 0 and self.group_replanners:
            self._order_cur = self._order_next

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev_ctx(self):
        """Make the environment's device current for a native call; nothing to do (and nothing to pay per step) when it
        already is."""
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        return _NO_CTX if torch.cuda.current_device() == idx else torch.cuda.device(self.device)

    def step_batched(self, actions: torch.Tensor, n_steps: int = 1):
        """actions: int32 [E, A, 2] ordered (agent_id, index into last_tasks_info) pairs, agent_id = -1
        terminates an env's list; or int32 [E, A] with one index per agent id (-1 = no action), applied
        in ascending agent id.  Returns (reward f64[E], terminated u8[E], truncated u8[E])."""
        if actions.dim() == 2:
            E, A = actions.shape
            ids = torch.arange(A, device=actions.device, dtype=torch.int32).expand(E, A)
            valid = actions >= 0
            order = torch.argsort((~valid).to(torch.int8), dim=1, stable=True)
            a_sorted = torch.gather(torch.where(valid, ids, torch.full_like(ids, -1)), 1, order)
            i_sorted = torch.gather(actions.to(torch.int32), 1, order)
            actions = torch.stack([a_sorted, i_sorted], dim=2)
        actions = actions.to(device=self.device, dtype=torch.int32).contiguous()
        self._order_args()
        with self._dev_ctx():
            rc = self.lib.dll.muav_step(C.byref(self.cfg), self.records.data_ptr(), self.tapes.data_ptr(),
                                        actions.data_ptr(), None, C.byref(self._out), self._tok_ref(), self.n_envs, n_steps,
                                        self._stream())
        self._order_commit(rc, n_steps)
        _lib.check(rc, "muav_step")
        self.launches += 1
        return self.reward, self.terminated, self.truncated

    def step_host(self, h_actions, h_reward, h_terminated, h_truncated, n_steps: int = 1, hint: Optional[AllocSpec] = None):
        """muav_step_host: ordered actions int32 [E, A, 2] in HOST memory in, reward f64 [E] / terminated u8 [E] /
        truncated u8 [E] in HOST memory out (copies and the final synchronisation happen inside the call).
        `hint`: the AllocSpec whose replan rule the caller's allocator follows (launch-slot grouping only)."""
        ptr = lambda x: x.data_ptr() if isinstance(x, torch.Tensor) else x.ctypes.data
        O = None
        if hint is not None:
            O = _lib.MuavAllocOpts()
            O.mode = 0
            O.order_hint_mode = hint.mode
            O.replan_interval, O.event_mask, O.planner = hint.replan_interval, hint.event_mask, hint.planner
        self._order_args()
        with self._dev_ctx():
            rc = self.lib.dll.muav_ctx_step_host(self._host_ctx(), self.records.data_ptr(), self.tapes.data_ptr(),
                                                 ptr(h_actions), None if O is None else C.byref(O), self._tok_ref(),
                                                 ptr(h_reward), ptr(h_terminated), ptr(h_truncated), n_steps, self._stream(),
                                                 self._out.d_env_order, self._out.d_env_order_next)
        self._order_commit(rc, n_steps)
        _lib.check(rc, "muav_ctx_step_host")
        self.launches += 1

    def allocate_host(self, spec: "AllocSpec", h_actions_out, edge_scores: Optional[torch.Tensor] = None,
                      priorities: Optional[torch.Tensor] = None, reserved: Optional[torch.Tensor] = None,
                      task_order: Optional[torch.Tensor] = None, plan_pri: Optional[torch.Tensor] = None,
                      plan_commit: Optional[torch.Tensor] = None):
        """allocate() with the ordered action list delivered to HOST memory int32 [E, A, 2] (the allocator handing its
        decision to the caller, HungarianAllocator.py:72-208): kernel, one D2H copy and the synchronisation in one call."""
        ptr = lambda x: x.data_ptr() if isinstance(x, torch.Tensor) else x.ctypes.data
        O, keep = self._alloc_opts(spec, edge_scores, priorities, reserved, task_order, plan_pri, plan_commit)
        cur = self._order_cur
        self._out.d_env_order = self._order[cur].data_ptr() if (self.group_replanners and cur >= 0) else None
        self._out.d_env_order_next = None
        with self._dev_ctx():
            rc = self.lib.dll.muav_ctx_allocate_host(self._host_ctx(), self.records.data_ptr(), C.byref(O),
                                                     C.byref(self._out), ptr(h_actions_out), self._stream())
        _lib.check(rc, "muav_ctx_allocate_host")
        self.launches += 1

    def _alloc_opts(self, spec, edge_scores, priorities, reserved, task_order=None, plan_pri=None, plan_commit=None):
        # the same spec over the same tensors every step (a rollout loop): reuse the filled struct
        key = (id(spec), spec.mode, spec.replan_interval, spec.event_mask, spec.planner, spec.commit_fraction,
               spec.commit_threshold, spec.use_visibility, spec.pair_tokens, spec.max_coord, spec.max_tasks_per_agent) + tuple(
            (x.data_ptr(), x.dtype, tuple(x.shape), x.device.index, x.is_contiguous()) if isinstance(x, torch.Tensor) else None
            for x in (edge_scores, priorities, reserved, task_order, plan_pri, plan_commit))
        cached = getattr(self, "_opts_cache", None)
        if cached is not None and cached[0] == key:
            # a copy: callers may fill further fields (the facade backend does) without touching the cached struct
            return _lib.MuavAllocOpts.from_buffer_copy(cached[1]), list(cached[2])
        O, keep = self._alloc_opts_build(spec, edge_scores, priorities, reserved, task_order, plan_pri, plan_commit)
        # the cache holds the argument tensors themselves too, so that the data pointers in the key cannot be reused
        keep = keep + [x for x in (edge_scores, priorities, reserved, task_order, plan_pri, plan_commit)
                       if isinstance(x, torch.Tensor)]
        self._opts_cache = (key, _lib.MuavAllocOpts.from_buffer_copy(O), keep)
        return O, list(keep)

    def _alloc_opts_build(self, spec, edge_scores, priorities, reserved, task_order=None, plan_pri=None, plan_commit=None):
        O = _lib.MuavAllocOpts()
        O.mode = spec.mode
        O.replan_interval = spec.replan_interval
        O.event_mask = spec.event_mask
        O.use_visibility = int(spec.use_visibility)
        O.pair_tokens = int(spec.pair_tokens)
        O.max_coord = spec.max_coord
        O.planner = spec.planner
        O.commit_fraction = spec.commit_fraction
        keep = []
        if edge_scores is not None:
            es = edge_scores.to(device=self.device, dtype=torch.float32).contiguous()
            keep.append(es)
            O.score_rows, O.score_cols = int(es.shape[1]), int(es.shape[2])
            O.d_edge_scores = es.data_ptr()
        if priorities is not None:
            pr = priorities.to(device=self.device, dtype=torch.float64).contiguous()
            keep.append(pr)
            O.d_priorities = pr.data_ptr()
        if reserved is not None:
            rs = reserved.to(device=self.device, dtype=torch.uint8).contiguous()
            keep.append(rs)
            O.d_reserved = rs.data_ptr()
        if task_order is not None:
            to = task_order.to(device=self.device, dtype=torch.int32).contiguous()
            if to.shape != (self.n_envs, max(self.cfg.id_cap, self.cfg.task_cap)):
                raise ValueError("task_order must be [n_envs, id_cap]")
            keep.append(to)
            O.d_task_order = to.data_ptr()
        O.commit_threshold = spec.commit_threshold
        if spec.planner in (6, 7) and spec.max_tasks_per_agent > 1:
            if spec.max_tasks_per_agent > 4:
                raise ValueError("the device market allocators build bundles of at most 4 tasks per agent")
            O.max_tasks_per_agent = spec.max_tasks_per_agent
            need = self.n_agents * spec.max_tasks_per_agent
            bp = getattr(self, "bundle_pairs", None)
            if bp is None or bp.shape[1] != need:
                self.bundle_pairs = torch.zeros(self.n_envs, need, dtype=torch.int32, device=self.device)
                self.n_bundle_pairs = torch.zeros(self.n_envs, dtype=torch.int32, device=self.device)
            O.d_bundle_pairs, O.d_n_bundle_pairs = self.bundle_pairs.data_ptr(), self.n_bundle_pairs.data_ptr()
        if spec.planner == 7:
            O.d_cbba_seed = self.seeds.data_ptr()
        if spec.planner == 3:
            if plan_pri is None or plan_commit is None:
                raise ValueError("planner 3 needs plan_pri [E, max_tasks] and plan_commit [E, max_agents]")
            pp = plan_pri.to(device=self.device, dtype=torch.float32).contiguous()
            pc = plan_commit.to(device=self.device, dtype=torch.float32).contiguous()
            keep += [pp, pc]
            O.score_cols, O.score_rows = int(pp.shape[1]), int(pc.shape[1])
            O.d_plan_pri, O.d_plan_commit = pp.data_ptr(), pc.data_ptr()
        if spec.planner == 4 and (edge_scores is None or task_order is None):
            raise ValueError("planner 4 needs edge_scores [E, max_agents, max_tasks] and task_order from tokens_escort()")
        return O, keep

    def step_allocated(self, spec: AllocSpec, n_steps: int = 1, edge_scores: Optional[torch.Tensor] = None,
                       priorities: Optional[torch.Tensor] = None, reserved: Optional[torch.Tensor] = None,
                       task_order: Optional[torch.Tensor] = None, plan_pri: Optional[torch.Tensor] = None,
                       plan_commit: Optional[torch.Tensor] = None):
        """n_steps fused (allocate -> step) iterations per environment, state resident in shared memory."""
        O, keep = self._alloc_opts(spec, edge_scores, priorities, reserved, task_order, plan_pri, plan_commit)
        self._order_args()
        with self._dev_ctx():
            rc = self.lib.dll.muav_rollout(C.byref(self.cfg), self.records.data_ptr(), self.tapes.data_ptr(), C.byref(O),
                                           C.byref(self._out), self._tok_ref(), self.n_envs, n_steps, self._stream())
        self._order_commit(rc, n_steps)
        _lib.check(rc, "muav_rollout")
        self.launches += 1
        return self.reward, self.terminated, self.truncated

    def allocate(self, spec: AllocSpec, edge_scores: Optional[torch.Tensor] = None,
                 priorities: Optional[torch.Tensor] = None, reserved: Optional[torch.Tensor] = None,
                 actions_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Allocator only: fills self.pairs / self.n_pairs and returns the ordered action tensor
        int32 [E, A, 2] that step_batched accepts (allocate_tasks + _apply_assign)."""
        O, keep = self._alloc_opts(spec, edge_scores, priorities, reserved)
        if actions_out is None:
            actions_out = torch.empty(self.n_envs, self.n_agents, 2, dtype=torch.int32, device=self.device)
        # read-only use of the current launch-slot order (the allocator kernel profits from the grouping too)
        cur = self._order_cur
        self._out.d_env_order = self._order[cur].data_ptr() if (self.group_replanners and cur >= 0) else None
        self._out.d_env_order_next = None
        with self._dev_ctx():
            rc = self.lib.dll.muav_allocate(C.byref(self.cfg), self.records.data_ptr(), C.byref(O), C.byref(self._out),
                                            actions_out.data_ptr(), self.n_envs, self._stream())
        _lib.check(rc, "muav_allocate")
        self.launches += 1
        return actions_out

    # ------------------------------------------------------------------ views
    def metrics(self) -> torch.Tensor:
        out = torch.empty(self.n_envs, _lib.N_METRICS, dtype=torch.float64, device=self.device)
        rc = self.lib.dll.muav_metrics(C.byref(self.cfg), self.records.data_ptr(), out.data_ptr(), self.n_envs,
                                       self._stream())
        _lib.check(rc, "muav_metrics")
        self.launches += 1
        return out

    def metric_dict(self, e: int) -> dict:
        m = self.metrics()[e].cpu().numpy()
        return dict(zip(self.lib.metric_names(), m.tolist()))

    def tokens_pair(self, max_tasks=32, max_agents=16):
        E, dev = self.n_envs, self.device
        tf = torch.empty(E, max_tasks, 13, dtype=torch.float32, device=dev)
        tm = torch.empty(E, max_tasks, dtype=torch.uint8, device=dev)
        af = torch.empty(E, max_agents, 12, dtype=torch.float32, device=dev)
        am = torch.empty(E, max_agents, dtype=torch.uint8, device=dev)
        ev = torch.empty(E, max_agents, max_tasks, dtype=torch.float32, device=dev)
        ids = torch.empty(E, max_tasks, dtype=torch.int32, device=dev)
        rc = self.lib.dll.muav_tokens_pair(C.byref(self.cfg), self.records.data_ptr(), max_tasks, max_agents,
                                           tf.data_ptr(), tm.data_ptr(), af.data_ptr(), am.data_ptr(), ev.data_ptr(),
                                           ids.data_ptr(), E, self._stream())
        _lib.check(rc, "muav_tokens_pair")
        self.launches += 1
        return {"task_feats": tf, "task_mask": tm.bool(), "agent_feats": af, "agent_mask": am.bool(),
                "edge_valid": ev, "task_ids": ids}

    def pair_mask(self, tok: dict, require_valid: bool) -> torch.Tensor:
        """Mask [E, max_agents, max_tasks] of the allocator pairs of the last allocate / step call over the token grid of
        `tok`: the imitation target _expert_mask (require_valid=True, train_pair_cost.py:53-70) or PairCostHybrid's
        _selected_mask (False).  Call before the state moves on (row = i-th live agent of the current state)."""
        E = self.n_envs
        MA, MT = tok["edge_valid"].shape[1], tok["edge_valid"].shape[2]
        mask = torch.empty(E, MA, MT, dtype=torch.float32, device=self.device)
        rc = self.lib.dll.muav_pair_mask(C.byref(self.cfg), self.records.data_ptr(), self.pairs.data_ptr(),
                                         self.n_pairs.data_ptr(), tok["task_ids"].data_ptr(), tok["edge_valid"].data_ptr(),
                                         MT, MA, int(require_valid), mask.data_ptr(), E, self._stream())
        _lib.check(rc, "muav_pair_mask")
        self.launches += 1
        return mask

    def tokens_context(self, max_tasks=32, max_agents=16, raw=False):
        """build_context_pair_tokens(env, raw) for every environment (ContextPairHybrid.py:33-78): pair tokens (the
        per-entity `raw` variant has 9 / 11 features) plus the context vector [E, 8] ([E, 1] when raw)."""
        E, dev = self.n_envs, self.device
        tf = torch.empty(E, max_tasks, 9 if raw else 13, dtype=torch.float32, device=dev)
        tm = torch.empty(E, max_tasks, dtype=torch.uint8, device=dev)
        af = torch.empty(E, max_agents, 11 if raw else 12, dtype=torch.float32, device=dev)
        am = torch.empty(E, max_agents, dtype=torch.uint8, device=dev)
        ev = torch.empty(E, max_agents, max_tasks, dtype=torch.float32, device=dev)
        ids = torch.empty(E, max_tasks, dtype=torch.int32, device=dev)
        ctx = torch.empty(E, 1 if raw else 8, dtype=torch.float32, device=dev)
        rc = self.lib.dll.muav_tokens_context(C.byref(self.cfg), self.records.data_ptr(), max_tasks, max_agents, int(raw),
                                              tf.data_ptr(), tm.data_ptr(), af.data_ptr(), am.data_ptr(), ev.data_ptr(),
                                              ids.data_ptr(), ctx.data_ptr(), E, self._stream())
        _lib.check(rc, "muav_tokens_context")
        self.launches += 1
        return {"task_feats": tf, "task_mask": tm.bool(), "agent_feats": af, "agent_mask": am.bool(),
                "edge_valid": ev, "task_ids": ids, "context": ctx}

    def tokens_commit(self, max_tasks=32, max_agents=16):
        """enrich_commit_tokens(build_att_tokens(env)) for every environment (AttentionCommit.py:49-62)."""
        E, dev = self.n_envs, self.device
        tf = torch.empty(E, max_tasks, 13, dtype=torch.float32, device=dev)
        tm = torch.empty(E, max_tasks, dtype=torch.uint8, device=dev)
        af = torch.empty(E, max_agents, 13, dtype=torch.float32, device=dev)
        am = torch.empty(E, max_agents, dtype=torch.uint8, device=dev)
        ids = torch.empty(E, max_tasks, dtype=torch.int32, device=dev)
        rc = self.lib.dll.muav_tokens_commit(C.byref(self.cfg), self.records.data_ptr(), max_tasks, max_agents,
                                             tf.data_ptr(), tm.data_ptr(), af.data_ptr(), am.data_ptr(), ids.data_ptr(),
                                             E, self._stream())
        _lib.check(rc, "muav_tokens_commit")
        self.launches += 1
        return {"task_feats": tf, "task_mask": tm.bool(), "agent_feats": af, "agent_mask": am.bool(), "task_ids": ids}

    def tokens_escort(self, max_tasks=48, max_agents=16):
        """build_escort_tokens(env) for every environment (AttentionEscort.py:76-241).  `task_order` [E, id_cap] is the
        `tasks` argument / score layout that AllocSpec.att_escort() expects."""
        E, dev = self.n_envs, self.device
        tf = torch.empty(E, max_tasks, 22, dtype=torch.float32, device=dev)
        tm = torch.empty(E, max_tasks, dtype=torch.uint8, device=dev)
        af = torch.empty(E, max_agents, 16, dtype=torch.float32, device=dev)
        am = torch.empty(E, max_agents, dtype=torch.uint8, device=dev)
        ev = torch.empty(E, max_agents, max_tasks, dtype=torch.float32, device=dev)
        ids = torch.empty(E, max_tasks, dtype=torch.int32, device=dev)
        order = torch.empty(E, max(self.cfg.id_cap, self.cfg.task_cap), dtype=torch.int32, device=dev)
        rc = self.lib.dll.muav_tokens_escort(C.byref(self.cfg), self.records.data_ptr(), max_tasks, max_agents,
                                             tf.data_ptr(), tm.data_ptr(), af.data_ptr(), am.data_ptr(), ev.data_ptr(),
                                             ids.data_ptr(), order.data_ptr(), E, self._stream())
        _lib.check(rc, "muav_tokens_escort")
        self.launches += 1
        return {"task_feats": tf, "task_mask": tm.bool(), "agent_feats": af, "agent_mask": am.bool(),
                "edge_valid": ev, "task_ids": ids, "task_order": order}

    def observe(self, max_rows: Optional[int] = None):
        """Observation tensors of _generate_observations (DroneEnv.py:468-492); see include/muav.h."""
        E, A, dev = self.n_envs, self.n_agents, self.device
        mr = int(max_rows or self.cfg.max_tasks)
        ti = torch.empty(E, mr, 21, dtype=torch.float64, device=dev)
        pm = torch.empty(E, mr, dtype=torch.uint8, device=dev)
        lm = torch.empty(E, A, mr, dtype=torch.uint8, device=dev)
        ao = torch.empty(E, A, 9, dtype=torch.float64, device=dev)
        ef = torch.empty(E, 5, dtype=torch.float32, device=dev)
        nr = torch.empty(E, dtype=torch.int32, device=dev)
        rc = self.lib.dll.muav_observe(C.byref(self.cfg), self.records.data_ptr(), mr, ti.data_ptr(), pm.data_ptr(),
                                       lm.data_ptr(), ao.data_ptr(), ef.data_ptr(), nr.data_ptr(), E, self._stream())
        _lib.check(rc, "muav_observe")
        self.launches += 1
        return {"tasks_info": ti, "mask": pm.bool(), "legal_mask": lm.bool(), "agent_obs": ao, "event_flags": ef,
                "n_rows": nr}

    def record_host(self, e: int) -> np.ndarray:
        buf = np.empty(self.record_bytes, dtype=np.uint8)
        rc = self.lib.dll.muav_snapshot(C.byref(self.cfg), self.records.data_ptr(), int(e), buf.ctypes.data, self._stream())
        _lib.check(rc, "muav_snapshot")
        return buf

    def snapshot(self, e: int) -> dict:
        return self.codec.snapshot(self.record_host(e))

    def error_flags(self) -> torch.Tensor:
        off, cnt, dt = self.codec.F["hi"]
        idx = self.codec.extra["ERRFLAGS"]
        hi = self.records[:, off:off + cnt * 4].view(torch.int32)
        return hi[:, idx]

    def header_int(self, name: str) -> torch.Tensor:
        off, cnt, dt = self.codec.F["hi"]
        hi = self.records[:, off:off + cnt * 4].view(torch.int32)
        return hi[:, self.lib.header_index(name)]

    def events_of(self, e: int) -> list:
        return _state.decode_events(int(self.n_events[e].item()), self.events[e].cpu().numpy())

    def bundle_pairs_of(self, e: int) -> list:
        """The whole plan of the last Performance-Impact launch with bundles: [agent id, task id] of every path entry in the
        order allocate_tasks returns them (pairs_of(e) holds the first task of every path, what the step used)."""
        n = int(self.n_bundle_pairs[e].item())
        p = self.bundle_pairs[e, :n].cpu().numpy()
        return [[int(x) >> 16, int(x) & 0xFFFF] for x in p]

    def pairs_of(self, e: int) -> list:
        n = int(self.n_pairs[e].item())
        p = self.pairs[e, :n].cpu().numpy()
        return [[int(x) >> 16, int(x) & 0xFFFF] for x in p]
