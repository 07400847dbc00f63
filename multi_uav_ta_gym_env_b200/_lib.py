"""ctypes binding of the C ABI in include/muav.h (libmuav_b200.so, built in-tree by
__graft_entry__.build()).  There is no CPU fallback: if the CUDA library is missing the
import of the product path fails loudly."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .config import CAP_TABLE, ENGAGE_RANGE, MAX_SPEEDS, RW_KEYS, TASK_DURATION, TASK_TYPES, UAV_TYPES

MAX_GROUPS = 8
N_METRICS = 30
PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CUDA_LIB_PATH = os.environ.get("MUAV_LIB_OVERRIDE") or os.path.join(PKG_DIR, "libmuav_b200.so")  # override: tuning builds only


class MuavConfig(C.Structure):
    _fields_ = [
        ("n_agents", C.c_int32), ("task_cap", C.c_int32), ("n_threats", C.c_int32), ("queue_cap", C.c_int32),
        ("event_cap", C.c_int32), ("n_obstacles", C.c_int32), ("n_groups", C.c_int32),
        ("n_tasks_cfg", C.c_int32), ("max_tasks", C.c_int32), ("max_time_steps", C.c_int32),
        ("multiple_tasks_per_agent", C.c_int32), ("early_terminate", C.c_int32), ("capability_mask", C.c_int32),
        ("saturate_mask", C.c_int32),
        ("hard_windows", C.c_int32), ("burst_mode", C.c_int32), ("dual_region_bursts", C.c_int32),
        ("share_knowledge", C.c_int32), ("escort_enabled", C.c_int32),
        ("threat_delay", C.c_int32), ("window_length", C.c_int32), ("burst_size", C.c_int32),
        ("commit_horizon", C.c_int32),
        ("escort_required_agents", C.c_int32), ("escort_type_mask", C.c_int32),
        ("tape_words", C.c_int32 * 3),
        ("group_start", C.c_int32 * (MAX_GROUPS + 1)),
        ("duration", C.c_int32 * 6),
        ("id_cap", C.c_int32),
        ("arrival_rate", C.c_double), ("sense_radius", C.c_double), ("miss_penalty", C.c_double),
        ("on_time_bonus", C.c_double), ("dynamic_idle_penalty", C.c_double), ("reassign_penalty", C.c_double),
        ("escort_radius", C.c_double), ("escort_requirement", C.c_double),
        ("escort_intercept_radius", C.c_double), ("mutual_support_radius", C.c_double),
        ("threat_gen_prob", C.c_double), ("threat_wide", C.c_double), ("max_coord", C.c_double),
        ("area_w", C.c_double), ("area_h", C.c_double), ("base_x", C.c_double), ("base_y", C.c_double),
        ("contact_line", C.c_double),
        ("rw", C.c_double * 8),
        ("speed", C.c_double * 7),
        ("engage", C.c_double * 7),
        ("cap_table", (C.c_double * 6) * 7),
    ]


class MuavAllocOpts(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("replan_interval", C.c_int32), ("event_mask", C.c_int32),
        ("use_visibility", C.c_int32), ("pair_tokens", C.c_int32), ("score_rows", C.c_int32),
        ("score_cols", C.c_int32), ("score_f64", C.c_int32),
        ("planner", C.c_int32), ("order_hint_mode", C.c_int32),
        ("commit_fraction", C.c_double),
        ("max_coord", C.c_double),
        ("d_edge_scores", C.c_void_p), ("d_priorities", C.c_void_p), ("d_reserved", C.c_void_p),
        ("d_task_order", C.c_void_p),
        ("d_plan_pri", C.c_void_p), ("d_plan_commit", C.c_void_p), ("commit_threshold", C.c_double),
        ("d_cbba_seed", C.c_void_p),
        ("max_tasks_per_agent", C.c_int32), ("reserved0", C.c_int32),
        ("d_bundle_pairs", C.c_void_p), ("d_n_bundle_pairs", C.c_void_p),
    ]


class MuavStepOut(C.Structure):
    _fields_ = [
        ("d_reward", C.c_void_p), ("d_terminated", C.c_void_p), ("d_truncated", C.c_void_p),
        ("d_n_events", C.c_void_p), ("d_events", C.c_void_p), ("d_n_pairs", C.c_void_p),
        ("d_pairs", C.c_void_p), ("d_n_open", C.c_void_p),
        ("d_env_order", C.c_void_p), ("d_env_order_next", C.c_void_p), ("d_actions_ws", C.c_void_p),
    ]


class MuavTokenOut(C.Structure):
    _fields_ = [
        ("d_task_feats", C.c_void_p), ("d_task_mask", C.c_void_p), ("d_agent_feats", C.c_void_p),
        ("d_agent_mask", C.c_void_p), ("d_edge_valid", C.c_void_p), ("d_task_ids", C.c_void_p), ("d_need", C.c_void_p),
        ("max_tasks", C.c_int32), ("max_agents", C.c_int32), ("interval", C.c_int32), ("event_mask", C.c_int32),
        ("d_context", C.c_void_p), ("agent_feat_dim", C.c_int32), ("d_task_order", C.c_void_p),
    ]


class MuavAttPairOffsets(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "agent_proj_w", "agent_proj_b", "task_proj_w", "task_proj_b", "type_embed",
        "enc_in_w", "enc_in_b", "enc_out_w", "enc_out_b", "enc_l1_w", "enc_l1_b", "enc_l2_w", "enc_l2_b",
        "enc_n1_w", "enc_n1_b", "enc_n2_w", "enc_n2_b",
        "a2t_in_w", "a2t_in_b", "a2t_out_w", "a2t_out_b", "t2a_in_w", "t2a_in_b", "t2a_out_w", "t2a_out_b",
        "head1_w", "head1_b", "head2_w", "head2_b", "head3_w", "head3_b", "ctx_proj_w", "ctx_proj_b", "has_context")]


class MuavAttCommitOffsets(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in ("agent_proj_w", "agent_proj_b", "task_proj_w", "task_proj_b", "type_embed")]
                + [(n, C.c_int32 * 2) for n in ("enc_in_w", "enc_in_b", "enc_out_w", "enc_out_b", "enc_l1_w", "enc_l1_b",
                                                "enc_l2_w", "enc_l2_b", "enc_n1_w", "enc_n1_b", "enc_n2_w", "enc_n2_b")]
                + [(n, C.c_int32) for n in ("priority_w", "priority_b", "commit_w", "commit_b")])


class MuavAttCoalOffsets(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in ("agent_proj_w", "agent_proj_b", "task_proj_w", "task_proj_b", "type_embed")]
                + [(n, C.c_int32 * 2) for n in ("enc_in_w", "enc_in_b", "enc_out_w", "enc_out_b", "enc_l1_w", "enc_l1_b",
                                                "enc_l2_w", "enc_l2_b", "enc_n1_w", "enc_n1_b", "enc_n2_w", "enc_n2_b")]
                + [(n, C.c_int32) for n in ("a2t_in_w", "a2t_in_b", "a2t_out_w", "a2t_out_b", "t2a_in_w", "t2a_in_b",
                                            "t2a_out_w", "t2a_out_b", "head1_w", "head1_b", "head2_w", "head2_b", "head3_w",
                                            "head3_b")])


# every symbol include/muav.h declares
ABI_SYMBOLS = [
    "muav_version", "muav_config_size", "muav_record_bytes", "muav_scratch_bytes", "muav_hot_bytes", "muav_num_fields",
    "muav_field_info", "muav_header_index", "muav_step", "muav_allocate", "muav_step_host", "muav_lsap",
    "muav_avoid_obstacles", "muav_metric_name", "muav_metrics", "muav_tokens_pair", "muav_tokens_commit", "muav_tokens_escort", "muav_tokens_context", "muav_pair_mask", "muav_observe",
    "muav_att_pair_scores", "muav_att_context_pair_scores", "muav_rollout", "muav_state_bytes", "muav_tape_bytes", "muav_reset_upload", "muav_snapshot",
    "muav_ctx_create", "muav_ctx_destroy", "muav_ctx_step_host", "muav_ctx_allocate_host", "muav_att_commit_vectors",
    "muav_att_pair_tc_floats", "muav_att_pair_tc_pack", "muav_att_pair_scores_tc", "muav_att_coalition_scores",
    "muav_att_commit_tc_floats", "muav_att_commit_tc_pack", "muav_att_commit_vectors_tc",
]


class Lib:
    """Layout queries shared by the CUDA library and the tests' host-check build."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise RuntimeError(f"native library not found: {path} (run `python -c 'import __graft_entry__ as g; g.build()'`)")
        self.path = path
        self.dll = C.CDLL(path)
        d = self.dll
        d.muav_version.restype = C.c_char_p
        d.muav_config_size.restype = C.c_size_t
        d.muav_record_bytes.restype = C.c_size_t
        d.muav_record_bytes.argtypes = [C.POINTER(MuavConfig)]
        d.muav_scratch_bytes.restype = C.c_size_t
        d.muav_scratch_bytes.argtypes = [C.POINTER(MuavConfig)]
        d.muav_hot_bytes.restype = C.c_size_t
        d.muav_hot_bytes.argtypes = [C.POINTER(MuavConfig)]
        d.muav_num_fields.restype = C.c_int
        d.muav_field_info.restype = C.c_int
        d.muav_field_info.argtypes = [C.POINTER(MuavConfig), C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int64),
                                      C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
        d.muav_header_index.restype = C.c_int
        d.muav_header_index.argtypes = [C.c_char_p]
        if d.muav_config_size() != C.sizeof(MuavConfig):
            raise RuntimeError(f"muav_config ABI mismatch: C {d.muav_config_size()} vs ctypes {C.sizeof(MuavConfig)}")

    def record_bytes(self, cfg: MuavConfig) -> int:
        return int(self.dll.muav_record_bytes(C.byref(cfg)))

    def hot_bytes(self, cfg: MuavConfig) -> int:
        return int(self.dll.muav_hot_bytes(C.byref(cfg)))

    def scratch_bytes(self, cfg: MuavConfig) -> int:
        return int(self.dll.muav_scratch_bytes(C.byref(cfg)))

    def fields(self, cfg: MuavConfig):
        """{name: (offset, count, numpy dtype)}"""
        out = {}
        name = C.c_char_p()
        off = C.c_int64()
        cnt = C.c_int64()
        esz = C.c_int32()
        unsigned = {"k_tbl_lo", "k_tbl_hi", "known", "open_mask"}
        for i in range(self.dll.muav_num_fields()):
            rc = self.dll.muav_field_info(C.byref(cfg), i, C.byref(name), C.byref(off), C.byref(cnt), C.byref(esz))
            if rc != 0:
                raise RuntimeError("muav_field_info failed")
            n = name.value.decode()
            if esz.value == 8:
                dt = np.float64
            elif esz.value == 4:
                dt = np.uint32 if n in unsigned else np.int32
            else:
                dt = np.int16
            out[n] = (int(off.value), int(cnt.value), np.dtype(dt))
        return out

    def header_index(self, name: str) -> int:
        i = int(self.dll.muav_header_index(name.encode()))
        if i < 0:
            raise KeyError(name)
        return i


_cuda_lib = None


def cuda_lib() -> "CudaLib":
    """The product library.  Raises if it has not been built -- there is no fallback."""
    global _cuda_lib
    if _cuda_lib is None:
        _cuda_lib = CudaLib(CUDA_LIB_PATH)
    return _cuda_lib


class CudaLib(Lib):
    def __init__(self, path):
        super().__init__(path)
        d = self.dll
        P = C.c_void_p
        d.muav_step.restype = C.c_int
        d.muav_step.argtypes = [C.POINTER(MuavConfig), P, P, P, C.POINTER(MuavAllocOpts), C.POINTER(MuavStepOut),
                                C.POINTER(MuavTokenOut), C.c_int, C.c_int, P]
        d.muav_allocate.restype = C.c_int
        d.muav_allocate.argtypes = [C.POINTER(MuavConfig), P, C.POINTER(MuavAllocOpts), C.POINTER(MuavStepOut), P,
                                    C.c_int, P]
        d.muav_step_host.restype = C.c_int
        d.muav_step_host.argtypes = [C.POINTER(MuavConfig), P, P, P, C.POINTER(MuavAllocOpts), C.POINTER(MuavTokenOut),
                                     P, P, P, C.c_int, C.c_int, P, P, P]
        d.muav_ctx_create.restype = C.c_int
        d.muav_ctx_create.argtypes = [C.POINTER(MuavConfig), C.c_int, C.c_int, C.POINTER(P)]
        d.muav_ctx_destroy.restype = None
        d.muav_ctx_destroy.argtypes = [P]
        d.muav_ctx_step_host.restype = C.c_int
        d.muav_ctx_step_host.argtypes = [P, P, P, P, C.POINTER(MuavAllocOpts), C.POINTER(MuavTokenOut), P, P, P, C.c_int, P,
                                         P, P]
        d.muav_ctx_allocate_host.restype = C.c_int
        d.muav_ctx_allocate_host.argtypes = [P, P, C.POINTER(MuavAllocOpts), C.POINTER(MuavStepOut), P, P]
        d.muav_lsap.restype = C.c_int
        d.muav_lsap.argtypes = [P, P, P, C.c_int, C.c_int, P, C.c_int, P]
        d.muav_avoid_obstacles.restype = C.c_int
        d.muav_avoid_obstacles.argtypes = [P, P, P, C.c_int, P, C.c_int, P]
        d.muav_metric_name.restype = C.c_char_p
        d.muav_metric_name.argtypes = [C.c_int]
        d.muav_metrics.restype = C.c_int
        d.muav_metrics.argtypes = [C.POINTER(MuavConfig), P, P, C.c_int, P]
        d.muav_tokens_pair.restype = C.c_int
        d.muav_tokens_pair.argtypes = [C.POINTER(MuavConfig), P, C.c_int, C.c_int, P, P, P, P, P, P, C.c_int, P]
        d.muav_tokens_commit.restype = C.c_int
        d.muav_tokens_commit.argtypes = [C.POINTER(MuavConfig), P, C.c_int, C.c_int, P, P, P, P, P, C.c_int, P]
        d.muav_rollout.restype = C.c_int
        d.muav_rollout.argtypes = [C.POINTER(MuavConfig), P, P, C.POINTER(MuavAllocOpts), C.POINTER(MuavStepOut), P, C.c_int,
                                   C.c_int, P]
        d.muav_state_bytes.restype = C.c_size_t
        d.muav_state_bytes.argtypes = [C.POINTER(MuavConfig), C.c_int]
        d.muav_tape_bytes.restype = C.c_size_t
        d.muav_tape_bytes.argtypes = [C.POINTER(MuavConfig), C.c_int]
        d.muav_reset_upload.restype = C.c_int
        d.muav_reset_upload.argtypes = [C.POINTER(MuavConfig), P, P, P, P, C.c_int, P]
        d.muav_snapshot.restype = C.c_int
        d.muav_snapshot.argtypes = [C.POINTER(MuavConfig), P, C.c_int, P, P]
        d.muav_pair_mask.restype = C.c_int
        d.muav_pair_mask.argtypes = [C.POINTER(MuavConfig), P, P, P, P, P, C.c_int, C.c_int, C.c_int, P, C.c_int, P]
        d.muav_tokens_context.restype = C.c_int
        d.muav_tokens_context.argtypes = [C.POINTER(MuavConfig), P, C.c_int, C.c_int, C.c_int, P, P, P, P, P, P, P, C.c_int, P]
        d.muav_tokens_escort.restype = C.c_int
        d.muav_tokens_escort.argtypes = [C.POINTER(MuavConfig), P, C.c_int, C.c_int, P, P, P, P, P, P, P, C.c_int, P]
        d.muav_att_pair_scores.restype = C.c_int
        d.muav_att_pair_scores.argtypes = [P, C.POINTER(MuavAttPairOffsets), P, P, P, P, P, P, P, C.c_int, C.c_int,
                                           C.c_int, C.c_float, P, P]
        d.muav_att_context_pair_scores.restype = C.c_int
        d.muav_att_context_pair_scores.argtypes = [P, C.POINTER(MuavAttPairOffsets), P, P, P, P, P, P, P, P, C.c_int, C.c_int,
                                                   C.c_int, C.c_float, P, P]
        d.muav_att_pair_tc_floats.restype = C.c_int64
        d.muav_att_pair_tc_floats.argtypes = []
        d.muav_att_pair_tc_pack.restype = C.c_int
        d.muav_att_pair_tc_pack.argtypes = [P, C.POINTER(MuavAttPairOffsets), P, P]
        d.muav_att_pair_scores_tc.restype = C.c_int
        d.muav_att_pair_scores_tc.argtypes = [P, C.POINTER(MuavAttPairOffsets), P, P, P, P, P, P, P, P, P, C.c_int, C.c_int,
                                              C.c_int, C.c_float, P, P]
        d.muav_att_commit_tc_floats.restype = C.c_int64
        d.muav_att_commit_tc_floats.argtypes = []
        d.muav_att_commit_tc_pack.restype = C.c_int
        d.muav_att_commit_tc_pack.argtypes = [P, C.POINTER(MuavAttCommitOffsets), P, P]
        d.muav_att_commit_vectors_tc.restype = C.c_int
        d.muav_att_commit_vectors_tc.argtypes = [P, C.POINTER(MuavAttCommitOffsets), P, P, P, P, P, P, P, C.c_int, C.c_int,
                                                 C.c_int, P, P, P]
        d.muav_att_coalition_scores.restype = C.c_int
        d.muav_att_coalition_scores.argtypes = [P, C.POINTER(MuavAttCoalOffsets), P, P, P, P, P, P, P, C.c_int, C.c_int,
                                                C.c_int, P, P]
        d.muav_att_commit_vectors.restype = C.c_int
        d.muav_att_commit_vectors.argtypes = [P, C.POINTER(MuavAttCommitOffsets), P, P, P, P, P, P, C.c_int, C.c_int, C.c_int,
                                              P, P, P]
        d.muav_observe.restype = C.c_int
        d.muav_observe.argtypes = [C.POINTER(MuavConfig), P, C.c_int, P, P, P, P, P, P, C.c_int, P]

    def metric_names(self):
        return [self.dll.muav_metric_name(i).decode() for i in range(N_METRICS)]


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed with code {rc}" + (f" (cudaError {-1000 - rc})" if rc <= -1000 else ""))


def build_config(opts, task_cap=None, queue_cap=16, event_cap=None, id_cap=None) -> MuavConfig:
    """agentEnvOptions -> muav_config.  Derived constants follow MultiUAVEnv.__init__
    (mUAV_TA/DroneEnv.py:73-323) and the entity constructors."""
    g = lambda n, d=None: getattr(opts, n, d)
    cfg = MuavConfig()
    agents = dict(g("agents"))
    tasks = dict(g("tasks"))
    threats = [tuple(x) for x in (g("threats_list") or [])]
    n_agents = sum(agents.values())
    n_threats = sum(c for _, c in threats)
    n_tasks_cfg = sum(tasks.values()) + 1
    max_tasks = n_tasks_cfg + 28
    if len(threats) > MAX_GROUPS:
        raise ValueError("too many threat groups")
    if not g("multiple_agents_per_task", True):
        raise NotImplementedError("multiple_agents_per_task=False is a dead branch in the reference (DroneEnv.py:935)")
    escort = bool(g("escort_enabled", False))
    # Task ids: arrivals stop at max_tasks-1 (DroneEnv.py:1652), every threat adds one Int task, escorts are re-created
    # every step a recon sits on a Rec (~250 tasks per WPS_escort episode).
    base = max(max_tasks - 1, sum(tasks.values()) + len(threats)) + n_threats
    if id_cap is None:
        id_cap = base + (448 if escort else 0)   # 8192 WPS_escort seeds: up to ~390 ids in one episode
        id_cap = (id_cap + 31) // 32 * 32
    # Task slots: open tasks plus closed ones that a queue / the escort map / a live threat still references.
    if task_cap == "all":
        # one slot per id plus the head-room below which the step starts recycling (agents + threats + 2 new tasks per
        # step): nothing is ever recycled -- the full task history of env.tasks, for the E = 1 facade
        task_cap = (id_cap + n_agents + n_threats + 2 + 15) // 16 * 16
    if task_cap is None:
        task_cap = base + (24 if escort else 0)
        task_cap = (task_cap + 15) // 16 * 16
    id_cap = max(int(id_cap), int(task_cap))
    if event_cap is None:
        event_cap = max(32, 4 * n_agents + 2 * n_threats + 8)
    cfg.n_agents = n_agents
    cfg.task_cap = int(task_cap)
    cfg.id_cap = int(id_cap)
    cfg.n_threats = n_threats
    cfg.queue_cap = int(queue_cap)
    cfg.event_cap = int(event_cap)
    cfg.n_obstacles = int(g("num_obstacles", 0))
    cfg.n_groups = len(threats)
    cfg.n_tasks_cfg = n_tasks_cfg
    cfg.max_tasks = max_tasks
    cfg.max_time_steps = int(g("max_time_steps", 150))
    cfg.multiple_tasks_per_agent = int(bool(g("multiple_tasks_per_agent", False)))
    cfg.early_terminate = int(bool(g("early_terminate", False)))
    cfg.capability_mask = int(bool(g("capability_mask", False)))
    cfg.saturate_mask = int(bool(g("saturate_mask", False)))
    cfg.hard_windows = int(bool(g("hard_windows", False)))
    cfg.burst_mode = int(bool(g("burst_mode", False)))
    cfg.dual_region_bursts = int(bool(g("dual_region_bursts", False)))
    cfg.share_knowledge = int(bool(g("share_knowledge", True)))
    cfg.escort_enabled = int(escort)
    cfg.threat_delay = int(g("threat_delay", 0) or 0)
    cfg.window_length = int(g("window_length", 30) or 30)
    cfg.burst_size = int(g("burst_size", 3) or 3)
    cfg.commit_horizon = int(g("commit_horizon", 0) or 0)
    ereq = float(g("escort_requirement", 1.2) or 1.2)
    cfg.escort_required_agents = max(2, int(np.ceil(ereq)))
    mask = 0
    for name in tuple(g("escort_agent_types", ("F1", "F2")) or ("F1", "F2")):
        mask |= 1 << UAV_TYPES.index(name)
    cfg.escort_type_mask = mask
    cfg.tape_words[0] = max(512, 16 * n_threats + 128)
    cfg.tape_words[1] = 2048
    cfg.tape_words[2] = 128
    acc = 0
    for i in range(MAX_GROUPS + 1):
        cfg.group_start[i] = acc
        if i < len(threats):
            acc += threats[i][1]
    for i, tt in enumerate(TASK_TYPES):
        cfg.duration[i] = TASK_DURATION[tt]
    cfg.arrival_rate = float(g("arrival_rate", 0.0) or 0.0)
    cfg.sense_radius = float(g("sense_radius", 0.0) or 0.0)
    cfg.miss_penalty = float(g("miss_penalty", 25.0) or 0.0)
    cfg.on_time_bonus = float(g("on_time_bonus", 10.0) or 0.0)
    cfg.dynamic_idle_penalty = float(g("dynamic_idle_penalty", 0.0) or 0.0)
    cfg.reassign_penalty = float(g("reassign_penalty", 0.0) or 0.0)
    cfg.escort_radius = float(g("escort_radius", 70.0) or 70.0)
    cfg.escort_requirement = ereq
    cfg.escort_intercept_radius = float(g("escort_intercept_radius", 100.0) or 100.0)
    cfg.mutual_support_radius = float(g("mutual_support_radius", 80.0) or 80.0)
    fr = g("simulation_frame_rate", 0.01)
    cfg.threat_gen_prob = 0.7 / fr * 0.02
    cfg.threat_wide = 1200 / 10
    cfg.max_coord = 1200.0
    cfg.area_w = 1200.0
    cfg.area_h = 700.0
    cfg.base_x = 400.0
    cfg.base_y = 680.0
    cfg.contact_line = 550.0
    rw = g("reward_weights", None) or {}
    for i, k in enumerate(RW_KEYS):
        cfg.rw[i] = float(rw.get(k, 0.0))
    for i, ut in enumerate(UAV_TYPES):
        cfg.speed[i] = MAX_SPEEDS[ut] / fr * 0.02
        cfg.engage[i] = ENGAGE_RANGE[ut]
        for j in range(6):
            cfg.cap_table[i][j] = CAP_TABLE[ut][j]
    return cfg
