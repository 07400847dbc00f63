"""Canonical per-env snapshot of the REFERENCE's live objects (DroneEnv.MultiUAVEnv).

The same dictionary of numpy arrays is produced by
  * this extractor (reference objects -> arrays),
  * oracle/sim.py  (oracle state is already in this form),
  * multi_uav_ta_gym_env_b200.state.unpack_snapshot (CUDA record -> arrays),
so parity is `digest(a) == digest(b)` or field-by-field equality.

Conventions: task index k = task.id - 1 (ids are allocated monotonically and
appended in the same order, DroneEnv.py:325-328,649,692,1627,1656,1895);
queue entries / references hold task *ids* with 0 = idle/none; agent index =
UAV.id = position in env.agents_obj (DroneEnv.py:605-610); threat index =
Threat.id (creation order, DroneEnv.py:714-729).
"""
from __future__ import annotations

import hashlib

import numpy as np

Q_CAP = 16

UAV_TYPES = ["R1", "R2", "E1", "F1", "F2", "T1", "T2"]
TASK_TYPES = ["Hold", "Rec", "Att", "Def", "Int", "Det"]
EVENT_TAGS = ["Reset_Allocation", "Agent_Fail", "New_Threat", "Escort_Created", "Escort_Retired"]

# order matters: digest() walks this list
INT_SCALARS = [
    "t", "n_tasks", "n_thr_active", "n_reallocations", "n_task_switches", "n_arrivals",
    "pending_reset", "n_missed", "n_on_time", "n_windowed", "idle_reserve_steps",
    "burst_toggle", "escort_requests", "escort_completed", "escort_failed",
    "escort_required_steps", "escort_covered_steps", "protection_breaches",
    "threats_intercepted", "recon_losses", "escort_losses", "mutual_support",
    "protected_rec_completed", "n_reached", "conclusion_time",
]
F64_SCALARS = ["F_Reward", "total_distance"]
AGENT_FIELDS = [
    "a_pos", "a_state", "a_task_start", "a_fail_event", "a_type", "a_caps", "a_ammo",
    "a_nft", "a_nfp", "a_re_eval", "a_last_task", "a_commit_until", "a_qlen", "a_queue",
    "a_dist", "a_escort",
]
TASK_FIELDS = [
    "k_pos", "k_type", "k_status", "k_cur", "k_alloc", "k_done_ti", "k_org_ti",
    "k_init_time", "k_done_time", "k_created_at", "k_deadline", "k_counted",
    "k_final_quality", "k_kind", "k_required_agents", "k_elig", "k_threat",
    "k_prot_agent", "k_prot_task", "k_reveal_t", "k_det_time", "k_tbl_mask", "k_reached",
]
THREAT_FIELDS = [
    "h_pos", "h_status", "h_type", "h_group", "h_ammo", "h_target", "h_mission", "h_task",
    "h_det_task", "h_spawned", "h_order",
]
OTHER_FIELDS = ["known", "events"]
ALL_FIELDS = INT_SCALARS + F64_SCALARS + AGENT_FIELDS + TASK_FIELDS + THREAT_FIELDS + OTHER_FIELDS


def _aid(agent):
    return -1 if agent is None else int(agent.id)


def snapshot(env) -> dict:
    A = len(env.agents_obj)
    T = len(env.tasks)
    s = {}
    s["t"] = int(env.time_steps)
    s["n_tasks"] = T
    s["n_thr_active"] = len(env.threats)
    s["n_reallocations"] = int(env.n_reallocations)
    s["n_task_switches"] = int(env.n_task_switches)
    s["n_arrivals"] = int(env.n_arrivals)
    s["pending_reset"] = int(bool(env._pending_reset))
    s["n_missed"] = int(env.n_missed_windows)
    s["n_on_time"] = int(env.n_on_time)
    s["n_windowed"] = int(env.n_windowed_tasks)
    s["idle_reserve_steps"] = int(env._idle_reserve_steps)
    s["burst_toggle"] = int(env._burst_region_toggle)
    s["escort_requests"] = int(env.escort_requests)
    s["escort_completed"] = int(env.escort_completed)
    s["escort_failed"] = int(env.escort_failed)
    s["escort_required_steps"] = int(env.escort_required_steps)
    s["escort_covered_steps"] = int(env.escort_covered_steps)
    s["protection_breaches"] = int(env.protection_breaches)
    s["threats_intercepted"] = int(env.threats_intercepted)
    s["recon_losses"] = int(env.recon_losses)
    s["escort_losses"] = int(env.escort_losses)
    s["mutual_support"] = int(env.mutual_support_engagements)
    s["protected_rec_completed"] = int(env.protected_rec_completed)
    s["n_reached"] = len(env.reached_tasks)
    s["conclusion_time"] = int(env.conclusion_time)
    s["F_Reward"] = float(env.F_Reward)
    s["total_distance"] = float(env.total_distance)

    # ---- agents
    a_pos = np.zeros((A, 2)); a_nfp = np.zeros((A, 2)); a_caps = np.zeros((A, 6))
    a_nft = np.zeros(A); a_dist = np.zeros(A)
    ai = {k: np.zeros(A, np.int64) for k in (
        "a_state", "a_task_start", "a_fail_event", "a_type", "a_ammo", "a_re_eval",
        "a_last_task", "a_commit_until", "a_qlen", "a_escort")}
    a_queue = np.zeros((A, Q_CAP), np.int64)
    for i, a in enumerate(env.agents_obj):
        assert a.id == i
        a_pos[i] = np.asarray(a.position, dtype=np.float64)
        a_nfp[i] = np.asarray(a.next_free_position, dtype=np.float64)
        a_caps[i] = a.currentCap2Task
        a_nft[i] = float(a.next_free_time)
        a_dist[i] = float(env.agent_distances[i])
        ai["a_state"][i] = a.state
        ai["a_task_start"][i] = a.task_start
        ai["a_fail_event"][i] = a.fail_event
        ai["a_type"][i] = UAV_TYPES.index(a.type)
        ai["a_ammo"][i] = a.attackCap
        ai["a_re_eval"][i] = int(bool(a.re_eval))
        ai["a_last_task"][i] = -1 if a.last_task is None else int(a.last_task.id)
        ai["a_commit_until"][i] = int(a.commit_until)
        q = [int(t.id) for t in a.tasks]
        if 0 in q:
            # idle only ever appears as the sole queue entry
            assert q == [0], q
            q = []
        assert len(q) <= Q_CAP, q
        ai["a_qlen"][i] = len(q)
        a_queue[i, : len(q)] = q
        esc = env._escort_by_recon.get(a.name)
        ai["a_escort"][i] = 0 if esc is None else int(esc.id)
    s.update(a_pos=a_pos, a_nfp=a_nfp, a_caps=a_caps, a_nft=a_nft, a_dist=a_dist, a_queue=a_queue, **ai)

    # ---- tasks
    k_pos = np.zeros((T, 2)); k_cur = np.zeros((T, 6)); k_alloc = np.zeros((T, 6))
    kf = {k: np.zeros(T) for k in ("k_done_ti", "k_org_ti", "k_init_time", "k_done_time", "k_final_quality")}
    ki = {k: np.zeros(T, np.int64) for k in (
        "k_type", "k_status", "k_created_at", "k_deadline", "k_counted", "k_kind",
        "k_required_agents", "k_elig", "k_threat", "k_prot_agent", "k_prot_task",
        "k_reveal_t", "k_tbl_mask", "k_reached")}
    k_det_time = np.full((T, A), -1.0)
    name_to_id = {a.name: a.id for a in env.agents_obj}
    reveal = {}
    for rt, tid in env.pending_reveals:
        assert tid not in reveal
        reveal[tid] = rt
    for k, t in enumerate(env.tasks):
        assert t.id == k + 1, (t.id, k)
        ti = t.typeIdx
        k_pos[k] = np.asarray(t.position, dtype=np.float64)
        k_cur[k] = t.currentReqs
        k_alloc[k] = t.allocatedReqs
        kf["k_done_ti"][k] = t.doneReqs[ti]
        kf["k_org_ti"][k] = t.orgReqs[ti]
        kf["k_init_time"][k] = float(t.initTime)
        kf["k_done_time"][k] = float(t.doneTime)
        kf["k_final_quality"][k] = float(t.final_quality)
        ki["k_type"][k] = ti
        ki["k_status"][k] = t.status
        ki["k_created_at"][k] = int(t.created_at or 0)
        dl = getattr(t, "hard_deadline", None)
        ki["k_deadline"][k] = -1 if dl is None else int(dl)
        ki["k_counted"][k] = int(bool(getattr(t, "_wps_outcome_counted", False)))
        ki["k_kind"][k] = 1 if t.kind == "Escort" else 0
        ki["k_required_agents"][k] = int(t.required_agents or 0)
        el = t.eligible_agent_types
        if el is None:
            ki["k_elig"][k] = 0
        else:
            m = 0
            for name in el:
                m |= 1 << UAV_TYPES.index(name)
            ki["k_elig"][k] = m
        thr = t.relative_threat
        ki["k_threat"][k] = -1 if thr is None else int(thr.id)
        ki["k_prot_agent"][k] = _aid(t.protected_agent)
        ki["k_prot_task"][k] = 0 if t.protected_task is None else int(t.protected_task.id)
        ki["k_reveal_t"][k] = reveal.get(t.id, -1)
        for aid, det in t.allocationDetails.items():
            k_det_time[k, aid] = float(det[1])
        m = 0
        if t.id < len(env.allocation_table):
            for name in env.allocation_table[t.id]:
                m |= 1 << name_to_id[name]
        ki["k_tbl_mask"][k] = np.uint64(m).view(np.int64) if m >= 2**63 else m
        ki["k_reached"][k] = int(t.id in env.reached_tasks)
    s.update(k_pos=k_pos, k_cur=k_cur, k_alloc=k_alloc, k_det_time=k_det_time, **kf, **ki)

    # ---- threats (all groups + active, indexed by id)
    allthr = {}
    for g in env.threats_groups:
        for th in g:
            allthr[th.id] = (th, 0)
    for th in env.threats:
        allthr[th.id] = (th, 1)
    H = len(allthr)
    h_pos = np.zeros((H, 2))
    hi = {k: np.zeros(H, np.int64) for k in (
        "h_status", "h_type", "h_group", "h_ammo", "h_target", "h_mission", "h_task",
        "h_det_task", "h_spawned")}
    for hid in range(H):
        th, sp = allthr[hid]
        h_pos[hid] = np.asarray(th.position, dtype=np.float64)
        hi["h_status"][hid] = th.status
        hi["h_type"][hid] = UAV_TYPES.index(th.threat_type)
        hi["h_group"][hid] = th.threat_group
        hi["h_ammo"][hid] = th.attackCap
        hi["h_target"][hid] = _aid(th.target_agent)
        hi["h_mission"][hid] = _aid(th.mission_target_agent)
        hi["h_task"][hid] = 0 if th.relative_task is None else int(th.relative_task.id)
        hi["h_det_task"][hid] = int(th.relative_detect_task.id)
        hi["h_spawned"][hid] = sp
    s.update(h_pos=h_pos, **hi)
    s["h_order"] = np.asarray([th.id for th in env.threats], np.int64)

    known = np.zeros((A, T), np.int64)
    for a in env.agents_obj:
        for tid in env.agent_known_tasks.get(a.name, ()):
            known[a.id, tid - 1] = 1
    s["known"] = known
    ev = [[EVENT_TAGS.index(e[0]), int(e[1])] for e in env.event_list]
    s["events"] = np.asarray(ev, np.int64).reshape(-1, 2)
    return s


DEAD_BLANK = {"k_pos": 0.0, "k_cur": 0.0, "k_alloc": 0.0, "k_done_ti": 0.0, "k_org_ti": 0.0, "k_init_time": 0.0,
              "k_done_time": 0.0, "k_created_at": 0, "k_deadline": 0, "k_counted": 0, "k_final_quality": 0.0,
              "k_kind": 0, "k_required_agents": 0, "k_elig": 0, "k_threat": 0, "k_prot_agent": 0, "k_prot_task": 0,
              "k_tbl_mask": 0, "k_reached": 0}


def dead_tasks(s: dict) -> np.ndarray:
    """Closed tasks that nothing refers to any more: no agent queue or last_task (the switch penalty reads the old
    head's type / position, DroneEnv.py:852,859), not an entry of _escort_by_recon (:1977-2000 keeps visiting stale
    entries), not the task of a threat that is not destroyed (update_threats keeps writing its position, :1740).
    Closed tasks never reopen (:1460), so every other field of such a task is dead data in the reference; the
    batched state recycles their storage."""
    status = np.asarray(s["k_status"])
    T = len(status)
    ref = np.zeros(T + 1, dtype=bool)
    qlen = np.asarray(s["a_qlen"])
    queue = np.asarray(s["a_queue"])
    for a in range(len(qlen)):
        for tid in queue[a, : qlen[a]]:
            ref[int(tid)] = True
    for tid in np.asarray(s["a_last_task"]):
        if tid > 0:
            ref[int(tid)] = True
    for tid in np.asarray(s["a_escort"]):
        if tid > 0:
            ref[int(tid)] = True
    h_status, h_task = np.asarray(s["h_status"]), np.asarray(s["h_task"])
    for hid in np.asarray(s["h_order"]):
        if h_status[hid] != 2 and h_task[hid] > 0:
            ref[int(h_task[hid])] = True
    return (status == 2) & ~ref[1:]


def canonicalize(s: dict) -> dict:
    """(1) allocationDetails of a CLOSED task are dead data (Task.removeAgentCap is a no-op once status == 2,
    DroneEnvComponents.py:282): the reference keeps stale entries, the batched state does not.
    (2) every per-task field of a dead task (see dead_tasks) is blanked."""
    out = dict(s)
    det = np.array(s["k_det_time"], dtype=np.float64, copy=True)
    closed = np.asarray(s["k_status"]) == 2
    det[closed, :] = -1.0
    out["k_det_time"] = det
    dead = dead_tasks(s)
    for name, blank in DEAD_BLANK.items():
        v = np.array(s[name], copy=True)
        v[dead] = blank
        out[name] = v
    out.pop("k_has_slot", None)
    return out


def digest(s: dict, skip=()) -> int:
    """64-bit digest of a canonical snapshot (exact bits of every field)."""
    s = canonicalize(s)
    h = hashlib.blake2b(digest_size=8)
    for name in ALL_FIELDS:
        if name in skip:
            continue
        v = s[name]
        if isinstance(v, (int, np.integer)):
            h.update(np.int64(v).tobytes())
        elif isinstance(v, float):
            h.update(np.float64(v).tobytes())
        else:
            v = np.ascontiguousarray(v)
            if v.dtype.kind == "f":
                v = v.astype(np.float64)
                v = v + 0.0  # canonicalise -0.0 -> +0.0
            else:
                v = v.astype(np.int64)
            h.update(name.encode())
            h.update(np.asarray(v.shape, np.int64).tobytes())
            h.update(v.tobytes())
    return int.from_bytes(h.digest(), "little")


def diff(a: dict, b: dict, skip=()):
    """Human-readable list of differing fields (for debugging)."""
    a = canonicalize(a)
    b = canonicalize(b)
    out = []
    for name in ALL_FIELDS:
        if name in skip:
            continue
        va, vb = a[name], b[name]
        if isinstance(va, (int, float, np.integer, np.floating)):
            if va != vb:
                out.append(f"{name}: {va!r} != {vb!r}")
            continue
        va = np.asarray(va); vb = np.asarray(vb)
        if va.shape != vb.shape:
            out.append(f"{name}: shape {va.shape} != {vb.shape}")
            continue
        bad = np.argwhere(va != vb)
        if len(bad):
            i = tuple(bad[0])
            out.append(f"{name}: {len(bad)} diffs, first at {i}: {va[i]!r} != {vb[i]!r}")
    return out
