"""Decode an environment record (host copy) into named arrays: the canonical snapshot that the
parity tests diff against the oracle / reference (schema: tests/golden/refsnap.py), and the
source of the object proxies of the single-env facade."""
from __future__ import annotations

import numpy as np

HDR_INT_FIELDS = {
    "t": "T", "n_tasks": "N_TASKS", "n_thr_active": "N_ACTIVE", "n_reallocations": "N_REALLOC",
    "n_task_switches": "N_SWITCH", "n_arrivals": "N_ARRIVALS", "pending_reset": "PENDING_RESET",
    "n_missed": "N_MISSED", "n_on_time": "N_ON_TIME", "n_windowed": "N_WINDOWED",
    "idle_reserve_steps": "IDLE_RESERVE", "burst_toggle": "BURST_TOGGLE", "escort_requests": "ESC_REQUESTS",
    "escort_completed": "ESC_COMPLETED", "escort_failed": "ESC_FAILED", "escort_required_steps": "ESC_REQ_STEPS",
    "escort_covered_steps": "ESC_COV_STEPS", "protection_breaches": "BREACHES",
    "threats_intercepted": "INTERCEPTED", "recon_losses": "RECON_LOSSES", "escort_losses": "ESCORT_LOSSES",
    "mutual_support": "MUTUAL", "protected_rec_completed": "PROT_REC_DONE", "n_reached": "N_REACHED",
    "conclusion_time": "CONCLUSION",
}


class RecordCodec:
    def __init__(self, lib, cfg):
        self.lib = lib
        self.cfg = cfg
        self.F = lib.fields(cfg)
        self.record_bytes = lib.record_bytes(cfg)
        self.H = {k: lib.header_index(v) for k, v in HDR_INT_FIELDS.items()}
        self.extra = {n: lib.header_index(n) for n in (
            "F_REWARD", "TOTAL_DIST", "NORM_FACTOR", "LAST_REWARD", "N_EVENTS", "ERRFLAGS", "DONE", "N_OPEN",
            "CUR_AGENT", "CUR_TGT", "CUR_MISSION", "LAST_PLAN_STEP", "N_REPLANS", "N_CALLS", "EV_TAGMASK", "N_LSAP")}

    def field(self, rec_row: np.ndarray, name: str) -> np.ndarray:
        off, cnt, dt = self.F[name]
        return rec_row[off:off + cnt * dt.itemsize].view(dt)

    def header(self, rec_row, name):
        if name in ("F_REWARD", "TOTAL_DIST", "NORM_FACTOR", "LAST_REWARD"):
            return float(self.field(rec_row, "hf")[self.extra[name]])
        return int(self.field(rec_row, "hi")[self.extra[name]])

    def snapshot(self, rec_row: np.ndarray) -> dict:
        """Canonical snapshot (same keys and conventions as tests/golden/refsnap.snapshot)."""
        cfg = self.cfg
        A, TC, HC, QC = cfg.n_agents, cfg.task_cap, cfg.n_threats, cfg.queue_cap
        IC = max(cfg.id_cap, TC)
        KW = (IC + 31) // 32
        f = lambda n: self.field(rec_row, n)
        hi = f("hi")
        hf = f("hf")
        s = {k: int(hi[i]) for k, i in self.H.items()}
        T = s["n_tasks"]
        s["F_Reward"] = float(hf[self.extra["F_REWARD"]])
        s["total_distance"] = float(hf[self.extra["TOTAL_DIST"]])
        i64 = lambda v: np.asarray(v, dtype=np.int64)
        s["a_pos"] = np.stack([f("a_posx"), f("a_posy")], axis=1).astype(np.float64)
        s["a_nfp"] = np.stack([f("a_nfpx"), f("a_nfpy")], axis=1).astype(np.float64)
        s["a_state"] = i64(f("a_state"))
        s["a_task_start"] = i64(f("a_task_start"))
        s["a_fail_event"] = i64(f("a_fail_event"))
        s["a_type"] = i64(f("a_type"))
        s["a_caps"] = f("a_caps").reshape(6, A).T.copy()
        s["a_ammo"] = i64(f("a_ammo"))
        s["a_nft"] = f("a_nft").astype(np.float64)
        s["a_re_eval"] = i64(f("a_re_eval"))
        s["a_last_task"] = i64(f("a_last_task"))
        s["a_commit_until"] = i64(f("a_commit"))
        qlen = i64(f("a_qlen"))
        s["a_qlen"] = qlen
        q = f("a_queue").reshape(QC, A).T.astype(np.int64)
        qt = f("a_qtime").reshape(QC, A).T
        aq = np.zeros((A, 16), np.int64)
        for a in range(A):
            aq[a, : qlen[a]] = q[a, : qlen[a]]
        s["a_queue"] = aq
        s["a_dist"] = f("a_dist").astype(np.float64)
        s["a_escort"] = i64(f("a_escort"))
        # per-task fields live in slots; tasks that gave their slot back (closed, unreferenced) read as blanks,
        # which is also how the canonical snapshot treats them (tests/golden/refsnap.canonicalize)
        slot = i64(f("k_slot")[:T])
        has = slot >= 0
        sl = np.where(has, slot, 0)

        def per_task(name, blank=0):
            v = f(name)[sl].astype(np.float64 if f(name).dtype.kind == "f" else np.int64)
            v[~has] = blank
            return v

        s["k_pos"] = np.stack([per_task("k_posx"), per_task("k_posy")], axis=1)
        s["k_type"] = i64(f("k_type")[:T])
        s["k_status"] = i64(f("k_status")[:T])
        cur = f("k_cur").reshape(6, TC)[:, sl].T.copy()
        alloc = f("k_alloc").reshape(6, TC)[:, sl].T.copy()
        cur[~has] = 0.0
        alloc[~has] = 0.0
        s["k_cur"] = cur
        s["k_alloc"] = alloc
        s["k_done_ti"] = per_task("k_done_ti")
        s["k_org_ti"] = per_task("k_org_ti")
        s["k_init_time"] = per_task("k_init")
        s["k_done_time"] = per_task("k_dtime")
        s["k_created_at"] = per_task("k_created")
        s["k_deadline"] = per_task("k_deadline")
        s["k_counted"] = per_task("k_counted")
        s["k_final_quality"] = per_task("k_fq").astype(np.float64)
        s["k_kind"] = per_task("k_kind")
        s["k_required_agents"] = per_task("k_req_agents")
        s["k_elig"] = per_task("k_elig")
        s["k_threat"] = per_task("k_threat")
        s["k_prot_agent"] = per_task("k_prot_agent")
        s["k_prot_task"] = per_task("k_prot_task")
        s["k_reveal_t"] = i64(f("k_reveal")[:T])
        det = np.full((T, A), -1.0)
        for a in range(A):
            for q_i in range(qlen[a]):
                tid = int(q[a, q_i])
                if tid > 0 and s["k_status"][tid - 1] != 2:
                    det[tid - 1, a] = qt[a, q_i]
        s["k_det_time"] = det
        tbl = (f("k_tbl_lo")[sl].astype(np.uint64) | (f("k_tbl_hi")[sl].astype(np.uint64) << np.uint64(32)))
        tbl[~has] = 0
        s["k_tbl_mask"] = tbl.view(np.int64)
        s["k_reached"] = per_task("k_reached")
        s["k_det_frozen"] = per_task("k_det_frozen")  # facade only: len(allocationDetails) of closed tasks
        s["k_has_slot"] = has.astype(np.int64)
        s["h_pos"] = np.stack([f("h_posx"), f("h_posy")], axis=1).astype(np.float64).reshape(HC, 2)
        for n in ("h_status", "h_type", "h_group", "h_ammo", "h_target", "h_mission", "h_task", "h_det_task", "h_spawned",
                  "h_intercept"):
            s[n] = i64(f(n))
        s["h_order"] = i64(f("h_order")[: s["n_thr_active"]])
        kn = f("known").reshape(KW, A)
        known = np.zeros((A, T), np.int64)
        for k in range(T):
            known[:, k] = (kn[k >> 5, :] >> np.uint32(k & 31)) & 1
        s["known"] = known
        nev = int(hi[self.extra["N_EVENTS"]])
        ev = f("events")[:nev].astype(np.int64)
        s["events"] = np.stack([ev & 0xFF, (ev >> 8) - 1], axis=1).reshape(-1, 2)
        return s

    def open_task_ids(self, rec_row):
        """env.last_tasks_info as task ids."""
        IC = max(self.cfg.id_cap, self.cfg.task_cap)
        om = self.field(rec_row, "open_mask")
        return [k + 1 for k in range(IC) if (int(om[k >> 5]) >> (k & 31)) & 1]


def decode_events(n_events: int, events_row) -> list:
    from .config import EVENT_TAGS

    out = []
    for i in range(int(n_events)):
        ev = int(events_row[i])
        out.append([EVENT_TAGS[ev & 0xFF], (ev >> 8) - 1])
    return out
