"""Host-side scenario generation (MultiUAVEnv.reset, mUAV_TA/DroneEnv.py:522-762) and packing
of the per-environment records that the CUDA step kernel consumes.

Reset is not on the hot path (1 of 151 calls per episode): it runs on the host and drives
CPython's `random.Random` exactly as the reference does (four MT19937 streams derived from the
seed, DroneEnv.py:535-538), then freezes the three in-episode streams into raw 32-bit word
tapes that the device consumes with cursors.
"""
from __future__ import annotations

import random
import sys

import numpy as np

from .config import BASE, CONTACT_LINE, FAIL_TABLE, GAME_AREA, TASK_TYPES, UAV_TYPES, CAP_TABLE

MAX_INT = sys.maxsize
AREA_W, AREA_H = GAME_AREA


def _tape(gen: random.Random, n_words: int) -> np.ndarray:
    """Next n_words raw MT19937 outputs of `gen` (not advanced).  getrandbits(32*n) assembles the
    words little-endian in generation order (Modules/_randommodule.c)."""
    clone = random.Random()
    clone.setstate(gen.getstate())
    big = clone.getrandbits(32 * n_words)
    return np.frombuffer(big.to_bytes(4 * n_words, "little"), dtype="<u4").copy()


def _norm2(x, y):
    v = np.array([x, y], dtype=np.float64)
    return float(np.sqrt(v.dot(v)))


def _random_position(gen, obstacles=None, min_distance=20, own_range=3, contact_line=False, mission_area=None):
    """random_position (DroneEnv.py:1371-1410)."""
    limit_line = CONTACT_LINE if contact_line else 0
    for _ in range(100):
        if mission_area is not None:
            tlx, tly, w, h = mission_area
            x = gen.uniform(tlx, tlx + w)
            y = gen.uniform(tly, tly + h)
        else:
            x = gen.uniform(own_range + min_distance, AREA_W - own_range - min_distance)
            y = gen.uniform(own_range + min_distance,
                            AREA_H - own_range - min_distance - ((AREA_H - limit_line) if limit_line != 0 else 0))
        if obstacles is None:
            return x, y
        ok = True
        for ox, oy, osz in obstacles:
            if _norm2(x - ox, y - oy) - own_range < osz + min_distance:
                ok = False
                break
        if ok:
            return x, y
    raise ValueError("Error to build a valid scenario: no space left for obstacles")


class Scenario:
    """One environment's initial conditions."""

    __slots__ = ("seed", "agent_names", "agent_type", "agent_pos", "fail_event", "mission_areas", "task_pos",
                 "task_type", "task_org", "det_task_of_group", "threat_pos", "threat_type", "threat_group",
                 "obstacles", "tapes", "reward_norm_factor")


def generate_scenario(opts, seed: int, tape_words) -> Scenario:
    s = Scenario()
    s.seed = seed
    gA = random.Random(seed)
    gO = random.Random(gA.randint(0, MAX_INT))
    gT = random.Random(gA.randint(0, MAX_INT))
    gM = random.Random(gA.randint(0, MAX_INT))
    agents = dict(opts.agents)
    tasks = dict(opts.tasks)
    threats = [tuple(x) for x in (opts.threats_list or [])]
    A = sum(agents.values())
    mt = opts.max_time_steps

    s.obstacles = []
    for _ in range(int(getattr(opts, "num_obstacles", 0) or 0)):
        size = gO.randint(30, 100)
        p = _random_position(gO, obstacles=s.obstacles, own_range=size, contact_line=True)
        s.obstacles.append((p[0], p[1], float(size)))

    ids = list(range(A))
    gA.shuffle(ids)
    s.agent_names = [None] * A
    s.agent_type = [0] * A
    s.agent_pos = [None] * A
    for tname, n in agents.items():
        for i in range(n):
            aid = ids.pop(0)
            pos = _random_position(gA, obstacles=s.obstacles) if opts.random_init_pos else BASE
            s.agent_names[aid] = f"{tname[0:2]}_agent{i}"
            s.agent_type[aid] = UAV_TYPES.index(tname)
            s.agent_pos[aid] = (float(pos[0]), float(pos[1]))
    s.fail_event = [-1] * A
    for aid in range(A):
        if gA.random() < opts.fail_rate * FAIL_TABLE[UAV_TYPES[s.agent_type[aid]]]:
            s.fail_event[aid] = gA.randint(1, 1000 if mt == -1 else mt)

    s.mission_areas = []
    for _ in range(3):
        w = AREA_W * gM.randint(10, 20) / 100
        h = AREA_H * gM.randint(10, 20) / 100
        c = _random_position(gM, min_distance=max(w, h))
        s.mission_areas.append((c[0] - w / 2, c[1] - w / 2, w, w))  # SquareArea(center, w, w), DroneEnv.py:629-633

    s.task_pos, s.task_type, s.task_org = [], [], []
    hold_n = 0
    for ttype, n in tasks.items():
        for _ in range(n):
            mission = gM.choice(s.mission_areas)
            if ttype != "Hold":
                pos = _random_position(gT, obstacles=s.obstacles, contact_line=True, mission_area=mission)
            else:
                pos = (int((hold_n + 1) * AREA_W / 5), int(AREA_H / 4))
                hold_n += 1
            s.task_pos.append((float(pos[0]), float(pos[1])))
            s.task_type.append(TASK_TYPES.index(ttype))
            s.task_org.append(1.0)
    poss = 0
    for v in s.task_org:
        poss += v
    s.reward_norm_factor = (poss * 1 + poss) / 1000

    s.det_task_of_group, s.threat_pos, s.threat_type, s.threat_group = [], [], [], []
    wide = AREA_W / 10
    for ng, (gtype, count) in enumerate(threats):
        gx = gA.randint(int(0 + wide), int(AREA_W - wide))
        s.task_pos.append((float(gx), AREA_H / 5))
        s.task_type.append(TASK_TYPES.index("Det"))
        s.task_org.append(float(count))
        s.det_task_of_group.append(len(s.task_pos))  # task id
        for _ in range(count):
            sx = gA.randint(int(gx - wide), int(gx + wide))
            s.threat_pos.append((float(sx), 0.0))
            s.threat_type.append(UAV_TYPES.index(gtype))
            s.threat_group.append(ng)
    s.tapes = np.concatenate([_tape(gA, tape_words[0]), _tape(gT, tape_words[1]), _tape(gM, tape_words[2])])
    return s


def pack_records(lib, cfg, scenarios) -> tuple[np.ndarray, np.ndarray]:
    """Scenarios -> (records uint8 [E, record_bytes], tapes uint32 [E, tape_stride])."""
    E = len(scenarios)
    rb = lib.record_bytes(cfg)
    F = lib.fields(cfg)
    rec = np.zeros((E, rb), dtype=np.uint8)
    A, TC, HC, QC = cfg.n_agents, cfg.task_cap, cfg.n_threats, cfg.queue_cap
    IC = max(cfg.id_cap, TC)
    KW = (IC + 31) // 32

    def view(name):
        off, cnt, dt = F[name]
        return rec[:, off:off + cnt * dt.itemsize].view(dt)

    hi = view("hi")
    hf = view("hf")
    H = lib.header_index
    n_tasks0 = len(scenarios[0].task_pos)
    if n_tasks0 > TC:
        raise ValueError("task_cap too small")
    hi[:, H("N_TASKS")] = n_tasks0
    hi[:, H("N_SLOTS_USED")] = n_tasks0
    # initial tasks occupy slots 0..n0-1 (identity id -> slot map); every other id has no slot yet
    view("k_slot")[:] = -1
    view("k_slot")[:, :n_tasks0] = np.arange(n_tasks0)
    view("s_used")[:, :n_tasks0] = np.arange(1, n_tasks0 + 1)
    hi[:, H("CONCLUSION")] = cfg.max_time_steps + 1
    hi[:, H("LAST_PLAN_STEP")] = -10**9
    hi[:, H("N_OPEN")] = n_tasks0
    for g in range(8):
        hi[:, H(f"GROUP_NEXT{g}")] = cfg.group_start[g]
    hf[:, H("NORM_FACTOR")] = [s.reward_norm_factor for s in scenarios]
    m = np.array([[v for area in s.mission_areas for v in area] for s in scenarios], dtype=np.float64)
    hf[:, H("M0_X"):H("M0_X") + 12] = m

    apos = np.array([s.agent_pos for s in scenarios], dtype=np.float64)  # [E, A, 2]
    view("a_posx")[:] = apos[:, :, 0]
    view("a_posy")[:] = apos[:, :, 1]
    view("a_nfpx")[:] = apos[:, :, 0]
    view("a_nfpy")[:] = apos[:, :, 1]
    atype = np.array([s.agent_type for s in scenarios], dtype=np.int64)  # [E, A]
    view("a_type")[:] = atype
    captab = np.array([CAP_TABLE[u] for u in UAV_TYPES], dtype=np.float64)  # [7, 6]
    caps = captab[atype]  # [E, A, 6]
    view("a_caps")[:] = caps.transpose(0, 2, 1).reshape(E, 6 * A)
    ammo = np.where((atype == UAV_TYPES.index("F1")) | (atype == UAV_TYPES.index("F2")), 10, 0)
    view("a_ammo")[:] = ammo
    ranks = np.zeros((E, A), dtype=np.int64)
    for e, sc in enumerate(scenarios):
        order = sorted(range(A), key=lambda a: sc.agent_names[a])
        for r, a in enumerate(order):
            ranks[e, a] = r
    view("a_name_rank")[:] = ranks  # string order of the agent names (tie-break of UrgencyCommit's lock ranking)
    view("a_task_start")[:] = -1
    view("a_fail_event")[:] = np.array([s.fail_event for s in scenarios], dtype=np.int64)
    view("a_last_task")[:] = -1

    tpos = np.array([s.task_pos for s in scenarios], dtype=np.float64)  # [E, n0, 2]
    ttype = np.array([s.task_type for s in scenarios], dtype=np.int64)
    torg = np.array([s.task_org for s in scenarios], dtype=np.float64)
    view("k_posx")[:, :n_tasks0] = tpos[:, :, 0]
    view("k_posy")[:, :n_tasks0] = tpos[:, :, 1]
    view("k_type")[:, :n_tasks0] = ttype
    view("k_org_ti")[:, :n_tasks0] = torg
    kcur = view("k_cur").reshape(E, 6, TC)
    ee, kk = np.meshgrid(np.arange(E), np.arange(n_tasks0), indexing="ij")
    kcur[ee, ttype, kk] = torg
    view("k_cur_ti")[:, :n_tasks0] = torg   # hot copy of component [task type] of the requirement vector (muav_layout.h)
    view("k_init")[:] = -1.0
    view("k_dtime")[:] = -1.0
    view("k_deadline")[:] = -1
    view("k_reveal")[:] = -1
    view("k_fq")[:] = -1
    view("k_threat")[:] = -1
    view("k_prot_agent")[:] = -1

    if HC > 0:
        hpos = np.array([s.threat_pos for s in scenarios], dtype=np.float64)
        view("h_posx")[:] = hpos[:, :, 0]
        view("h_posy")[:] = hpos[:, :, 1]
        view("h_type")[:] = np.array([s.threat_type for s in scenarios], dtype=np.int64)
        hgroup = np.array([s.threat_group for s in scenarios], dtype=np.int64)
        view("h_group")[:] = hgroup
        det = np.array([s.det_task_of_group for s in scenarios], dtype=np.int64)  # [E, G]
        view("h_det_task")[:] = np.take_along_axis(det, hgroup, axis=1)
        view("h_status")[:] = 1
        view("h_ammo")[:] = 4
        view("h_target")[:] = -1
        view("h_mission")[:] = -1
        view("h_intercept")[:] = -1
    if cfg.n_obstacles > 0:
        view("obst")[:] = np.array([s.obstacles for s in scenarios], dtype=np.float64).reshape(E, -1)

    # static / initial tasks are known to everyone (DroneEnv.py:757-758); all of them are open
    mask_words = np.zeros(KW, dtype=np.uint32)
    for k in range(n_tasks0):
        mask_words[k >> 5] |= np.uint32(1 << (k & 31))
    view("known").reshape(E, KW, A)[:] = mask_words[None, :, None]
    view("open_mask")[:] = mask_words[None, :]
    tapes = np.stack([s.tapes for s in scenarios]).astype(np.uint32)
    return rec, tapes
