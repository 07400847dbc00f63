#!/usr/bin/env python
"""Benchmark of the batched WPS step path (BASELINE.json metric: WPS_hard env-steps/sec).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores

A "step" is one pass of the hot path over the whole batch: pair tokens -> Att-Pair scorer (torch,
random init, only for environments whose replan rule fires) -> fused Local-Hungarian allocate +
env step kernel, for 4096 WPS_hard environments per GPU (BASELINE config 2).  Environments are
independent, so N GPUs run N shards (weak scaling) and meet only in one NCCL all-reduce of the
episode metric vector every 150 steps.

Timed numbers:
  value     env-steps/s, state resident in HBM, every step bracketed by CUDA events on the launching
            stream, L2 flushed (256 MB write) between timed steps because the 46 MB state would
            otherwise stay resident in the 126 MB L2; max over ranks.
  e2e       same metric through the host-buffer path: allocator decisions are copied to pinned host
            memory, fed back through muav_step_host (H2D actions, D2H reward/terminated/truncated).
  roofline  algorithmic bytes of the step kernel / its measured launch time, against the measured
            HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the oracle port (oracle/, the CPU restatement of the reference's Python path) on all
            host cores, one env process per core, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CASE = "WPS_hard"
ENVS_PER_GPU = 4096
EPISODE = 150
HYBRID_INTERVAL = 15
METRIC = "WPS_hard env-steps/sec (batched, 1/2/4/8 B200) vs ref CPU; % HBM roofline"
# roofline.traffic comes from profiles/step_kernel_traffic.json, written by tools/ncu_summary.py from an `ncu --set full`
# capture together with a hash of the kernel sources it was taken on; a stale capture (sources changed since) reads as null
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
# the unmodified reference (DroneEnv + HungarianAllocator incl. observation building) measured in
# the survey container, env-steps/s per core on WPS_hard; the oracle port used here skips the observation dicts
REFERENCE_PER_CORE_SURVEY = "920-1560 env-steps/s/core (BASELINE.md section 2, unmodified reference incl. observations)"


def kernel_source_hash():
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "multi_uav_ta_gym_env_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.startswith("muav_scorer"):
            continue
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def ncu_traffic(workload, envs):
    """(bytes per launch | None, note) for the step kernel of this workload from the committed ncu summary."""
    try:
        rec = json.load(open(TRAFFIC_FILE))
    except Exception:
        return None, "no ncu capture committed (profiles/step_kernel_traffic.json)"
    ent = rec.get(f"{workload}:{envs}")
    if not ent:
        return None, f"no ncu capture for {workload} with {envs} environments"
    if ent.get("source_hash") != kernel_source_hash():
        return None, f"ncu capture {ent.get('report')} is from other kernel sources ({ent.get('source_hash')}): stale"
    return float(ent["dram_bytes_per_launch"]), (f"dram__bytes_read.sum + dram__bytes_write.sum per launch, {ent.get('report')}"
                                                  f" ({ent.get('note', '')})")


# ----------------------------------------------------------------------------- CPU arm (oracle port)
def _cpu_worker(args):
    """One env process: the reference's Att-Pair episode loop (wps_eval.py:226-230, PairCostHybrid.plan)
    restated on the oracle: tokens -> AttPairNet (CPU torch, 1 thread) -> Hungarian -> env.step."""
    wid, min_steps, min_seconds, warmup = args
    import torch

    torch.set_num_threads(1)
    from multi_uav_ta_gym_env_b200.config import wps_config
    from multi_uav_ta_gym_env_b200.scorers import AttPairNet
    from oracle.hungarian import OracleHungarian, apply_assign
    from oracle.sim import OracleEnv
    from oracle import tokens as otok

    torch.manual_seed(0)
    net = AttPairNet().eval()
    cfg = wps_config(CASE)

    def episode(seed, budget):
        env = OracleEnv(cfg).reset(seed)
        hung = OracleHungarian(20, env.max_coord)
        n = 0
        while n < budget and env.t < EPISODE:
            pairs = []
            if otok.hybrid_should_replan(env, env.last_events, HYBRID_INTERVAL):
                tok = otok.build_pair_tokens(env, 32, 16)
                if tok["agent_mask"].all() or tok["task_mask"].all():
                    # nobody left to assign or nothing left to do: the allocator returns [] whatever the scores are
                    # (and torch's padded-encoder fast path rejects a batch whose tokens are all padding)
                    pairs = []
                else:
                    with torch.no_grad():
                        logits, _ = net(torch.from_numpy(tok["task_feats"])[None], torch.from_numpy(tok["task_mask"])[None],
                                        torch.from_numpy(tok["agent_feats"])[None],
                                        torch.from_numpy(tok["agent_mask"])[None])
                    scores = (torch.tanh(logits[0]) * 0.35).numpy() * tok["edge_valid"]
                    pairs = otok.pair_plan(env, hung, scores)
            env.step(apply_assign(env, pairs))
            n += 1
        return n

    seed = 10_000 * (wid + 1)
    done = 0
    while done < warmup:
        done += episode(seed, warmup - done)
        seed += 1
    t0 = time.perf_counter()
    steps = 0
    while steps < min_steps or (time.perf_counter() - t0) < min_seconds:
        steps += episode(seed, 10**9)
        seed += 1
    return steps, time.perf_counter() - t0


def run_cpu_arm(min_steps, warmup, min_seconds=12.0, procs=None):
    import multiprocessing as mp

    procs = procs or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        res = pool.map(_cpu_worker, [(w, min_steps, min_seconds, warmup) for w in range(procs)])
    total = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    return {"value": total / wall, "unit": "env-steps/s", "cores": procs, "kind": "port",
            "sample": f"{procs} oracle env processes x >= {min_seconds:.0f} s of WPS_hard Att-Pair episodes "
                      f"({total} env-steps, torch CPU scorer 1 thread/process)",
            "env_steps": total, "seconds": wall}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.samples = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", os.environ.get("MUAV_SMI_PERIOD_MS", "100")], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def mark(self):
        return len(self.samples)

    def wait_for(self, n, keep_busy, timeout_s=3.0):
        """Keep the GPU under the same load (keep_busy() enqueues a few more untimed steps) until n samples exist."""
        t0 = time.perf_counter()
        while self.proc is not None and len(self.samples) < n and time.perf_counter() - t0 < timeout_s:
            keep_busy()

    def stop(self, first=0):
        """Summary of the samples taken from index `first` on (the timed region); when the region was too short for
        three samples, every sample since start() is used (the warm-up runs the same load)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        use = self.samples[first:] if len(self.samples) - first >= 3 else self.samples
        for s in use:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- GPU arm
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from multi_uav_ta_gym_env_b200 import AllocSpec, BatchedMultiUAVEnv, sharding, wps_config
    from multi_uav_ta_gym_env_b200.scorers import AttContextPairNet, AttPairNet, FusedAttPairScorer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = run_cpu_arm(EPISODE, 30, min_seconds=args.cpu_seconds)  # before CUDA is initialised (fork)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sampler = ClockSampler(local)   # started before any GPU work so that it is warm when the timed region begins
    sampler.start()
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # workload: the default is BASELINE config 2; the others are BASELINE configs 3-5 (extra lines for profiles/)
    wl = args.workload
    use_scorer = wl in ("hard_pair", "attn_context")
    use_commit_net = wl == "commit_att"
    use_coal_net = wl == "escort_att"
    if wl == "attn_context":
        # the paper's primary method: Att-ContextPair on WPS_attn (12 agents, dual-front bursts, private knowledge)
        case_name, cfg, spec = "WPS_attn", wps_config("WPS_attn"), AllocSpec.pair_hybrid(HYBRID_INTERVAL)
        desc = "Local-Hungarian + random-init Att-ContextPair edge scores, hybrid replan rule t%15/events"
    elif wl == "hard_pair":
        case_name, cfg, spec = CASE, wps_config(CASE), AllocSpec.pair_hybrid(HYBRID_INTERVAL)
        desc = "Local-Hungarian + random-init Att-Pair edge scores, hybrid replan rule t%15/events"
    elif wl == "hard_local":
        case_name, cfg, spec = CASE, wps_config(CASE), AllocSpec.local_hungarian(20)
        desc = "Local-Hungarian interval 20 (no scorer)"
    elif wl == "hard_pi":
        case_name, cfg, spec = CASE, wps_config(CASE), AllocSpec.performance_impact(20)
        desc = "Local-PI market allocator (PerformanceImpact, max_tasks_per_agent=1) interval 20 (no scorer)"
    elif wl == "hard_cbba":
        case_name, cfg, spec = CASE, wps_config(CASE), AllocSpec.cbba_replan(20)
        desc = "Local-CBBA-Replan market allocator (CBBAReplan, max_tasks_per_agent=1) interval 20 (no scorer)"
    elif wl == "escort_cbba":
        case_name, cfg, spec = "WPS_escort", wps_config("WPS_escort"), AllocSpec.cbba_replan(12)
        desc = "Local-CBBA-Coalition market allocator interval 12 with visibility map"
    elif wl == "escort_pi":
        case_name, cfg, spec = "WPS_escort", wps_config("WPS_escort"), AllocSpec.performance_impact(12)
        desc = "Local-PI-Coalition market allocator interval 12 with visibility map"
    elif wl == "commit_urgency":
        case_name, cfg, spec = "WPS_commit", wps_config("WPS_commit"), AllocSpec.urgency_commit(HYBRID_INTERVAL)
        desc = "UrgencyCommit planner on the device (commit locks, rematch penalty), hybrid replan rule"
    elif wl == "commit_att":
        case_name, cfg, spec = "WPS_commit", wps_config("WPS_commit"), AllocSpec.att_commit(HYBRID_INTERVAL)
        desc = ("Att-Commit: random-init AttCommitNet (fused forward kernel on fused commit tokens) -> "
                "AttentionCommit._plan_from_scores on the device, hybrid replan rule")
    elif wl == "escort_att":
        case_name, cfg, spec = "WPS_escort", wps_config("WPS_escort"), AllocSpec.att_escort(12)
        desc = ("Att-Coalition: random-init AttCoalitionNet (fused forward kernel on fused escort tokens) -> "
                "AttentionEscort._plan_from_scores on the device, replan rule t%12/any event")
    elif wl == "escort_coalition":
        case_name, cfg, spec = "WPS_escort", wps_config("WPS_escort"), AllocSpec.coalition_hungarian(12)
        desc = "Coalition-Hungarian interval 12 with visibility map"
    elif wl.startswith("burst_x"):
        from multi_uav_ta_gym_env_b200 import burst_scaled_spec
        k = int(wl[7:])
        case_name, cfg, spec = f"WPS_burst x{k}", wps_config(burst_scaled_spec(k)), AllocSpec.local_hungarian(20)
        desc = f"agents/tasks/threats scaled x{k}, Local-Hungarian interval 20"
    else:
        raise SystemExit(f"unknown workload {wl}")
    E = args.envs
    scaling = "weak"
    if args.global_envs > 0:
        # strong scaling: a fixed total sharded by environment index (BASELINE configs 3-5)
        if args.global_envs % world:
            raise SystemExit("--global-envs must be a multiple of the number of GPUs")
        E = args.global_envs // world
        scaling = "strong"
    # Task slots per environment.  The library default is the provable bound (every task that could ever be created: 48
    # for WPS_hard); with slot recycling at most 24 are alive at once over 4096 seeds x 150 steps, so the WPS_hard
    # workloads run with 32 slots (16 instead of 12 resident environments per SM, step kernel 0.112 -> 0.099 ms,
    # profiles/r02_step_kernel.md).  A slot overflow would set the sticky ERR_TASK_OVERFLOW bit: `error_flags` in the
    # output line must be (and is) 0.  --task-cap -1 runs the library default.
    if args.task_cap > 0:
        task_cap = args.task_cap
    elif args.task_cap == 0 and wl in ("hard_pair", "hard_local", "hard_pi", "hard_cbba"):
        task_cap = 32
    else:
        task_cap = None
    seeds = list(sharding.shard_range(E, rank))
    if args.unique_seeds > 0:
        # big sweeps: scenario generation on the host is the slow part (10 ms per 64-agent environment), so the shard
        # cycles through a bounded set of distinct seeds (stated in config.workload)
        seeds = [s % args.unique_seeds for s in seeds]
    env = BatchedMultiUAVEnv(cfg, E, device=dev, task_cap=task_cap).reset(seeds)
    torch.manual_seed(0)
    scores = tok = scorer = None
    if use_scorer:
        ctx = wl == "attn_context"
        net = (AttContextPairNet() if ctx else AttPairNet()).to(dev).eval()
        scores = torch.zeros(E, 16, 32, dtype=torch.float32, device=dev)
        # the step kernel emits the pair tokens of every env that will replan before the next step
        tok = env.enable_fused_tokens(32, 16, HYBRID_INTERVAL, 0b111, context=ctx)
        scorer = FusedAttPairScorer(net, dev)   # hand-written fused forward (csrc/muav_scorer.cu)
    plan_kw = {}
    if use_commit_net:
        from multi_uav_ta_gym_env_b200.scorers import AttCommitNet, FusedAttCommitScorer
        cnet = AttCommitNet().to(dev).eval()
        tok = env.enable_fused_tokens(32, 16, HYBRID_INTERVAL, 0b111, commit=True)
        commit_scorer = FusedAttCommitScorer(cnet, dev)
        plan_kw = {"plan_pri": torch.zeros(E, 32, dtype=torch.float32, device=dev),
                   "plan_commit": torch.zeros(E, 16, dtype=torch.float32, device=dev)}
    if use_coal_net:
        from multi_uav_ta_gym_env_b200.scorers import AttCoalitionNet, FusedAttCoalitionScorer
        tok = env.enable_fused_tokens(48, 16, 12, 0x1F, escort=True)
        coal_scorer = FusedAttCoalitionScorer(AttCoalitionNet().to(dev).eval(), dev)
        scores = torch.zeros(E, 16, 48, dtype=torch.float32, device=dev)
        plan_kw = {"task_order": tok["task_order"]}
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    metric_acc = torch.zeros(32, dtype=torch.float64, device=dev)
    names = env.lib.metric_names()
    launches = {"n": 0}

    def score_step(t):
        """Att-Pair scorer for the environments whose hybrid replan rule fires at time t (wps_eval.py:64-73)."""
        if use_commit_net:
            if t == 0:
                env.refresh_fused_tokens()
                launches["n"] += 1
            commit_scorer.vectors(tok, plan_kw["plan_pri"], plan_kw["plan_commit"], use_need=True)
            launches["n"] += 1
            return
        if use_coal_net:
            if t == 0:
                env.refresh_fused_tokens()
                launches["n"] += 1
            coal_scorer.score(tok, scores, use_need=True)
            launches["n"] += 1
            return
        if not use_scorer:
            return
        if t == 0:
            env.refresh_fused_tokens()  # episode start: standalone token kernel for all envs
            launches["n"] += 1
        # every environment whose replan rule fires (need flag set by the step kernel) is scored; the others are
        # skipped on the device, so there is no host synchronisation in the loop
        scorer.score(tok, scores, use_need=True)
        launches["n"] += 1

    err_acc = torch.zeros(1, dtype=torch.int32, device=dev)

    def episode_end():
        # capacity / tape overflow bits are sticky per episode: collect them before the records are rewound
        err_acc.copy_(torch.maximum(err_acc, env.error_flags().abs().max().to(torch.int32).reshape(1)))
        m = env.metrics()
        launches["n"] += 1
        vec = sharding.metric_vector(m, names)
        sharding.allreduce_metric_vector(vec)  # the path's only collective: end-of-episode metric sums (NCCL/NVLink)
        metric_acc[: vec.numel()].add_(vec)
        env.restore()

    def device_step(t):
        score_step(t)
        env.step_allocated(spec, 1, edge_scores=scores, **plan_kw)
        launches["n"] += 1
        if t + 1 == EPISODE:
            episode_end()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- warm-up (then rewind so that the timed region starts at an episode boundary)
    for w in range(max(args.warmup, 3)):
        device_step(w % EPISODE)
    env.restore()
    metric_acc.zero_()
    barrier()

    # ---- timed: K steps, each bracketed by CUDA events, L2 flushed between steps
    K = args.steps
    first_sample = sampler.mark()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches["n"] = 0
    barrier()
    t_enq = time.perf_counter()
    for k in range(K):
        t = k % EPISODE
        flush_buf.fill_(k & 0xFF)
        ev[k][0].record()
        score_step(t)
        ev[k][1].record()
        env.step_allocated(spec, 1, edge_scores=scores, **plan_kw)
        launches["n"] += 1
        ev[k][2].record()
        if t + 1 == EPISODE:
            episode_end()
    enqueue_ms = (time.perf_counter() - t_enq) * 1e3 / K   # host time to enqueue one step (the GPU must not wait for it)
    barrier()
    wall_ms = (time.perf_counter() - t_enq) * 1e3 / K      # wall clock per step including the L2 flush between steps
    gpu_launches = launches["n"]                           # our kernels launched inside the timed region
    err_flags = max(int(env.error_flags().abs().max().item()), int(err_acc.item()))  # overflow bits: must stay 0
    if sampler.mark() - first_sample < 3:
        # a 20-step timed region lasts a few milliseconds: continue the identical load, untimed, until the sampler has
        # three readings taken under it
        def more():
            for t in range(15):
                flush_buf.fill_(t)
                score_step(1 + t)
                env.step_allocated(spec, 1, edge_scores=scores, **plan_kw)
            torch.cuda.synchronize(dev)
        sampler.wait_for(first_sample + 3, more)
        env.restore()
    clocks = sampler.stop(first_sample)
    clocks["window"] = "timed region + identical untimed continuation until 3 samples"
    step_ms = sum(a.elapsed_time(c) for a, b, c in ev)
    kern_ms = sum(b.elapsed_time(c) for a, b, c in ev)
    t_all = torch.tensor([step_ms, kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    step_ms, kern_ms = t_all.tolist()
    value = world * E * K / (step_ms / 1e3)

    # episode statistics of the line: the episodes that ended inside the timed window only (none in a 20-step run), not the
    # probe calls below, which read the metrics of freshly reset environments
    metric_acc_timed = metric_acc.clone()

    # ---- the episode end (metrics kernel + the path's only collective + rewind), timed on its own: a 20-step run never
    # reaches it, so the collective is measured explicitly
    env.restore()
    vec_probe = sharding.metric_vector(env.metrics(), names)
    for _ in range(3):
        sharding.allreduce_metric_vector(vec_probe)
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    NCOLL = 20
    c0.record()
    for _ in range(NCOLL):
        sharding.allreduce_metric_vector(vec_probe)
    c1.record()
    barrier()
    coll_us = torch.tensor([c0.elapsed_time(c1) * 1e3 / NCOLL], dtype=torch.float64, device=dev)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(5):
        episode_end()
    p1.record()
    barrier()
    ep_end_ms = torch.tensor([p0.elapsed_time(p1) / 5], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(coll_us, op=dist.ReduceOp.MAX)
        dist.all_reduce(ep_end_ms, op=dist.ReduceOp.MAX)
    coll_us, ep_end_ms = coll_us.item(), ep_end_ms.item()

    # ---- end-to-end through the host-buffer entry points (muav_ctx_allocate_host / muav_ctx_step_host): the allocator's
    # decision goes to pinned host memory, comes back as host actions, reward / terminated / truncated land in host memory
    env.restore()
    A = env.n_agents
    h_act = torch.empty(E, A, 2, dtype=torch.int32).pin_memory()
    h_rew = torch.empty(E, dtype=torch.float64).pin_memory()
    h_term = torch.empty(E, dtype=torch.uint8).pin_memory()
    h_trunc = torch.empty(E, dtype=torch.uint8).pin_memory()

    def host_step(t):
        score_step(t)
        env.allocate_host(spec, h_act, edge_scores=scores, **plan_kw)   # kernel + D2H + sync: the caller holds the decision
        env.step_host(h_act, h_rew, h_term, h_trunc, 1, hint=spec)   # H2D + step + packed D2H + sync
        if t + 1 == EPISODE:
            episode_end()

    for w in range(max(args.warmup, 3)):   # warm the host-buffer path too (module load of the allocator-only kernel, ctx)
        host_step(w % EPISODE)
    env.restore()
    Ke = min(K, 2 * EPISODE)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(Ke):
        host_step(k % EPISODE)
    e1.record()
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * E * Ke / (e2e_ms.item() / 1e3)
    err_flags = max(err_flags, int(err_acc.item()), int(env.error_flags().abs().max().item()))
    act_bytes = E * A * 2 * 4
    metric_acc = metric_acc_timed

    if rank == 0:
        rb = env.record_bytes
        out_bytes = 8 + 1 + 1 + 4 + 4 + 4
        b_rec = 2 * rb + out_bytes     # what this build actually moves per env-step: record in + out, step outputs
        b_alg, b_alg_src = survey_b_alg(wl, A, env)
        peak, peak_src = hbm_peak()
        achieved = E * b_alg / (kern_ms / K / 1e3) / 1e9
        traffic, traffic_note = ncu_traffic(wl, E)
        ms_step = step_ms / K
        stats = sharding.summarize(metric_acc[: len(sharding.METRIC_VECTOR)].cpu())
        line = {
            "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms / K, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{case_name} ({A} agents), {E} envs per GPU, {desc}, seeds = env index"
                                   + (f" mod {args.unique_seeds}" if args.unique_seeds > 0 else ""),
                       "envs_per_gpu": E, "global_envs": world * E, "parallelism": f"env-shard x{world}",
                       "l2": f"flushed between timed steps (256 MB write); state {E * rb / 1e6:.0f} MB vs 126 MB L2",
                       "record_bytes": rb, "task_slots": int(env.task_cap), "agent_steps_per_s": value * A,
                       "host_enqueue_ms_per_step": enqueue_ms, "wall_ms_per_step_incl_flush": wall_ms,
                       "episode_end_ms": ep_end_ms,
                       "env_steps_per_s_incl_episode_end": world * E * EPISODE / ((EPISODE * ms_step + ep_end_ms) / 1e3)},
            "collective_us": coll_us,
            "collective": f"end-of-episode all-reduce of the {len(sharding.METRIC_VECTOR)}-double metric vector, "
                          f"{'NCCL over NVLink' if world > 1 else 'single rank (no-op)'}, mean of {NCOLL} after 3 warm-ups",
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": act_bytes,
                    "d2h_bytes_per_step": act_bytes + E * 10, "steps": Ke},
            "gpu_launches": gpu_launches,
            "error_flags": err_flags,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_note": traffic_note,
                         "algorithmic_bytes_per_launch": E * b_alg,
                         "kernel": "muav_step_kernel", "bytes_per_env_step": b_alg, "bytes_per_env_step_source": b_alg_src,
                         "record_io_bytes_per_env_step": b_rec,
                         "record_io_gbs": E * b_rec / (kern_ms / K / 1e3) / 1e9,
                         "kernel_ms_per_launch": kern_ms / K, "peak_source": peak_src,
                         "kernel_share_of_step": kern_ms / step_ms},
            "episode_stats": {k: stats[k] for k in ("episodes", "mean_S_WPS", "sd_S_WPS", "mean_n_on_time",
                                                    "mean_n_missed_windows", "mean_on_time_rate", "mean_Kills")},
        }
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_baseline_record(cpu_base)
        emit_line(line)
    if world > 1:
        dist.destroy_process_group()


def survey_b_alg(wl, A, env):
    """Algorithmic bytes per env-step, SURVEY.md 8(d) (fp64 parity build): the stated figures for the registered
    configs (WPS_hard 14.4 KB, WPS_commit 21 KB, WPS_escort 25 KB), else the section's own formula
    B_alg = 2 S_env + O_env with agent 104 B, live task 100 + 8 A B, threat 32 B, known bitmask 4 A ceil(Tcap / 32) B,
    pending reveals 4 x 48 B, scalars 128 B, RNG tape 44 B per step, O_env = 4 A + 16 B."""
    stated = {"hard_pair": 14400, "hard_local": 14400, "commit_urgency": 21000, "commit_att": 21000, "escort_coalition": 25000, "escort_att": 25000,
              "hard_pi": 14400, "escort_pi": 25000, "hard_cbba": 14400, "escort_cbba": 25000}
    if wl in stated:
        return stated[wl], "SURVEY.md 8(d), stated figure"
    H = int(env.cfg.n_threats)
    t_live = int(env.cfg.max_tasks) + H
    s_env = 104 * A + (100 + 8 * A) * t_live + 32 * H + 4 * A * ((t_live + 31) // 32) + 4 * 48 + 128
    return 2 * s_env + 44 + 4 * A + 16, "SURVEY.md 8(d), formula"


def cpu_baseline_record(res):
    rec = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
    rec["per_core"] = res["value"] / max(res["cores"], 1)
    rec["reference_per_core_survey"] = REFERENCE_PER_CORE_SURVEY
    return rec


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = run_cpu_arm(max(args.steps, 1), max(args.warmup, 3), min_seconds=args.cpu_seconds)
    value = res["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": int(args.gpus),
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 / value if value else None,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{CASE} (8 agents), one env process per host core, Local-Hungarian + random-init "
                               f"Att-Pair edge scores (reference algorithm restated in oracle/; the Python reference "
                               f"itself cannot travel to the GPU box)"},
        "cpu_baseline": cpu_baseline_record(res),
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_line(line)


class _QuietStdout:
    """Route fd 1 to stderr while the run is in progress (NCCL prints its version banner to stdout), so that the
    process emits exactly ONE line on stdout: the JSON result."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.saved, (text + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


_OUT = None


def emit_line(obj):
    text = json.dumps(obj)
    if _OUT is not None:
        _OUT.emit(text)
    else:
        print(text, flush=True)


def main():
    global _OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1500)
    ap.add_argument("--warmup", type=int, default=150)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="environments per GPU")
    ap.add_argument("--workload", default="hard_pair",
                    help="hard_pair (default, BASELINE config 2) | hard_local | commit_urgency | escort_coalition | burst_xK "
                         "| attn_context | hard_pi | escort_pi | hard_cbba | escort_cbba | commit_att | escort_att")
    ap.add_argument("--task-cap", type=int, default=0,
                    help="task slots per environment (0 = workload default: 32 for WPS_hard, else the library's bound; "
                         "-1 = always the library's provable bound)")
    ap.add_argument("--unique-seeds", type=int, default=0,
                    help="cycle through this many distinct scenario seeds (0 = one seed per environment)")
    ap.add_argument("--global-envs", type=int, default=0,
                    help="strong scaling: total environments sharded across the GPUs (overrides --envs)")
    ap.add_argument("--cpu-seconds", type=float, default=60.0,
                    help="timed seconds of the CPU baseline per worker (BASELINE.md section 3: >= 60)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    with _QuietStdout() as q:
        _OUT = q
        if args.impl == "reference":
            run_reference_arm(args)
        else:
            run_gpu_arm(args)
        _OUT = None


if __name__ == "__main__":
    main()
