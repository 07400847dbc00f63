"""Batched evaluation driver with the row schemas of the reference's experiment scripts:

    per-episode scores     run_wps_episode's return dict          experiments/wps_eval.py:276-290
    summary row            main()                                   experiments/wps_eval.py:549-592
    episodes CSV           --episodes-out                           experiments/wps_eval.py:626-647

One environment per evaluation seed (seed = episode index, as in the reference loop), all of them stepped together:
the allocators that need no network (Global / Local / Coalition Hungarian, Urgency-Pair / -Commit / -Coalition) are one
150-step launch; the learned hybrids take a `score_fn` and run step by step on the device.

    python -m multi_uav_ta_gym_env_b200.evaluate --case WPS_hard --algorithms Local-Hungarian Urgency-Pair --episodes 100
"""
from __future__ import annotations

import argparse
import csv
import time
from typing import Callable, Dict, List, Optional

import numpy as np
import torch

from .batched_env import AllocSpec, BatchedMultiUAVEnv
from .config import CASE_SPECS, wps_config

HYBRID_INTERVAL = 15   # experiments/wps_eval.py:64-73
ESCORT_INTERVAL = 12   # experiments/escort_eval.py:52-58, 85-93

DEVICE_ALGORITHMS: Dict[str, Callable[[], AllocSpec]] = {
    "Global-Hungarian": lambda: AllocSpec.global_hungarian(20),            # wps_eval.py:117-122
    "Local-Hungarian": lambda: AllocSpec.local_hungarian(20),              # wps_eval.py:123-133
    "Local-CBBA-Replan": lambda: AllocSpec.cbba_replan(20),                # wps_eval.py:134-146 (reference under PYTHONHASHSEED=0)
    "Local-CBBA-Coalition": lambda: AllocSpec.cbba_replan(ESCORT_INTERVAL),  # escort_eval.py:149-161
    "Local-PI": lambda: AllocSpec.performance_impact(20),                  # wps_eval.py:147-159
    "Local-PI-Coalition": lambda: AllocSpec.performance_impact(ESCORT_INTERVAL),   # escort_eval.py:162-174
    "Urgency-Pair": lambda: AllocSpec.urgency_pair(HYBRID_INTERVAL),       # wps_eval.py:232-236
    "Urgency-Commit": lambda: AllocSpec.urgency_commit(HYBRID_INTERVAL),   # wps_eval.py:207-213
    "Coalition-Hungarian": lambda: AllocSpec.coalition_hungarian(ESCORT_INTERVAL),   # escort_eval.py:137-148
    "Global-Coalition": lambda: AllocSpec(1, ESCORT_INTERVAL, 0x1F, False, False),   # escort_eval.py:124-136
    "Urgency-Coalition": lambda: AllocSpec.urgency_coalition(ESCORT_INTERVAL),       # escort_eval.py:175-179
}
# learned hybrids with a fused forward kernel and in-step token emission: algorithm -> (tokens, planner spec)
LEARNED_ALGORITHMS = {
    "Att-Pair": ("pair", lambda: AllocSpec.pair_hybrid(HYBRID_INTERVAL)),               # wps_eval.py:239-255
    "Att-ContextPair": ("context", lambda: AllocSpec.pair_hybrid(HYBRID_INTERVAL)),     # wps_eval.py:256-272
    "Att-Commit": ("commit", lambda: AllocSpec.att_commit(HYBRID_INTERVAL)),            # wps_eval.py:222-238
    "Att-Coalition": ("escort", lambda: AllocSpec.att_escort(ESCORT_INTERVAL)),         # escort_eval.py:180-196
}
SCORE_KEYS = ("F_Reward", "S_WPS", "on_time_rate", "n_missed_windows", "n_on_time", "n_windowed_tasks",
              "reserve_idle_fraction", "makespan", "total_distance", "n_task_switches")


def _run_learned(env, algorithm: str, net, n_steps: int):
    """A learned hybrid end to end on the device: the step kernel emits the tokens of the environments that replan next,
    the fused forward kernel scores exactly those (need flags, no host synchronisation), the planner half runs inside the
    next step's kernel."""
    from . import scorers

    kind, make_spec = LEARNED_ALGORITHMS[algorithm]
    spec = make_spec()
    E, dev = env.n_envs, env.device
    if kind in ("pair", "context"):
        tok = env.enable_fused_tokens(32, 16, HYBRID_INTERVAL, 0b111, context=kind == "context")
        fused = scorers.FusedAttPairScorer(net, dev)
        scores = torch.zeros(E, 16, 32, dtype=torch.float32, device=dev)
        env.refresh_fused_tokens()
        for _ in range(n_steps):
            fused.score(tok, scores, use_need=True)
            env.step_allocated(spec, 1, edge_scores=scores)
    elif kind == "commit":
        tok = env.enable_fused_tokens(32, 16, HYBRID_INTERVAL, 0b111, commit=True)
        fused = scorers.FusedAttCommitScorer(net, dev)
        pri = torch.zeros(E, 32, dtype=torch.float32, device=dev)
        com = torch.zeros(E, 16, dtype=torch.float32, device=dev)
        env.refresh_fused_tokens()
        for _ in range(n_steps):
            fused.vectors(tok, pri, com, use_need=True)
            env.step_allocated(spec, 1, plan_pri=pri, plan_commit=com)
    else:
        tok = env.enable_fused_tokens(48, 16, ESCORT_INTERVAL, 0x1F, escort=True)
        fused = scorers.FusedAttCoalitionScorer(net, dev)
        scores = torch.zeros(E, 16, 48, dtype=torch.float32, device=dev)
        env.refresh_fused_tokens()
        for _ in range(n_steps):
            fused.score(tok, scores, use_need=True)
            env.step_allocated(spec, 1, edge_scores=scores, task_order=tok["task_order"])


def run_episodes(case: str, algorithm: str, episodes: int, device="cuda:0", score_fn: Optional[Callable] = None,
                 tokens: str = "pair", net=None) -> List[dict]:
    """Scores of seeds 0..episodes-1, one dict per seed with run_wps_episode's keys (decision_ms_mean is the device time
    of the whole batch per step divided by the batch size).  Learned hybrids (LEARNED_ALGORITHMS): pass the network as
    `net` (AttPairNet / AttContextPairNet / AttCommitNet / AttCoalitionNet with the default shapes) to run them on the
    fused forward kernels, or any `score_fn(tokens) -> scores` for the pair hybrids in PyTorch."""
    cfg = wps_config(case)
    env = BatchedMultiUAVEnv(cfg, episodes, device=device).reset(range(episodes))
    n_steps = int(cfg.max_time_steps)
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    if algorithm in DEVICE_ALGORITHMS:
        env.step_allocated(DEVICE_ALGORITHMS[algorithm](), n_steps)
    elif algorithm in LEARNED_ALGORITHMS and net is not None:
        _run_learned(env, algorithm, net, n_steps)
    else:
        if score_fn is None:
            raise ValueError(f"{algorithm}: pass score_fn(tokens) -> scores (a learned pair scorer)")
        spec = AllocSpec.pair_hybrid(HYBRID_INTERVAL)
        for _ in range(n_steps):
            tok = env.tokens_context(32, 16) if tokens == "context" else env.tokens_pair(32, 16)
            env.step_allocated(spec, 1, edge_scores=score_fn(tok))
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / n_steps / episodes
    names = env.lib.metric_names()
    m = env.metrics().cpu().numpy()
    replans = env.header_int("N_REPLANS").cpu().numpy()
    flags = env.error_flags().cpu().numpy()
    if np.any(flags != 0):
        raise RuntimeError(f"capacity overflow in environments {np.nonzero(flags)[0][:8].tolist()}")
    out = []
    for e in range(episodes):
        row = {k: float(m[e, names.index(k)]) for k in SCORE_KEYS}
        row["decision_ms_mean"] = float(ms)
        row["algo_replans"] = float(replans[e])
        row["max_coord"] = float(env.max_coord)
        row["S_ESC"] = float(m[e, names.index("S_ESC")])  # escort_eval.py's score (not a wps_eval column)
        out.append(row)
    return out


def bootstrap_ci_diff(a, b, n_boot=2000, alpha=0.05):
    """Paired bootstrap CI of mean(a - b) (experiments/wps_eval.py:294-310: RandomState(0), 2000 resamples)."""
    d = np.asarray(a, np.float64) - np.asarray(b, np.float64)
    n = len(d)
    if n == 0:
        return 0.0, 0.0, 0.0
    rng = np.random.RandomState(0)
    means = [float(d[rng.randint(0, n, n)].mean()) for _ in range(n_boot)]
    return float(d.mean()), float(np.percentile(means, 100 * alpha / 2)), float(np.percentile(means, 100 * (1 - alpha / 2)))


def summary_row(exp: str, case: str, algorithm: str, scores: List[dict], seconds: float,
                local_scores: Optional[List[dict]] = None) -> dict:
    col = lambda k: [s[k] for s in scores]
    row = {
        "exp": exp, "case": case, "label": CASE_SPECS[case].get("label", case), "algorithm": algorithm,
        "episodes": len(scores),
        "mean_S_WPS": float(np.mean(col("S_WPS"))), "std_S_WPS": float(np.std(col("S_WPS"))),
        "mean_on_time_rate": float(np.mean(col("on_time_rate"))), "std_on_time_rate": float(np.std(col("on_time_rate"))),
        "mean_missed_windows": float(np.mean(col("n_missed_windows"))), "mean_on_time": float(np.mean(col("n_on_time"))),
        "mean_F_Reward": float(np.mean(col("F_Reward"))), "std_F_Reward": float(np.std(col("F_Reward"))),
        "mean_total_distance": float(np.mean(col("total_distance"))), "mean_makespan": float(np.mean(col("makespan"))),
        "mean_reserve_idle": float(np.mean(col("reserve_idle_fraction"))),
        "mean_decision_ms": float(np.mean(col("decision_ms_mean"))), "mean_algo_replans": float(np.mean(col("algo_replans"))),
        "seconds": round(seconds, 2),
    }
    d_s = d_ot = (0.0, 0.0, 0.0)
    if local_scores is not None and algorithm != "Local-Hungarian":
        d_s = bootstrap_ci_diff(col("S_WPS"), [s["S_WPS"] for s in local_scores])
        d_ot = bootstrap_ci_diff(col("on_time_rate"), [s["on_time_rate"] for s in local_scores])
    row.update({"delta_S_WPS_vs_LocalH": d_s[0], "delta_S_WPS_ci_lo": d_s[1], "delta_S_WPS_ci_hi": d_s[2],
                "delta_on_time_vs_LocalH": d_ot[0], "delta_on_time_ci_lo": d_ot[1], "delta_on_time_ci_hi": d_ot[2]})
    return row


def episode_rows(exp: str, case: str, algorithm: str, scores: List[dict]) -> List[dict]:
    return [{"exp": exp, "case": case, "algorithm": algorithm, "seed": seed, "S_WPS": s["S_WPS"], "n_on_time": s["n_on_time"],
             "n_missed_windows": s["n_missed_windows"], "total_distance": s["total_distance"],
             "max_coord": s.get("max_coord", 1000.0), "on_time_rate": s["on_time_rate"],
             "reserve_idle_fraction": s.get("reserve_idle_fraction", 0.0)} for seed, s in enumerate(scores)]


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="WPS_hard")
    ap.add_argument("--algorithms", nargs="+", default=["Local-Hungarian", "Global-Hungarian", "Urgency-Pair"])
    ap.add_argument("--episodes", type=int, default=100)
    ap.add_argument("--exp", default="WPS")
    ap.add_argument("--out", default="wps_eval.csv")
    ap.add_argument("--episodes-out", default=None)
    ap.add_argument("--device", default="cuda:0")
    args = ap.parse_args(argv)
    per, rows = {}, []
    for algo in args.algorithms:
        t = time.time()
        per[algo] = run_episodes(args.case, algo, args.episodes, args.device)
        rows.append((algo, time.time() - t))
    table = [summary_row(args.exp, args.case, a, per[a], sec, per.get("Local-Hungarian")) for a, sec in rows]
    with open(args.out, "w", newline="", encoding="utf-8") as f:
        w = csv.DictWriter(f, fieldnames=list(table[0].keys()))
        w.writeheader()
        w.writerows(table)
    if args.episodes_out:
        ep = [r for a in per for r in episode_rows(args.exp, args.case, a, per[a])]
        with open(args.episodes_out, "w", newline="", encoding="utf-8") as f:
            w = csv.DictWriter(f, fieldnames=list(ep[0].keys()))
            w.writeheader()
            w.writerows(ep)
    for r in table:
        print(f"[{args.exp}] {r['case']} {r['algorithm']}: S_WPS={r['mean_S_WPS']:.1f}+/-{r['std_S_WPS']:.1f} "
              f"on_time={r['mean_on_time_rate']:.2f} miss={r['mean_missed_windows']:.1f} F={r['mean_F_Reward']:.1f} "
              f"({r['seconds']:.1f}s)")


if __name__ == "__main__":
    main()
