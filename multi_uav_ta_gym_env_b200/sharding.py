"""Environment sharding across ranks (one process per GPU) and the path's only collective.

Environments never interact (each owns its MT19937 streams), so rank r of W simply owns the contiguous
index block [r*E, (r+1)*E) and no data moves during a rollout.  At the end of an episode batch every rank
contributes one small float64 vector of metric sums; torch.distributed all-reduces it (NCCL over
NVLink / NVSwitch on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import torch

METRIC_VECTOR = ["episodes", "S_WPS", "S_WPS_sq", "n_on_time", "n_missed_windows", "Kills", "Losses", "total_distance",
                 "n_task_switches", "S_ESC", "S_ESC_sq", "on_time_rate", "n_windowed_tasks", "n_arrivals",
                 "protected_rec_completed", "recon_losses"]


def shard_range(n_envs_per_rank: int, rank: int):
    """Global env indices (== reset seeds) owned by `rank`."""
    return range(rank * n_envs_per_rank, (rank + 1) * n_envs_per_rank)


def metric_vector(metrics: torch.Tensor, names) -> torch.Tensor:
    """[E, M] per-env metrics -> float64 vector of sums in METRIC_VECTOR order."""
    col = lambda n: metrics[:, names.index(n)]
    v = torch.zeros(len(METRIC_VECTOR), dtype=torch.float64, device=metrics.device)
    v[0] = metrics.shape[0]
    for i, n in enumerate(METRIC_VECTOR):
        if n == "episodes":
            continue
        if n.endswith("_sq"):
            v[i] = (col(n[:-3]) ** 2).sum()
        else:
            v[i] = col(n).sum()
    return v


def allreduce_metric_vector(v: torch.Tensor, group=None) -> torch.Tensor:
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
    return v


def summarize(v: torch.Tensor) -> dict:
    n = max(float(v[0]), 1.0)
    out = {"episodes": float(v[0])}
    for i, name in enumerate(METRIC_VECTOR):
        if name == "episodes" or name.endswith("_sq"):
            continue
        out["mean_" + name] = float(v[i]) / n
    for base in ("S_WPS", "S_ESC"):
        m = out["mean_" + base]
        sq = float(v[METRIC_VECTOR.index(base + "_sq")]) / n
        out["sd_" + base] = max(sq - m * m, 0.0) ** 0.5
    return out
