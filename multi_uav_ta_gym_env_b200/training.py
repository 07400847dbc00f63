"""Training loops of the learned hybrids on the batched environment.

The reference trains one environment at a time and keeps its replay buffers as Python lists of numpy dicts
(`PairCostHybrid.push/update`, `AttentionCommit.push/update`, `AttentionEscort.push/update`).  Here a step of E
environments yields up to E transitions at once, the buffer is a ring of device tensors, and one update consumes a
random mini-batch of it.  The LOSSES are the reference's, term by term:

    pair_il_loss     PairCostHybrid._il_update      TaskAllocation/Hybrid/PairCostHybrid.py:336-369
    pair_rl_loss     PairCostHybrid.update          TaskAllocation/Hybrid/PairCostHybrid.py:410-470
    commit_loss      AttentionCommit.update         TaskAllocation/Hybrid/AttentionCommit.py:198-241
    escort_loss      AttentionEscort.update         TaskAllocation/Hybrid/AttentionEscort.py:550-634

and the episode drivers follow experiments/train_pair_cost.py:96-156, train_att_commit.py:28-75, train_escort.py:28-82
(replan cadence, reward = delta score / 20, exploration noise, target-network refresh, gradient clipping).
tests/test_training_reference.py checks every loss against the reference's own update() on identical batches.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F
from torch import nn

from .collectors import CommitCollector, EscortCollector, ILCollector, RLCollector

TOKEN_KEYS = ("task_feats", "task_mask", "agent_feats", "agent_mask", "edge_valid")


class ReplayBuffer:
    """Ring buffer of transitions stored as stacked device tensors (max_buffer of the reference policies: 50 000)."""

    def __init__(self, capacity: int = 50_000, device="cuda"):
        self.capacity, self.device = int(capacity), torch.device(device)
        self.store: Dict[str, torch.Tensor] = {}
        self.size = 0
        self.head = 0

    def __len__(self):
        return self.size

    def push(self, batch: Dict[str, torch.Tensor], rows: Optional[torch.Tensor] = None):
        """batch: {name: tensor [E, ...]}; rows: bool / uint8 [E] selecting the transitions to keep."""
        if rows is not None:
            idx = rows.bool().nonzero().flatten()
            if idx.numel() == 0:
                return 0
            batch = {k: v.index_select(0, idx) for k, v in batch.items()}
        n = next(iter(batch.values())).shape[0]
        if n > self.capacity:
            batch = {k: v[-self.capacity:] for k, v in batch.items()}
            n = self.capacity
        if not self.store:
            self.store = {k: torch.zeros((self.capacity,) + tuple(v.shape[1:]), dtype=v.dtype, device=self.device)
                          for k, v in batch.items()}
        pos = (self.head + torch.arange(n, device=self.device)) % self.capacity
        for k, v in batch.items():
            self.store[k].index_copy_(0, pos, v.to(self.device))
        self.head = (self.head + n) % self.capacity
        self.size = min(self.capacity, self.size + n)
        return n

    def sample(self, batch_size: int, generator: Optional[torch.Generator] = None) -> Dict[str, torch.Tensor]:
        bs = min(batch_size, self.size)
        idx = torch.randperm(self.size, device=self.device, generator=generator)[:bs]
        return {k: v.index_select(0, idx) for k, v in self.store.items()}


def flatten_transition(tr: dict, extra=()) -> Dict[str, torch.Tensor]:
    """collector transition -> flat {name: tensor [E, ...]} for ReplayBuffer.push (token dicts get a prefix)."""
    out = {}
    for k in TOKEN_KEYS:
        if k in tr["tokens"]:
            out["tok." + k] = tr["tokens"][k]
            out["next." + k] = tr["next_tokens"][k]
    out["reward"] = tr["reward"].to(torch.float32)
    out["done"] = tr["done"].to(torch.float32)
    for k in extra:
        out[k] = tr[k]
    return out


def _tok(batch, prefix):
    return (batch[prefix + "task_feats"], batch[prefix + "task_mask"].bool(), batch[prefix + "agent_feats"],
            batch[prefix + "agent_mask"].bool())


# ---------------------------------------------------------------------------------------------- losses
def pair_il_loss(net: nn.Module, tokens: dict, expert_mask: torch.Tensor) -> torch.Tensor:
    """Batched BCE on visible edges, positives re-weighted per sample (PairCostHybrid._il_update :336-357)."""
    logits, _ = net(tokens["task_feats"], tokens["task_mask"].bool(), tokens["agent_feats"], tokens["agent_mask"].bool())
    edge_valid = tokens["edge_valid"].to(torch.float32)
    target = expert_mask.to(torch.float32)
    logits = logits.clamp(-8.0, 8.0)
    bce = F.binary_cross_entropy_with_logits(logits, target, reduction="none")
    pos = (target * edge_valid).sum(dim=(1, 2)).clamp(min=1.0)
    neg = ((1.0 - target) * edge_valid).sum(dim=(1, 2)).clamp(min=1.0)
    ratio = (neg / pos).view(-1, 1, 1)
    w = edge_valid * (target * ratio + (1.0 - target))
    denom = edge_valid.sum(dim=(1, 2)).clamp(min=1.0)
    return ((bce * w).sum(dim=(1, 2)) / denom).mean()


def _actor_critic_loss(logits, values, next_values, batch, edge_valid, scores, gamma, explore_std, value_coef, entropy_coef):
    rewards, dones = batch["reward"], batch["done"]
    selected, noise = batch["selected"], batch["noise"]
    with torch.no_grad():
        target_v = rewards + gamma * next_values * (1.0 - dones)
        advantage = (target_v - values).detach().clamp(-5.0, 5.0)
    std = max(explore_std * 0.5, 0.05)
    sel_count = selected.sum(dim=(1, 2)).clamp(min=1.0)
    log_prob = (-0.5 * ((noise / std) ** 2) * selected).sum(dim=(1, 2)) / sel_count
    selected_score = (scores * selected).sum(dim=(1, 2)) / sel_count
    policy_term = log_prob * advantage + 0.5 * selected_score * advantage
    ent = -(scores.clamp(1e-6, 1 - 1e-6) * torch.log(scores.clamp(1e-6, 1 - 1e-6)))
    entropy = (ent * edge_valid).sum(dim=(1, 2)) / edge_valid.sum(dim=(1, 2)).clamp(min=1.0)
    value_loss = F.mse_loss(values, target_v)
    return -policy_term.mean() + value_coef * value_loss - entropy_coef * entropy.mean()


def pair_rl_loss(net, target_net, batch, gamma=0.95, explore_std=0.15, value_coef=0.5, entropy_coef=0.01):
    """Advantage actor-critic on the selected edges (PairCostHybrid.update :443-461)."""
    logits, values = net(*_tok(batch, "tok."))
    with torch.no_grad():
        _, next_values = target_net(*_tok(batch, "next."))
    scores = torch.sigmoid(logits.clamp(-8, 8))
    return _actor_critic_loss(logits, values, next_values, batch, batch["tok.edge_valid"], scores, gamma, explore_std,
                              value_coef, entropy_coef)


def escort_loss(net, target_net, batch, gamma=0.95, explore_std=0.35, value_coef=0.5, entropy_coef=0.01):
    """AttentionEscort.update (:597-625): as the pair loss, with unclamped logits in the sigmoid and the entropy masked
    by edge_valid AND the padding masks."""
    tf, tm, af, am = _tok(batch, "tok.")
    logits, values = net(tf, tm, af, am)
    with torch.no_grad():
        _, next_values = target_net(*_tok(batch, "next."))
    scores = torch.sigmoid(logits)
    pad = (~am).unsqueeze(2) * (~tm).unsqueeze(1)
    ev = batch["tok.edge_valid"]
    # the reference multiplies the entropy by edge_valid * pad but normalises by edge_valid alone
    rewards, dones = batch["reward"], batch["done"]
    selected, noise = batch["selected"], batch["noise"]
    with torch.no_grad():
        target_v = rewards + gamma * next_values * (1.0 - dones)
        advantage = (target_v - values).detach().clamp(-5.0, 5.0)
    std = max(explore_std * 0.5, 0.05)
    sel_count = selected.sum(dim=(1, 2)).clamp(min=1.0)
    log_prob = (-0.5 * ((noise / std) ** 2) * selected).sum(dim=(1, 2)) / sel_count
    selected_score = (scores * selected).sum(dim=(1, 2)) / sel_count
    policy_term = log_prob * advantage + 0.5 * selected_score * advantage
    ent = -(scores.clamp(1e-6, 1 - 1e-6) * torch.log(scores.clamp(1e-6, 1 - 1e-6)))
    entropy = (ent * ev * pad).sum(dim=(1, 2)) / ev.sum(dim=(1, 2)).clamp(min=1.0)
    value_loss = F.mse_loss(values, target_v)
    return -policy_term.mean() + value_coef * value_loss - entropy_coef * entropy.mean()


def commit_loss(net, target_net, batch, gamma=0.95):
    """AttentionCommit.update (:218-233): TD on the pooled (priority, commit) value + regression to the taken vectors."""
    tf, tm, af, am = _tok(batch, "tok.")
    pri_pred, com_pred = net(tf, tm, af, am)
    n_tasks = (~tm).float().sum(dim=1).clamp(min=1.0)
    n_agents = (~am).float().sum(dim=1).clamp(min=1.0)
    value = pri_pred.sum(dim=1) / n_tasks + 0.5 * com_pred.sum(dim=1) / n_agents
    with torch.no_grad():
        ntf, ntm, naf, nam = _tok(batch, "next.")
        n_pri, n_com = target_net(ntf, ntm, naf, nam)
        n_nt = (~ntm).float().sum(dim=1).clamp(min=1.0)
        n_na = (~nam).float().sum(dim=1).clamp(min=1.0)
        n_value = n_pri.sum(dim=1) / n_nt + 0.5 * n_com.sum(dim=1) / n_na
        target = batch["reward"] + gamma * (1.0 - batch["done"]) * n_value
    loss_v = F.mse_loss(value, target)
    valid_t, valid_a = (~tm).float(), (~am).float()
    loss_pri = ((pri_pred - batch["pri"]) ** 2 * valid_t).sum() / valid_t.sum().clamp(min=1.0)
    loss_com = ((com_pred - batch["com"]) ** 2 * valid_a).sum() / valid_a.sum().clamp(min=1.0)
    return loss_v + 0.5 * loss_pri + 0.5 * loss_com


# ---------------------------------------------------------------------------------------------- update steps
class Learner:
    """Optimiser state shared by the three trainers: Adam, gradient clipping, periodic target refresh (the policies'
    __init__ / update of the reference: clip 5.0 and refresh every 40 updates for Att-Commit, 1.0 / 20 for the actor-critics)."""

    def __init__(self, net: nn.Module, lr=1e-3, clip=1.0, target_every=20):
        import copy

        self.net = net
        self.target = copy.deepcopy(net).eval()
        self.optim = torch.optim.Adam(net.parameters(), lr=lr)
        self.lr, self.clip, self.target_every = lr, clip, target_every
        self.n_updates = 0

    def step(self, loss: torch.Tensor, refresh_target=True) -> float:
        self.optim.zero_grad()
        loss.backward()
        nn.utils.clip_grad_norm_(self.net.parameters(), self.clip)
        self.optim.step()
        self.n_updates += 1
        if refresh_target and self.n_updates % self.target_every == 0:
            self.target.load_state_dict(self.net.state_dict())
        return float(loss.item())


def train_pair_il(env, net, episodes=1, batch_rows=1024, lr=1e-3, il_warmup=50, log: Optional[Callable] = None):
    """Imitation of the Global-Hungarian teacher (run_il_episode, train_pair_cost.py:96-131): every step the rows whose
    replan rule fired and whose expert mask is not empty are one mini-batch (split into chunks of batch_rows)."""
    col = ILCollector(env)
    learner = Learner(net, lr=lr, clip=5.0)
    losses = []
    for ep in range(episodes):
        col.reset()
        for _ in range(int(env.cfg.max_time_steps)):
            tokens, mask, planned = col.step()
            rows = (planned.bool() & (mask.sum(dim=(1, 2)) > 0)).nonzero().flatten()
            for i in range(0, rows.numel(), batch_rows):
                r = rows[i:i + batch_rows]
                net.train()
                sub = {k: v.index_select(0, r) for k, v in tokens.items() if k in TOKEN_KEYS}
                scale = min(1.0, (learner.n_updates + 1) / max(il_warmup, 1))     # linear warm-up (:359-361)
                for g in learner.optim.param_groups:
                    g["lr"] = lr * scale
                losses.append(learner.step(pair_il_loss(net, sub, mask.index_select(0, r)), refresh_target=False))
        if log:
            log(ep, losses[-1] if losses else None)
    return losses


def _noisy_pair_scores(net, explore, explore_std, clamp):
    """PairCostHybrid.act (:266-278): Gaussian noise of std explore_std on the valid edges, tanh, clamp."""
    def act(tokens):
        with torch.no_grad():
            logits, _ = net(tokens["task_feats"], tokens["task_mask"], tokens["agent_feats"], tokens["agent_mask"])
        noise = torch.randn_like(logits) * explore_std * tokens["edge_valid"] if explore else torch.zeros_like(logits)
        scores = torch.tanh(logits + noise) * clamp * tokens["edge_valid"]
        act.noise, act.logits = noise, logits
        return scores
    return act


def train_pair_rl(env, net, episodes=1, batch_size=64, updates_per_step=1, lr=1e-3, gamma=0.95, explore_std=0.15,
                  explore=True, score_clamp=0.35, capacity=50_000, log: Optional[Callable] = None):
    """run_rl_episode (train_pair_cost.py:134-159) on the batch: noisy scores -> Local-Hungarian -> transition."""
    col = RLCollector(env)
    learner = Learner(net, lr=lr, clip=1.0, target_every=20)
    buf = ReplayBuffer(capacity, env.device)
    losses = []
    for ep in range(episodes):
        col.reset()
        act = _noisy_pair_scores(net, explore, explore_std, score_clamp)
        for _ in range(int(env.cfg.max_time_steps)):
            net.eval()
            tr = col.step(act)
            tr["noise"], tr["logits"] = act.noise, act.logits
            buf.push(flatten_transition(tr, ("selected", "noise")), tr["planned"])
            for _ in range(updates_per_step):
                if len(buf) >= min(batch_size, 16):
                    net.train()
                    losses.append(learner.step(pair_rl_loss(net, learner.target, buf.sample(batch_size), gamma, explore_std)))
        if log:
            log(ep, losses[-1] if losses else None)
    return losses


def train_att_commit(env, net, episodes=1, batch_size=64, updates_per_step=1, lr=1e-3, gamma=0.95, eps=0.45,
                     capacity=50_000, log: Optional[Callable] = None):
    """experiments/train_att_commit.py:28-75 on the batch (cadence 12, Gaussian exploration 0.2 with probability eps)."""
    col = CommitCollector(env)
    learner = Learner(net, lr=lr, clip=5.0, target_every=40)
    buf = ReplayBuffer(capacity, env.device)
    losses = []

    def act(tokens):
        with torch.no_grad():
            pri, com = net(tokens["task_feats"], tokens["task_mask"], tokens["agent_feats"], tokens["agent_mask"])
        explore = (torch.rand(pri.shape[0], 1, device=pri.device) < eps).float()
        pri = torch.where(explore.bool(), (pri + torch.randn_like(pri) * 0.2).clamp(0.0, 1.0), pri)
        com = torch.where(explore.bool(), (com + torch.randn_like(com) * 0.2).clamp(0.0, 1.0), com)
        return pri, com

    for ep in range(episodes):
        col.reset()
        for _ in range(int(env.cfg.max_time_steps)):
            net.eval()
            tr = col.step(act)
            buf.push(flatten_transition(tr, ("pri", "com")), tr["planned"])
            for _ in range(updates_per_step):
                if len(buf) >= batch_size:
                    net.train()
                    losses.append(learner.step(commit_loss(net, learner.target, buf.sample(batch_size), gamma)))
        if log:
            log(ep, losses[-1] if losses else None)
    return losses


def train_escort(env, net, episodes=1, batch_size=64, update_every=4, lr=1e-3, gamma=0.95, explore_std=0.35, eps=0.3,
                 capacity=50_000, log: Optional[Callable] = None):
    """experiments/train_escort.py:28-82 on the batch (cadence 12, every event tag, reward delta S_ESC / 20, one update per
    four pushes in the reference: here `update_every` environment steps)."""
    col = EscortCollector(env)
    learner = Learner(net, lr=lr, clip=1.0, target_every=20)
    buf = ReplayBuffer(capacity, env.device)
    losses = []

    def act(tokens):
        with torch.no_grad():
            logits, _ = net(tokens["task_feats"], tokens["task_mask"], tokens["agent_feats"], tokens["agent_mask"])
        noise = torch.zeros_like(logits)
        if eps > 0:
            noise = torch.randn_like(logits) * (explore_std * max(eps, 0.05)) * tokens["edge_valid"]
        scores = torch.sigmoid((logits + noise).clamp(-20.0, 20.0)) * tokens["edge_valid"]
        scores = scores * (~tokens["agent_mask"]).unsqueeze(2) * (~tokens["task_mask"]).unsqueeze(1)
        return scores, noise, logits

    step_no = 0
    for ep in range(episodes):
        col.reset()
        for _ in range(int(env.cfg.max_time_steps)):
            net.eval()
            tr = col.step(act)
            buf.push(flatten_transition(tr, ("selected", "noise")), tr["planned"])
            step_no += 1
            if step_no % update_every == 0 and len(buf) >= min(batch_size, 16):
                net.train()
                losses.append(learner.step(escort_loss(net, learner.target, buf.sample(batch_size), gamma, explore_std)))
        if log:
            log(ep, losses[-1] if losses else None)
    return losses
