"""Learned pair scorers (PyTorch, by design: BASELINE.json north_star keeps the scorers in torch).

AttPairNet / MLPPairNet reproduce the architectures of TaskAllocation/Hybrid/PairCostHybrid.py:89-196
(same sub-module construction order and shapes, so `torch.manual_seed(s)` followed by construction
yields the reference's random-init weights).  `pair_scores` is PairCostHybrid.act without exploration
noise (:266-278): scores = tanh(logits) * 0.35 * edge_valid.
"""
from __future__ import annotations

from typing import Optional

import os

import torch
import torch.nn as nn

TASK_FEAT_DIM = 13
AGENT_FEAT_DIM = 12
SCORE_CLAMP = 0.35


class AttPairNet(nn.Module):
    def __init__(self, max_tasks=32, max_agents=16, d_model=64, nhead=4, n_layers=2, dropout=0.1,
                 task_feat_dim=TASK_FEAT_DIM, agent_feat_dim=AGENT_FEAT_DIM, ff_mult=2):
        super().__init__()
        self.max_tasks, self.max_agents = max_tasks, max_agents
        self.task_proj = nn.Linear(task_feat_dim, d_model)
        self.agent_proj = nn.Linear(agent_feat_dim, d_model)
        self.type_embed = nn.Embedding(2, d_model)
        layer = nn.TransformerEncoderLayer(d_model=d_model, nhead=nhead, dim_feedforward=d_model * ff_mult,
                                           batch_first=True, dropout=dropout)
        self.self_encoder = nn.TransformerEncoder(layer, num_layers=max(1, n_layers - 1))
        self.cross_a2t = nn.MultiheadAttention(d_model, nhead, dropout=dropout, batch_first=True)
        self.cross_t2a = nn.MultiheadAttention(d_model, nhead, dropout=dropout, batch_first=True)
        self.pair_head = nn.Sequential(nn.Linear(d_model * 3, d_model), nn.ReLU(), nn.Linear(d_model, d_model // 2),
                                       nn.ReLU(), nn.Linear(d_model // 2, 1))
        self.value_head = nn.Sequential(nn.Linear(d_model, d_model // 2), nn.ReLU(), nn.Linear(d_model // 2, 1))

    def forward(self, task_feats, task_mask, agent_feats, agent_mask):
        t_emb = self.task_proj(task_feats) + self.type_embed.weight[1]
        a_emb = self.agent_proj(agent_feats) + self.type_embed.weight[0]
        tokens = torch.cat([a_emb, t_emb], dim=1)
        pad_mask = torch.cat([agent_mask, task_mask], dim=1)
        h = self.self_encoder(tokens, src_key_padding_mask=pad_mask)
        a_h = h[:, : self.max_agents, :]
        t_h = h[:, self.max_agents:, :]
        a_ctx, _ = self.cross_a2t(a_h, t_h, t_h, key_padding_mask=task_mask, need_weights=False)
        t_ctx, _ = self.cross_t2a(t_h, a_h, a_h, key_padding_mask=agent_mask, need_weights=False)
        a_h = a_h + a_ctx
        t_h = t_h + t_ctx
        a_exp = a_h.unsqueeze(2).expand(-1, -1, self.max_tasks, -1)
        t_exp = t_h.unsqueeze(1).expand(-1, self.max_agents, -1, -1)
        pair = torch.cat([a_exp, t_exp, a_exp * t_exp], dim=-1)
        logits = self.pair_head(pair).squeeze(-1)
        logits = logits.masked_fill(agent_mask.unsqueeze(2), -1e9)
        logits = logits.masked_fill(task_mask.unsqueeze(1), -1e9)
        valid = (~pad_mask).unsqueeze(-1).float()
        pooled = (h * valid).sum(dim=1) / valid.sum(dim=1).clamp(min=1.0)
        value = self.value_head(pooled).squeeze(-1)
        return logits, value


class MLPPairNet(nn.Module):
    def __init__(self, max_tasks=32, max_agents=16, hidden=128, d_model=64, task_feat_dim=TASK_FEAT_DIM,
                 agent_feat_dim=AGENT_FEAT_DIM, **_):
        super().__init__()
        self.max_tasks, self.max_agents = max_tasks, max_agents
        in_dim = task_feat_dim + agent_feat_dim
        self.pair_mlp = nn.Sequential(nn.Linear(in_dim, hidden), nn.ReLU(), nn.Linear(hidden, hidden), nn.ReLU(),
                                      nn.Linear(hidden, 1))
        self.value_mlp = nn.Sequential(nn.Linear(in_dim, hidden), nn.ReLU(), nn.Linear(hidden, 1))

    def forward(self, task_feats, task_mask, agent_feats, agent_mask):
        a = agent_feats.size(1)
        t = task_feats.size(1)
        a_exp = agent_feats.unsqueeze(2).expand(-1, -1, t, -1)
        t_exp = task_feats.unsqueeze(1).expand(-1, a, -1, -1)
        logits = self.pair_mlp(torch.cat([a_exp, t_exp], dim=-1)).squeeze(-1)
        logits = logits.masked_fill(agent_mask.unsqueeze(2), -1e9)
        logits = logits.masked_fill(task_mask.unsqueeze(1), -1e9)
        am = (~agent_mask).float().unsqueeze(-1)
        tm = (~task_mask).float().unsqueeze(-1)
        a_pool = (agent_feats * am).sum(1) / am.sum(1).clamp(min=1.0)
        t_pool = (task_feats * tm).sum(1) / tm.sum(1).clamp(min=1.0)
        value = self.value_mlp(torch.cat([a_pool, t_pool], dim=-1)).squeeze(-1)
        return logits, value


CONTEXT_DIM = 8  # build_context_summary (ContextPairHybrid.py:23-31); 1 in raw mode


class AttContextPairNet(nn.Module):
    """AttContextPairNet (ContextPairHybrid.py:81-151): Att-Pair plus a context vector (projected summary + pooled encoder
    output) appended to every pair."""

    def __init__(self, max_tasks=32, max_agents=16, d_model=64, nhead=4, n_layers=2, dropout=0.1,
                 task_feat_dim=TASK_FEAT_DIM, agent_feat_dim=AGENT_FEAT_DIM, context_dim=CONTEXT_DIM):
        super().__init__()
        self.max_tasks, self.max_agents, self.d_model = max_tasks, max_agents, d_model
        self.task_proj = nn.Linear(task_feat_dim, d_model)
        self.agent_proj = nn.Linear(agent_feat_dim, d_model)
        self.ctx_proj = nn.Linear(context_dim, d_model)
        self.type_embed = nn.Embedding(2, d_model)
        layer = nn.TransformerEncoderLayer(d_model=d_model, nhead=nhead, dim_feedforward=d_model * 2,
                                           batch_first=True, dropout=dropout)
        self.self_encoder = nn.TransformerEncoder(layer, num_layers=max(1, n_layers - 1))
        self.cross_a2t = nn.MultiheadAttention(d_model, nhead, dropout=dropout, batch_first=True)
        self.cross_t2a = nn.MultiheadAttention(d_model, nhead, dropout=dropout, batch_first=True)
        self.pair_head = nn.Sequential(nn.Linear(d_model * 4, d_model), nn.ReLU(), nn.Linear(d_model, d_model // 2),
                                       nn.ReLU(), nn.Linear(d_model // 2, 1))
        self.value_head = nn.Sequential(nn.Linear(d_model * 2, d_model), nn.ReLU(), nn.Linear(d_model, 1))

    def forward(self, task_feats, task_mask, agent_feats, agent_mask, context):
        t_emb = self.task_proj(task_feats) + self.type_embed.weight[1]
        a_emb = self.agent_proj(agent_feats) + self.type_embed.weight[0]
        tokens = torch.cat([a_emb, t_emb], dim=1)
        pad_mask = torch.cat([agent_mask, task_mask], dim=1)
        h = self.self_encoder(tokens, src_key_padding_mask=pad_mask)
        a_h = h[:, : self.max_agents, :]
        t_h = h[:, self.max_agents:, :]
        a_ctx, _ = self.cross_a2t(a_h, t_h, t_h, key_padding_mask=task_mask, need_weights=False)
        t_ctx, _ = self.cross_t2a(t_h, a_h, a_h, key_padding_mask=agent_mask, need_weights=False)
        a_h = a_h + a_ctx
        t_h = t_h + t_ctx
        valid = (~pad_mask).unsqueeze(-1).float()
        pooled = (h * valid).sum(dim=1) / valid.sum(dim=1).clamp(min=1.0)
        ctx = self.ctx_proj(context) + pooled
        ctx_exp = ctx.unsqueeze(1).unsqueeze(2).expand(-1, self.max_agents, self.max_tasks, -1)
        a_exp = a_h.unsqueeze(2).expand(-1, -1, self.max_tasks, -1)
        t_exp = t_h.unsqueeze(1).expand(-1, self.max_agents, -1, -1)
        logits = self.pair_head(torch.cat([a_exp, t_exp, a_exp * t_exp, ctx_exp], dim=-1)).squeeze(-1)
        logits = logits.masked_fill(agent_mask.unsqueeze(2), -1e9)
        logits = logits.masked_fill(task_mask.unsqueeze(1), -1e9)
        value = self.value_head(torch.cat([pooled, ctx], dim=-1)).squeeze(-1)
        return logits, value


class MLPContextPairNet(nn.Module):
    """MLPContextPairNet (ContextPairHybrid.py:154-207): the matched control without attention."""

    def __init__(self, max_tasks=32, max_agents=16, hidden=192, d_model=64, task_feat_dim=TASK_FEAT_DIM,
                 agent_feat_dim=AGENT_FEAT_DIM, context_dim=CONTEXT_DIM, **_):
        super().__init__()
        self.max_tasks, self.max_agents = max_tasks, max_agents
        in_pair = task_feat_dim + agent_feat_dim + task_feat_dim + agent_feat_dim + context_dim
        self.ctx_mlp = nn.Sequential(nn.Linear(task_feat_dim + agent_feat_dim + context_dim, hidden), nn.ReLU(),
                                     nn.Linear(hidden, hidden), nn.ReLU())
        self.pair_mlp = nn.Sequential(nn.Linear(in_pair, hidden), nn.ReLU(), nn.Linear(hidden, hidden), nn.ReLU(),
                                      nn.Linear(hidden, 1))
        self.value_mlp = nn.Sequential(nn.Linear(hidden, hidden // 2), nn.ReLU(), nn.Linear(hidden // 2, 1))

    def forward(self, task_feats, task_mask, agent_feats, agent_mask, context):
        am = (~agent_mask).float().unsqueeze(-1)
        tm = (~task_mask).float().unsqueeze(-1)
        a_pool = (agent_feats * am).sum(1) / am.sum(1).clamp(min=1.0)
        t_pool = (task_feats * tm).sum(1) / tm.sum(1).clamp(min=1.0)
        ctx_h = self.ctx_mlp(torch.cat([a_pool, t_pool, context], dim=-1))
        b, a, _ = agent_feats.shape
        t = task_feats.size(1)
        a_exp = agent_feats.unsqueeze(2).expand(-1, -1, t, -1)
        t_exp = task_feats.unsqueeze(1).expand(-1, a, -1, -1)
        a_p = a_pool.unsqueeze(1).unsqueeze(2).expand(-1, a, t, -1)
        t_p = t_pool.unsqueeze(1).unsqueeze(2).expand(-1, a, t, -1)
        c_exp = context.unsqueeze(1).unsqueeze(2).expand(-1, a, t, -1)
        logits = self.pair_mlp(torch.cat([a_exp, t_exp, a_p, t_p, c_exp], dim=-1)).squeeze(-1)
        logits = logits.masked_fill(agent_mask.unsqueeze(2), -1e9)
        logits = logits.masked_fill(task_mask.unsqueeze(1), -1e9)
        return logits, self.value_mlp(ctx_h).squeeze(-1)


class BipartiteMPLayer(nn.Module):
    """One agent <-> task message-passing step restricted to the valid edges (GNNPairHybrid.py:23-57)."""

    def __init__(self, d_model=64, msg_hidden=96):
        super().__init__()
        self.msg_a2t = nn.Sequential(nn.Linear(d_model * 2, msg_hidden), nn.ReLU(), nn.Linear(msg_hidden, d_model))
        self.msg_t2a = nn.Sequential(nn.Linear(d_model * 2, msg_hidden), nn.ReLU(), nn.Linear(msg_hidden, d_model))
        self.norm_a = nn.LayerNorm(d_model)
        self.norm_t = nn.LayerNorm(d_model)

    def forward(self, a_h, t_h, edge_valid):
        a_exp = a_h.unsqueeze(2).expand(-1, -1, t_h.size(1), -1)
        t_exp = t_h.unsqueeze(1).expand(-1, a_h.size(1), -1, -1)
        pair = torch.cat([a_exp, t_exp], dim=-1)
        w = edge_valid.unsqueeze(-1)
        msg_t = self.msg_a2t(pair) * w
        t_h = self.norm_t(t_h + msg_t.sum(dim=1) / w.sum(dim=1).clamp(min=1e-6))
        msg_a = self.msg_t2a(pair) * w
        a_h = self.norm_a(a_h + msg_a.sum(dim=2) / w.sum(dim=2).clamp(min=1e-6))
        return a_h, t_h


class GNNContextPairNet(nn.Module):
    """GNNContextPairNet (GNNPairHybrid.py:60-122): bipartite GNN edge scorer with the context-biased pair head."""

    def __init__(self, max_tasks=32, max_agents=16, d_model=64, n_layers=2, task_feat_dim=TASK_FEAT_DIM,
                 agent_feat_dim=AGENT_FEAT_DIM, context_dim=CONTEXT_DIM, **_):
        super().__init__()
        self.max_tasks, self.max_agents, self.d_model = max_tasks, max_agents, d_model
        self.task_proj = nn.Linear(task_feat_dim, d_model)
        self.agent_proj = nn.Linear(agent_feat_dim, d_model)
        self.ctx_proj = nn.Linear(context_dim, d_model)
        self.type_embed = nn.Embedding(2, d_model)
        self.layers = nn.ModuleList([BipartiteMPLayer(d_model) for _ in range(max(1, n_layers))])
        self.pair_head = nn.Sequential(nn.Linear(d_model * 4, d_model), nn.ReLU(), nn.Linear(d_model, d_model // 2),
                                       nn.ReLU(), nn.Linear(d_model // 2, 1))
        self.value_head = nn.Sequential(nn.Linear(d_model * 2, d_model), nn.ReLU(), nn.Linear(d_model, 1))

    def forward(self, task_feats, task_mask, agent_feats, agent_mask, context, edge_valid=None):
        a_h = self.agent_proj(agent_feats) + self.type_embed.weight[0]
        t_h = self.task_proj(task_feats) + self.type_embed.weight[1]
        pad = (~agent_mask).unsqueeze(2).float() * (~task_mask).unsqueeze(1).float()
        edge_valid = pad if edge_valid is None else edge_valid.float() * pad
        for layer in self.layers:
            a_h, t_h = layer(a_h, t_h, edge_valid)
        am = (~agent_mask).float().unsqueeze(-1)
        tm = (~task_mask).float().unsqueeze(-1)
        a_pool = (a_h * am).sum(1) / am.sum(1).clamp(min=1.0)
        t_pool = (t_h * tm).sum(1) / tm.sum(1).clamp(min=1.0)
        pooled = 0.5 * (a_pool + t_pool)
        ctx = self.ctx_proj(context) + pooled
        ctx_exp = ctx.unsqueeze(1).unsqueeze(2).expand(-1, self.max_agents, self.max_tasks, -1)
        a_exp = a_h.unsqueeze(2).expand(-1, -1, self.max_tasks, -1)
        t_exp = t_h.unsqueeze(1).expand(-1, self.max_agents, -1, -1)
        logits = self.pair_head(torch.cat([a_exp, t_exp, a_exp * t_exp, ctx_exp], dim=-1)).squeeze(-1)
        logits = logits.masked_fill(agent_mask.unsqueeze(2), -1e9)
        logits = logits.masked_fill(task_mask.unsqueeze(1), -1e9)
        logits = logits.masked_fill(edge_valid < 0.5, -1e9)
        value = self.value_head(torch.cat([pooled, ctx], dim=-1)).squeeze(-1)
        return logits, value


@torch.no_grad()
def context_pair_scores(net: nn.Module, tok: dict, score_clamp: float = SCORE_CLAMP) -> torch.Tensor:
    """tokens (BatchedMultiUAVEnv.tokens_context) -> edge scores f32 [B, max_agents, max_tasks]: ContextPairHybrid.act /
    GNNContextPairHybrid.act without exploration (ContextPairHybrid.py:235-246, GNNPairHybrid.py:160-171); feed to
    AllocSpec.pair_hybrid().  The GNN network also takes the valid-edge mask."""
    args = [tok["task_feats"], tok["task_mask"], tok["agent_feats"], tok["agent_mask"], tok["context"]]
    if isinstance(net, GNNContextPairNet):
        args.append(tok["edge_valid"])
    logits, _ = net(*args)
    return torch.tanh(logits) * score_clamp * tok["edge_valid"]


TASK_FEAT_DIM_E = 22   # build_escort_tokens task features (AttentionEscort.py:23)
AGENT_FEAT_DIM_E = 16  # build_escort_tokens agent features (AttentionEscort.py:25)
AGENT_FEAT_DIM_C = AGENT_FEAT_DIM + 1  # enrich_commit_tokens (AttentionCommit.py:65)


class AttCoalitionNet(AttPairNet):
    """AttCoalitionNet (AttentionEscort.py:244-330): the Att-Pair architecture at d_model 128, three layers, feed-forward
    4 d_model, over escort tokens [48, 22] / [16, 16].  Same construction order as the reference class."""

    def __init__(self, max_tasks=48, max_agents=16, d_model=128, nhead=4, n_layers=3, dropout=0.1):
        super().__init__(max_tasks, max_agents, d_model, nhead, n_layers, dropout, TASK_FEAT_DIM_E, AGENT_FEAT_DIM_E,
                         ff_mult=4)


class MLPCoalitionNet(nn.Module):
    """MLPCoalitionNet (AttentionEscort.py:333-375)."""

    def __init__(self, max_tasks=48, max_agents=16, hidden=256, **_):
        super().__init__()
        self.max_tasks, self.max_agents = max_tasks, max_agents
        in_dim = TASK_FEAT_DIM_E + AGENT_FEAT_DIM_E
        self.pair_mlp = nn.Sequential(nn.Linear(in_dim, hidden), nn.ReLU(), nn.Linear(hidden, hidden), nn.ReLU(),
                                      nn.Linear(hidden, 1))
        self.value_mlp = nn.Sequential(nn.Linear(max_tasks * TASK_FEAT_DIM_E + max_agents * AGENT_FEAT_DIM_E, hidden),
                                       nn.ReLU(), nn.Linear(hidden, 1))

    def forward(self, task_feats, task_mask, agent_feats, agent_mask):
        b, a, _ = agent_feats.shape
        t = task_feats.size(1)
        a_exp = agent_feats.unsqueeze(2).expand(-1, -1, t, -1)
        t_exp = task_feats.unsqueeze(1).expand(-1, a, -1, -1)
        logits = self.pair_mlp(torch.cat([a_exp, t_exp], dim=-1)).squeeze(-1)
        logits = logits.masked_fill(agent_mask.unsqueeze(2), -1e9)
        logits = logits.masked_fill(task_mask.unsqueeze(1), -1e9)
        flat = torch.cat([task_feats.reshape(b, -1), agent_feats.reshape(b, -1)], dim=1)
        return logits, self.value_mlp(flat).squeeze(-1)


class AttCommitNet(nn.Module):
    """AttCommitNet (AttentionCommit.py:68-101): joint encoder over agent + task tokens, sigmoid priority head per task
    and commit head per agent."""

    def __init__(self, max_tasks=32, max_agents=16, d_model=64, nhead=4, n_layers=2):
        super().__init__()
        self.max_tasks, self.max_agents = max_tasks, max_agents
        self.task_proj = nn.Linear(TASK_FEAT_DIM, d_model)
        self.agent_proj = nn.Linear(AGENT_FEAT_DIM_C, d_model)
        self.type_embed = nn.Embedding(2, d_model)
        enc_layer = nn.TransformerEncoderLayer(d_model=d_model, nhead=nhead, dim_feedforward=d_model * 2,
                                               batch_first=True, dropout=0.1)
        self.encoder = nn.TransformerEncoder(enc_layer, num_layers=n_layers)
        self.priority_head = nn.Linear(d_model, 1)
        self.commit_head = nn.Linear(d_model, 1)

    def forward(self, task_feats, task_mask, agent_feats, agent_mask):
        t_emb = self.task_proj(task_feats) + self.type_embed.weight[1]
        a_emb = self.agent_proj(agent_feats) + self.type_embed.weight[0]
        tokens = torch.cat([a_emb, t_emb], dim=1)
        pad_mask = torch.cat([agent_mask, task_mask], dim=1)
        h = self.encoder(tokens, src_key_padding_mask=pad_mask)
        a_h = h[:, : self.max_agents, :]
        t_h = h[:, self.max_agents:, :]
        priorities = torch.sigmoid(self.priority_head(t_h).squeeze(-1)).masked_fill(task_mask, 0.0)
        commits = torch.sigmoid(self.commit_head(a_h).squeeze(-1)).masked_fill(agent_mask, 0.0)
        return priorities, commits


class MLPCommitNet(nn.Module):
    """MLPCommitNet (AttentionCommit.py:104-129)."""

    def __init__(self, max_tasks=32, max_agents=16, hidden=128):
        super().__init__()
        self.max_tasks, self.max_agents = max_tasks, max_agents
        in_dim = max_tasks * TASK_FEAT_DIM + max_agents * AGENT_FEAT_DIM_C
        self.backbone = nn.Sequential(nn.Linear(in_dim, hidden), nn.ReLU(), nn.Linear(hidden, hidden), nn.ReLU())
        self.priority_head = nn.Linear(hidden, max_tasks)
        self.commit_head = nn.Linear(hidden, max_agents)

    def forward(self, task_feats, task_mask, agent_feats, agent_mask):
        b = task_feats.size(0)
        h = self.backbone(torch.cat([task_feats.reshape(b, -1), agent_feats.reshape(b, -1)], dim=1))
        priorities = torch.sigmoid(self.priority_head(h)).masked_fill(task_mask, 0.0)
        commits = torch.sigmoid(self.commit_head(h)).masked_fill(agent_mask, 0.0)
        return priorities, commits


@torch.no_grad()
def commit_vectors(net: nn.Module, tok: dict):
    """tokens (BatchedMultiUAVEnv.tokens_commit) -> (priorities f32 [B, max_tasks], commits f32 [B, max_agents]):
    AttentionCommit.act without exploration (AttentionCommit.py:167-175); feed to AllocSpec.att_commit()."""
    return net(tok["task_feats"], tok["task_mask"], tok["agent_feats"], tok["agent_mask"])


@torch.no_grad()
def coalition_scores(net: nn.Module, tok: dict) -> torch.Tensor:
    """tokens (BatchedMultiUAVEnv.tokens_escort) -> edge scores f32 [B, max_agents, max_tasks]: AttentionEscort.act
    without exploration (AttentionEscort.py:449-466): sigmoid of the logits clipped to +-20, zero on invalid edges
    and on padded rows / columns; feed to AllocSpec.att_escort() together with tok["task_order"]."""
    logits, _ = net(tok["task_feats"], tok["task_mask"], tok["agent_feats"], tok["agent_mask"])
    scores = 1.0 / (1.0 + torch.exp(-logits.clamp(-20.0, 20.0)))
    scores = scores * tok["edge_valid"]
    return scores * (~tok["agent_mask"]).unsqueeze(2) * (~tok["task_mask"]).unsqueeze(1)


@torch.no_grad()
def att_pair_logits_split(net: AttPairNet, task_feats, task_mask, agent_feats, agent_mask):
    """AttPairNet.forward with the first pair-head layer split by input block:
    W1 [a; t; a*t] = Wa a + Wt t + Wat (a*t).  Same function and parameters as the reference forward
    (PairCostHybrid.py:130-151); the [B, A, T, 192] concatenation is never materialised and two thirds of
    the first layer are evaluated per token instead of per pair.  fp32 results differ from the
    concatenated GEMM only by summation order (tests: <= 2e-5 absolute on logits)."""
    t_emb = net.task_proj(task_feats) + net.type_embed.weight[1]
    a_emb = net.agent_proj(agent_feats) + net.type_embed.weight[0]
    tokens = torch.cat([a_emb, t_emb], dim=1)
    pad_mask = torch.cat([agent_mask, task_mask], dim=1)
    h = net.self_encoder(tokens, src_key_padding_mask=pad_mask)
    n_a = agent_feats.shape[1]  # may be smaller than net.max_agents: padded agent rows carry no information
    a_h = h[:, :n_a, :]
    t_h = h[:, n_a:, :]
    a_ctx, _ = net.cross_a2t(a_h, t_h, t_h, key_padding_mask=task_mask, need_weights=False)
    t_ctx, _ = net.cross_t2a(t_h, a_h, a_h, key_padding_mask=agent_mask, need_weights=False)
    a_h = a_h + a_ctx
    t_h = t_h + t_ctx
    d = a_h.shape[-1]
    l1, l2, l3 = net.pair_head[0], net.pair_head[2], net.pair_head[4]
    wa, wt, wat = l1.weight[:, :d], l1.weight[:, d:2 * d], l1.weight[:, 2 * d:]
    ha = a_h @ wa.t()
    ht = t_h @ wt.t() + l1.bias
    prod = a_h.unsqueeze(2) * t_h.unsqueeze(1)
    h1 = torch.relu_(prod @ wat.t() + ha.unsqueeze(2) + ht.unsqueeze(1))
    h2 = torch.relu_(l2(h1))
    logits = l3(h2).squeeze(-1)
    logits = logits.masked_fill(agent_mask.unsqueeze(2), -1e9)
    logits = logits.masked_fill(task_mask.unsqueeze(1), -1e9)
    return logits


@torch.no_grad()
def pair_scores_fast(net: AttPairNet, tok: dict, score_clamp: float = SCORE_CLAMP) -> torch.Tensor:
    logits = att_pair_logits_split(net, tok["task_feats"], tok["task_mask"], tok["agent_feats"], tok["agent_mask"])
    return torch.tanh(logits) * score_clamp * tok["edge_valid"]


@torch.no_grad()
def pair_scores(net: nn.Module, tok: dict, score_clamp: float = SCORE_CLAMP) -> torch.Tensor:
    """tokens (BatchedMultiUAVEnv.tokens_pair) -> edge scores f32 [B, max_agents, max_tasks]."""
    logits, _ = net(tok["task_feats"], tok["task_mask"], tok["agent_feats"], tok["agent_mask"])
    return torch.tanh(logits) * score_clamp * tok["edge_valid"]


class GraphedPairScorer:
    """Att-Pair scoring of a (sub)batch of environments as one CUDA-graph replay.

    The eager forward is ~80 small launches and is CPU-launch-bound for the few hundred environments
    that replan on an ordinary step; the graph removes that overhead.  Sub-batches are padded to a
    bucket size (stale rows in the pad are computed and discarded).  The network, its parameters and
    the fp32 arithmetic are unchanged; nested-tensor packing of the padding mask is disabled because it
    is data dependent (padded rows are masked out downstream either way).
    """

    def __init__(self, net: AttPairNet, n_envs: int, device, buckets=None, max_tasks: int = 32, max_agents: int = 16,
                 live_agents: Optional[int] = None):
        """live_agents: upper bound of live agents (the fleet size).  Agent rows beyond it are always padding
        (masked keys, masked logits, zero edge_valid), so they are dropped from the forward: the valid
        outputs are unchanged up to fp32 summation order."""
        self.n_a = min(max_agents, live_agents) if live_agents else max_agents
        if buckets is None:
            buckets = list(range(128, 1537, 128)) + [2048, 3072]
        self.net = net.eval()
        if hasattr(net.self_encoder, "use_nested_tensor"):
            net.self_encoder.use_nested_tensor = False
        self.device = device
        self.buckets = sorted(b for b in buckets if b < n_envs) + [n_envs]
        self.n_envs = n_envs
        self.pool = None
        self.state = {}
        for B in self.buckets:
            bufs = {
                "task_feats": torch.zeros(B, max_tasks, TASK_FEAT_DIM, device=device),
                "task_mask_u8": torch.ones(B, max_tasks, dtype=torch.uint8, device=device),
                "agent_feats": torch.zeros(B, max_agents, AGENT_FEAT_DIM, device=device),
                "agent_mask_u8": torch.ones(B, max_agents, dtype=torch.uint8, device=device),
                "out_live": torch.zeros(B, self.n_a, max_tasks, device=device),
                "edge_valid": torch.zeros(B, max_agents, max_tasks, device=device),
                "out": torch.zeros(B, max_agents, max_tasks, device=device),
            }
            bufs["task_mask_u8"][:, 0] = 0
            bufs["agent_mask_u8"][:, 0] = 0
            self.state[B] = {"bufs": bufs, "graph": None}

    def _forward(self, b):
        n = self.n_a
        tok = {"task_feats": b["task_feats"], "task_mask": b["task_mask_u8"].bool(),
               "agent_feats": b["agent_feats"][:, :n], "agent_mask": b["agent_mask_u8"][:, :n].bool(),
               "edge_valid": b["edge_valid"][:, :n]}
        b["out"][:, :n].copy_(pair_scores_fast(self.net, tok))

    def _graph(self, B):
        st = self.state[B]
        if st["graph"] is None:
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(s):
                for _ in range(3):
                    self._forward(st["bufs"])
            torch.cuda.current_stream(self.device).wait_stream(s)
            g = torch.cuda.CUDAGraph()
            if self.pool is None:
                self.pool = torch.cuda.graph_pool_handle()
            with torch.cuda.graph(g, pool=self.pool):  # graphs are never replayed concurrently: share memory
                self._forward(st["bufs"])
            st["graph"] = g
        return st["graph"]

    def warm(self):
        for B in self.buckets:
            self._graph(B)

    @torch.no_grad()
    def score_all(self, tok: dict, scores_out: torch.Tensor):
        B = self.n_envs
        b = self.state[B]["bufs"]
        for k in ("task_feats", "task_mask_u8", "agent_feats", "agent_mask_u8", "edge_valid"):
            b[k].copy_(tok[k])
        self._graph(B).replay()
        scores_out.copy_(b["out"])

    @torch.no_grad()
    def score_subset(self, tok: dict, idx: torch.Tensor, scores_out: torch.Tensor):
        n = int(idx.numel())
        if n == 0:
            return
        B = next(x for x in self.buckets if x >= n)
        b = self.state[B]["bufs"]
        for k in ("task_feats", "task_mask_u8", "agent_feats", "agent_mask_u8", "edge_valid"):
            torch.index_select(tok[k], 0, idx, out=b[k][:n])
        self._graph(B).replay()
        scores_out.index_copy_(0, idx, b["out"][:n])


class FusedAttPairScorer:
    """AttPairNet forward as one CUDA kernel (csrc/muav_scorer.cu, C ABI muav_att_pair_scores): three environments
    per CTA packed token by token, only live agents / valid task columns are tokens, scores are written straight into the
    [E, max_agents, max_tasks] tensor the allocator reads (no gather / scatter, no intermediates in HBM).
    Parameters are the module's own (packed once); arithmetic is fp32 like the module."""

    _ORDER = [
        ("agent_proj_w", "agent_proj.weight"), ("agent_proj_b", "agent_proj.bias"),
        ("task_proj_w", "task_proj.weight"), ("task_proj_b", "task_proj.bias"), ("type_embed", "type_embed.weight"),
        ("enc_in_w", "self_encoder.layers.0.self_attn.in_proj_weight"),
        ("enc_in_b", "self_encoder.layers.0.self_attn.in_proj_bias"),
        ("enc_out_w", "self_encoder.layers.0.self_attn.out_proj.weight"),
        ("enc_out_b", "self_encoder.layers.0.self_attn.out_proj.bias"),
        ("enc_l1_w", "self_encoder.layers.0.linear1.weight"), ("enc_l1_b", "self_encoder.layers.0.linear1.bias"),
        ("enc_l2_w", "self_encoder.layers.0.linear2.weight"), ("enc_l2_b", "self_encoder.layers.0.linear2.bias"),
        ("enc_n1_w", "self_encoder.layers.0.norm1.weight"), ("enc_n1_b", "self_encoder.layers.0.norm1.bias"),
        ("enc_n2_w", "self_encoder.layers.0.norm2.weight"), ("enc_n2_b", "self_encoder.layers.0.norm2.bias"),
        ("a2t_in_w", "cross_a2t.in_proj_weight"), ("a2t_in_b", "cross_a2t.in_proj_bias"),
        ("a2t_out_w", "cross_a2t.out_proj.weight"), ("a2t_out_b", "cross_a2t.out_proj.bias"),
        ("t2a_in_w", "cross_t2a.in_proj_weight"), ("t2a_in_b", "cross_t2a.in_proj_bias"),
        ("t2a_out_w", "cross_t2a.out_proj.weight"), ("t2a_out_b", "cross_t2a.out_proj.bias"),
        ("head1_w", "pair_head.0.weight"), ("head1_b", "pair_head.0.bias"), ("head2_w", "pair_head.2.weight"),
        ("head2_b", "pair_head.2.bias"), ("head3_w", "pair_head.4.weight"), ("head3_b", "pair_head.4.bias"),
    ]

    def __init__(self, net, device, score_clamp: float = SCORE_CLAMP):
        """net: AttPairNet or AttContextPairNet (the context vector then comes from tok["context"])."""
        import ctypes as C

        from . import _lib

        sd = net.state_dict()
        self.has_context = "ctx_proj.weight" in sd
        if len(net.self_encoder.layers) != 1 or sd["task_proj.weight"].shape != (64, TASK_FEAT_DIM) \
                or sd["self_encoder.layers.0.linear1.weight"].shape != (128, 64) \
                or net.cross_a2t.num_heads != 4:
            raise ValueError("the fused kernel implements the default AttPairNet (d_model 64, 4 heads, 2 layers)")
        self.lib = _lib.cuda_lib()
        self.offsets = _lib.MuavAttPairOffsets()
        chunks, pos = [], 0
        order = list(self._ORDER)
        if self.has_context:
            if sd["pair_head.0.weight"].shape != (64, 256) or sd["ctx_proj.weight"].shape != (64, 8):
                raise ValueError("the fused kernel implements the default AttContextPairNet (context_dim 8)")
            order += [("ctx_proj_w", "ctx_proj.weight"), ("ctx_proj_b", "ctx_proj.bias")]
        self.offsets.has_context = int(self.has_context)
        for field, key in order:
            t = sd[key].detach().to(torch.float32)
            if t.dim() == 2 and field != "type_embed":
                t = t.t().contiguous()  # the kernel streams W^T ([in][out]) rows as float4
            t = t.reshape(-1)
            pos = (pos + 3) // 4 * 4  # keep every tensor 16-byte aligned
            setattr(self.offsets, field, pos)
            chunks.append((pos, t))
            pos += t.numel()
        buf = torch.zeros(pos, dtype=torch.float32)
        for p, t in chunks:
            buf[p:p + t.numel()] = t.cpu()
        self.params = buf.to(device)
        self.device = device
        self.clamp = float(score_clamp)
        self._C = C
        # tensor-core path (csrc/muav_scorer_tc.cu): the linear layers' weights re-packed once as TF32
        # hi / lo planes in the MMA's shared-memory layout.  MUAV_SCORER_TC=0 selects the FP32-pipe kernel (A/B runs).
        self.tcw = None
        if os.environ.get("MUAV_SCORER_TC", "1") != "0":
            with torch.cuda.device(device):
                self.tcw = torch.empty(int(self.lib.dll.muav_att_pair_tc_floats()), dtype=torch.float32, device=device)
                rc = self.lib.dll.muav_att_pair_tc_pack(self.params.data_ptr(), C.byref(self.offsets), self.tcw.data_ptr(),
                                                        C.c_void_p(torch.cuda.current_stream(device).cuda_stream))
            if rc != 0:
                raise RuntimeError(f"muav_att_pair_tc_pack failed: {rc}")

    @torch.no_grad()
    def score(self, tok: dict, scores_out: torch.Tensor, idx: Optional[torch.Tensor] = None, use_need: bool = False):
        """tok: fused token tensors of BatchedMultiUAVEnv.enable_fused_tokens (masks as uint8); idx: int32 env indices
        or None for all environments; use_need: skip (on the device) every environment whose tok["need"] flag is 0.
        Rows of environments not scored are left untouched."""
        C = self._C
        E, MA, MT = scores_out.shape
        n = E if idx is None else int(idx.numel())
        if n == 0:
            return
        if idx is not None and idx.dtype != torch.int32:
            idx = idx.to(torch.int32)
        if self.tcw is not None:
            rc = self.lib.dll.muav_att_pair_scores_tc(
                self.params.data_ptr(), C.byref(self.offsets), self.tcw.data_ptr(), tok["task_feats"].data_ptr(),
                tok["task_mask_u8"].data_ptr(), tok["agent_feats"].data_ptr(), tok["agent_mask_u8"].data_ptr(),
                tok["edge_valid"].data_ptr(), tok["context"].data_ptr() if self.has_context else None,
                None if idx is None else idx.data_ptr(),
                tok["need"].data_ptr() if use_need else None, n, MT, MA, C.c_float(self.clamp), scores_out.data_ptr(),
                C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
            if rc != 0:
                raise RuntimeError(f"muav_att_pair_scores_tc failed: {rc}")
            return
        rc = self.lib.dll.muav_att_context_pair_scores(
            self.params.data_ptr(), C.byref(self.offsets), tok["task_feats"].data_ptr(), tok["task_mask_u8"].data_ptr(),
            tok["agent_feats"].data_ptr(), tok["agent_mask_u8"].data_ptr(), tok["edge_valid"].data_ptr(),
            tok["context"].data_ptr() if self.has_context else None,
            None if idx is None else idx.data_ptr(), tok["need"].data_ptr() if use_need else None, n, MT, MA,
            C.c_float(self.clamp), scores_out.data_ptr(),
            C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"muav_att_context_pair_scores failed: {rc}")


class FusedAttCommitScorer:
    """AttCommitNet forward as one CUDA kernel (csrc/muav_scorer.cu att_commit_kernel, C ABI muav_att_commit_vectors): the
    pair kernel's packing (three environments per 64-token pass, only valid tokens) with two encoder layers and the
    priority / commit heads.  Outputs feed AllocSpec.att_commit() (plan_pri / plan_commit)."""

    def __init__(self, net, device):
        import ctypes as C

        from . import _lib

        sd = net.state_dict()
        if len(net.encoder.layers) != 2 or sd["task_proj.weight"].shape != (64, TASK_FEAT_DIM) \
                or sd["agent_proj.weight"].shape != (64, 13) or sd["encoder.layers.0.linear1.weight"].shape != (128, 64):
            raise ValueError("the fused kernel implements the default AttCommitNet (d_model 64, 4 heads, 2 layers)")
        self.lib = _lib.cuda_lib()
        self.offsets = _lib.MuavAttCommitOffsets()
        chunks, pos = [], 0

        def put(t, transpose):
            nonlocal pos
            t = t.detach().to(torch.float32)
            if transpose:
                t = t.t().contiguous()
            t = t.reshape(-1)
            pos = (pos + 3) // 4 * 4
            at = pos
            chunks.append((at, t))
            pos += t.numel()
            return at

        o = self.offsets
        o.agent_proj_w, o.agent_proj_b = put(sd["agent_proj.weight"], True), put(sd["agent_proj.bias"], False)
        o.task_proj_w, o.task_proj_b = put(sd["task_proj.weight"], True), put(sd["task_proj.bias"], False)
        o.type_embed = put(sd["type_embed.weight"], False)
        for l in range(2):
            pre = f"encoder.layers.{l}."
            o.enc_in_w[l], o.enc_in_b[l] = put(sd[pre + "self_attn.in_proj_weight"], True), put(sd[pre + "self_attn.in_proj_bias"], False)
            o.enc_out_w[l], o.enc_out_b[l] = put(sd[pre + "self_attn.out_proj.weight"], True), put(sd[pre + "self_attn.out_proj.bias"], False)
            o.enc_l1_w[l], o.enc_l1_b[l] = put(sd[pre + "linear1.weight"], True), put(sd[pre + "linear1.bias"], False)
            o.enc_l2_w[l], o.enc_l2_b[l] = put(sd[pre + "linear2.weight"], True), put(sd[pre + "linear2.bias"], False)
            o.enc_n1_w[l], o.enc_n1_b[l] = put(sd[pre + "norm1.weight"], False), put(sd[pre + "norm1.bias"], False)
            o.enc_n2_w[l], o.enc_n2_b[l] = put(sd[pre + "norm2.weight"], False), put(sd[pre + "norm2.bias"], False)
        o.priority_w, o.priority_b = put(sd["priority_head.weight"], False), put(sd["priority_head.bias"], False)
        o.commit_w, o.commit_b = put(sd["commit_head.weight"], False), put(sd["commit_head.bias"], False)
        buf = torch.zeros(pos, dtype=torch.float32)
        for p, t in chunks:
            buf[p:p + t.numel()] = t.cpu()
        self.params = buf.to(device)
        self.device = device
        self._C = C
        # tensor-core path (csrc/muav_scorer_tc.cu); MUAV_SCORER_TC=0 selects the FP32-pipe kernel (A/B runs)
        self.tcw = None
        if os.environ.get("MUAV_SCORER_TC", "1") != "0":
            with torch.cuda.device(device):
                self.tcw = torch.empty(int(self.lib.dll.muav_att_commit_tc_floats()), dtype=torch.float32, device=device)
                rc = self.lib.dll.muav_att_commit_tc_pack(self.params.data_ptr(), C.byref(self.offsets), self.tcw.data_ptr(),
                                                          C.c_void_p(torch.cuda.current_stream(device).cuda_stream))
            if rc != 0:
                raise RuntimeError(f"muav_att_commit_tc_pack failed: {rc}")

    @torch.no_grad()
    def vectors(self, tok: dict, pri_out: torch.Tensor, com_out: torch.Tensor, idx: Optional[torch.Tensor] = None,
                use_need: bool = False):
        """tok: fused commit tokens (enable_fused_tokens(commit=True); masks uint8) or a tokens_commit() dict; rows of
        environments that are not scored are left untouched."""
        C = self._C
        E, MT = pri_out.shape
        MA = com_out.shape[1]
        tm = tok["task_mask_u8"] if "task_mask_u8" in tok else tok["task_mask"].to(torch.uint8)
        am = tok["agent_mask_u8"] if "agent_mask_u8" in tok else tok["agent_mask"].to(torch.uint8)
        n = E if idx is None else int(idx.numel())
        if n == 0:
            return
        if idx is not None and idx.dtype != torch.int32:
            idx = idx.to(torch.int32)
        if self.tcw is not None:
            rc = self.lib.dll.muav_att_commit_vectors_tc(
                self.params.data_ptr(), C.byref(self.offsets), self.tcw.data_ptr(), tok["task_feats"].data_ptr(),
                tm.data_ptr(), tok["agent_feats"].data_ptr(), am.data_ptr(), None if idx is None else idx.data_ptr(),
                tok["need"].data_ptr() if use_need else None, n, MT, MA, pri_out.data_ptr(), com_out.data_ptr(),
                C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
            if rc != 0:
                raise RuntimeError(f"muav_att_commit_vectors_tc failed: {rc}")
            return
        rc = self.lib.dll.muav_att_commit_vectors(
            self.params.data_ptr(), C.byref(self.offsets), tok["task_feats"].data_ptr(), tm.data_ptr(),
            tok["agent_feats"].data_ptr(), am.data_ptr(), None if idx is None else idx.data_ptr(),
            tok["need"].data_ptr() if use_need else None, n, MT, MA, pri_out.data_ptr(), com_out.data_ptr(),
            C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"muav_att_commit_vectors failed: {rc}")


class FusedAttCoalitionScorer:
    """AttCoalitionNet forward as one CUDA kernel (csrc/muav_scorer.cu, att_pair_kernel at d_model 128 / two encoder layers /
    feed-forward 512; C ABI muav_att_coalition_scores): escort tokens -> sigmoid(clip(logits)) * edge_valid, the edge scores
    of AllocSpec.att_escort().  Same packing rules as the pair kernel (only live agents / valid task columns are tokens)."""

    def __init__(self, net, device):
        import ctypes as C

        from . import _lib

        sd = net.state_dict()
        if len(net.self_encoder.layers) != 2 or sd["task_proj.weight"].shape != (128, TASK_FEAT_DIM_E) \
                or sd["agent_proj.weight"].shape != (128, AGENT_FEAT_DIM_E) \
                or sd["self_encoder.layers.0.linear1.weight"].shape != (512, 128) or net.cross_a2t.num_heads != 4:
            raise ValueError("the fused kernel implements the default AttCoalitionNet (d_model 128, 4 heads, 3 layers)")
        self.lib = _lib.cuda_lib()
        self.offsets = _lib.MuavAttCoalOffsets()
        chunks, pos = [], 0

        def put(t, transpose):
            nonlocal pos
            t = t.detach().to(torch.float32)
            if transpose:
                t = t.t().contiguous()
            t = t.reshape(-1)
            pos = (pos + 3) // 4 * 4
            at = pos
            chunks.append((at, t))
            pos += t.numel()
            return at

        o = self.offsets
        o.agent_proj_w, o.agent_proj_b = put(sd["agent_proj.weight"], True), put(sd["agent_proj.bias"], False)
        o.task_proj_w, o.task_proj_b = put(sd["task_proj.weight"], True), put(sd["task_proj.bias"], False)
        o.type_embed = put(sd["type_embed.weight"], False)
        for l in range(2):
            pre = f"self_encoder.layers.{l}."
            o.enc_in_w[l], o.enc_in_b[l] = put(sd[pre + "self_attn.in_proj_weight"], True), put(sd[pre + "self_attn.in_proj_bias"], False)
            o.enc_out_w[l], o.enc_out_b[l] = put(sd[pre + "self_attn.out_proj.weight"], True), put(sd[pre + "self_attn.out_proj.bias"], False)
            o.enc_l1_w[l], o.enc_l1_b[l] = put(sd[pre + "linear1.weight"], True), put(sd[pre + "linear1.bias"], False)
            o.enc_l2_w[l], o.enc_l2_b[l] = put(sd[pre + "linear2.weight"], True), put(sd[pre + "linear2.bias"], False)
            o.enc_n1_w[l], o.enc_n1_b[l] = put(sd[pre + "norm1.weight"], False), put(sd[pre + "norm1.bias"], False)
            o.enc_n2_w[l], o.enc_n2_b[l] = put(sd[pre + "norm2.weight"], False), put(sd[pre + "norm2.bias"], False)
        for name, pre in (("a2t", "cross_a2t."), ("t2a", "cross_t2a.")):
            setattr(o, name + "_in_w", put(sd[pre + "in_proj_weight"], True))
            setattr(o, name + "_in_b", put(sd[pre + "in_proj_bias"], False))
            setattr(o, name + "_out_w", put(sd[pre + "out_proj.weight"], True))
            setattr(o, name + "_out_b", put(sd[pre + "out_proj.bias"], False))
        o.head1_w, o.head1_b = put(sd["pair_head.0.weight"], True), put(sd["pair_head.0.bias"], False)
        o.head2_w, o.head2_b = put(sd["pair_head.2.weight"], True), put(sd["pair_head.2.bias"], False)
        o.head3_w, o.head3_b = put(sd["pair_head.4.weight"], True), put(sd["pair_head.4.bias"], False)
        buf = torch.zeros(pos, dtype=torch.float32)
        for p, t in chunks:
            buf[p:p + t.numel()] = t.cpu()
        self.params = buf.to(device)
        self.device = device
        self._C = C

    @torch.no_grad()
    def score(self, tok: dict, scores_out: torch.Tensor, idx: Optional[torch.Tensor] = None, use_need: bool = False):
        """tok: escort tokens (tokens_escort() dict or the fused escort token tensors); rows of environments that are not
        scored are left untouched."""
        C = self._C
        E, MA, MT = scores_out.shape
        tm = tok["task_mask_u8"] if "task_mask_u8" in tok else tok["task_mask"].to(torch.uint8)
        am = tok["agent_mask_u8"] if "agent_mask_u8" in tok else tok["agent_mask"].to(torch.uint8)
        n = E if idx is None else int(idx.numel())
        if n == 0:
            return
        if idx is not None and idx.dtype != torch.int32:
            idx = idx.to(torch.int32)
        rc = self.lib.dll.muav_att_coalition_scores(
            self.params.data_ptr(), C.byref(self.offsets), tok["task_feats"].data_ptr(), tm.data_ptr(),
            tok["agent_feats"].data_ptr(), am.data_ptr(), tok["edge_valid"].data_ptr(),
            None if idx is None else idx.data_ptr(), tok["need"].data_ptr() if use_need else None, n, MT, MA,
            scores_out.data_ptr(), C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"muav_att_coalition_scores failed: {rc}")
