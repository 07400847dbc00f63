"""Learned pair scorers (PyTorch, by design: BASELINE.json north_star keeps the scorers in torch).

AttPairNet / MLPPairNet reproduce the architectures of TaskAllocation/Hybrid/PairCostHybrid.py:89-196
(same sub-module construction order and shapes, so `torch.manual_seed(s)` followed by construction
yields the reference's random-init weights).  `pair_scores` is PairCostHybrid.act without exploration
noise (:266-278): scores = tanh(logits) * 0.35 * edge_valid.
"""
from __future__ import annotations

import torch
import torch.nn as nn

TASK_FEAT_DIM = 13
AGENT_FEAT_DIM = 12
SCORE_CLAMP = 0.35


class AttPairNet(nn.Module):
    def __init__(self, max_tasks=32, max_agents=16, d_model=64, nhead=4, n_layers=2, dropout=0.1,
                 task_feat_dim=TASK_FEAT_DIM, agent_feat_dim=AGENT_FEAT_DIM):
        super().__init__()
        self.max_tasks, self.max_agents = max_tasks, max_agents
        self.task_proj = nn.Linear(task_feat_dim, d_model)
        self.agent_proj = nn.Linear(agent_feat_dim, d_model)
        self.type_embed = nn.Embedding(2, d_model)
        layer = nn.TransformerEncoderLayer(d_model=d_model, nhead=nhead, dim_feedforward=d_model * 2,
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


@torch.no_grad()
def pair_scores(net: nn.Module, tok: dict, score_clamp: float = SCORE_CLAMP) -> torch.Tensor:
    """tokens (BatchedMultiUAVEnv.tokens_pair) -> edge scores f32 [B, max_agents, max_tasks]."""
    logits, _ = net(tok["task_feats"], tok["task_mask"], tok["agent_feats"], tok["agent_mask"])
    return torch.tanh(logits) * score_clamp * tok["edge_valid"]
