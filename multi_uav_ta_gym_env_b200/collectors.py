"""Batched data collectors for the pair-cost trainers.

    ILCollector  <- run_il_episode   experiments/train_pair_cost.py:96-131: the Global-Hungarian teacher (no visibility
                    mask, force=True) plans whenever the replan rule fires; the sample is (tokens of the state the teacher
                    saw, _expert_mask of its pairs through the valid edges); the rollout follows the teacher.
    RLCollector  <- run_rl_episode   experiments/train_pair_cost.py:134-159: the policy's scores drive the visibility-masked
                    Local-Hungarian; the transition is (tokens, scores, selected mask, (S_WPS' - S_WPS) / 20, next tokens,
                    done) for the environments that planned.

Everything stays on the device: tokens come out of the step kernel (fused emission), the teacher / policy plan is the
fused allocator, masks come from muav_pair_mask, S_WPS from muav_metrics.  The network update itself is the caller's
(PyTorch), exactly as in the reference trainers.
"""
from __future__ import annotations

import torch

from .batched_env import AllocSpec, BatchedMultiUAVEnv, HYBRID_EVENTS

TRAIN_INTERVAL = 20  # _should_replan of train_pair_cost.py:34-43 (the evaluation drivers use 15)


class _Base:
    def __init__(self, env: BatchedMultiUAVEnv, max_tasks=32, max_agents=16, interval=TRAIN_INTERVAL):
        self.env = env
        self.interval = interval
        self.tok = env.enable_fused_tokens(max_tasks, max_agents, interval, HYBRID_EVENTS)
        self._s_wps = env.lib.metric_names().index("S_WPS")
        self.reset()

    def reset(self):
        """Rewind every environment to its reset state (the trainers call env.reset() per episode)."""
        self.env.restore()
        self.env.refresh_fused_tokens()

    def _snapshot_tokens(self):
        t = self.tok
        return {"task_feats": t["task_feats"].clone(), "task_mask": t["task_mask_u8"].bool(),
                "agent_feats": t["agent_feats"].clone(), "agent_mask": t["agent_mask_u8"].bool(),
                "edge_valid": t["edge_valid"].clone(), "task_ids": t["task_ids"].clone()}

    def s_wps(self) -> torch.Tensor:
        return self.env.metrics()[:, self._s_wps]


class ILCollector(_Base):
    def step(self):
        """One environment step under the teacher.  Returns (tokens, expert_mask [E, A, T], planned u8 [E]): rows with
        planned == 1 are the imitation samples of this step (the trainer also skips empty masks, :121)."""
        env = self.env
        planned = self.tok["need"].clone()
        tokens = self._snapshot_tokens()
        teacher = AllocSpec(2, self.interval, HYBRID_EVENTS, False, False)   # Global-Hungarian: no visibility mask
        actions = env.allocate(teacher)
        mask = env.pair_mask(tokens, require_valid=True)
        env.step_batched(actions)
        return tokens, mask, planned


class RLCollector(_Base):
    def step(self, score_fn):
        """score_fn(tokens) -> edge scores f32 [E, A, T] (e.g. scorers.pair_scores with exploration noise added by the
        caller).  Returns a transition dict for the environments with planned == 1."""
        env = self.env
        planned = self.tok["need"].clone()
        tokens = self._snapshot_tokens()
        s_prev = self.s_wps()
        scores = score_fn(tokens)
        policy = AllocSpec(2, self.interval, HYBRID_EVENTS, True, True)      # PairCostHybrid.plan on Local-Hungarian
        actions = env.allocate(policy, edge_scores=scores)
        selected = env.pair_mask(tokens, require_valid=False)
        env.step_batched(actions)
        # tensor / tensor is an IEEE division (tensor / python scalar is a multiplication by the rounded reciprocal on CUDA)
        reward = (self.s_wps() - s_prev) / torch.full_like(s_prev, 20.0)
        done = (env.terminated | env.truncated).clone()
        # next_tok = policy.build_tokens(env) for every transition: the fused emission only covers environments that
        # replan next, so the standalone token kernel builds all of them
        mt, ma = tokens["edge_valid"].shape[2], tokens["edge_valid"].shape[1]
        return {"tokens": tokens, "scores": scores, "selected": selected, "reward": reward, "planned": planned,
                "next_tokens": env.tokens_pair(mt, ma), "done": done}


def _planned_mask(env: BatchedMultiUAVEnv, interval: int, event_mask: int) -> torch.Tensor:
    """`should` of the commit / escort trainers (train_att_commit.py:36-45, train_escort.py:35-50): t == 0, t % 12 == 0
    or an event with one of the listed tags since the last step; evaluated for every environment on the device."""
    t = env.header_int("T")
    tags = env.header_int("EV_TAGMASK")
    done = env.header_int("DONE")
    return ((t == 0) | (t % interval == 0) | ((tags & event_mask) != 0)) & (done == 0)


class CommitCollector:
    """run_episode of experiments/train_att_commit.py:28-75 for a whole batch: tokens (enrich_commit_tokens), the
    policy's (priorities, commits) drive AttentionCommit._plan_from_scores on the device (AllocSpec.att_commit, cadence 12),
    reward (S_WPS' - S_WPS) / 20, next tokens, done.  Transitions are the rows with planned == 1."""

    INTERVAL = 12

    def __init__(self, env: BatchedMultiUAVEnv, max_tasks=32, max_agents=16):
        self.env, self.mt, self.ma = env, max_tasks, max_agents
        self._s_wps = env.lib.metric_names().index("S_WPS")
        self.spec = AllocSpec.att_commit(self.INTERVAL)
        self.reset()

    def reset(self):
        self.env.restore()

    def score(self) -> torch.Tensor:
        return self.env.metrics()[:, self._s_wps]

    def step(self, act_fn):
        """act_fn(tokens) -> (priorities f32 [E, max_tasks], commits f32 [E, max_agents]) (exploration is the caller's)."""
        env = self.env
        planned = _planned_mask(env, self.INTERVAL, HYBRID_EVENTS)
        tokens = env.tokens_commit(self.mt, self.ma)
        s_prev = self.score()
        pri, com = act_fn(tokens)
        env.step_allocated(self.spec, 1, plan_pri=pri, plan_commit=com)
        reward = (self.score() - s_prev) / torch.full_like(s_prev, 20.0)
        done = (env.terminated | env.truncated).clone()
        return {"tokens": tokens, "pri": pri, "com": com, "reward": reward, "planned": planned.to(torch.uint8),
                "next_tokens": env.tokens_commit(self.mt, self.ma), "done": done}


class EscortCollector:
    """run_episode of experiments/train_escort.py:28-82 for a whole batch: build_escort_tokens, the policy's edge scores
    (sigmoid of noisy logits on valid edges, AttentionEscort.act :444-470) drive AttentionEscort._plan_from_scores on the
    device (AllocSpec.att_escort, cadence 12, every event tag), reward (S_ESC' - S_ESC) / 20, selected-edge mask of the
    plan, next tokens, done."""

    INTERVAL = 12

    def __init__(self, env: BatchedMultiUAVEnv, max_tasks=48, max_agents=16):
        self.env, self.mt, self.ma = env, max_tasks, max_agents
        self._s_esc = env.lib.metric_names().index("S_ESC")
        self.spec = AllocSpec.att_escort(self.INTERVAL)
        self.reset()

    def reset(self):
        self.env.restore()

    def score(self) -> torch.Tensor:
        return self.env.metrics()[:, self._s_esc]

    def step(self, act_fn):
        """act_fn(tokens) -> (scores [E, A, T], noise [E, A, T], logits [E, A, T])."""
        env = self.env
        planned = _planned_mask(env, self.INTERVAL, 0x1F)
        tokens = env.tokens_escort(self.mt, self.ma)
        s_prev = self.score()
        scores, noise, logits = act_fn(tokens)
        env.step_allocated(self.spec, 1, edge_scores=scores, task_order=tokens["task_order"])
        selected = env.pair_mask(tokens, require_valid=False)   # AttentionEscort._selected_mask == PairCostHybrid's
        reward = (self.score() - s_prev) / torch.full_like(s_prev, 20.0)
        done = (env.terminated | env.truncated).clone()
        return {"tokens": tokens, "scores": scores, "noise": noise, "logits": logits, "selected": selected,
                "reward": reward, "planned": planned.to(torch.uint8), "next_tokens": env.tokens_escort(self.mt, self.ma),
                "done": done}
