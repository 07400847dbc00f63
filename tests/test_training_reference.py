"""multi_uav_ta_gym_env_b200.training against the reference policies' own update() (authoring container only).

Every loss function of training.py is evaluated on a synthetic batch and compared with the value the reference's
update() returns for the same transitions and the same network weights (PairCostHybrid._il_update / update,
AttentionCommit.update, AttentionEscort.update).  Dropout is switched off on both sides (the reference samples its
mini-batch in random order, so dropout masks could not line up); the buffers hold exactly one mini-batch, so the random
choice is a permutation and the (mean) losses agree up to float32 summation order."""
import numpy as np
import pytest
import torch

import refshim

pytestmark = pytest.mark.skipif(not refshim.reference_available(), reason="reference tree not present")

B = 64


def _freeze_mode(net):
    net.eval()
    net.train = lambda *a, **k: net   # update() calls net.train(): keep dropout off
    return net


def _tokens(rng, mt, ma, tdim, adim):
    tf = rng.random((B, mt, tdim), dtype=np.float32)
    af = rng.random((B, ma, adim), dtype=np.float32)
    tm = np.zeros((B, mt), bool)
    am = np.zeros((B, ma), bool)
    for b in range(B):
        tm[b, 4 + b % (mt - 6):] = True
        am[b, 3 + b % (ma - 4):] = True
    ev = (rng.random((B, ma, mt)) > 0.3).astype(np.float32) * (~am)[:, :, None] * (~tm)[:, None, :]
    return {"task_feats": tf, "task_mask": tm, "agent_feats": af, "agent_mask": am, "edge_valid": ev.astype(np.float32)}


def _flat(tok, nxt, extra):
    out = {}
    for k, v in tok.items():
        out["tok." + k] = torch.from_numpy(v)
        out["next." + k] = torch.from_numpy(nxt[k])
    out.update({k: torch.from_numpy(np.asarray(v, dtype=np.float32)) for k, v in extra.items()})
    return out


def _row(tok, b):
    return {k: v[b] for k, v in tok.items()}


def test_pair_losses_match_reference_updates():
    refshim.install()
    from TaskAllocation.Hybrid.PairCostHybrid import PairCostHybrid
    from multi_uav_ta_gym_env_b200 import scorers as S, training as T

    rng = np.random.default_rng(0)
    torch.manual_seed(0)
    pol = PairCostHybrid(use_attention=True)
    _freeze_mode(pol.net), _freeze_mode(pol.target)
    mine, mine_t = S.AttPairNet().eval(), S.AttPairNet().eval()
    mine.load_state_dict(pol.net.state_dict())
    mine_t.load_state_dict(pol.target.state_dict())
    tok, nxt = _tokens(rng, 32, 16, 13, 12), _tokens(rng, 32, 16, 13, 12)
    # imitation
    expert = ((rng.random((B, 16, 32)) > 0.9) * tok["edge_valid"]).astype(np.float32)
    want = pol._il_update([_row(tok, b) for b in range(B)], [expert[b] for b in range(B)])
    got = T.pair_il_loss(mine, {k: torch.from_numpy(v) for k, v in tok.items()}, torch.from_numpy(expert))
    assert abs(float(got) - want) < 1e-5 * max(1.0, abs(want))
    # actor-critic (weights moved by the imitation step above: reload)
    mine.load_state_dict(pol.net.state_dict())
    selected = ((rng.random((B, 16, 32)) > 0.9) * tok["edge_valid"]).astype(np.float32)
    noise = (rng.standard_normal((B, 16, 32)) * 0.15).astype(np.float32) * tok["edge_valid"]
    reward = rng.standard_normal(B).astype(np.float32)
    done = (rng.random(B) > 0.9)
    pol.buffer = []
    for b in range(B):
        pol.push(_row(tok, b), np.zeros((16, 32)), noise[b], np.zeros((16, 32)), selected[b], float(reward[b]), _row(nxt, b), bool(done[b]))
    want = pol.update(batch_size=B)
    batch = _flat(tok, nxt, {"selected": selected, "noise": noise, "reward": reward, "done": done.astype(np.float32)})
    got = T.pair_rl_loss(mine, mine_t, batch, gamma=pol.gamma, explore_std=pol.explore_std, value_coef=pol.value_coef,
                         entropy_coef=pol.entropy_coef)
    assert abs(float(got) - want) < 2e-5 * max(1.0, abs(want)), (float(got), want)


def test_commit_loss_matches_reference_update():
    refshim.install()
    from TaskAllocation.Hybrid.AttentionCommit import AttentionCommit
    from multi_uav_ta_gym_env_b200 import scorers as S, training as T

    rng = np.random.default_rng(1)
    torch.manual_seed(1)
    pol = AttentionCommit(use_attention=True)
    _freeze_mode(pol.net), _freeze_mode(pol.target)
    mine, mine_t = S.AttCommitNet().eval(), S.AttCommitNet().eval()
    mine.load_state_dict(pol.net.state_dict())
    mine_t.load_state_dict(pol.target.state_dict())
    tok, nxt = _tokens(rng, 32, 16, 13, 13), _tokens(rng, 32, 16, 13, 13)
    pri = rng.random((B, 32), dtype=np.float32)
    com = rng.random((B, 16), dtype=np.float32)
    reward = rng.standard_normal(B).astype(np.float32)
    done = (rng.random(B) > 0.9)
    pol.buffer = []
    for b in range(B):
        pol.push(_row(tok, b), pri[b], com[b], float(reward[b]), _row(nxt, b), float(done[b]))
    want = pol.update(batch_size=B)
    batch = _flat(tok, nxt, {"pri": pri, "com": com, "reward": reward, "done": done.astype(np.float32)})
    got = T.commit_loss(mine, mine_t, batch, gamma=pol.gamma)
    assert abs(float(got) - want) < 2e-5 * max(1.0, abs(want)), (float(got), want)


def test_escort_loss_matches_reference_update():
    refshim.install()
    from TaskAllocation.Hybrid.AttentionEscort import AttentionEscort
    from multi_uav_ta_gym_env_b200 import scorers as S, training as T

    rng = np.random.default_rng(2)
    torch.manual_seed(2)
    pol = AttentionEscort(use_attention=True)
    _freeze_mode(pol.net), _freeze_mode(pol.target)
    mine, mine_t = S.AttCoalitionNet().eval(), S.AttCoalitionNet().eval()
    mine.load_state_dict(pol.net.state_dict())
    mine_t.load_state_dict(pol.target.state_dict())
    mt, ma = pol.max_tasks, pol.max_agents
    tok, nxt = _tokens(rng, mt, ma, 22, 16), _tokens(rng, mt, ma, 22, 16)
    selected = ((rng.random((B, ma, mt)) > 0.9) * tok["edge_valid"]).astype(np.float32)
    noise = (rng.standard_normal((B, ma, mt)) * 0.1).astype(np.float32) * tok["edge_valid"]
    reward = rng.standard_normal(B).astype(np.float32)
    done = (rng.random(B) > 0.9)
    pol.buffer = []
    for b in range(B):
        pol.push(_row(tok, b), np.zeros((ma, mt)), noise[b], np.zeros((ma, mt)), selected[b], float(reward[b]), _row(nxt, b), bool(done[b]))
    want = pol.update(batch_size=B)
    batch = _flat(tok, nxt, {"selected": selected, "noise": noise, "reward": reward, "done": done.astype(np.float32)})
    got = T.escort_loss(mine, mine_t, batch, gamma=pol.gamma, explore_std=pol.explore_std, value_coef=pol.value_coef,
                        entropy_coef=pol.entropy_coef)
    assert abs(float(got) - want) < 2e-5 * max(1.0, abs(want)), (float(got), want)


def test_replay_buffer_ring_and_sampling():
    from multi_uav_ta_gym_env_b200.training import ReplayBuffer

    buf = ReplayBuffer(capacity=10, device="cpu")
    x = torch.arange(8, dtype=torch.float32)
    assert buf.push({"x": x, "y": x.view(8, 1) * 2}, rows=torch.tensor([1, 0, 1, 1, 0, 1, 1, 1], dtype=torch.uint8)) == 6
    assert len(buf) == 6
    buf.push({"x": x + 100, "y": (x + 100).view(8, 1) * 2})
    assert len(buf) == 10                        # wrapped around: the four oldest rows were overwritten
    s = buf.sample(64)
    assert s["x"].shape[0] == 10 and torch.equal(s["y"][:, 0], s["x"] * 2)
    kept = set(s["x"].tolist())
    assert {100.0 + i for i in range(8)} <= kept and len(kept & {0.0, 2.0, 3.0, 5.0}) == 0
