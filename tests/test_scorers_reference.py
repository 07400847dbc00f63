"""multi_uav_ta_gym_env_b200.scorers against the reference network classes (authoring container only): the same
torch seed followed by construction gives identical parameters, and the forward passes agree exactly on CPU.
Pins AttPairNet / MLPPairNet (PairCostHybrid.py:89-196), AttCommitNet / MLPCommitNet (AttentionCommit.py:68-129) and
AttCoalitionNet / MLPCoalitionNet (AttentionEscort.py:244-375)."""
import numpy as np
import pytest
import torch

import refshim

pytestmark = pytest.mark.skipif(not refshim.reference_available(), reason="reference tree not present")


def _pairs():
    refshim.install()
    from TaskAllocation.Hybrid import AttentionCommit as RC
    from TaskAllocation.Hybrid import ContextPairHybrid as RX
    from TaskAllocation.Hybrid import GNNPairHybrid as RG
    from TaskAllocation.Hybrid import AttentionEscort as RE
    from TaskAllocation.Hybrid import PairCostHybrid as RP
    from multi_uav_ta_gym_env_b200 import scorers as S

    return [
        ("att_pair", RP.AttPairNet, S.AttPairNet, (32, 13), (16, 12)),
        ("mlp_pair", RP.MLPPairNet, S.MLPPairNet, (32, 13), (16, 12)),
        ("att_commit", RC.AttCommitNet, S.AttCommitNet, (32, 13), (16, 13)),
        ("mlp_commit", RC.MLPCommitNet, S.MLPCommitNet, (32, 13), (16, 13)),
        ("att_coalition", RE.AttCoalitionNet, S.AttCoalitionNet, (48, 22), (16, 16)),
        ("mlp_coalition", RE.MLPCoalitionNet, S.MLPCoalitionNet, (48, 22), (16, 16)),
        ("att_context", RX.AttContextPairNet, S.AttContextPairNet, (32, 13), (16, 12)),
        ("mlp_context", RX.MLPContextPairNet, S.MLPContextPairNet, (32, 13), (16, 12)),
        ("gnn_context", RG.GNNContextPairNet, S.GNNContextPairNet, (32, 13), (16, 12)),
    ]


@pytest.mark.parametrize("idx", range(9))
def test_same_seed_same_parameters_same_forward(idx):
    name, Ref, Mine, tshape, ashape = _pairs()[idx]
    torch.manual_seed(7)
    ref = Ref().eval()
    torch.manual_seed(7)
    mine = Mine().eval()
    rs, ms = ref.state_dict(), mine.state_dict()
    assert list(rs.keys()) == list(ms.keys()), name
    for k in rs:
        assert torch.equal(rs[k], ms[k]), (name, k)
    g = torch.Generator().manual_seed(3)
    B = 5
    tf = torch.rand(B, *tshape, generator=g)
    af = torch.rand(B, *ashape, generator=g)
    tm = torch.zeros(B, tshape[0], dtype=torch.bool)
    am = torch.zeros(B, ashape[0], dtype=torch.bool)
    for b in range(B):
        tm[b, 5 + 3 * b:] = True
        am[b, 6 + b:] = True
    extra = (torch.rand(B, 8, generator=g),) if "context" in name else ()
    if name == "gnn_context":
        extra += ((torch.rand(B, ashape[0], tshape[0], generator=g) > 0.4).float(),)
    with torch.no_grad():
        ro = ref(tf, tm, af, am, *extra)
        mo = mine(tf, tm, af, am, *extra)
    for r, m in zip(ro, mo):
        assert torch.equal(r, m), name


def test_coalition_scores_follow_act():
    """scorers.coalition_scores == AttentionEscort.act(explore=False) (AttentionEscort.py:449-466)."""
    refshim.install()
    from TaskAllocation.Hybrid.AttentionEscort import AttentionEscort
    from multi_uav_ta_gym_env_b200 import scorers as S

    torch.manual_seed(11)
    pol = AttentionEscort(device="cpu", use_attention=False)
    g = torch.Generator().manual_seed(5)
    tok = {"task_feats": torch.rand(48, 22, generator=g).numpy(), "agent_feats": torch.rand(16, 16, generator=g).numpy(),
           "task_mask": np.arange(48) >= 20, "agent_mask": np.arange(16) >= 9,
           "edge_valid": (torch.rand(16, 48, generator=g) > 0.3).float().numpy()}
    want, _, _ = pol.act(tok, explore=False)
    mine = S.MLPCoalitionNet(hidden=max(128, 128 * 2))
    mine.load_state_dict(pol.net.state_dict())
    tt = {k: torch.as_tensor(v)[None] for k, v in tok.items()}
    got = S.coalition_scores(mine.eval(), tt)[0].numpy()
    assert np.allclose(got, want, atol=1e-6, rtol=0)


def test_registered_scenarios_match_the_reference():
    """config.CASE_SPECS / wps_config against experiments/paper_scenarios.CASE_SPECS + paper_eval.make_config."""
    refshim.install()
    from experiments.paper_scenarios import CASE_SPECS as REF
    from multi_uav_ta_gym_env_b200 import config as mine

    assert set(REF) <= set(mine.CASE_SPECS)
    for case in REF:
        want = {k: v for k, v in REF[case].items() if k != "label"}
        got = {k: v for k, v in mine.CASE_SPECS[case].items() if k != "label"}
        assert want == got, case
        a, b = refshim.wps_config(case), mine.wps_config(case)
        da = {k: v for k, v in vars(a).items() if not k.startswith("_")}
        db = {k: v for k, v in vars(b).items() if not k.startswith("_")}
        assert da == db, (case, {k: (da.get(k), db.get(k)) for k in set(da) | set(db) if da.get(k) != db.get(k)})


def test_training_masks_follow_the_reference_trainers():
    """oracle.tokens.pair_mask (the checker of the batched collectors) against _expert_mask of
    experiments/train_pair_cost.py:53-70 and PairCostHybrid._selected_mask (:293-306) on a live reference episode with
    the Global-Hungarian teacher."""
    import sys
    import types

    refshim.install()
    for name in ("tianshou", "tianshou.data", "TaskAllocation.RL_Policies.Tianshou_Policy"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["tianshou"].__path__ = []
    sys.modules["tianshou.data"].Batch = dict
    sys.modules["TaskAllocation.RL_Policies.Tianshou_Policy"]._get_model = lambda *a, **k: None
    import experiments.train_pair_cost as T
    from mUAV_TA.DroneEnv import MultiUAVEnv as RefEnv
    from TaskAllocation.Hybrid.PairCostHybrid import PairCostHybrid
    from TaskAllocation.OptimizationBased.HungarianAllocator import HungarianAllocator as RefHung
    from oracle import tokens as otok
    from oracle.sim import OracleEnv

    cfg = refshim.wps_config("WPS_hard")
    ref = RefEnv(cfg)
    obs, info = ref.reset(seed=9)
    orc = OracleEnv(cfg).reset(9)
    hung = RefHung(20, ref.max_coord)
    pol = PairCostHybrid(use_attention=False, device="cpu")
    n = 0
    for t in range(150):
        events = list(info.get("events") or []) if isinstance(info, dict) else []
        actions = {}
        if T._should_replan(ref, events):
            expert = hung.allocate_tasks(ref.get_live_agents(), T._open_tasks(ref), time_step=ref.time_steps, events=events,
                                         force=True)
            tok = pol.build_tokens(ref)
            otk = otok.build_pair_tokens(orc, 32, 16)
            pairs = [(ref.agent_by_name[nm].id, task.id) for nm, task in expert]
            assert np.array_equal(T._expert_mask(tok, expert), otok.pair_mask(otk, pairs, True)), t
            assert np.array_equal(pol._selected_mask(tok, expert), otok.pair_mask(otk, pairs, False)), t
            n += int(T._expert_mask(tok, expert).sum() > 0)
            actions = T._apply_assign(ref, expert)
        obs, _, _, _, info = ref.step(actions)
        orc.step([(ref.agent_by_name[nm].id, i) for nm, i in actions.items()])
    assert n > 5
