import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from multi_uav_ta_gym_env_b200 import AllocSpec, BatchedMultiUAVEnv, wps_config
from multi_uav_ta_gym_env_b200.scorers import AttPairNet, FusedAttPairScorer
dev = torch.device("cuda:0")
E = 4096
cfg = wps_config("WPS_hard")
spec = AllocSpec.pair_hybrid(15)
for order in (True, False):
    env = BatchedMultiUAVEnv(cfg, E, device=dev).reset(range(E))
    env.group_replanners = order
    torch.manual_seed(0)
    net = AttPairNet().to(dev).eval()
    scorer = FusedAttPairScorer(net, dev)
    scores = torch.zeros(E, 16, 32, dtype=torch.float32, device=dev)
    tok = env.enable_fused_tokens(32, 16, 15, 0b111)
    for ep in range(3):
        env.restore(); env.refresh_fused_tokens()
        for t in range(150):
            scorer.score(tok, scores, use_need=True)
            env.step_allocated(spec, 1, edge_scores=scores)
            if t % 10 == 9 or t == 149:
                ef = env.error_flags()
                nz = (ef != 0).nonzero().flatten()
                if nz.numel():
                    print("order", order, "ep", ep, "t", t, "n_err", nz.numel(), "envs", nz[:8].tolist(), "flags", ef[nz[:8]].tolist())
                    break
    print("order", order, "done; T", env.header_int("T")[:4].tolist())
