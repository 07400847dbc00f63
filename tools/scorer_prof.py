"""Development aid: phase timing of the Att-Pair scorer forward on real WPS_hard tokens."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_uav_ta_gym_env_b200 import AllocSpec, BatchedMultiUAVEnv, wps_config  # noqa: E402
from multi_uav_ta_gym_env_b200.scorers import AttPairNet, GraphedPairScorer, pair_scores, pair_scores_fast  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = BatchedMultiUAVEnv(wps_config("WPS_hard"), E).reset(range(E))
env.step_allocated(AllocSpec.local_hungarian(20), n_steps=60)
tok = env.enable_fused_tokens(32, 16, 15, 0b111)
env.refresh_fused_tokens()
torch.manual_seed(0)
net = AttPairNet().cuda().eval()
net.self_encoder.use_nested_tensor = False
t = {"task_feats": tok["task_feats"], "task_mask": tok["task_mask_u8"].bool(), "agent_feats": tok["agent_feats"],
     "agent_mask": tok["agent_mask_u8"].bool(), "edge_valid": tok["edge_valid"]}


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


with torch.no_grad():
    print("eager reference-shaped forward  ms", timeit(lambda: pair_scores(net, t)))
    print("eager split-head forward        ms", timeit(lambda: pair_scores_fast(net, t)))

    def enc():
        t_emb = net.task_proj(t["task_feats"]) + net.type_embed.weight[1]
        a_emb = net.agent_proj(t["agent_feats"]) + net.type_embed.weight[0]
        tokens = torch.cat([a_emb, t_emb], dim=1)
        pad = torch.cat([t["agent_mask"], t["task_mask"]], dim=1)
        return net.self_encoder(tokens, src_key_padding_mask=pad)

    print("  proj + self encoder           ms", timeit(enc))
    h = enc()
    a_h, t_h = h[:, :16], h[:, 16:]

    def cross():
        a_ctx, _ = net.cross_a2t(a_h, t_h, t_h, key_padding_mask=t["task_mask"], need_weights=False)
        t_ctx, _ = net.cross_t2a(t_h, a_h, a_h, key_padding_mask=t["agent_mask"], need_weights=False)
        return a_h + a_ctx, t_h + t_ctx

    print("  cross attention x2            ms", timeit(cross))
    a2, t2 = cross()
    l1, l2, l3 = net.pair_head[0], net.pair_head[2], net.pair_head[4]

    def head():
        d = 64
        wa, wt, wat = l1.weight[:, :d], l1.weight[:, d:2 * d], l1.weight[:, 2 * d:]
        ha = a2 @ wa.t()
        ht = t2 @ wt.t() + l1.bias
        prod = a2.unsqueeze(2) * t2.unsqueeze(1)
        h1 = torch.relu_(prod @ wat.t() + ha.unsqueeze(2) + ht.unsqueeze(1))
        h2 = torch.relu_(l2(h1))
        return l3(h2).squeeze(-1)

    print("  split pair head               ms", timeit(head))
    sc = GraphedPairScorer(net, E, torch.device("cuda"), buckets=[512, 1024])
    out = torch.zeros(E, 16, 32, device="cuda")
    print("graph replay full batch         ms", timeit(lambda: sc.score_all(tok, out)))
    for n in (256, 512, 1024):
        idx = torch.arange(n, device="cuda")
        print(f"graph replay subset {n:5d}       ms", timeit(lambda: sc.score_subset(tok, idx, out)))
    need = tok["need"]
    print("need fraction after refresh", float(need.float().mean()))
