"""Development aid: host-side (enqueue) cost per call of the hot-path entry points, with a cProfile breakdown.
python tools/host_overhead.py"""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_uav_ta_gym_env_b200 import AllocSpec, BatchedMultiUAVEnv, wps_config  # noqa: E402
from multi_uav_ta_gym_env_b200.scorers import AttPairNet, FusedAttPairScorer  # noqa: E402

E = 4096
dev = torch.device("cuda:0")
env = BatchedMultiUAVEnv(wps_config("WPS_hard"), E, device=dev, task_cap=32).reset(range(E))
spec = AllocSpec.pair_hybrid(15)
scores = torch.zeros(E, 16, 32, device=dev)
tok = env.enable_fused_tokens(32, 16, 15, 0b111)
env.refresh_fused_tokens()
scorer = FusedAttPairScorer(AttPairNet().to(dev).eval(), dev)
A = env.n_agents
h_act = torch.empty(E, A, 2, dtype=torch.int32).pin_memory()
h_rew = torch.empty(E, dtype=torch.float64).pin_memory()
h_term = torch.empty(E, dtype=torch.uint8).pin_memory()
h_trunc = torch.empty(E, dtype=torch.uint8).pin_memory()


def dev_step():
    scorer.score(tok, scores, use_need=True)
    env.step_allocated(spec, 1, edge_scores=scores)


def host_step():
    scorer.score(tok, scores, use_need=True)
    env.allocate_host(spec, h_act, edge_scores=scores)
    env.step_host(h_act, h_rew, h_term, h_trunc, 1, hint=spec)


for name, fn, n in (("device loop (enqueue only)", dev_step, 100), ("host-buffer loop (synchronous)", host_step, 100)):
    env.restore() if hasattr(env, "_saved") else None
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name}: {1e6 * (t1 - t0) / n:.1f} us of host time per step (+ {1e6 * (t2 - t1) / n:.1f} us drain)")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(n):
        fn()
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(18)
