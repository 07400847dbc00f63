"""Development aid: Att-Pair scorer kernel (FP32-pipe bound) overlapped with the step kernel (latency bound) by splitting
the batch into independent parts on side streams, replayed as CUDA graphs (two graphs: the launch-slot order buffers
alternate).    python tools/overlap_prof.py [envs] [parts] [graph 0/1] [flush 0/1]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_uav_ta_gym_env_b200 import AllocSpec, BatchedMultiUAVEnv, wps_config  # noqa: E402
from multi_uav_ta_gym_env_b200.scorers import AttPairNet, FusedAttPairScorer  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
parts = int(sys.argv[2]) if len(sys.argv) > 2 else 2
use_graph = (sys.argv[3] if len(sys.argv) > 3 else "1") == "1"
flush = (sys.argv[4] if len(sys.argv) > 4 else "0") == "1"
tc = int(os.environ["KB_TASK_CAP"]) if "KB_TASK_CAP" in os.environ else None
dev = torch.device("cuda:0")
cfg = wps_config("WPS_hard")
spec = AllocSpec.pair_hybrid(15)
torch.manual_seed(0)
net = AttPairNet().to(dev).eval()
scorer = FusedAttPairScorer(net, dev)
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


class Part:
    def __init__(self, lo, hi):
        self.stream = torch.cuda.Stream()
        self.env = BatchedMultiUAVEnv(cfg, hi - lo, device=dev, task_cap=tc).reset(range(lo, hi))
        self.scores = torch.zeros(hi - lo, 16, 32, dtype=torch.float32, device=dev)
        self.tok = self.env.enable_fused_tokens(32, 16, 15, 0b111)

    def work(self):
        scorer.score(self.tok, self.scores, use_need=True)
        self.env.step_allocated(spec, 1, edge_scores=self.scores)


bounds = [E * i // parts for i in range(parts + 1)]
ps = [Part(bounds[i], bounds[i + 1]) for i in range(parts)]
main = torch.cuda.Stream()


def step_all():
    cur = torch.cuda.current_stream()
    for p in ps:
        p.stream.wait_stream(cur)
        with torch.cuda.stream(p.stream):
            p.work()
    for p in ps:
        cur.wait_stream(p.stream)


def episode_start():
    for p in ps:
        p.env.restore()
        p.env.refresh_fused_tokens()


with torch.cuda.stream(main):
    episode_start()
    step_all()       # initialises the order buffers
    step_all()
    torch.cuda.synchronize()
    graphs = []
    if use_graph:
        for _ in range(2):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=main):
                step_all()
            graphs.append(g)

    def run(K):
        torch.cuda.synchronize()
        tot = 0.0
        evs = []
        for k in range(K):
            if k % 150 == 0:
                episode_start()
            if flush:
                flush_buf.fill_(k & 0xFF)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if use_graph:
                graphs[k & 1].replay()
            else:
                step_all()
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs) / K

    run(30)
    ms = run(600)
    err = max(int(p.env.error_flags().abs().max().item()) for p in ps)
print(f"envs {E} parts {parts} graph {use_graph} flush {flush} task_cap {ps[0].env.task_cap}: {ms:.4f} ms per step, "
      f"{E / ms / 1e3:.2f} M env-steps/s, err {err}")
