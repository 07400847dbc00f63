"""Development aid: time / profile the fused Att-Pair scorer kernel on real WPS_hard tokens (full batch)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_uav_ta_gym_env_b200 import AllocSpec, BatchedMultiUAVEnv, wps_config  # noqa: E402
from multi_uav_ta_gym_env_b200.scorers import AttPairNet, FusedAttPairScorer  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = BatchedMultiUAVEnv(wps_config("WPS_hard"), E).reset(range(E))
env.step_allocated(AllocSpec.local_hungarian(20), n_steps=60)
tok = env.enable_fused_tokens(32, 16, 15, 0b111)
env.refresh_fused_tokens()
torch.manual_seed(0)
net = AttPairNet().cuda().eval()
fused = FusedAttPairScorer(net, torch.device("cuda"))
out = torch.zeros(E, 16, 32, device="cuda")
for _ in range(3):
    fused.score(tok, out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    fused.score(tok, out)
b.record()
torch.cuda.synchronize()
print("fused scorer full batch ms", a.elapsed_time(b) / 10, "valid tasks mean", float((tok["task_mask_u8"] == 0).sum(1).float().mean()))
