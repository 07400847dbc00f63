"""Kernel micro-benchmark (development aid): times muav_step_kernel alone with CUDA events.
    python tools/kbench.py [case] [envs]
Prints per-launch time for (a) 150 single-step fused Local-Hungarian launches, L2 flushed between
launches, (b) the same without flush, (c) one 150-step resident launch."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_uav_ta_gym_env_b200 import AllocSpec, BatchedMultiUAVEnv, wps_config  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "WPS_hard"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
interval = 12 if case == "WPS_escort" else 20
if case.startswith("burst_x"):
    from multi_uav_ta_gym_env_b200 import burst_scaled_spec
    cfg = wps_config(burst_scaled_spec(int(case[7:])))
else:
    cfg = wps_config(case)
tc = int(os.environ["KB_TASK_CAP"]) if "KB_TASK_CAP" in os.environ else None
env = BatchedMultiUAVEnv(cfg, E, task_cap=tc).reset(range(E))
spec = AllocSpec(1, interval, 0x1F, True, False)
if len(sys.argv) > 3 and sys.argv[3] == "urgency_commit":
    spec = AllocSpec.urgency_commit(15)
if len(sys.argv) > 3 and sys.argv[3] == "urgency_coalition":
    spec = AllocSpec.urgency_coalition(12)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = {"case": case, "envs": E, "record_bytes": env.record_bytes, "agents": env.n_agents, "task_cap": env.task_cap}


def run(flush_l2, nsteps=150):
    env.restore()
    evs = []
    for t in range(nsteps):
        if flush_l2:
            flush.fill_(t & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        env.step_allocated(spec, 1)
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in evs]
    return sum(ms) / len(ms), ms


for _ in range(2):
    run(False, 20)
m, ms = run(True)
res["single_step_flush_ms"] = m
res["by_phase_ms"] = [sum(ms[i:i + 30]) / 30 for i in range(0, 150, 30)]
m, _ = run(False)
res["single_step_noflush_ms"] = m
env.restore()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
env.step_allocated(spec, 150)
b.record()
torch.cuda.synchronize()
res["resident_150_ms_per_step"] = a.elapsed_time(b) / 150
res["env_steps_per_s_single_flush"] = E / (res["single_step_flush_ms"] / 1e3)
res["env_steps_per_s_resident"] = E / (res["resident_150_ms_per_step"] / 1e3)
res["hbm_frac_single_flush"] = E * (2 * env.record_bytes + 22) / (res["single_step_flush_ms"] / 1e3) / 6553.9e9
assert int(env.error_flags().abs().max().item()) == 0
print(json.dumps(res))
