"""Development aid: plain Local-Hungarian rollouts with several steps per launch (muav_rollout keeps the state of an
environment in shared memory across the steps of one launch).  python tools/rollout_prof.py [n_envs]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_uav_ta_gym_env_b200 import AllocSpec, BatchedMultiUAVEnv, wps_config  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = BatchedMultiUAVEnv(wps_config("WPS_hard"), E, task_cap=32).reset(range(E))
spec = AllocSpec.local_hungarian(20)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for n in (1, 2, 5, 15, 50, 150):
    tot = 0.0
    reps = 3
    for r in range(reps):
        env.restore()
        env.step_allocated(spec, 1)      # warm
        env.restore()
        torch.cuda.synchronize()
        ms = 0.0
        done = 0
        while done < 150:
            k = min(n, 150 - done)
            flush.fill_(done & 0xFF)      # cold L2 at the start of every launch, as in bench.py
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            env.step_allocated(spec, k)
            b.record()
            torch.cuda.synchronize()
            ms += a.elapsed_time(b)
            done += k
        tot += ms
    ms = tot / reps
    print(f"{n:4d} steps per launch: episode of 150 steps in {ms:.3f} ms = {E * 150 / ms / 1e3:.2f} M env-steps/s "
          f"({E * 150 * 8 / ms / 1e3:.0f} M agent-steps/s)", flush=True)
assert int(env.error_flags().abs().max().item()) == 0
