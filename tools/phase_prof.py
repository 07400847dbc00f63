"""Development aid: per-phase cycle attribution of muav_step_kernel (needs libmuav_b200_phase.so built with
-DMUAV_PHASE_TIMING; swaps it in for this process only)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_uav_ta_gym_env_b200 import _lib  # noqa: E402

_lib.CUDA_LIB_PATH = os.path.join(_lib.PKG_DIR, "libmuav_b200_phase.so")
import torch  # noqa: E402
from multi_uav_ta_gym_env_b200 import AllocSpec, BatchedMultiUAVEnv, wps_config  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "WPS_hard"
E = 4096
env = BatchedMultiUAVEnv(wps_config(case), E).reset(range(E))
spec = AllocSpec.local_hungarian(20)
for _ in range(150):
    env.step_allocated(spec, 1)
torch.cuda.synchronize()
os.environ["MUAV_PHASE_DUMP"] = "1"
env.step_allocated(spec, 0) if False else None
# one more launch triggers the dump of the accumulated counters (then resets them)
env.restore()
env.step_allocated(spec, 1)
torch.cuda.synchronize()
