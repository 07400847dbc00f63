"""Development aid: per-phase cycle attribution of muav_step_kernel (needs libmuav_b200_phase.so built with
-DMUAV_PHASE_TIMING; swaps it in for this process only).  Build it next to the product library (after build()):

    F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -fmad=false"
    nvcc $F -DMUAV_PHASE_TIMING -c -o build/muav_kernels_phase.o multi_uav_ta_gym_env_b200/csrc/muav_kernels.cu
    nvcc -shared -o multi_uav_ta_gym_env_b200/libmuav_b200_phase.so build/muav_kernels_phase.o build/muav_scorer.o \
         build/muav_step_lean.o build/muav_step_lean_escort.o

and run with MUAV_NO_LEAN=1 (the counters live in the general instantiation).  Slots: 0 allocator, 1 event drain +
releaseAllTasks, 2 actions, 3 agent FSM, 4 distance / time penalty, 6 arrivals + escorts, 7 sensing, 8 reveal / expiry /
open scan, 9 rewards + slot recycling, 15 whole kernel (profiles/r01_step_kernel_ncu_v3.md has a measured breakdown)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_uav_ta_gym_env_b200 import _lib  # noqa: E402

os.environ.setdefault("MUAV_NO_LEAN", "1")
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
