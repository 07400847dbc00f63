"""Importable stand-in for the reference's PyO3 module `core_sim` (core_sim/src/lib.rs:10-18): the one
function with behaviour, SimCore.avoid_obstacles (core_sim/src/sim_core.rs:24-59), evaluated by the CUDA
library through the C ABI (muav_avoid_obstacles).  Inside the step kernel the same code is a device function;
this entry point exists so that callers of the FFI keep working (`import ...core_sim as core_sim`)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class SimCore:
    def __init__(self, max_time_steps: int):
        self.time_steps = 0
        self.max_time_steps = max_time_steps

    @staticmethod
    def avoid_obstacles(agent_pos, obstacles, movement):
        """agent_pos [2], obstacles [[x, y, r], ...], movement [2] -> [ax, ay]."""
        out = SimCore.avoid_obstacles_batch([agent_pos], obstacles, [movement])
        return [float(out[0][0]), float(out[0][1])]

    @staticmethod
    def avoid_obstacles_batch(positions, obstacles, movements):
        lib = _lib.cuda_lib()
        dev = torch.device("cuda")
        pos = torch.as_tensor(positions, dtype=torch.float64, device=dev).reshape(-1, 2).contiguous()
        mv = torch.as_tensor(movements, dtype=torch.float64, device=dev).reshape(-1, 2).contiguous()
        n_obs = len(obstacles)
        ob = torch.as_tensor(obstacles, dtype=torch.float64, device=dev).reshape(-1, 3).contiguous() if n_obs else \
            torch.zeros(1, 3, dtype=torch.float64, device=dev)
        out = torch.zeros_like(pos)
        rc = lib.dll.muav_avoid_obstacles(pos.data_ptr(), mv.data_ptr(), ob.data_ptr(), n_obs, out.data_ptr(),
                                          pos.shape[0], C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "muav_avoid_obstacles")
        return out.cpu().tolist()
