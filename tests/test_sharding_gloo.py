"""N > 1 host logic on CPU: two gloo ranks each run their shard of environments (CPU build of the kernel
core) and all-reduce the episode metric vector; the result must equal the single-process total."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard_metrics(rank, per_rank):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ctypes as C
    from helpers import HostCheck, alloc_opts_for
    from multi_uav_ta_gym_env_b200 import sharding, wps_config

    hc = HostCheck()
    env = hc.make(wps_config("WPS_hard"), list(sharding.shard_range(per_rank, rank)))
    env.step_alloc(alloc_opts_for("local_hungarian"), n_steps=150)
    d = hc.lib.dll
    d.hostcheck_metrics.restype = C.c_int
    d.hostcheck_metrics.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    m = np.zeros((per_rank, 30))
    assert d.hostcheck_metrics(C.byref(env.cfg), env.rec.ctypes.data, m.ctypes.data, per_rank) == 0
    d.muav_metric_name.restype = C.c_char_p
    d.muav_metric_name.argtypes = [C.c_int]
    names = [d.muav_metric_name(i).decode() for i in range(30)]
    return sharding.metric_vector(torch.from_numpy(m), names)


def _worker(rank, world, port, per_rank, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multi_uav_ta_gym_env_b200 import sharding

    v = _shard_metrics(rank, per_rank)
    v = sharding.allreduce_metric_vector(v)
    dist.barrier()
    if rank == 0:
        q.put(v.numpy().tolist())
    dist.destroy_process_group()


def test_two_rank_metric_allreduce_equals_single_process():
    per_rank, world = 6, 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, _free_port_once(), per_rank, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    got = np.array(q.get())
    want = sum(_shard_metrics(r, per_rank) for r in range(world)).numpy()
    assert np.array_equal(got, want)
    from multi_uav_ta_gym_env_b200 import sharding

    s = sharding.summarize(torch.from_numpy(got))
    assert s["episodes"] == 12 and -400 < s["mean_S_WPS"] < 0


_PORT = []


def _free_port_once():
    if not _PORT:
        _PORT.append(_free_port())
    return _PORT[0]
