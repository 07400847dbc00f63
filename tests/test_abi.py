"""The C ABI library loads without a GPU and exports every symbol include/muav.h declares."""
import ctypes as C
import os
import re

import numpy as np

from multi_uav_ta_gym_env_b200 import _lib, config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_header_symbols():
    import __graft_entry__ as ge

    ge.build_cuda()
    lib = _lib.cuda_lib()
    hdr = open(os.path.join(ROOT, "include", "muav.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|void|size_t|const char\*)\s+(muav_[a-z_]+)\s*\(", hdr, flags=re.M))
    assert declared == set(_lib.ABI_SYMBOLS)
    for sym in declared:
        assert hasattr(lib.dll, sym), sym
    assert lib.dll.muav_config_size() == C.sizeof(_lib.MuavConfig)
    assert lib.dll.muav_version().startswith(b"muav_b200")
    assert lib.metric_names()[4] == "S_WPS"


def test_layout_is_consistent():
    lib = _lib.cuda_lib()
    for case in ("WPS_hard", "WPS_commit", "WPS_escort", "WPS_attn_XL"):
        cfg = _lib.build_config(config.wps_config(case))
        F = lib.fields(cfg)
        rb = lib.record_bytes(cfg)
        assert rb % 16 == 0
        spans = sorted((off, off + cnt * dt.itemsize, dt.itemsize) for off, cnt, dt in F.values())
        for (a0, a1, sz), (b0, _, _) in zip(spans, spans[1:]):
            assert a1 <= b0 and a0 % sz == 0
        assert spans[-1][1] <= rb
        assert lib.scratch_bytes(cfg) >= 8 * cfg.n_agents * cfg.task_cap
    assert lib.header_index("T") == 0
    assert lib.header_index("F_REWARD") == 0


def test_bad_arguments_are_rejected_without_gpu():
    lib = _lib.cuda_lib()
    cfg = _lib.build_config(config.wps_config("WPS_hard"))
    assert lib.dll.muav_step(None, None, None, None, None, None, None, 1, 1, None) == -22
    assert lib.dll.muav_step(C.byref(cfg), None, None, None, None, None, None, 1, 1, None) == -22
    bad = _lib.build_config(config.wps_config("WPS_hard"))
    bad.n_agents = 1000
    assert lib.dll.muav_metrics(C.byref(bad), None, None, 1, None) == -22
    assert lib.dll.muav_lsap(None, None, None, 4, 4, None, 1, None) == -22


def test_product_has_no_cpu_fallback():
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from multi_uav_ta_gym_env_b200 import BatchedMultiUAVEnv

    with pytest.raises(RuntimeError):
        BatchedMultiUAVEnv(config.wps_config("WPS_hard"), 2)
