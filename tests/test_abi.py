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


def _header_structs():
    """typedef struct NAME { ... } NAME; blocks of include/muav.h -> {name: [(field, array length or 0, is pointer)]}"""
    hdr = open(os.path.join(ROOT, "include", "muav.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    out = {}
    for m in re.finditer(r"typedef struct (\w+) \{(.*?)\} \1;", hdr, flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            head, *rest = decl.split(",")
            head = re.sub(r"\[[^\]]*\]", "[]", head)   # array extents may be expressions
            toks = head.replace("*", " * ").split()
            ptr = "*" in toks
            names = [toks[-1]] + [re.sub(r"\[[^\]]*\]", "[]", r).replace("*", " ").strip() for r in rest]
            for n in names:
                fields.append((n.split("[")[0], 1 if "[" in n else 0, ptr))
        out[m.group(1)] = fields
    return out


def test_ctypes_mirrors_follow_the_header_structs():
    """Field names and order of every ctypes mirror in _lib.py equal the struct in include/muav.h (pointers are c_void_p,
    arrays are ctypes arrays): a field added on one side only would silently shift everything behind it."""
    structs = _header_structs()
    mirrors = {"muav_token_out": _lib.MuavTokenOut, "muav_attpair_offsets": _lib.MuavAttPairOffsets,
               "muav_attcommit_offsets": _lib.MuavAttCommitOffsets, "muav_attcoal_offsets": _lib.MuavAttCoalOffsets,
               "muav_alloc_opts": _lib.MuavAllocOpts, "muav_step_out": _lib.MuavStepOut, "muav_config": _lib.MuavConfig}
    for name, cls in mirrors.items():
        assert name in structs, name
        want = structs[name]
        got = list(cls._fields_)
        assert [f[0] for f in got] == [w[0] for w in want], (name, [f[0] for f in got], [w[0] for w in want])
        for (fname, ftype), (_, alen, ptr) in zip(got, want):
            if ptr:
                assert ftype is C.c_void_p, (name, fname)
            if alen:
                assert issubclass(ftype, C.Array), (name, fname)
