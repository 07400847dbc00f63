"""oracle/lsap.py and the kernel's LSAP (CPU build) against SciPy's linear_sum_assignment, the
third-party solver the reference calls (HungarianAllocator.py:181): exact (row_ind, col_ind)."""
import numpy as np
import pytest

from oracle.lsap import lsap

scipy_opt = pytest.importorskip("scipy.optimize")


def corpus(n, seed=0, rmax=15, cmax=26):
    rng = np.random.default_rng(seed)
    out = []
    for it in range(n):
        nr = int(rng.integers(1, rmax))
        nc = int(rng.integers(1, cmax))
        kind = it % 7
        if kind == 0:
            c = rng.uniform(0, 1, (nr, nc))
        elif kind == 1:
            c = rng.integers(0, 4, (nr, nc)).astype(float)
        elif kind == 2:
            c = rng.uniform(-1, 1, (nr, nc))
            c[rng.random((nr, nc)) < 0.5] = 1e6
        elif kind == 3:
            c = rng.uniform(0, 1, (nr, nc))
            if nr > 1:
                c[rng.integers(0, nr)] = c[0]
        elif kind == 4:
            c = np.full((nr, nc), 1e6)
        elif kind == 5:
            c = np.round(rng.uniform(0, 1, (nr, nc)), 1)
        else:
            c = np.zeros((nr, nc))
        out.append(c)
    return out


def test_oracle_lsap_matches_scipy():
    for c in corpus(1500):
        r, cc = scipy_opt.linear_sum_assignment(c)
        r2, c2 = lsap(c.ravel().tolist(), *c.shape)
        assert list(r) == r2 and list(cc) == c2


def test_probes_from_survey():
    assert lsap([1.0] * 6, 2, 3) == ([0, 1], [0, 1])
    assert lsap([1.0] * 6, 3, 2) == ([0, 1], [0, 1])
    c = np.array([[5, 9], [1, 2], [7, 7], [1, 2]], float)
    r, cc = scipy_opt.linear_sum_assignment(c)
    assert (list(r), list(cc)) == tuple(lsap(c.ravel().tolist(), 4, 2))


def test_kernel_lsap_cpu_build_matches_scipy(hostcheck):
    mats = corpus(1500, seed=1)
    B = len(mats)
    cost = np.zeros((B, 14, 25))
    nr = np.zeros(B, np.int32)
    nc = np.zeros(B, np.int32)
    for b, c in enumerate(mats):
        nr[b], nc[b] = c.shape
        cost[b, : c.shape[0], : c.shape[1]] = c
    got = hostcheck.lsap(cost, nr, nc)
    for b, c in enumerate(mats):
        r, cc = scipy_opt.linear_sum_assignment(c)
        want = np.full(14, -1)
        want[r] = cc
        assert list(got[b]) == list(want), b


def test_performance_impact_slot_key_order_is_the_string_order(hostcheck):
    """PerformanceImpact breaks ties between equal (IPI, agent) candidates by the slot key STRING "<id>#c<k>" / "<id>#r<k>"
    (MarketBased/PerformanceImpact.py:141-143 compares tuples that end in the key; keys from CBBA.py:46-65): the device
    allocator's integer rule must order task ids exactly like Python orders those strings."""
    import ctypes as C

    f = hostcheck.lib.dll.hostcheck_slot_id_less
    f.restype, f.argtypes = C.c_int, [C.c_int, C.c_int]
    ids = list(range(1, 130)) + [199, 200, 201, 999, 1000, 1001, 1099, 1100, 1999, 2000, 2047]
    for x in ids:
        for y in ids:
            if x != y:
                assert bool(f(x, y)) == (f"{x}#c0" < f"{y}#c0"), (x, y)
                assert (f"{x}#c0" < f"{y}#r1") == (f"{x}#" < f"{y}#")   # the id part decides between different tasks
