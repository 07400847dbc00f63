"""TEST INFRASTRUCTURE (oracle) -- not part of the product path.

Rectangular linear-sum-assignment, restating the algorithm behind
`scipy.optimize.linear_sum_assignment` (SciPy `_lsap` C++ module; Crouse 2016
shortest augmenting path).  The reference calls it at
TaskAllocation/OptimizationBased/HungarianAllocator.py:181 (import :8-11);
SciPy is unpinned there (README.md:30); the container has SciPy 1.18.1 and the
restatement is pinned by differential tests against it (tests/test_lsap_oracle.py)
-- exact equality of (row_ind, col_ind), ties included.
"""
from __future__ import annotations

INF = float("inf")


def lsap(cost, nr: int, nc: int):
    """cost: row-major list of nr*nc floats. Returns (row_ind, col_ind) lists."""
    transposed = nc < nr
    if transposed:
        ct = [0.0] * (nr * nc)
        for i in range(nr):
            for j in range(nc):
                ct[j * nr + i] = cost[i * nc + j]
        cost = ct
        nr, nc = nc, nr
    u = [0.0] * nr
    v = [0.0] * nc
    path = [-1] * nc
    col4row = [-1] * nr
    row4col = [-1] * nc
    for cur in range(nr):
        min_val = 0.0
        remaining = [nc - it - 1 for it in range(nc)]
        num_remaining = nc
        SR = [False] * nr
        SC = [False] * nc
        spc = [INF] * nc
        i = cur
        sink = -1
        while sink == -1:
            index = -1
            lowest = INF
            SR[i] = True
            for it in range(num_remaining):
                j = remaining[it]
                r = min_val + cost[i * nc + j] - u[i] - v[j]
                if r < spc[j]:
                    path[j] = i
                    spc[j] = r
                if spc[j] < lowest or (spc[j] == lowest and row4col[j] == -1):
                    lowest = spc[j]
                    index = it
            min_val = lowest
            if min_val == INF:
                raise ValueError("cost matrix is infeasible")
            j = remaining[index]
            if row4col[j] == -1:
                sink = j
            else:
                i = row4col[j]
            SC[j] = True
            num_remaining -= 1
            remaining[index] = remaining[num_remaining]
        u[cur] += min_val
        for i in range(nr):
            if SR[i] and i != cur:
                u[i] += min_val - spc[col4row[i]]
        for j in range(nc):
            if SC[j]:
                v[j] -= min_val - spc[j]
        j = sink
        while True:
            i = path[j]
            row4col[j] = i
            col4row[i], j = j, col4row[i]
            if i == cur:
                break
    if transposed:
        pairs = sorted((col4row[vv], vv) for vv in range(nr))
        return [p[0] for p in pairs], [p[1] for p in pairs]
    return list(range(nr)), list(col4row)
