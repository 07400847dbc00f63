"""TEST INFRASTRUCTURE (oracle) -- not part of the product path.

Bit-exact float64 primitives that mirror what NumPy does inside the reference
(SURVEY.md Appendix F):

  norm2(x, y)      == np.linalg.norm(np.array([x, y]))          (1-D: BLAS ddot -> FMA on 2nd term)
                      used at mUAV_TA/DroneEnv.py:859-860,992,1015,1056,1116,1220-1222,1521,1701,1738,1760,
                      DroneEnvComponents.py:64, HungarianAllocator.py:59
  norm2_rows(x, y) == np.linalg.norm(M, axis=1)[row]            (ufunc mul + add.reduce, no FMA)
                      used only at DroneEnv.py:1135
  np_sum(v)        == np.sum(v) for a contiguous float64 vector, n <= 128
                      (NumPy pairwise_sum: 8 running sums for n >= 8); DroneEnv.py:1138
"""
from __future__ import annotations

import ctypes
import ctypes.util
import math

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.fma.restype = ctypes.c_double
_libm.fma.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_double]


def fma(a: float, b: float, c: float) -> float:
    return _libm.fma(a, b, c)


def norm2(x: float, y: float) -> float:
    return math.sqrt(fma(y, y, x * x))


def norm2_rows(x: float, y: float) -> float:
    return math.sqrt(x * x + y * y)


def _pairwise(a):
    n = len(a)
    if n < 8:
        res = 0.0
        for v in a:
            res += v
        return res
    # n <= 128 (PW_BLOCKSIZE): 8 running sums, then the tail
    r = list(a[:8])
    i = 8
    lim = n - (n % 8)
    while i < lim:
        for j in range(8):
            r[j] += a[i + j]
        i += 8
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    while i < n:
        res += a[i]
        i += 1
    return res


def np_sum(v) -> float:
    v = [float(x) for x in v]
    if not v:
        return 0.0
    if len(v) > 128:
        raise ValueError("np_sum restatement covers n <= 128")
    return _pairwise(v)
