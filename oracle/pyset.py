"""TEST INFRASTRUCTURE (oracle) -- not part of the product path.

CPython's `str` hash and `set` iteration order, restated, so that the one place where the reference depends on them
(CBBA.allocate_tasks builds `ordered_keys = list(remaining)` from a set of slot-key strings before it shuffles it,
TaskAllocation/MarketBased/CBBA.py:116,128) can be pinned.  With PYTHONHASHSEED=0 the interpreter's hash secret is all
zeros and the order is a pure function of the keys and of the set's insertion / removal history:

  * str hash  = SipHash-1-3 (Python/pyhash.c, `siphash13`, k0 = k1 = 0) over the string's UTF-8 / ASCII bytes,
                reinterpreted as a signed 64-bit integer, -1 mapped to -2;
  * set       = open addressing table (Objects/setobject.c, CPython 3.12): 8 slots initially, LINEAR_PROBES = 9,
                perturbation shift 5, growth when fill * 5 >= mask * 3 to the first power of two above 4 * used
                (2 * used above 50 000 entries); `discard` leaves a dummy that still counts as fill; iteration walks the
                table in slot order.

tests/test_pyset.py compares both with the running interpreter started under PYTHONHASHSEED=0.
"""
from __future__ import annotations

M64 = (1 << 64) - 1


def _rotl(x, b):
    return ((x << b) | (x >> (64 - b))) & M64


def siphash13(data: bytes, k0: int = 0, k1: int = 0) -> int:
    v0 = k0 ^ 0x736F6D6570736575
    v1 = k1 ^ 0x646F72616E646F6D
    v2 = k0 ^ 0x6C7967656E657261
    v3 = k1 ^ 0x7465646279746573

    def rnd():
        nonlocal v0, v1, v2, v3
        v0 = (v0 + v1) & M64
        v2 = (v2 + v3) & M64
        v1 = _rotl(v1, 13) ^ v0
        v3 = _rotl(v3, 16) ^ v2
        v0 = _rotl(v0, 32)
        v2 = (v2 + v1) & M64
        v0 = (v0 + v3) & M64
        v1 = _rotl(v1, 17) ^ v2
        v3 = _rotl(v3, 21) ^ v0
        v2 = _rotl(v2, 32)

    n = len(data)
    b = (n << 56) & M64
    i = 0
    while n - i >= 8:
        mi = int.from_bytes(data[i:i + 8], "little")
        v3 ^= mi
        rnd()
        v0 ^= mi
        i += 8
    t = int.from_bytes(data[i:], "little") if i < n else 0
    b |= t
    v3 ^= b
    rnd()
    v0 ^= b
    v2 ^= 0xFF
    rnd()
    rnd()
    rnd()
    return (v0 ^ v1) ^ (v2 ^ v3)


def str_hash(s: str) -> int:
    """hash(s) of CPython >= 3.11 under PYTHONHASHSEED=0 (as an unsigned 64-bit value; 0 for the empty string)."""
    data = s.encode("utf-8")
    if not data:
        return 0
    h = siphash13(data)
    if h == M64:      # (Py_hash_t)-1 is reserved
        h = M64 - 1
    return h


class PySet:
    """The part of CPython's set that CBBA uses: construction from a list, discard, `in`, iteration order."""

    MINSIZE = 8
    LINEAR_PROBES = 9
    PERTURB_SHIFT = 5
    _DUMMY = object()

    def __init__(self, keys=()):
        self.mask = self.MINSIZE - 1
        self.table = [None] * self.MINSIZE      # None = empty, _DUMMY = deleted, else (key, hash)
        self.fill = 0
        self.used = 0
        for k in keys:
            self.add(k)

    def __len__(self):
        return self.used

    def __bool__(self):
        return self.used > 0

    def _lookup(self, key, h):
        """Slot index of `key` or None (set_lookkey)."""
        mask = self.mask
        perturb = h
        i = h & mask
        while True:
            probes = self.LINEAR_PROBES if i + self.LINEAR_PROBES <= mask else 0
            for j in range(probes + 1):
                e = self.table[i + j]
                if e is None:
                    return None
                if e is not self._DUMMY and e[1] == h and e[0] == key:
                    return i + j
            perturb >>= self.PERTURB_SHIFT
            i = (i * 5 + 1 + perturb) & mask

    def __contains__(self, key):
        return self._lookup(key, str_hash(key)) is not None

    def _insert_clean(self, table, mask, entry):
        h = entry[1]
        perturb = h
        i = h & mask
        while True:
            if table[i] is None:
                table[i] = entry
                return
            if i + self.LINEAR_PROBES <= mask:
                for j in range(1, self.LINEAR_PROBES + 1):
                    if table[i + j] is None:
                        table[i + j] = entry
                        return
            perturb >>= self.PERTURB_SHIFT
            i = (i * 5 + 1 + perturb) & mask

    def _resize(self, minused):
        newsize = self.MINSIZE
        while newsize <= minused:
            newsize <<= 1
        old = self.table
        table = [None] * newsize
        mask = newsize - 1
        for e in old:
            if e is not None and e is not self._DUMMY:
                self._insert_clean(table, mask, e)
        self.table, self.mask = table, mask
        self.fill = self.used

    def add(self, key):
        """set_add_entry: the first free / dummy slot on the probe sequence unless the key is found first."""
        h = str_hash(key)
        mask = self.mask
        perturb = h
        i = h & mask
        freeslot = None
        while True:
            probes = self.LINEAR_PROBES if i + self.LINEAR_PROBES <= mask else 0
            found_unused = None
            for j in range(probes + 1):
                e = self.table[i + j]
                if e is None:
                    found_unused = i + j
                    break
                if e is self._DUMMY:
                    if freeslot is None:
                        freeslot = i + j
                elif e[1] == h and e[0] == key:
                    return
            if found_unused is not None:
                if freeslot is not None:
                    self.table[freeslot] = (key, h)     # reuse a dummy: fill unchanged
                    self.used += 1
                    return
                self.table[found_unused] = (key, h)
                self.fill += 1
                self.used += 1
                if self.fill * 5 >= mask * 3:
                    self._resize(self.used * 2 if self.used > 50000 else self.used * 4)
                return
            perturb >>= self.PERTURB_SHIFT
            i = (i * 5 + 1 + perturb) & mask

    def discard(self, key):
        i = self._lookup(key, str_hash(key))
        if i is not None:
            self.table[i] = self._DUMMY
            self.used -= 1

    def __iter__(self):
        for e in self.table:
            if e is not None and e is not self._DUMMY:
                yield e[0]
