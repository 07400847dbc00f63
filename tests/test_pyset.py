"""oracle/pyset.py (CPython's str hash and set order under PYTHONHASHSEED=0) against the running interpreter started
with PYTHONHASHSEED=0: hashes of slot-key strings, and the iteration order of sets after the construction / discard
histories CBBA produces (CBBA.py:106-196: set(slot_keys), discard of committed keys, list(remaining) every round)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PROBE = r"""
import json, random, sys
rng = random.Random(int(sys.argv[1]))
out = {"hash": {}, "orders": []}
keys_all = [f"{t}#{c}{k}" for t in range(1, 400) for c in "rc" for k in range(4)]
for k in rng.sample(keys_all, 300) + ["", "a", "abcdefgh", "abcdefghi", "12#r0"]:
    out["hash"][k] = hash(k) & ((1 << 64) - 1)
for trial in range(300):
    n = rng.randint(1, 70)
    keys = rng.sample(keys_all, n)
    s = set(keys)
    hist = [list(s)]
    rem = list(keys)
    rng.shuffle(rem)
    while rem:
        for _ in range(rng.randint(1, 6)):
            if rem:
                s.discard(rem.pop())
        hist.append(list(s))
    out["orders"].append({"keys": keys, "removed": [k for k in keys], "hist": hist, "seed": trial})
json.dump(out, sys.stdout)
"""


def _cpython(seed):
    env = dict(os.environ, PYTHONHASHSEED="0")
    res = subprocess.run([sys.executable, "-c", PROBE, str(seed)], env=env, capture_output=True, text=True, check=True)
    return json.loads(res.stdout)


def test_str_hash_and_set_order_match_cpython_with_hashseed_zero():
    import random
    from oracle.pyset import PySet, str_hash

    for seed in (0, 1):
        ref = _cpython(seed)
        for k, h in ref["hash"].items():
            assert str_hash(k) == h, k
        # replay the same histories: the probe's removal order is reproducible from its seed
        rng = random.Random(seed)
        keys_all = [f"{t}#{c}{k}" for t in range(1, 400) for c in "rc" for k in range(4)]
        rng.sample(keys_all, 300)
        for tr in ref["orders"]:
            n = rng.randint(1, 70)
            keys = rng.sample(keys_all, n)
            assert keys == tr["keys"]
            s = PySet(keys)
            hist = [list(s)]
            rem = list(keys)
            rng.shuffle(rem)
            while rem:
                for _ in range(rng.randint(1, 6)):
                    if rem:
                        s.discard(rem.pop())
                hist.append(list(s))
            assert hist == tr["hist"], tr["seed"]
