"""TEST INFRASTRUCTURE (oracle) -- not part of the product path.

CPython `random.Random` draw rules restated over a tape of raw MT19937 32-bit
outputs.  The in-episode consumers in the reference are
  rndAgentGen : DroneEnv.py:1607 (random), :1620-1622 (uniform), :1813 (random)
  rndTgtGen   : DroneEnv.py:1650 (random), :1655 (choice of 2), :1661-1665, :1384-1385 (uniform)
  rndMissionGen: DroneEnv.py:1657 (choice of 3)
CPython rules (Lib/random.py, Modules/_randommodule.c):
  random()        = ((a >> 5) * 2**26 + (b >> 6)) / 2**53, two words a, b
  uniform(a, b)   = a + (b - a) * random()
  _randbelow(n)   : k = n.bit_length(); r = word >> (32 - k) until r < n   (k <= 32)
  choice(seq)     = seq[_randbelow(len(seq))]
"""
from __future__ import annotations

import random


def make_tape(gen: random.Random, n_words: int):
    """Raw 32-bit outputs the generator WOULD produce next (generator is not advanced)."""
    clone = random.Random()
    clone.setstate(gen.getstate())
    return [clone.getrandbits(32) for _ in range(n_words)]


class TapeRng:
    def __init__(self, words, cursor: int = 0):
        self.words = words
        self.cursor = cursor
        self.overflow = False

    def _word(self) -> int:
        if self.cursor >= len(self.words):
            self.overflow = True
            raise IndexError("RNG tape exhausted")
        w = self.words[self.cursor]
        self.cursor += 1
        return int(w)

    def random(self) -> float:
        a = self._word() >> 5
        b = self._word() >> 6
        return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0)

    def uniform(self, a: float, b: float) -> float:
        return a + (b - a) * self.random()

    def randbelow(self, n: int) -> int:
        k = n.bit_length()
        r = self._word() >> (32 - k)
        while r >= n:
            r = self._word() >> (32 - k)
        return r
