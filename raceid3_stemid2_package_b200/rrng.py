"""R's default random number stream, for the host side of `clustexp` (reference R/RaceID_utils.R:111-112,132-133:
`set.seed(rseed)`, `sample(...)`, and `fpc::clusterboot(seed=rseed)`'s `sample(n, n, replace=TRUE)`).

The GPU library draws no random numbers; it only ever receives index vectors.  To stay a drop-in for
`clustexp(rseed=...)` the host must therefore produce the index vectors R would: Mersenne-Twister seeded through
R's LCG scrambling (RNG.c: RNG_Init / FixupSeeds / MT_genrand) and, for R >= 3.6, `sample()` by rejection sampling
on ceil(log2(n)) random bits taken 16 at a time (R_unif_index / rbits).

PARITY UNPINNED against a live R (none in this image); pinned by R's well-known known-answer values in
tests/test_host_and_lib.py::test_rrng_known_answers (set.seed(42); sample(1:10) etc.).
"""
from __future__ import annotations

import math

import numpy as np

_I2_32M1 = 2.328306437080797e-10  # = 1/(2^32 - 1), RNG.c fixup()
_MT_SCALE = 2.3283064365386963e-10  # = 2^-32, MT_genrand()


class RRng:
    def __init__(self, seed: int | None = None):
        self._bg = np.random.MT19937()
        if seed is not None:
            self.set_seed(seed)

    def set_seed(self, seed: int) -> None:
        s = np.uint32(int(seed) & 0xFFFFFFFF)
        a, one = np.uint32(69069), np.uint32(1)
        with np.errstate(over="ignore"):
            for _ in range(50):  # RNG_Init: initial scrambling
                s = a * s + one
            key = np.empty(625, dtype=np.uint32)
            for j in range(625):
                s = a * s + one
                key[j] = s
        # FixupSeeds: i_seed[0] is mti and is forced to N=624; i_seed[1..624] is the MT state
        self._bg.state = {"bit_generator": "MT19937", "state": {"key": key[1:].copy(), "pos": 624}}

    def raw(self, n: int) -> np.ndarray:
        return self._bg.random_raw(n).astype(np.uint64)

    def unif_rand(self, n: int = 1) -> np.ndarray:
        u = self.raw(n).astype(np.float64) * _MT_SCALE
        u[u <= 0.0] = 0.5 * _I2_32M1
        u[(1.0 - u) <= 0.0] = 1.0 - 0.5 * _I2_32M1
        return u

    # R_unif_index for a whole vector: each attempt consumes a fixed number of 32-bit outputs, so attempts can be
    # generated in bulk and accepted in stream order.
    def _unif_index(self, dn: int, count: int) -> np.ndarray:
        if dn <= 0:
            return np.zeros(count, dtype=np.int64)
        bits = int(math.ceil(math.log2(dn)))
        words = bits // 16 + 1
        mask = (1 << bits) - 1
        out = np.empty(count, dtype=np.int64)
        got = 0
        state0 = None
        while got < count:
            need = count - got
            attempts = int(need * (2.0 ** bits / dn) * 1.1) + 16
            state0 = self._bg.state
            r = self.raw(attempts * words).reshape(attempts, words) >> np.uint64(16)  # floor(unif*65536)
            v = np.zeros(attempts, dtype=np.uint64)
            for w in range(words):
                v = v * np.uint64(65536) + r[:, w]
            v &= np.uint64(mask)
            ok = np.nonzero(v < np.uint64(dn))[0]
            if len(ok) >= need:
                used_attempts = int(ok[need - 1]) + 1
                out[got:] = v[ok[:need]].astype(np.int64)
                # rewind: only `used_attempts` attempts were really consumed
                self._bg.state = state0
                if used_attempts * words:
                    self.raw(used_attempts * words)
                got = count
            else:
                out[got:got + len(ok)] = v[ok].astype(np.int64)
                got += len(ok)
        return out

    def sample_int(self, n: int, size: int | None = None, replace: bool = False) -> np.ndarray:
        """R's sample.int(n, size, replace); returns 1-based indices like R."""
        if size is None:
            size = n
        if replace or size < 2:
            return self._unif_index(n, size) + 1
        if size > n:
            raise ValueError("cannot take a sample larger than the population when 'replace = FALSE'")
        # do_sample without replacement: partial Fisher-Yates with a shrinking population; the bit width follows
        # the shrinking n, so draw one index at a time.
        x = np.arange(n, dtype=np.int64)
        y = np.empty(size, dtype=np.int64)
        nn = n
        for i in range(size):
            j = int(self._unif_index(nn, 1)[0])
            y[i] = x[j] + 1
            nn -= 1
            x[j] = x[nn]
        return y


def r_unique(v: np.ndarray) -> np.ndarray:
    """R's unique(): first-occurrence order."""
    _, first = np.unique(v, return_index=True)
    return v[np.sort(first)]
