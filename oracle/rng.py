"""MATLAB ``rand`` stream for the oracle (test infrastructure only).

The reference never calls ``rng``; MATLAB's start-up generator is mt19937ar with seed 0, which
is ``init_genrand(5489)`` producing 53-bit doubles (``genrand_res53``).  The stream is global
state shared by ``AMG/mis_set.m:31,35`` and ``Hybrid_AMG.m:40,69`` and consumed in program
order, so it is a module-level object here too.
"""
import ctypes

import numpy as np

from . import _ck


class MatlabRand:
    """mt19937ar / genrand_res53 stream; ``rand(k)`` returns the next k doubles."""

    def __init__(self, seed=5489):
        self._buf = ctypes.create_string_buffer(_ck.lib().orc_mt_sizeof())
        self.seed = seed
        self.drawn = 0
        _ck.lib().orc_mt_init(ctypes.addressof(self._buf), seed)

    def reset(self, seed=5489):
        self.seed = seed
        self.drawn = 0
        _ck.lib().orc_mt_init(ctypes.addressof(self._buf), seed)

    def rand(self, count):
        count = int(count)
        out = np.empty(count, dtype=np.float64)
        if count:
            _ck.lib().orc_mt_rand(ctypes.addressof(self._buf), count, out)
        self.drawn += count
        return out


GLOBAL_STREAM = MatlabRand()


def rand(count):
    return GLOBAL_STREAM.rand(count)


def rng_reset(seed=5489):
    GLOBAL_STREAM.reset(seed)
